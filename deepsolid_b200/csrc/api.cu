// C ABI of deepsolid_b200: context, parameter upload, workspace, and the launch
// sequences of log psi / local energy / Metropolis step.  See include/deepsolid_b200.h.
#include "../../include/deepsolid_b200.h"
#include "kernels.cuh"
#include "ozaki.cuh"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>

static thread_local char g_err[1024] = "";

void ds_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* ds_last_error(void) { return g_err; }
extern "C" int ds_version(void) { return 100; }

namespace {

struct DevBuf {
    double* p = nullptr;
    size_t n = 0;   // doubles
};

struct Region {
    const char* name;
    double* p;
    size_t n;
};

struct Workspace {
    double* base = nullptr;
    size_t cap = 0;     // doubles
    size_t used = 0;
    std::vector<Region> regions;
    void reset() { used = 0; regions.clear(); }
    double* take(const char* name, size_t n) {
        n = (n + 31) & ~size_t(31);     // 256-byte granules keep every buffer 16B aligned for cp.async
        double* p = base ? base + used : nullptr;
        used += n;
        regions.push_back(Region{name, p, n});
        return p;
    }
};

struct ProfEvent {
    cudaEvent_t a, b;
    double flops;
};

}  // namespace

struct ds_ctx {
    int device = 0;
    DsSys sys;
    EwaldDev ew;
    int npar[2] = {0, 0}, npar_max = 0, off_s[2] = {0, 0}, n_s[2] = {0, 0};
    int nbuf = 2;
    bool params_set = false;
    std::vector<void*> owned;           // device allocations freed at destroy
    // parameters (device)
    double* B_am[DS_MAX_LAYERS] = {};   // [(C + 2Pl) x H] own + pair-mean rows
    double* B_g[DS_MAX_LAYERS] = {};    // [2C x H] spin-mean rows
    double* bias1[DS_MAX_LAYERS] = {};  // [H]
    double* Wp[DS_MAX_LAYERS] = {};     // pair-stream weights
    double* bp[DS_MAX_LAYERS] = {};
    double* Worb[2] = {};               // [H x 2 npar_s] (use_last: [K1 x 2 npar_s], own + pair-mean rows), columns interleaved (re, im)
    double* WorbG[2] = {};              // use_last: [2H x 2 npar_s] spin-mean rows of the orbital projection
    double* borb[2] = {};               // [2 npar_s] orbital bias (bias_orbitals=True), interleaved like Worb; else null
    double* gborb[2] = {};
    bool bias_orb = false;
    // int8 digits of the transposed weights for the tcgen05 path (ozaki.cuh): [N][OZ_S][K] + scales [N]
    signed char* Wd_am[DS_MAX_LAYERS] = {};
    double* sb_am[DS_MAX_LAYERS] = {};
    signed char* Wd_g[DS_MAX_LAYERS] = {};   // spin-mean block B_g (layers >= 1, K = 2H)
    double* sb_g[DS_MAX_LAYERS] = {};
    bool use_i8_means = true;                // shared-mean GEMM GOUT = GIN.B_g on the int8 path as well
    bool use_i8_value = true;                // value / Laplacian rows of layers >= 1 on the int8 path as well
    signed char* Wd_orb[2] = {};
    double* sb_orb[2] = {};
    bool use_i8 = true;                 // Jacobian-sweep GEMMs on tcgen05 (false: fp64 DMMA kernels)
    bool i8_ok = false;                 // stream widths are multiples of the tcgen05 K block
    bool use_l0_kernel = true;          // layer-0 Jacobian rows by the streaming kernel (false: DMMA GEMM)
    bool use_slice_means = true;        // digits + spin-channel means of a layer's Jacobian rows in one pass
    // layers >= 1: the GEMM epilogue writes the next operand's digits (OZ_JACD) and no fp64 Jacobian reaches HBM.  Off by
    // default: in-call A/B on B200 (profiles/r2_fused_ab.log) 7.26-7.30 k vs 7.38-7.43 k local energies/s -- digit formation
    // costs the same issue slots in the epilogue as in the separate HBM-bound pass; DS_FUSED_DIGITS=1 selects it
    bool use_fused_digits = false;
    double* env_pi[2] = {};
    double* env_sigma[2] = {};
    double* klist[2] = {};
    // transposed weights and gradient accumulators of the parameter-gradient path (ds_logpsi_vjp)
    double* B_amT[DS_MAX_LAYERS] = {};  // [H x (C + 2Pl)]
    double* B_gT[DS_MAX_LAYERS] = {};   // [H x 2C]
    double* WorbT[2] = {};              // [2 npar_s x H]  (use_last: [2 npar_s x K1])
    double* WorbGT[2] = {};             // use_last: [2 npar_s x 2H]
    double* gWorbG[2] = {};             // use_last: gradient of the spin-mean rows [2H x 2 npar_s]
    bool transposes_ready = false;
    double* gB_am[DS_MAX_LAYERS] = {};
    double* gB_g[DS_MAX_LAYERS] = {};
    double* gbias1[DS_MAX_LAYERS] = {};
    double* gWp[DS_MAX_LAYERS] = {};
    double* gbp[DS_MAX_LAYERS] = {};
    double* gWorb[2] = {};
    double* genv_pi[2] = {};
    double* genv_sigma[2] = {};
    // workspace
    Workspace ws;
    // bytes: 169 walkers per chunk at 54 electrons -> 4096 walkers run as 25 chunks of 164.  The step is power-capped on
    // B200 (sw_power_cap at ~1 kW) and the chunk length changes the clocks the governor settles at: 25-26 chunks run at
    // 1.77-1.84 GHz, 17 chunks (24 GiB) at 1.71-1.74 GHz, 27 at 1.69 GHz (profiles/r2_chunk_sweep.log: +2.5-4.5 % on two boxes)
    size_t ws_limit = size_t(16) << 30;
    // mcmc scratch
    DevBuf mc_x2, mc_lp, mc_lp2, host_stage;
    // instrumentation
    long long launches = 0;
    bool prof_on = false;
    std::vector<ProfEvent> prof;
    cudaEvent_t tot_a = nullptr, tot_b = nullptr;
    double tot_ms = 0.0;
    int dbg_stop_layer = -1;
    // optional outputs of the sweep: d log|psi| / dx and d phase / dx of the current chunk (ds_logpsi_grad_x)
    double* gx_abs = nullptr;
    double* gx_phase = nullptr;
    const double* cot_mats = nullptr;       // cotangent of the orbital matrices of the current chunk (ds_orbitals_vjp)
    // Kronecker-factor statistics (ds_kfac_factors): raw accumulators, allocated on first use
    bool fact_on = false;
    double* fA1[DS_MAX_LAYERS] = {};        // [(3C+2Pl+2)^2] Gram matrix of the one-electron layer inputs (x, 1, 0)
    double* fG1[DS_MAX_LAYERS] = {};        // [H^2] Gram matrix of the pre-activation cotangents, both passes
    double* fAo[2] = {};                    // [(H+2)^2]
    double* fGo[2] = {};                    // [(2 npar_s)^2], (re, im) interleaved
    double* fAp[DS_MAX_LAYERS] = {};        // [32*32] pair layers
    double* fAps[DS_MAX_LAYERS] = {};       // [32]
    double* fGp[DS_MAX_LAYERS] = {};        // [32*32]
    double* fenv_pi[2][2] = {};             // [pass][spin] gradients of sum_w log|psi_w| (pass 0) / sum_w phase_w (pass 1)
    double* fenv_sigma[2][2] = {};
    double* fones = nullptr;                // [fones_n] ones then [fones_n] zeros: unit cotangents
    size_t fones_n = 0;
    // layout of the last local-energy chunk (for ds_debug_buffer)
    std::vector<Region> last_regions;
    double* scratch8 = nullptr;             // packed statistics of ds_stats_allreduce
    long long last_chunk = 0;               // walkers per chunk of the last batched call
};

double* ds_ctx_scratch8(ds_ctx* c) {
    if (!c->scratch8) {
        void* p = nullptr;
        if (cudaMalloc(&p, 8 * sizeof(double)) != cudaSuccess) return nullptr;
        c->owned.push_back(p);
        c->scratch8 = (double*)p;
    }
    return c->scratch8;
}
void ds_ctx_count_launches(ds_ctx* c, int n) { c->launches += n; }

namespace {

// Slater matrix blocks: two n_s x n_s matrices per determinant, or one N x N matrix with full_det
inline int nblk_of(const ds_ctx* c) { return ds_nblk(c->sys.d); }
inline size_t blk_n(const ds_ctx* c, int b) { return (size_t)ds_blk_n(c->sys.d, b); }
inline size_t mats_per_walker(const ds_ctx* c) {
    size_t t = 0;
    for (int b = 0; b < nblk_of(c); ++b) t += (size_t)c->sys.d.D * blk_n(c, b) * blk_n(c, b) * 2;
    return t;
}

template <typename T>
int dev_alloc(ds_ctx* c, T** out, size_t count) {
    void* p = nullptr;
    DS_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    c->owned.push_back(p);
    *out = (T*)p;
    return 0;
}

int upload(ds_ctx* c, double** out, const double* host, size_t count) {
    if (int rc = dev_alloc(c, out, count)) return rc;
    DS_CUDA_CHECK(cudaMemcpy(*out, host, count * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

void inv3(const double* a, double* o) {
    double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    double id = 1.0 / det;
    o[0] = (a[4] * a[8] - a[5] * a[7]) * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = (a[5] * a[6] - a[3] * a[8]) * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = (a[3] * a[7] - a[4] * a[6]) * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

void fill_lattice(DsLattice& L, const double* lat, const double* AV, const double* BV) {
    memcpy(L.lat, lat, 9 * sizeof(double));
    inv3(lat, L.inv);
    memcpy(L.AV, AV, 9 * sizeof(double));
    memcpy(L.BV, BV, 9 * sizeof(double));
    for (int l = 0; l < 3; ++l) {
        L.an2[l] = AV[l * 3] * AV[l * 3] + AV[l * 3 + 1] * AV[l * 3 + 1] + AV[l * 3 + 2] * AV[l * 3 + 2];
        for (int m = 0; m < 3; ++m)
            L.metric[l * 3 + m] = AV[l * 3] * AV[m * 3] + AV[l * 3 + 1] * AV[m * 3 + 1] + AV[l * 3 + 2] * AV[m * 3 + 2];
    }
}

struct Guard {
    int prev = -1;
    explicit Guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
    ~Guard() { int cur; cudaGetDevice(&cur); if (cur != prev && prev >= 0) cudaSetDevice(prev); }
};

int gemm(ds_ctx* c, const GemmParams& p, int mode, bool res, cudaStream_t st, bool profile_as_jac = false) {
    ProfEvent ev{};
    bool rec = c->prof_on && profile_as_jac;
    if (rec) {
        DS_CUDA_CHECK(cudaEventCreate(&ev.a));
        DS_CUDA_CHECK(cudaEventCreate(&ev.b));
        DS_CUDA_CHECK(cudaEventRecord(ev.a, st));
    }
    int rc = ds_launch_gemm(p, mode, res, st);
    if (rc) return rc;
    c->launches++;
    if (rec) {
        DS_CUDA_CHECK(cudaEventRecord(ev.b, st));
        ev.flops = 2.0 * (double)p.M * p.N * p.K;
        c->prof.push_back(ev);
    }
    return 0;
}

// Jacobian-sweep GEMM on the tcgen05 path: digits of the fp64 operand rows, then the int8 GEMM.
// Profiled like the DMMA kernels it replaces (flops = fp64-equivalent 2 M N K).
struct ProfScope {
    ds_ctx* c; cudaStream_t st; ProfEvent ev{}; bool rec;
    ProfScope(ds_ctx* c_, cudaStream_t st_, bool on, double flops) : c(c_), st(st_), rec(c_->prof_on && on) {
        if (rec) { cudaEventCreate(&ev.a); cudaEventCreate(&ev.b); cudaEventRecord(ev.a, st); ev.flops = flops; }
    }
    ~ProfScope() { if (rec) { cudaEventRecord(ev.b, st); c->prof.push_back(ev); } }
};

// per-walker workspace size in doubles
struct Layout {
    bool lap;
    int Wc;
    double *A0V, *A0L, *A0J;
    double *J[DS_MAX_LAYERS], *V[DS_MAX_LAYERS], *Lp[DS_MAX_LAYERS];
    double *T, *S, *GIN, *GOUT, *RAE, *ETAB, *YV, *YL, *YOWN;
    double *GOO[2];         // use_last_layer: shared spin-mean contribution to the raw orbital outputs, per spin [Wc*NDg x 2 npar_max]
    double *MAT[2], *LAPM[2], *DA[2];
    double *LOGDET, *TAU, *TRSQ, *TRLAP;
    double *AD, *SA;        // int8 digits of the current Jacobian operand (as bytes) and its row scales
    double *AD2, *SA2;      // second digit buffer: with fused digits the operands of consecutive layers alternate
    double *PMJ[DS_MAX_LAYERS];   // fused digits: compact pair-mean Jacobian rows [rows][2P] of layers >= 2
    double *SP;             // fused digits: per-8-row partial sums of zJ^2
    bool fused;
    double *GD, *GS;        // digits and scales of the spin-mean rows GIN (shared-mean GEMM on the int8 path)
    double *VD, *VS;        // digits and scales of the value / Laplacian rows of the current layer (per electron)
    // parameter-gradient path: per-layer activations kept by the forward, cotangent buffers of the reverse sweep
    bool grad;
    double *Tl[DS_MAX_LAYERS], *GINV[DS_MAX_LAYERS];
    double *XINV[2], *GYs[2], *GH[2], *GZ, *GZS, *GA, *GG, *GPM[DS_MAX_LAYERS];
    double *GYS[2];         // use_last: per-walker sums over the electrons of a spin channel of the orbital-output cotangents
    double *XF;             // factor statistics: explicit rows of a layer's input
    double *XF2;            // use_last: the spin-channel rows of the orbital projection's input, gathered from XF
};

// The fused-digit sweep (OZ_JACD) needs exactly two 128-channel blocks and at most 64 pair-mean columns.
inline bool fused_digits_on(const ds_ctx* c, bool lap) {
    const DsDims& d = c->sys.d;
    return lap && c->use_i8 && c->i8_ok && c->use_fused_digits && c->use_slice_means && d.H == 2 * OZ_TM && 2 * d.P <= 64 &&
           d.L >= 2 && c->dbg_stop_layer < 0 && !d.use_last;
}

void carve(ds_ctx* c, Workspace& ws, Layout& L, int Wc, bool lap, bool grad = false) {
    const DsDims& d = c->sys.d;
    const size_t W = (size_t)Wc, N = d.N;
    ws.reset();
    L.lap = lap; L.Wc = Wc; L.grad = grad;
    L.A0V = ws.take("A0V", W * N * d.K0);
    L.A0L = lap ? ws.take("A0L", W * N * d.K0) : nullptr;
    L.A0J = lap ? ws.take("A0J", W * N * d.NDp * d.K0) : nullptr;
    static const char* jn[] = {"J0", "J1", "J2", "J3"};
    static const char* vn[] = {"V0", "V1", "V2", "V3"};
    static const char* ln[] = {"L0", "L1", "L2", "L3"};
    const int nbuf = grad ? d.L : c->nbuf;          // the reverse sweep needs every layer's input
    L.fused = fused_digits_on(c, lap);
    for (int b = 0; b < nbuf; ++b) {
        L.V[b] = ws.take(vn[b], W * N * d.K1);
        L.Lp[b] = lap ? ws.take(ln[b], W * N * d.K1) : nullptr;
        // fused digits: only the layer-0 output exists as fp64 Jacobian rows
        L.J[b] = (lap && !(L.fused && b > 0)) ? ws.take(jn[b], W * N * d.NDp * d.K1) : nullptr;
    }
    L.T = ws.take("T", W * N * d.H);
    L.S = lap ? ws.take("S", W * N * d.H) : nullptr;
    const size_t cmax = (size_t)std::max(d.C0, d.H);
    L.GIN = ws.take("GIN", W * d.NDg * 2 * cmax);
    L.GOUT = ws.take("GOUT", W * d.NDg * d.H);
    for (int s = 0; s < 2; ++s)
        L.GOO[s] = d.use_last ? ws.take(s ? "GOO1" : "GOO0", W * (lap ? d.NDg : 1) * 2 * (size_t)c->npar_max) : nullptr;
    L.RAE = ws.take("RAE", W * N * d.A * DS_RAE_STRIDE);
    const size_t npm = c->npar_max;
    L.ETAB = ws.take("ETAB", W * N * 5 * npm * 2);
    L.YV = ws.take("YV", W * N * npm * 2);
    L.YL = lap ? ws.take("YL", W * N * npm * 2) : nullptr;
    L.YOWN = lap ? ws.take("YOWN", W * N * 3 * npm * 2) : nullptr;
    static const char* mn[] = {"MAT0", "MAT1"};
    static const char* lmn[] = {"LAPM0", "LAPM1"};
    static const char* dn[] = {"DA0", "DA1"};
    for (int s = 0; s < 2; ++s) {
        L.MAT[s] = L.LAPM[s] = L.DA[s] = nullptr;
        if (s >= nblk_of(c)) continue;
        size_t ns = blk_n(c, s);
        L.MAT[s] = ws.take(mn[s], W * d.D * ns * ns * 2);
        L.LAPM[s] = lap ? ws.take(lmn[s], W * d.D * ns * ns * 2) : nullptr;
        L.DA[s] = lap ? ws.take(dn[s], W * d.D * d.NDp * ns * ns * 2) : nullptr;
    }
    L.LOGDET = ws.take("LOGDET", W * 2 * d.D * 3);
    L.TAU = lap ? ws.take("TAU", W * 2 * d.D * d.NDp * 2) : nullptr;
    L.TRSQ = lap ? ws.take("TRSQ", W * 2 * d.D * 2) : nullptr;
    L.TRLAP = lap ? ws.take("TRLAP", W * 2 * d.D * 2) : nullptr;
    const bool i8 = lap && c->use_i8 && c->i8_ok;
    // (row pitch of the row-contiguous digit layout the fused-digit epilogue writes: rows rounded up to the 64-row tile)
    const size_t Rp = (W * N * d.NDp + 63) / 64 * 64;
    L.AD = i8 ? ws.take("AD", (Rp * OZ_S * d.K1 + 7) / 8) : nullptr;
    L.SA = i8 ? ws.take("SA", Rp) : nullptr;
    L.AD2 = L.fused ? ws.take("AD2", (Rp * OZ_S * d.K1 + 7) / 8) : nullptr;
    L.SA2 = L.fused ? ws.take("SA2", Rp) : nullptr;
    L.SP = i8 ? ws.take("SP", W * N * (d.NDp / 8) * d.H) : nullptr;
    static const char* pmn[] = {"PMJ0", "PMJ1", "PMJ2", "PMJ3"};
    for (int l = 0; l < DS_MAX_LAYERS; ++l)
        L.PMJ[l] = (L.fused && l >= 2 && l < d.L) ? ws.take(pmn[l], W * N * d.NDp * 2 * d.P) : nullptr;
    L.GD = i8 ? ws.take("GD", (W * d.NDg * OZ_S * 2 * d.H + 7) / 8) : nullptr;
    L.GS = i8 ? ws.take("GS", W * d.NDg) : nullptr;
    const bool i8v = c->use_i8 && c->i8_ok && c->use_i8_value;
    L.VD = i8v ? ws.take("VD", (W * N * OZ_S * d.K1 + 7) / 8) : nullptr;
    L.VS = i8v ? ws.take("VS", W * N) : nullptr;
    if (grad) {
        static const char* tn[] = {"T0", "T1", "T2", "T3"};
        static const char* gn[] = {"GINV0", "GINV1", "GINV2", "GINV3"};
        static const char* pn[] = {"GPM0", "GPM1", "GPM2", "GPM3"};
        for (int l = 0; l < d.L; ++l) {
            L.Tl[l] = ws.take(tn[l], W * N * d.H);
            L.GINV[l] = ws.take(gn[l], W * 2 * (size_t)((l == 0) ? d.C0 : d.H));
            L.GPM[l] = (l > 0) ? ws.take(pn[l], W * N * 2 * d.P) : nullptr;
        }
        if (d.use_last) {       // the orbital projection acts like one more (linear) layer on the symmetric features
            L.GINV[d.L] = ws.take(gn[d.L], W * 2 * (size_t)d.H);
            L.GPM[d.L] = ws.take(pn[d.L], W * N * 2 * d.P);
        }
        for (int s = 0; s < 2; ++s) {
            size_t ns = c->n_s[s];
            L.XINV[s] = (s < nblk_of(c)) ? ws.take(s ? "XINV1" : "XINV0", W * d.D * blk_n(c, s) * blk_n(c, s) * 2) : nullptr;
            L.GYs[s] = ws.take(s ? "GY1" : "GY0", W * ns * 2 * c->npar[s]);
            L.GYS[s] = d.use_last ? ws.take(s ? "GYS1" : "GYS0", W * 2 * (size_t)c->npar[s]) : nullptr;
        }
        L.GH[0] = ws.take("GH0", W * N * d.H);
        L.GH[1] = ws.take("GH1", W * N * d.H);
        L.GZ = ws.take("GZ", W * N * d.H);
        L.GZS = ws.take("GZS", W * d.H);
        L.GA = ws.take("GA", W * N * d.K1);
        L.GG = ws.take("GG", W * 2 * d.H);
        L.XF = c->fact_on ? ws.take("XF", W * N * (3 * cmax + 2 * (size_t)std::max(d.P, d.F) + 2)) : nullptr;
        L.XF2 = (c->fact_on && d.use_last) ? ws.take("XF2", W * N * (3 * cmax + 2 * (size_t)std::max(d.P, d.F) + 2)) : nullptr;
    }
}

int plan_chunk(ds_ctx* c, long long batch, bool lap, int* Wc_out, bool grad = false) {
    Workspace probe;                      // base == nullptr: sizes only
    Layout L;
    carve(c, probe, L, 1, lap, grad);
    size_t per_walker = probe.used + 64;  // doubles (granule slack)
    size_t limit = c->ws_limit / sizeof(double);
    long long Wmax = (long long)(limit / per_walker);
    if (Wmax < 1) {
        ds_set_error("workspace limit %zu bytes is below the %zu bytes one walker needs", c->ws_limit,
                     per_walker * sizeof(double));
        return DS_ERR_NOMEM;
    }
    // keep the row counts of the Jacobian GEMM within int-friendly grid sizes
    Wmax = std::min<long long>(Wmax, 1 << 15);
    if (const char* ev = getenv("DS_CHUNK_WALKERS")) { const long long v = atoll(ev); if (v >= 1) Wmax = std::min(Wmax, v); }   // tuning sweeps
    for (;;) {
        // equal chunks: 512 walkers at a 254-walker limit run as 3 x 171, not 254 + 254 + 4
        const long long n_chunks = (batch + Wmax - 1) / Wmax;
        const long long Wc = (batch + n_chunks - 1) / n_chunks;
        *Wc_out = (int)Wc;
        carve(c, probe, L, (int)Wc, lap, grad);
        const size_t need = probe.used;
        if (need <= c->ws.cap) return 0;
        if (c->ws.base) { DS_CUDA_CHECK(cudaFree(c->ws.base)); c->ws.base = nullptr; c->ws.cap = 0; }
        cudaError_t e = cudaMalloc((void**)&c->ws.base, need * sizeof(double));
        if (e == cudaSuccess) { c->ws.cap = need; return 0; }
        (void)cudaGetLastError();         // clear the sticky allocation error
        c->ws.base = nullptr;
        if (Wc <= 1) {
            ds_set_error("cannot allocate %zu bytes of workspace (one walker): %s", need * sizeof(double), cudaGetErrorString(e));
            return DS_ERR_NOMEM;
        }
        // the device is shared with the caller's allocator (torch, NCCL): retry with half the chunk
        Wmax = std::max<long long>(1, Wc / 2);
    }
}

// One chunk of walkers through the network.  lap=false: log psi only.
int grad_sweep(ds_ctx* c, Layout& Lo, const FeatParams& fp, SlaterBufs& sb, int Wc, const double* cot_abs,
               const double* cot_phase, cudaStream_t st, int fact_pass = -1);
int fact_sweep(ds_ctx* c, Layout& Lo, const FeatParams& fp, SlaterBufs& sb, int Wc, cudaStream_t st);

int run_chunk(ds_ctx* c, const double* X, int Wc, bool lap, double* log_abs, double* phase, double* ke_re,
              double* ke_im, double* mats_out, cudaStream_t st, const double* cot_abs = nullptr,
              const double* cot_phase = nullptr) {
    const DsSys& sys = c->sys;
    const DsDims& d = sys.d;
    Layout Lo;
    const bool grad = cot_abs != nullptr;
    carve(c, c->ws, Lo, Wc, lap, grad);
    c->last_regions = c->ws.regions;
    const int N = d.N, H = d.H, L = d.L;

    // buffer of the inputs of layer l >= 1 is (l-1); the last layer writes into buffer `outb(L-1)`
    auto inb = [&](int l) { return l - 1; };
    auto outb = [&](int l) { return (grad || d.use_last || l + 1 < L) ? l : ((L - 1 >= 2) ? 0 : 1); };

    FeatParams fp{};
    fp.X = X; fp.A0V = Lo.A0V; fp.A0L = Lo.A0L; fp.A0J = Lo.A0J; fp.RAE = Lo.RAE;
    const bool fused = Lo.fused;
    for (int l = 1; l < L; ++l) {
        fp.AV[l] = Lo.V[inb(l)]; fp.AL[l] = Lo.Lp[inb(l)];
        if (fused && l >= 2) { fp.AJ[l] = Lo.PMJ[l]; fp.ldj[l] = 2 * d.P; fp.joff[l] = 0; }      // compact: consumed by OZ_JACD
        else { fp.AJ[l] = Lo.J[inb(l)]; fp.ldj[l] = d.K1; fp.joff[l] = H; }
    }
    if (d.use_last) {       // level L of the pair stream feeds the orbital projection: pair-mean columns of the last layer's output
        fp.AV[L] = Lo.V[L - 1]; fp.AL[L] = Lo.Lp[L - 1]; fp.AJ[L] = Lo.J[L - 1]; fp.ldj[L] = d.K1; fp.joff[L] = H;
    }
    for (int l = 0; l < ds_pair_levels(d) - 1; ++l) { fp.Wp[l] = c->Wp[l]; fp.bp[l] = c->bp[l]; }
    if (int rc = ds_launch_features(sys, fp, Wc, lap, st)) return rc;
    c->launches++;

    const double *hV = nullptr, *hJ = nullptr, *hL = nullptr;
    for (int l = 0; l < L; ++l) {
        const int C = (l == 0) ? d.C0 : H;
        const int K = (l == 0) ? d.K0 : d.K1;
        const double* AV = (l == 0) ? Lo.A0V : Lo.V[inb(l)];
        const double* AL = (l == 0) ? Lo.A0L : Lo.Lp[inb(l)];
        const double* AJ = (l == 0) ? Lo.A0J : Lo.J[inb(l)];
        double* OV = Lo.V[outb(l)];
        double* OL = Lo.Lp[outb(l)];
        double* OJ = Lo.J[outb(l)];
        const bool res = (C == H);
        // fused digits (layers >= 1): the operand digits of consecutive layers alternate between AD and AD2; layer 1's
        // come from the fp64 rows the layer-0 kernel wrote, every later layer's from the previous GEMM's epilogue
        const bool fl = fused && l > 0;
        signed char* dig_in = reinterpret_cast<signed char*>(((l - 1) & 1) ? Lo.AD2 : Lo.AD);
        double* sa_in = ((l - 1) & 1) ? Lo.SA2 : Lo.SA;
        signed char* dig_out = reinterpret_cast<signed char*>((l & 1) ? Lo.AD2 : Lo.AD);
        double* sa_out = (l & 1) ? Lo.SA2 : Lo.SA;
        const bool i8_layer = lap && c->use_i8 && c->i8_ok && l > 0;
        // (the tcgen05 kernels write per-8-row partial sums of zJ^2, reduced in a fixed order; only the fp64 DMMA
        //  fallback accumulates into S with atomics and needs it cleared)
        if (lap && !i8_layer) DS_CUDA_CHECK(cudaMemsetAsync(Lo.S, 0, (size_t)Wc * N * H * sizeof(double), st));
        const bool fused_means = i8_layer && c->use_slice_means;
        const long long Rp = ((long long)Wc * N * d.NDp + 63) / 64 * 64;
        if (fl && l >= 2) {
            // spin-channel means of the Jacobian rows straight from their digits (no fp64 rows exist)
            if (int rc = ds_launch_means_digits(dig_in, sa_in, K, Rp, C, Wc, d.n_up, N, d.NDp, d.NDg, Lo.GIN, 2 * C, st)) return rc;
            c->launches++;
        } else if (fused_means) {
            // digits of the Jacobian rows and their spin-channel means in one pass over the fp64 rows
            if (int rc = ds_launch_slice_means(AJ, K, K, C, Wc, d.n_up, N, d.NDp, d.NDg, reinterpret_cast<signed char*>(Lo.AD),
                                               Lo.SA, Lo.GIN, 2 * C, st)) return rc;
            c->launches++;
        }
        if (int rc = ds_launch_means(d, Wc, C, AJ, K, AV, AL, K, Lo.GIN, 2 * C, lap, st, fused_means || fl)) return rc;
        c->launches++;
        {   // shared spin-mean contribution, once per walker and direction
            GemmParams g{};
            g.B = c->B_g[l]; g.ldb = H; g.N = H; g.K = 2 * C; g.rpg = 0;
            if (lap) {
                g.A = Lo.GIN; g.lda = 2 * C; g.M = (long long)Wc * d.NDg; g.C = Lo.GOUT; g.ldc = H;
            } else {    // only the value row d = NDp of every walker
                g.A = Lo.GIN + (size_t)d.NDp * 2 * C; g.lda = d.NDg * 2 * C; g.M = Wc;
                g.C = Lo.GOUT + (size_t)d.NDp * H; g.ldc = d.NDg * H;
            }
            if (lap && i8_layer && c->use_i8_means && c->Wd_g[l]) {
                // the same contraction as exact int8 slices (K = 2H): digits of the mean rows, then the tcgen05 GEMM;
                // AD/SA are free here (the layer's own digits are only formed after this point when not fused) --
                // with the fused digit+means pass they are already in use, so the mean digits go to a separate area
                const long long rows = (long long)Wc * d.NDg;
                if (int rc = ds_launch_slice_rows(Lo.GIN, 2 * C, rows, 2 * C, reinterpret_cast<signed char*>(Lo.GD), Lo.GS, st)) return rc;
                OzParams z{};
                z.Ad = reinterpret_cast<signed char*>(Lo.GD); z.sa = Lo.GS; z.rpg = rows; z.gstride = rows; z.goff = 0; z.n_groups = 1;
                z.Wd = c->Wd_g[l]; z.sb = c->sb_g[l]; z.N = H; z.K = 2 * C; z.C = Lo.GOUT; z.ldc = H;
                if (int rc = ds_launch_oz_gemm(z, OZ_PLAIN, false, st)) return rc;
                c->launches += 2;
            } else if (int rc = gemm(c, g, GEMM_PLAIN, false, st)) return rc;
        }
        if (grad)   // value rows of the spin means of this layer: operand of the mean-block weight gradient
            DS_CUDA_CHECK(cudaMemcpy2DAsync(Lo.GINV[l], (size_t)2 * C * sizeof(double), Lo.GIN + (size_t)d.NDp * 2 * C,
                                            (size_t)d.NDg * 2 * C * sizeof(double), (size_t)2 * C * sizeof(double), Wc,
                                            cudaMemcpyDeviceToDevice, st));
        GemmParams p{};
        p.B = c->B_am[l]; p.ldb = H; p.N = H; p.K = K; p.rpg = 0;
        p.G = Lo.GOUT; p.ldg = H; p.n_elec = N; p.NDp = d.NDp; p.NDg = d.NDg;
        p.T = grad ? Lo.Tl[l] : Lo.T; p.ldt = H; p.S = Lo.S; p.colbias = c->bias1[l];
        const bool i8_rows = Lo.VD && l > 0 && c->use_i8 && c->i8_ok;      // value / Laplacian rows on tcgen05 too
        auto oz_rows = [&](const double* Arows, double* Out, int mode) -> int {
            const long long rows = (long long)Wc * N;
            signed char* Vd = reinterpret_cast<signed char*>(Lo.VD);
            if (int rc = ds_launch_slice_rows(Arows, K, rows, K, Vd, Lo.VS, st)) return rc;
            OzParams o{};
            o.Ad = Vd; o.sa = Lo.VS; o.rpg = rows; o.gstride = rows; o.goff = 0; o.n_groups = 1;
            o.Wd = c->Wd_am[l]; o.sb = c->sb_am[l]; o.N = H; o.K = K;
            o.C = Out; o.ldc = d.K1; o.G = Lo.GOUT; o.ldg = H; o.n_elec = N; o.NDp = d.NDp; o.NDg = d.NDg;
            o.T = p.T; o.Tout = p.T; o.ldt = H; o.S = Lo.S; o.R = Arows; o.ldr = K; o.colbias = c->bias1[l];
            if (int rc = ds_launch_oz_gemm(o, mode, res, st)) return rc;
            c->launches += 2;
            return 0;
        };
        if (i8_rows) {
            if (int rc = oz_rows(AV, OV, OZ_VALUE)) return rc;
        } else {
            GemmParams v = p;
            v.A = AV; v.lda = K; v.M = (long long)Wc * N; v.C = OV; v.ldc = d.K1; v.R = AV; v.ldr = K;
            if (int rc = gemm(c, v, GEMM_VALUE, res, st)) return rc;
        }
        if (fl) {
            const long long rows = (long long)Wc * N * d.NDp;
            const bool last = (l == L - 1);
            {
                ProfScope ps(c, st, true, 2.0 * (double)Wc * N * d.ND * H * K);
                OzParams o{};
                o.Ad = dig_in; o.sa = sa_in; o.rpg = rows; o.gstride = rows; o.goff = 0; o.n_groups = 1;
                o.Wd = c->Wd_am[l]; o.sb = c->sb_am[l]; o.N = H; o.K = K;
                o.G = Lo.GOUT; o.ldg = H; o.n_elec = N; o.NDp = d.NDp; o.NDg = d.NDg;
                o.T = Lo.T; o.ldt = H; o.SP = Lo.SP;
                // layer 1 reads the K-major digits slice_means formed from the layer-0 kernel's fp64 rows (which are also
                // its residual); later layers read the row-contiguous digits the previous epilogue wrote
                o.bmn = (l >= 2) ? 1 : 0; o.Rp_in = Rp; o.R = AJ; o.ldr = K;
                o.Dout = dig_out; o.sa_out = sa_out; o.Rp_out = Rp;
                // the last layer feeds the orbital projection: own columns only
                o.Kout = last ? H : d.K1; o.PM = last ? nullptr : Lo.PMJ[l + 1]; o.npm = last ? 0 : 2 * d.P;
                if (int rc = ds_launch_oz_gemm(o, OZ_JACD, res, st)) return rc;
                c->launches++;
            }
            if (int rc = ds_launch_sp_reduce(Lo.SP, H, (long long)Wc * N, d.NDp / 8, H, Lo.S, st)) return rc;
            c->launches++;
        } else if (i8_layer) {
            const long long rows = (long long)Wc * N * d.NDp;
            signed char* Ad = reinterpret_cast<signed char*>(Lo.AD);
            if (!fused_means) {
                if (int rc = ds_launch_slice_rows(AJ, K, rows, K, Ad, Lo.SA, st)) return rc;
                c->launches++;
            }
            // the tcgen05 GEMM alone; algorithmic flops count the 3N real directions, not the NDp padded rows
            ProfScope ps(c, st, true, 2.0 * (double)Wc * N * d.ND * H * K);
            OzParams o{};
            o.Ad = Ad; o.sa = Lo.SA; o.rpg = rows; o.gstride = rows; o.goff = 0; o.n_groups = 1;
            o.Wd = c->Wd_am[l]; o.sb = c->sb_am[l]; o.N = H; o.K = K;
            o.C = OJ; o.ldc = d.K1; o.G = Lo.GOUT; o.ldg = H; o.n_elec = N; o.NDp = d.NDp; o.NDg = d.NDg;
            o.T = Lo.T; o.ldt = H; o.S = Lo.S; o.SP = Lo.SP; o.R = AJ; o.ldr = K;
            if (int rc = ds_launch_oz_gemm(o, OZ_JAC, res, st)) return rc;
            if (int rc = ds_launch_sp_reduce(Lo.SP, H, (long long)Wc * N, d.NDp / 8, H, Lo.S, st)) return rc;
            c->launches += 2;
        } else if (lap && l == 0 && !res && c->use_l0_kernel) {
            if (int rc = ds_launch_l0_jac(d, Wc, AJ, c->B_am[0], Lo.GOUT, H, Lo.T, H, Lo.S, OJ, d.K1, st)) return rc;
            c->launches++;
        } else if (lap) {
            GemmParams j = p;
            j.A = AJ; j.lda = K; j.M = (long long)Wc * N * d.NDp; j.C = OJ; j.ldc = d.K1; j.R = AJ; j.ldr = K;
            if (int rc = gemm(c, j, GEMM_JAC, res, st, /*profile*/ l > 0)) return rc;
        }
        if (lap && i8_rows) {
            if (int rc = oz_rows(AL, OL, OZ_LAP)) return rc;
        } else if (lap) {
            GemmParams q = p;
            q.A = AL; q.lda = K; q.M = (long long)Wc * N; q.C = OL; q.ldc = d.K1; q.R = AL; q.ldr = K;
            if (int rc = gemm(c, q, GEMM_LAP, res, st)) return rc;
        }
        hV = OV; hJ = OJ; hL = OL;
        if (c->dbg_stop_layer == l) return 0;
    }

    // ---- orbitals --------------------------------------------------------
    SlaterBufs sb{};
    sb.X = X; sb.RAE = Lo.RAE; sb.ETAB = Lo.ETAB; sb.YV = Lo.YV; sb.YL = Lo.YL; sb.YOWN = Lo.YOWN;
    for (int s = 0; s < 2; ++s) {
        sb.MAT[s] = Lo.MAT[s]; sb.LAPM[s] = Lo.LAPM[s]; sb.DA[s] = Lo.DA[s];
        sb.env_pi[s] = c->env_pi[s]; sb.env_sigma[s] = c->env_sigma[s]; sb.klist[s] = c->klist[s];
    }
    sb.LOGDET = Lo.LOGDET; sb.TAU = Lo.TAU; sb.TRSQ = Lo.TRSQ; sb.TRLAP = Lo.TRLAP;
    if (int rc = ds_launch_etab(sys, sb, Wc, c->npar_max, lap, st)) return rc;
    c->launches++;
    const bool i8 = lap && c->use_i8 && c->i8_ok;
    const long long jrows = (long long)Wc * N * d.NDp;
    // digits of the last layer's Jacobian rows (own columns), shared by both spins: written by the last layer's GEMM
    // epilogue on the fused path, else formed here from the fp64 rows
    signed char* orb_dig = reinterpret_cast<signed char*>((fused && ((L - 1) & 1)) ? Lo.AD2 : Lo.AD);
    double* orb_sa = (fused && ((L - 1) & 1)) ? Lo.SA2 : Lo.SA;
    // use_last_layer: the projection takes [own | pair means] per electron (K1 columns) plus the spin means of the last
    // layer, whose product with the mean rows of the weights is shared by the electrons of a walker (GOO)
    const int Korb = d.use_last ? d.K1 : H;
    if (d.use_last) {
        if (i8) {       // digits of [own | pair-mean] Jacobian rows and their spin-channel means in one pass
            if (int rc = ds_launch_slice_means(hJ, d.K1, d.K1, H, Wc, d.n_up, N, d.NDp, d.NDg, orb_dig, orb_sa, Lo.GIN, 2 * H, st)) return rc;
            c->launches++;
        }
        if (int rc = ds_launch_means(d, Wc, H, hJ, d.K1, hV, hL, d.K1, Lo.GIN, 2 * H, lap, st, i8)) return rc;
        c->launches++;
        if (grad)   // value rows of the spin means: operand of the gradient of the mean rows of the orbital weights
            DS_CUDA_CHECK(cudaMemcpy2DAsync(Lo.GINV[L], (size_t)2 * H * sizeof(double), Lo.GIN + (size_t)d.NDp * 2 * H,
                                            (size_t)d.NDg * 2 * H * sizeof(double), (size_t)2 * H * sizeof(double), Wc,
                                            cudaMemcpyDeviceToDevice, st));
        for (int s = 0; s < 2; ++s) {
            if (c->n_s[s] == 0) continue;
            GemmParams g{};
            g.B = c->WorbG[s]; g.ldb = 2 * c->npar[s]; g.N = 2 * c->npar[s]; g.K = 2 * H; g.rpg = 0;
            g.C = Lo.GOO[s]; g.ldc = 2 * c->npar_max;
            if (lap) { g.A = Lo.GIN; g.lda = 2 * H; g.M = (long long)Wc * d.NDg; }
            else { g.A = Lo.GIN + (size_t)d.NDp * 2 * H; g.lda = d.NDg * 2 * H; g.M = Wc; }   // the value row of every walker
            if (int rc = gemm(c, g, GEMM_PLAIN, false, st)) return rc;
        }
    } else if (i8 && !fused) {
        if (int rc = ds_launch_slice_rows(hJ, d.K1, jrows, H, orb_dig, orb_sa, st)) return rc;
        c->launches++;
    }
    for (int s = 0; s < 2; ++s) {
        const int ns = c->n_s[s];
        GemmParams o{};
        o.B = c->Worb[s]; o.ldb = 2 * c->npar[s]; o.N = 2 * c->npar[s]; o.K = Korb;
        o.lda = d.K1; o.cmap = 1; o.ldc = 2 * c->npar_max;
        o.rpg = ns; o.gstride = N; o.goff = c->off_s[s];
        o.M = (long long)Wc * ns;
        o.A = hV; o.C = Lo.YV;
        o.colbias = c->bias_orb ? c->borb[s] : nullptr;     // the bias enters the value only (network.py:538-541)
        if (int rc = gemm(c, o, GEMM_PLAIN, false, st)) return rc;
        o.colbias = nullptr;
        if (lap) {
            o.A = hL; o.C = Lo.YL;
            if (int rc = gemm(c, o, GEMM_PLAIN, false, st)) return rc;
        }
        if (i8) {
            ProfScope ps(c, st, true, 2.0 * (double)Wc * ns * d.ND * Korb * 2.0 * c->npar[s]);
            OzParams z{};
            z.Ad = orb_dig; z.sa = orb_sa;
            z.bmn = fused ? 1 : 0; z.Rp_in = ((long long)Wc * N * d.NDp + 63) / 64 * 64;
            z.rpg = (long long)ns * d.NDp; z.gstride = (long long)N * d.NDp; z.goff = (long long)c->off_s[s] * d.NDp;
            z.n_groups = Wc;
            z.Wd = c->Wd_orb[s]; z.sb = c->sb_orb[s]; z.N = 2 * c->npar[s]; z.K = Korb;
            z.n_elec = N; z.NDp = d.NDp; z.etab = Lo.ETAB; z.npar_max = c->npar_max;
            z.n_s = ns; z.off_s = c->off_s[s]; z.n_det = d.D; z.DA = Lo.DA[d.full_det ? 0 : s]; z.YOWN = Lo.YOWN;
            z.n_orb = ds_norb(d, s); z.n_rows_mat = ds_blk_n(d, d.full_det ? 0 : s); z.row0 = d.full_det ? c->off_s[s] : 0;
            if (int rc = ds_launch_oz_gemm(z, OZ_ORBJ, false, st)) return rc;
            c->launches++;
        } else if (lap) {
            GemmParams j = o;
            j.A = hJ; j.cmap = 0; j.C = nullptr;
            j.rpg = (long long)ns * d.NDp; j.gstride = (long long)N * d.NDp; j.goff = (long long)c->off_s[s] * d.NDp;
            j.M = (long long)Wc * ns * d.NDp;
            j.n_elec = N; j.NDp = d.NDp; j.etab = Lo.ETAB; j.npar_max = c->npar_max;
            j.n_s = ns; j.off_s = c->off_s[s]; j.n_det = d.D; j.DA = Lo.DA[d.full_det ? 0 : s]; j.YOWN = Lo.YOWN;
            j.n_orb = ds_norb(d, s); j.n_rows_mat = ds_blk_n(d, d.full_det ? 0 : s); j.row0 = d.full_det ? c->off_s[s] : 0;
            if (int rc = gemm(c, j, GEMM_ORBJ, false, st, /*profile*/ true)) return rc;
        }
    }
    if (d.use_last) {       // add the shared spin-mean contribution to the raw outputs (value, Laplacian, Jacobian rows)
        if (int rc = ds_launch_orb_mean_addend(sys, sb, Wc, c->npar_max, Lo.GOO[0], Lo.GOO[1], 2 * c->npar_max, lap, st)) return rc;
        c->launches++;
    }
    if (int rc = ds_launch_orb_assemble(sys, sb, Wc, c->npar_max, lap, st)) return rc;
    c->launches++;
    if (mats_out) {
        size_t off = 0;
        for (int s = 0; s < nblk_of(c); ++s) {
            size_t per = (size_t)d.D * blk_n(c, s) * blk_n(c, s) * 2;
            size_t tot = mats_per_walker(c);
            DS_CUDA_CHECK(cudaMemcpy2DAsync(mats_out + off, tot * sizeof(double), Lo.MAT[s], per * sizeof(double),
                                            per * sizeof(double), Wc, cudaMemcpyDeviceToDevice, st));
            off += per;
        }
        return 0;
    }
    if (grad) {
        FeatParams fpg = fp;
        if (c->fact_on) return fact_sweep(c, Lo, fpg, sb, Wc, st);
        return grad_sweep(c, Lo, fpg, sb, Wc, cot_abs, cot_phase, st);
    }
    if (int rc = ds_launch_det(sys, sb, Wc, lap, st)) return rc;
    c->launches++;
    if (int rc = ds_launch_combine(sys, sb, Wc, lap, log_abs, phase, ke_re, ke_im, st, lap ? c->gx_abs : nullptr,
                                   lap ? c->gx_phase : nullptr)) return rc;
    c->launches++;
    return 0;
}

// Reverse sweep of one chunk (forward activations are in the workspace): accumulates into the ctx gradient buffers.
// fact_pass >= 0: the sweep feeds the Kronecker-factor statistics instead of the parameter gradients (pass 0: unit
// cotangent on log|psi|, pass 1: on the phase): Gram matrices of the cotangents of every tagged layer output.
int grad_sweep(ds_ctx* c, Layout& Lo, const FeatParams& fp, SlaterBufs& sb, int Wc, const double* cot_abs,
               const double* cot_phase, cudaStream_t st, int fact_pass) {
    const bool fact = fact_pass >= 0;
    const DsSys& sys = c->sys;
    const DsDims& d = sys.d;
    const int N = d.N, H = d.H, L = d.L, P = d.P;
    GradBufs gb{};
    if (c->cot_mats) {      // eval_mats cotangent: no determinant involved
        gb.cot_mats = c->cot_mats;
        gb.cot_mats_stride = (long long)mats_per_walker(c);
    } else {
        sb.XINV[0] = Lo.XINV[0]; sb.XINV[1] = Lo.XINV[1];
        if (int rc = ds_launch_det_inverse(sys, sb, Wc, st)) return rc;
        c->launches++;
    }
    gb.cot_abs = cot_abs; gb.cot_phase = cot_phase;
    for (int s = 0; s < 2; ++s) {
        gb.GY[s] = Lo.GYs[s];
        gb.g_pi[s] = fact ? c->fenv_pi[fact_pass][s] : c->genv_pi[s];
        gb.g_sigma[s] = fact ? c->fenv_sigma[fact_pass][s] : c->genv_sigma[s];
    }
    if (int rc = ds_launch_orb_grad(sys, sb, gb, Wc, c->npar_max, st)) return rc;
    c->launches++;
    const double* hL = Lo.V[L - 1];                   // output of the last layer (own columns)
    double* GHcur = Lo.GH[0];
    double* GHnext = Lo.GH[1];
    const int Kam = d.use_last ? d.K1 : H;           // per-electron operand columns of the orbital projection
    if (d.use_last) DS_CUDA_CHECK(cudaMemsetAsync(Lo.GG, 0, (size_t)Wc * 2 * H * sizeof(double), st));
    for (int s = 0; s < 2; ++s) {
        const int ns = c->n_s[s], np2 = 2 * c->npar[s];
        GemmParams t{};                               // gWorb[s] += [hL | pair means]_s^T . GY_s
        t.A = hL; t.lda = d.K1; t.M = Kam; t.K = Wc * ns; t.rpg = ns; t.gstride = N; t.goff = c->off_s[s];
        t.B = Lo.GYs[s]; t.ldb = np2; t.N = np2; t.C = c->gWorb[s]; t.ldc = np2; t.accumulate = 1;
        if (fact) {                                   // fGo[s] += GY_s^T . GY_s
            t.A = Lo.GYs[s]; t.lda = np2; t.M = np2; t.rpg = 0; t.C = c->fGo[s];
        }
        if (ns > 0)
            if (int rc = gemm(c, t, GEMM_TN, false, st)) return rc;
        if (c->bias_orb && !fact) {
            if (int rc = ds_launch_colsum_add(Lo.GYs[s], np2, Wc * ns, np2, c->gborb[s], st)) return rc;
            c->launches++;
        }
        GemmParams g{};                               // cotangent of h_L (rows of spin s) = GY_s . Worb_s^T
        g.A = Lo.GYs[s]; g.lda = np2; g.M = (long long)Wc * ns; g.K = np2; g.no_amap = 1;
        g.B = c->WorbT[s]; g.ldb = Kam; g.N = Kam;
        // use_last: cotangent of [h_L | pair means of level L]; the spin-mean share is added by hin below
        g.C = d.use_last ? Lo.GA : GHcur; g.ldc = Kam; g.cmap = 1; g.rpg = ns; g.gstride = N; g.goff = c->off_s[s];
        if (ns > 0)
            if (int rc = gemm(c, g, GEMM_PLAIN, false, st)) return rc;
        if (d.use_last && ns > 0) {
            // per-walker sums of GY_s: cotangent of the shared spin-mean contribution GOO_s
            if (int rc = ds_launch_group_rowsum(Lo.GYs[s], np2, ns, Wc, np2, Lo.GYS[s], st)) return rc;
            c->launches++;
            if (!fact) {
                GemmParams m{};                       // gWorbG[s] += GINV_L^T . GYS_s
                m.A = Lo.GINV[L]; m.lda = 2 * H; m.M = 2 * H; m.K = Wc; m.rpg = 0;
                m.B = Lo.GYS[s]; m.ldb = np2; m.N = np2; m.C = c->gWorbG[s]; m.ldc = np2; m.accumulate = 1;
                if (int rc = gemm(c, m, GEMM_TN, false, st)) return rc;
            }
            GemmParams q{};                           // GG += GYS_s . WorbG_s^T : cotangent of the spin means of h_L
            q.A = Lo.GYS[s]; q.lda = np2; q.M = Wc; q.K = np2; q.rpg = 0;
            q.B = c->WorbGT[s]; q.ldb = 2 * H; q.N = 2 * H; q.C = Lo.GG; q.ldc = 2 * H; q.accumulate = 1;
            if (int rc = gemm(c, q, GEMM_PLAIN, false, st)) return rc;
        }
    }
    if (d.use_last) {
        // cotangent of h_L = own columns + the electron's share of the spin-mean cotangents; pair-mean columns -> level L
        gb.GA = Lo.GA; gb.lda = d.K1; gb.GG = Lo.GG; gb.ldgg = 2 * H; gb.GHin = GHcur; gb.GPM = Lo.GPM[L]; gb.GH = GHcur;
        if (int rc = ds_launch_hin(d, gb, Wc, H, d.K1, false, true, st)) return rc;
        c->launches++;
    }
    for (int l = L - 1; l >= 0; --l) {
        const int C = (l == 0) ? d.C0 : H;
        const int K = (l == 0) ? d.K0 : d.K1;
        const bool res = (C == H);
        const double* Ain = (l == 0) ? Lo.A0V : Lo.V[l - 1];
        gb.T = Lo.Tl[l]; gb.GH = GHcur; gb.GZ = Lo.GZ; gb.GZS = Lo.GZS; gb.g_bias = c->gbias1[l];
        if (int rc = ds_launch_gz(d, gb, Wc, res, st)) return rc;
        c->launches++;
        GemmParams t{};                               // own + pair-mean rows of the weight gradient
        t.A = Ain; t.lda = K; t.M = K; t.K = (long long)Wc * N; t.rpg = 0;
        t.B = Lo.GZ; t.ldb = H; t.N = H; t.C = c->gB_am[l]; t.ldc = H; t.accumulate = 1;
        if (fact) {                                   // fG1[l] += GZ^T . GZ
            t.A = Lo.GZ; t.lda = H; t.M = H; t.C = c->fG1[l];
        }
        if (int rc = gemm(c, t, GEMM_TN, false, st)) return rc;
        if (!fact) {
            GemmParams m{};                           // spin-mean rows
            m.A = Lo.GINV[l]; m.lda = 2 * C; m.M = 2 * C; m.K = Wc; m.rpg = 0;
            m.B = Lo.GZS; m.ldb = H; m.N = H; m.C = c->gB_g[l]; m.ldc = H; m.accumulate = 1;
            if (int rc = gemm(c, m, GEMM_TN, false, st)) return rc;
        }
        if (l == 0) break;                            // the layer-0 inputs are parameter-free features
        GemmParams a{};                               // GA = GZ . B_am^T
        a.A = Lo.GZ; a.lda = H; a.M = (long long)Wc * N; a.K = H; a.rpg = 0;
        a.B = c->B_amT[l]; a.ldb = K; a.N = K; a.C = Lo.GA; a.ldc = K;
        if (int rc = gemm(c, a, GEMM_PLAIN, false, st)) return rc;
        GemmParams q{};                               // GG = GZS . B_g^T
        q.A = Lo.GZS; q.lda = H; q.M = Wc; q.K = H; q.rpg = 0;
        q.B = c->B_gT[l]; q.ldb = 2 * C; q.N = 2 * C; q.C = Lo.GG; q.ldc = 2 * C;
        if (int rc = gemm(c, q, GEMM_PLAIN, false, st)) return rc;
        gb.GA = Lo.GA; gb.lda = K; gb.GG = Lo.GG; gb.ldgg = 2 * C; gb.GHin = GHnext; gb.GPM = Lo.GPM[l];
        if (int rc = ds_launch_hin(d, gb, Wc, C, K, res, true, st)) return rc;
        c->launches++;
        std::swap(GHcur, GHnext);
    }
    for (int l = 1; l < ds_pair_levels(d); ++l) gb.GPMl[l] = Lo.GPM[l];
    for (int l = 0; l < ds_pair_levels(d) - 1; ++l) { gb.g_Wp[l] = c->gWp[l]; gb.g_bp[l] = c->gbp[l]; gb.fact_G[l] = c->fGp[l]; }
    (void)P;
    if (int rc = ds_launch_pair_grad(sys, fp, gb, Wc, st, fact ? 2 : 0)) return rc;
    c->launches++;
    return 0;
}

// Kronecker-factor statistics of one chunk: Gram matrices of every tagged layer's inputs, then two reverse sweeps
// (unit cotangent on log|psi|, then on the phase) for the Gram matrices of the output cotangents.
int fact_sweep(ds_ctx* c, Layout& Lo, const FeatParams& fp, SlaterBufs& sb, int Wc, cudaStream_t st) {
    const DsDims& d = c->sys.d;
    const int N = d.N, H = d.H, L = d.L;
    for (int l = 0; l < L; ++l) {
        const int C = (l == 0) ? d.C0 : H;
        const int K = (l == 0) ? d.K0 : d.K1;
        const int ldx = 2 * C + K + 2;
        const double* Ain = (l == 0) ? Lo.A0V : Lo.V[l - 1];
        if (int rc = ds_launch_layer_input_rows(Ain, K, Lo.GINV[l], C, N, (long long)Wc * N, Lo.XF, st)) return rc;
        c->launches++;
        GemmParams t{};
        t.A = Lo.XF; t.lda = ldx; t.M = ldx; t.K = (long long)Wc * N; t.rpg = 0;
        t.B = Lo.XF; t.ldb = ldx; t.N = ldx; t.C = c->fA1[l]; t.ldc = ldx; t.accumulate = 1;
        if (int rc = gemm(c, t, GEMM_TN, false, st)) return rc;
    }
    if (d.use_last) {       // the projection's input = symmetric features of the last layer: [own | spin means | pair means | 1 | 0]
        if (int rc = ds_launch_layer_input_rows(Lo.V[L - 1], d.K1, Lo.GINV[L], H, N, (long long)Wc * N, Lo.XF, st)) return rc;
        c->launches++;
    }
    for (int s = 0; s < 2; ++s) {
        const int ns = c->n_s[s];
        if (ns == 0) continue;
        const int kin = d.use_last ? 3 * H + 2 * d.P : H;
        const double* rows_s = Lo.XF;
        if (d.use_last) {
            if (int rc = ds_launch_spin_rows(Lo.XF, kin + 2, kin, N, c->off_s[s], ns, Wc, Lo.XF2, st)) return rc;
            rows_s = Lo.XF2;
        } else if (int rc = ds_launch_spin_rows(Lo.V[L - 1], d.K1, H, N, c->off_s[s], ns, Wc, Lo.XF, st)) return rc;
        c->launches++;
        GemmParams t{};
        t.A = rows_s; t.lda = kin + 2; t.M = kin + 2; t.K = (long long)Wc * ns; t.rpg = 0;
        t.B = rows_s; t.ldb = kin + 2; t.N = kin + 2; t.C = c->fAo[s]; t.ldc = kin + 2; t.accumulate = 1;
        if (int rc = gemm(c, t, GEMM_TN, false, st)) return rc;
    }
    {
        GradBufs gb{};
        for (int l = 0; l < ds_pair_levels(d) - 1; ++l) { gb.fact_A[l] = c->fAp[l]; gb.fact_As[l] = c->fAps[l]; }
        if (int rc = ds_launch_pair_grad(c->sys, fp, gb, Wc, st, 1)) return rc;
        c->launches++;
    }
    DS_REQUIRE((size_t)Wc <= c->fones_n, "unit cotangent buffer too small");
    if (int rc = grad_sweep(c, Lo, fp, sb, Wc, c->fones, c->fones + c->fones_n, st, 0)) return rc;
    return grad_sweep(c, Lo, fp, sb, Wc, c->fones + c->fones_n, c->fones, st, 1);
}

int run_batched(ds_ctx* c, const double* X, long long batch, bool lap, double* log_abs, double* phase,
                double* ke_re, double* ke_im, double* mats_out, cudaStream_t st) {
    DS_REQUIRE(c && c->params_set, "parameters have not been set (ds_set_params)");
    DS_REQUIRE(batch >= 0, "negative batch");
    if (batch == 0) return 0;
    DS_REQUIRE(X != nullptr, "null walker pointer");
    int Wc = 0;
    if (int rc = plan_chunk(c, batch, lap, &Wc)) return rc;
    c->last_chunk = Wc;
    const int n3 = 3 * c->sys.d.N;
    const size_t mat_per = mats_per_walker(c);
    double* const gxa = c->gx_abs;
    double* const gxp = c->gx_phase;
    int rc = 0;
    for (long long w0 = 0; w0 < batch && !rc; w0 += Wc) {
        int wc = (int)std::min<long long>(Wc, batch - w0);
        c->gx_abs = gxa ? gxa + w0 * n3 : nullptr;
        c->gx_phase = gxp ? gxp + w0 * n3 : nullptr;
        rc = run_chunk(c, X + w0 * n3, wc, lap, log_abs ? log_abs + w0 : nullptr, phase ? phase + w0 : nullptr,
                       ke_re ? ke_re + w0 : nullptr, ke_im ? ke_im + w0 : nullptr,
                       mats_out ? mats_out + w0 * mat_per : nullptr, st);
    }
    c->gx_abs = gxa; c->gx_phase = gxp;
    return rc;
}

int ensure(ds_ctx* c, DevBuf& b, size_t n) {
    if (b.n >= n) return 0;
    if (b.p) DS_CUDA_CHECK(cudaFree(b.p));
    b.p = nullptr; b.n = 0;
    DS_CUDA_CHECK(cudaMalloc((void**)&b.p, n * sizeof(double)));
    b.n = n;
    (void)c;
    return 0;
}

}  // namespace

// ===========================================================================
extern "C" int ds_ctx_create(const ds_system_desc* sd, const ds_net_desc* nd, int device, ds_ctx** out) {
    DS_REQUIRE(sd && nd && out, "null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        ds_set_error("no CUDA device available (%s); deepsolid_b200 has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return DS_ERR_CUDA;
    }
    DS_REQUIRE(device >= 0 && device < ndev, "device %d out of range (0..%d)", device, ndev - 1);
    DS_REQUIRE(sd->n_up > 0 && sd->n_dn > 0,
               "both spin channels must be occupied (n_up=%d n_dn=%d): spin-polarised cells are not implemented",
               sd->n_up, sd->n_dn);
    DS_REQUIRE(sd->n_atoms_prim > 0 && sd->n_atoms_prim <= DS_MAX_ATOMS_PRIM,
               "primitive cell must have 1..%d atoms", DS_MAX_ATOMS_PRIM);
    DS_REQUIRE(nd->n_layers >= 2 && nd->n_layers <= DS_MAX_LAYERS, "n_layers must be in 2..%d", DS_MAX_LAYERS);
    DS_REQUIRE(!nd->use_last_layer || nd->n_layers < DS_MAX_LAYERS, "use_last_layer needs n_layers <= %d", DS_MAX_LAYERS - 1);
    DS_REQUIRE(nd->hidden_two >= 2 && nd->hidden_two <= 32 && nd->hidden_two % 2 == 0,
               "two-electron stream width must be even and <= 32 (got %d)", nd->hidden_two);
    DS_REQUIRE(nd->hidden_one >= 2 && nd->hidden_one % 2 == 0, "one-electron stream width must be even");
    DS_REQUIRE(nd->n_det >= 1, "need at least one determinant");
    DS_REQUIRE(sd->dist_kind >= 0 && sd->dist_kind <= 2, "dist_kind must be 0, 1 or 2");
    DS_REQUIRE(nd->distance_type == 0 || nd->distance_type == 1, "Unrecognized distance function.");
    DS_REQUIRE(nd->envelope_type >= 0 && nd->envelope_type <= 2, "envelope_type must be 0 (isotropic), 1 (diagonal) or 2 (full)");
    DS_REQUIRE(nd->envelope_type == 0 || nd->distance_type == 0,
               "diagonal / full envelopes act on 3-component relative vectors: they need distance_type='nu'");
    DS_REQUIRE(((nd->distance_type == 1) ? 7 : 4) * (sd->n_atoms_prim + 2) <= 32,
               "layer-0 operand rows wider than 32 columns are not supported (%d primitive-cell atoms with this distance_type)",
               sd->n_atoms_prim);
    DS_REQUIRE(sd->n_g >= 0 && sd->n_atoms_sim > 0, "bad Ewald table sizes");
    DS_CUDA_CHECK(cudaSetDevice(device));
    ds_ctx* c = new ds_ctx();
    c->device = device;
    if (const char* ev = getenv("DS_NO_I8")) c->use_i8 = atoi(ev) == 0;
    if (const char* ev = getenv("DS_L0_GEMM")) c->use_l0_kernel = atoi(ev) == 0;
    if (const char* ev = getenv("DS_NO_SLICE_MEANS")) c->use_slice_means = atoi(ev) == 0;
    if (const char* ev = getenv("DS_NO_I8_MEANS")) c->use_i8_means = atoi(ev) == 0;
    if (const char* ev = getenv("DS_NO_I8_VALUE")) c->use_i8_value = atoi(ev) == 0;
    if (const char* ev = getenv("DS_FUSED_DIGITS")) c->use_fused_digits = atoi(ev) != 0;
    if (const char* ev = getenv("DS_WS_GIB")) { double g = atof(ev); if (g >= 0.25) c->ws_limit = (size_t)(g * 1073741824.0); }
    DsDims& d = c->sys.d;
    d.n_up = sd->n_up; d.n_dn = sd->n_dn; d.N = sd->n_up + sd->n_dn; d.A = sd->n_atoms_prim;
    d.H = nd->hidden_one; d.P = nd->hidden_two; d.D = nd->n_det; d.L = nd->n_layers;
    d.ND = 3 * d.N; d.NDp = (d.ND + 7) / 8 * 8; d.NDg = d.NDp + 8;
    d.dist_type = nd->distance_type; d.F = (nd->distance_type == 1) ? 7 : 4;
    d.env_type = nd->envelope_type;
    c->bias_orb = nd->bias_orbitals != 0;
    d.full_det = nd->full_det != 0;
    d.use_last = nd->use_last_layer != 0;
    d.C0 = d.F * d.A; d.K0 = d.C0 + 2 * d.F; d.K1 = d.H + 2 * d.P;
    fill_lattice(c->sys.prim, sd->prim_latvec, sd->prim_AV, sd->prim_BV);
    fill_lattice(c->sys.sim, sd->sim_latvec, sd->sim_AV, sd->sim_BV);
    memcpy(c->sys.atoms, sd->prim_atoms, sizeof(double) * 3 * d.A);
    c->n_s[0] = d.n_up; c->n_s[1] = d.n_dn; c->off_s[0] = 0; c->off_s[1] = d.n_up;
    c->npar[0] = ds_norb(d, 0) * d.D; c->npar[1] = ds_norb(d, 1) * d.D; c->npar_max = std::max(c->npar[0], c->npar[1]);
    c->nbuf = d.use_last ? d.L : std::max(2, d.L - 1);      // use_last: every layer keeps its own output buffer

    EwaldDev& ew = c->ew;
    ew.n_elec = d.N; ew.n_atoms = sd->n_atoms_sim; ew.dist_kind = sd->dist_kind; ew.n_g = sd->n_g;
    memcpy(ew.lat, sd->sim_latvec, 9 * sizeof(double));
    inv3(sd->sim_latvec, ew.inv);
    ew.alpha = sd->alpha; ew.ee_const = sd->ee_const; ew.ei_const = sd->ei_const; ew.ii_total = sd->ii_total;
    int rc = 0;
    double* tmp = nullptr;
    rc |= upload(c, &tmp, sd->sim_atoms, 3 * (size_t)sd->n_atoms_sim); ew.atoms = tmp;
    rc |= upload(c, &tmp, sd->sim_charges, (size_t)sd->n_atoms_sim); ew.charges = tmp;
    rc |= upload(c, &tmp, sd->mi_shifts, 81); ew.mi_shifts = tmp;
    rc |= upload(c, &tmp, sd->lattice_displacements, 81); ew.disp = tmp;
    rc |= upload(c, &tmp, sd->gpoints, 3 * (size_t)sd->n_g); ew.gpoints = tmp;
    rc |= upload(c, &tmp, sd->gweight, (size_t)sd->n_g); ew.gweight = tmp;
    rc |= upload(c, &tmp, sd->ion_exp_re, (size_t)sd->n_g); ew.ion_re = tmp;
    rc |= upload(c, &tmp, sd->ion_exp_im, (size_t)sd->n_g); ew.ion_im = tmp;
    if (d.full_det) {       // every channel carries all N orbitals: k-points of both channels concatenated (network.py:453)
        std::vector<double> kcat(3 * (size_t)d.N);
        memcpy(kcat.data(), sd->klist_up, 3 * (size_t)d.n_up * sizeof(double));
        memcpy(kcat.data() + 3 * (size_t)d.n_up, sd->klist_dn, 3 * (size_t)d.n_dn * sizeof(double));
        rc |= upload(c, &c->klist[0], kcat.data(), kcat.size());
        rc |= upload(c, &c->klist[1], kcat.data(), kcat.size());
    } else {
        rc |= upload(c, &c->klist[0], sd->klist_up, 3 * (size_t)d.n_up);
        rc |= upload(c, &c->klist[1], sd->klist_dn, 3 * (size_t)d.n_dn);
    }
    if (rc) { ds_ctx_destroy(c); return DS_ERR_CUDA; }
    *out = c;
    return 0;
}

extern "C" int ds_ctx_destroy(ds_ctx* c) {
    if (!c) return 0;
    Guard g(c->device);
    cudaDeviceSynchronize();
    for (void* p : c->owned) cudaFree(p);
    if (c->ws.base) cudaFree(c->ws.base);
    for (DevBuf* b : {&c->mc_x2, &c->mc_lp, &c->mc_lp2, &c->host_stage}) if (b->p) cudaFree(b->p);
    for (auto& ev : c->prof) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    if (c->tot_a) cudaEventDestroy(c->tot_a);
    if (c->tot_b) cudaEventDestroy(c->tot_b);
    delete c;
    return 0;
}

extern "C" int ds_set_workspace_limit(ds_ctx* c, size_t bytes) {
    DS_REQUIRE(c, "null context");
    DS_REQUIRE(bytes >= (size_t(1) << 20), "workspace limit must be at least 1 MiB");
    c->ws_limit = bytes;
    return 0;
}

extern "C" int ds_set_params(ds_ctx* c, const double* const* leaves, const int64_t* sizes, int n_leaves) {
    DS_REQUIRE(c && leaves && sizes, "null argument");
    Guard g(c->device);
    const DsDims& d = c->sys.d;
    const int L = d.L, H = d.H, P = d.P;
    const int Lpair = d.use_last ? L : L - 1;        // pair layers
    const int Korb = d.use_last ? 3 * H + 2 * P : H; // rows of the orbital weights
    const int expect = 2 * L + 2 * Lpair + (c->bias_orb ? 4 : 2) + 4;
    DS_REQUIRE(n_leaves == expect, "expected %d parameter leaves for %d layers, got %d", expect, L, n_leaves);
    // expected sizes
    std::vector<int64_t> want;
    for (int l = 0; l < L; ++l) {
        int C = (l == 0) ? d.C0 : H, Pl = (l == 0) ? d.F : P;
        want.push_back((int64_t)(3 * C + 2 * Pl) * H);
        want.push_back(H);
    }
    for (int l = 0; l < Lpair; ++l) {
        want.push_back((int64_t)((l == 0) ? d.F : P) * P);
        want.push_back(P);
    }
    for (int s = 0; s < 2; ++s) {
        want.push_back((int64_t)Korb * 2 * c->npar[s]);
        if (c->bias_orb) want.push_back((int64_t)2 * c->npar[s]);
    }
    const int64_t sig_mult = (d.env_type == 0) ? 1 : (d.env_type == 1 ? 3 : 9);
    for (int s = 0; s < 2; ++s) { want.push_back((int64_t)d.A * c->npar[s]); want.push_back(sig_mult * d.A * c->npar[s]); }
    for (int i = 0; i < n_leaves; ++i)
        DS_REQUIRE(sizes[i] == want[i], "parameter leaf %d has %lld elements, expected %lld "
                   "(the leaves follow the net descriptor: layers, widths, use_last_layer, full_det, envelope, orbital bias)",
                   i, (long long)sizes[i], (long long)want[i]);
    // stage every leaf on the host (pointers may be host or device memory)
    std::vector<std::vector<double>> h(n_leaves);
    for (int i = 0; i < n_leaves; ++i) {
        h[i].resize((size_t)sizes[i]);
        DS_CUDA_CHECK(cudaMemcpy(h[i].data(), leaves[i], (size_t)sizes[i] * sizeof(double), cudaMemcpyDefault));
    }
    auto put = [&](double** dst, const std::vector<double>& v) -> int {
        if (!*dst) { if (int rc = dev_alloc(c, dst, v.size())) return rc; }
        DS_CUDA_CHECK(cudaMemcpy(*dst, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
        return 0;
    };
    int li = 0;
    for (int l = 0; l < L; ++l) {
        const int C = (l == 0) ? d.C0 : H, Pl = (l == 0) ? d.F : P;
        const std::vector<double>& W = h[li++];
        const std::vector<double>& b = h[li++];
        std::vector<double> am((size_t)(C + 2 * Pl) * H), gg((size_t)2 * C * H);
        std::copy(W.begin(), W.begin() + (size_t)C * H, am.begin());
        std::copy(W.begin() + (size_t)3 * C * H, W.end(), am.begin() + (size_t)C * H);
        std::copy(W.begin() + (size_t)C * H, W.begin() + (size_t)3 * C * H, gg.begin());
        if (int rc = put(&c->B_am[l], am)) return rc;
        if (int rc = put(&c->B_g[l], gg)) return rc;
        if (int rc = put(&c->bias1[l], b)) return rc;
    }
    for (int l = 0; l < Lpair; ++l) {
        if (int rc = put(&c->Wp[l], h[li++])) return rc;
        if (int rc = put(&c->bp[l], h[li++])) return rc;
    }
    for (int s = 0; s < 2; ++s) {
        const std::vector<double>& W = h[li++];
        const int np = c->npar[s];
        // rows of the reference leaf: own [0,H) (| spin means [H,3H) | pair means [3H,3H+2P) with use_last_layer);
        // device copies: own + pair-mean rows (the per-electron operand), spin-mean rows (shared per walker)
        const int Kam = d.use_last ? d.K1 : H;
        std::vector<double> wi((size_t)Kam * 2 * np);
        auto src_row = [&](int r) { return (!d.use_last || r < H) ? r : 3 * H + (r - H); };
        for (int r = 0; r < Kam; ++r)
            for (int p = 0; p < np; ++p) {
                wi[(size_t)r * 2 * np + 2 * p] = W[(size_t)src_row(r) * 2 * np + p];
                wi[(size_t)r * 2 * np + 2 * p + 1] = W[(size_t)src_row(r) * 2 * np + np + p];
            }
        if (int rc = put(&c->Worb[s], wi)) return rc;
        if (d.use_last) {
            std::vector<double> wg((size_t)2 * H * 2 * np);
            for (int r = 0; r < 2 * H; ++r)
                for (int p = 0; p < np; ++p) {
                    wg[(size_t)r * 2 * np + 2 * p] = W[(size_t)(H + r) * 2 * np + p];
                    wg[(size_t)r * 2 * np + 2 * p + 1] = W[(size_t)(H + r) * 2 * np + np + p];
                }
            if (int rc = put(&c->WorbG[s], wg)) return rc;
        }
        if (c->bias_orb) {
            const std::vector<double>& bv = h[li++];
            std::vector<double> bi((size_t)2 * np);
            for (int p = 0; p < np; ++p) { bi[2 * p] = bv[p]; bi[2 * p + 1] = bv[np + p]; }
            if (int rc = put(&c->borb[s], bi)) return rc;
        }
    }
    for (int s = 0; s < 2; ++s) {
        if (int rc = put(&c->env_pi[s], h[li++])) return rc;
        if (int rc = put(&c->env_sigma[s], h[li++])) return rc;
    }
    // int8 digits of the transposed weights for the tcgen05 path
    {
        auto digits = [&](const double* Bdev, int K, int Nn, signed char** Wd, double** sbp) -> int {
            double* bt = nullptr;
            DS_CUDA_CHECK(cudaMalloc((void**)&bt, (size_t)K * Nn * sizeof(double)));
            if (!*Wd) { if (int rc = dev_alloc(c, Wd, (size_t)Nn * OZ_S * K)) return rc; }
            if (!*sbp) { if (int rc = dev_alloc(c, sbp, (size_t)Nn)) return rc; }
            int rc = ds_launch_transpose(Bdev, K, Nn, bt, 0);
            if (!rc) rc = ds_launch_slice_rows(bt, K, Nn, K, *Wd, *sbp, 0);
            cudaError_t e = cudaStreamSynchronize(0);
            cudaFree(bt);
            if (rc) return rc;
            DS_CUDA_CHECK(e);
            return 0;
        };
        c->i8_ok = (H % OZ_BK == 0) && (d.K1 % OZ_BK == 0);
        if (c->i8_ok) {
            for (int l = 1; l < L; ++l)
                if (int rc = digits(c->B_am[l], d.K1, H, &c->Wd_am[l], &c->sb_am[l])) return rc;
            if (2 * H <= 512 && (2 * H) % OZ_BK == 0)
                for (int l = 1; l < L; ++l)
                    if (int rc = digits(c->B_g[l], 2 * H, H, &c->Wd_g[l], &c->sb_g[l])) return rc;
            for (int s = 0; s < 2; ++s)
                if (int rc = digits(c->Worb[s], d.use_last ? d.K1 : H, 2 * c->npar[s], &c->Wd_orb[s], &c->sb_orb[s])) return rc;
        }
    }
    // the re-layout kernels above ran on the legacy default stream: callers launch on their own (possibly non-blocking)
    // streams afterwards, so the new weights must be complete before this returns
    DS_CUDA_CHECK(cudaStreamSynchronize(0));
    c->params_set = true;
    c->transposes_ready = false;
    return 0;
}

extern "C" int ds_logpsi(ds_ctx* c, const double* x, int64_t batch, double* log_abs, double* phase, void* stream) {
    DS_REQUIRE(c, "null context");
    Guard g(c->device);
    return run_batched(c, x, batch, false, log_abs, phase, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

namespace {
// transposed weights and zeroed accumulators of the parameter-gradient path
int prepare_grad(ds_ctx* c, cudaStream_t st) {
    const DsDims& d = c->sys.d;
    const int L = d.L, H = d.H, P = d.P;
    auto need = [&](double** p, size_t n) -> int { if (!*p) return dev_alloc(c, p, n); return 0; };
    for (int l = 0; l < L; ++l) {
        const int C = (l == 0) ? d.C0 : H, Pl = (l == 0) ? d.F : P, K = C + 2 * Pl;
        if (int rc = need(&c->B_amT[l], (size_t)K * H)) return rc;
        if (int rc = need(&c->B_gT[l], (size_t)2 * C * H)) return rc;
        if (int rc = need(&c->gB_am[l], (size_t)K * H)) return rc;
        if (int rc = need(&c->gB_g[l], (size_t)2 * C * H)) return rc;
        if (int rc = need(&c->gbias1[l], (size_t)H)) return rc;
        DS_CUDA_CHECK(cudaMemsetAsync(c->gB_am[l], 0, (size_t)K * H * sizeof(double), st));
        DS_CUDA_CHECK(cudaMemsetAsync(c->gB_g[l], 0, (size_t)2 * C * H * sizeof(double), st));
        DS_CUDA_CHECK(cudaMemsetAsync(c->gbias1[l], 0, (size_t)H * sizeof(double), st));
        if (!c->transposes_ready) {
            if (int rc = ds_launch_transpose(c->B_am[l], K, H, c->B_amT[l], st)) return rc;
            if (int rc = ds_launch_transpose(c->B_g[l], 2 * C, H, c->B_gT[l], st)) return rc;
        }
    }
    for (int l = 0; l < ds_pair_levels(d) - 1; ++l) {
        const int pin = (l == 0) ? d.F : P;
        if (int rc = need(&c->gWp[l], (size_t)pin * P)) return rc;
        if (int rc = need(&c->gbp[l], (size_t)P)) return rc;
        DS_CUDA_CHECK(cudaMemsetAsync(c->gWp[l], 0, (size_t)pin * P * sizeof(double), st));
        DS_CUDA_CHECK(cudaMemsetAsync(c->gbp[l], 0, (size_t)P * sizeof(double), st));
    }
    for (int s = 0; s < 2; ++s) {
        const size_t np2 = 2 * (size_t)c->npar[s];
        const int Kam = d.use_last ? d.K1 : H;             // own (+ pair-mean) rows of the orbital weights
        if (int rc = need(&c->WorbT[s], np2 * Kam)) return rc;
        if (int rc = need(&c->gWorb[s], np2 * Kam)) return rc;
        if (d.use_last) {
            if (int rc = need(&c->WorbGT[s], np2 * 2 * H)) return rc;
            if (int rc = need(&c->gWorbG[s], np2 * 2 * H)) return rc;
            DS_CUDA_CHECK(cudaMemsetAsync(c->gWorbG[s], 0, np2 * 2 * H * sizeof(double), st));
            if (!c->transposes_ready)
                if (int rc = ds_launch_transpose(c->WorbG[s], 2 * H, (int)np2, c->WorbGT[s], st)) return rc;
        }
        if (c->bias_orb) {
            if (int rc = need(&c->gborb[s], np2)) return rc;
            DS_CUDA_CHECK(cudaMemsetAsync(c->gborb[s], 0, np2 * sizeof(double), st));
        }
        if (int rc = need(&c->genv_pi[s], (size_t)d.A * c->npar[s])) return rc;
        const size_t sig_mult = (d.env_type == 0) ? 1 : (d.env_type == 1 ? 3 : 9);
        if (int rc = need(&c->genv_sigma[s], sig_mult * d.A * c->npar[s])) return rc;
        DS_CUDA_CHECK(cudaMemsetAsync(c->gWorb[s], 0, np2 * Kam * sizeof(double), st));
        DS_CUDA_CHECK(cudaMemsetAsync(c->genv_pi[s], 0, (size_t)d.A * c->npar[s] * sizeof(double), st));
        DS_CUDA_CHECK(cudaMemsetAsync(c->genv_sigma[s], 0, sig_mult * d.A * c->npar[s] * sizeof(double), st));
        if (!c->transposes_ready)
            if (int rc = ds_launch_transpose(c->Worb[s], Kam, (int)np2, c->WorbT[s], st)) return rc;
    }
    c->transposes_ready = true;
    return 0;
}
}  // namespace

// Vector-Jacobian product of (log|psi|, phase) with respect to the parameters, summed over the batch:
//   grad_leaf = sum_w cot_abs[w] d log|psi_w| / d leaf + cot_phase[w] d phase_w / d leaf.
// `grads`: n_leaves device pointers in the leaf order and sizes of ds_set_params; overwritten.
static int vjp_impl(ds_ctx* c, const double* x, int64_t batch, const double* cot_abs, const double* cot_phase,
                    const double* cot_mats, double* const* grads, const int64_t* sizes, int n_leaves, void* stream);

extern "C" int ds_logpsi_vjp(ds_ctx* c, const double* x, int64_t batch, const double* cot_abs, const double* cot_phase,
                             double* const* grads, const int64_t* sizes, int n_leaves, void* stream) {
    DS_REQUIRE(batch == 0 || (cot_abs && cot_phase), "null argument");
    return vjp_impl(c, x, batch, cot_abs, cot_phase, nullptr, grads, sizes, n_leaves, stream);
}

// Pullback through method eval_mats (network.py:601-602; the pretraining loss of pretrain.py:70-89 differentiates the
// orbital matrices): cot_mats has the layout of ds_orbitals, grads = d/dparams sum cot_re Re(M) + cot_im Im(M).
extern "C" int ds_orbitals_vjp(ds_ctx* c, const double* x, int64_t batch, const double* cot_mats, double* const* grads,
                               const int64_t* sizes, int n_leaves, void* stream) {
    DS_REQUIRE(batch == 0 || cot_mats, "null argument");
    return vjp_impl(c, x, batch, nullptr, nullptr, cot_mats, grads, sizes, n_leaves, stream);
}

static int vjp_impl(ds_ctx* c, const double* x, int64_t batch, const double* cot_abs, const double* cot_phase,
                    const double* cot_mats, double* const* grads, const int64_t* sizes, int n_leaves, void* stream) {
    DS_REQUIRE(c && c->params_set, "parameters have not been set (ds_set_params)");
    DS_REQUIRE(grads && sizes, "null argument");
    DS_REQUIRE(batch >= 0, "negative batch");
    Guard g(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    const DsDims& d = c->sys.d;
    const int L = d.L, H = d.H, P = d.P;
    const int Lpair = d.use_last ? L : L - 1;        // pair layers
    const int Korb = d.use_last ? 3 * H + 2 * P : H; // rows of the orbital weights
    const int expect = 2 * L + 2 * Lpair + (c->bias_orb ? 4 : 2) + 4;
    DS_REQUIRE(n_leaves == expect, "expected %d gradient leaves for %d layers, got %d", expect, L, n_leaves);
    if (int rc = prepare_grad(c, st)) return rc;
    if (batch > 0) {
        DS_REQUIRE(x, "null argument");
        int Wc = 0;
        if (int rc = plan_chunk(c, batch, false, &Wc, true)) return rc;
        const int n3 = 3 * d.N;
        const long long mstride = (long long)mats_per_walker(c);
        static const double dummy = 0.0;           // grad mode of run_chunk is keyed on a non-null cotangent pointer
        int rc = 0;
        for (long long w0 = 0; w0 < batch && !rc; w0 += Wc) {
            int wc = (int)std::min<long long>(Wc, batch - w0);
            c->cot_mats = cot_mats ? cot_mats + w0 * mstride : nullptr;
            rc = run_chunk(c, x + w0 * n3, wc, false, nullptr, nullptr, nullptr, nullptr, nullptr, st,
                           cot_mats ? &dummy : cot_abs + w0, cot_mats ? &dummy : cot_phase + w0);
        }
        c->cot_mats = nullptr;
        if (rc) return rc;
    }
    // unpack into the leaf layout of the reference pytree
    int li = 0;
    for (int l = 0; l < L; ++l) {
        const int C = (l == 0) ? d.C0 : H, Pl = (l == 0) ? d.F : P;
        DS_REQUIRE(sizes[li] == (int64_t)(3 * C + 2 * Pl) * H && sizes[li + 1] == H, "gradient leaf %d has the wrong size", li);
        double* w = grads[li++];
        double* b = grads[li++];
        DS_CUDA_CHECK(cudaMemcpyAsync(w, c->gB_am[l], (size_t)C * H * sizeof(double), cudaMemcpyDeviceToDevice, st));
        DS_CUDA_CHECK(cudaMemcpyAsync(w + (size_t)C * H, c->gB_g[l], (size_t)2 * C * H * sizeof(double), cudaMemcpyDeviceToDevice, st));
        DS_CUDA_CHECK(cudaMemcpyAsync(w + (size_t)3 * C * H, c->gB_am[l] + (size_t)C * H, (size_t)2 * Pl * H * sizeof(double),
                                      cudaMemcpyDeviceToDevice, st));
        DS_CUDA_CHECK(cudaMemcpyAsync(b, c->gbias1[l], (size_t)H * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    for (int l = 0; l < Lpair; ++l) {
        const int pin = (l == 0) ? d.F : P;
        DS_REQUIRE(sizes[li] == (int64_t)pin * P && sizes[li + 1] == P, "gradient leaf %d has the wrong size", li);
        DS_CUDA_CHECK(cudaMemcpyAsync(grads[li++], c->gWp[l], (size_t)pin * P * sizeof(double), cudaMemcpyDeviceToDevice, st));
        DS_CUDA_CHECK(cudaMemcpyAsync(grads[li++], c->gbp[l], (size_t)P * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    for (int s = 0; s < 2; ++s) {
        DS_REQUIRE(sizes[li] == (int64_t)Korb * 2 * c->npar[s], "gradient leaf %d has the wrong size", li);
        {
            double* leaf = grads[li++];
            const size_t np2 = 2 * (size_t)c->npar[s];
            // reference rows: own [0,H) | spin means [H,3H) | pair means [3H,3H+2P)
            if (int rc = ds_launch_deinterleave(c->gWorb[s], leaf, H, c->npar[s], st)) return rc;
            if (d.use_last) {
                if (int rc = ds_launch_deinterleave(c->gWorbG[s], leaf + (size_t)H * np2, 2 * H, c->npar[s], st)) return rc;
                if (int rc = ds_launch_deinterleave(c->gWorb[s] + (size_t)H * np2, leaf + (size_t)3 * H * np2, 2 * P, c->npar[s], st)) return rc;
            }
        }
        if (c->bias_orb) {
            DS_REQUIRE(sizes[li] == (int64_t)2 * c->npar[s], "gradient leaf %d has the wrong size", li);
            if (int rc = ds_launch_deinterleave(c->gborb[s], grads[li++], 1, c->npar[s], st)) return rc;
        }
    }
    for (int s = 0; s < 2; ++s) {
        const int64_t sig_mult = (d.env_type == 0) ? 1 : (d.env_type == 1 ? 3 : 9);
        DS_REQUIRE(sizes[li] == (int64_t)d.A * c->npar[s] && sizes[li + 1] == sig_mult * sizes[li], "gradient leaf %d has the wrong size", li);
        DS_CUDA_CHECK(cudaMemcpyAsync(grads[li++], c->genv_pi[s], (size_t)d.A * c->npar[s] * sizeof(double), cudaMemcpyDeviceToDevice, st));
        DS_CUDA_CHECK(cudaMemcpyAsync(grads[li++], c->genv_sigma[s], (size_t)sig_mult * d.A * c->npar[s] * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

// Kronecker-factor statistics of the tagged dense layers, as the reference's KFAC estimator extracts them from
// total_energy_jvp (train.py:128-133 registers conj(log psi) as a normal predictive distribution; estimation mode
// fisher_exact, process.py:221, estimator.py:284-320): for every register_repeated_dense layer (network.py:443 - the
// one-electron layers, the pair layers, the orbital projections) the sums over all rows r (walker x electron / pair) of
//     a_out[layer] = sum_r (x_r, 1)(x_r, 1)^T                        [(in+1)^2]  (curvature_blocks.py:262-281)
//     g_out[layer] = sum_r ga_r ga_r^T + gp_r gp_r^T                 [out^2]
// with ga / gp the cotangents of the layer output y = x w + b for d(sum_w log|psi_w|) and d(sum_w phase_w).  The
// reference's dy = sqrt(2) (ga - i gp) (variance 0.5, loss_functions.py:593-597, vjp_rc.py), so its output factor is
// 2 g_out / rows; normalisation, EMA and the cross-device mean are the host's (deepsolid_b200/kfac.py).
// Layer order: single[0..L-1], double[0..L-2], orbital[spin 0, 1] (orbital output columns: re block | im block).
// env_abs / env_phase: gradients of sum_w log|psi_w| and sum_w phase_w with respect to (pi_0, sigma_0, pi_1, sigma_1)
// (the untagged envelope parameters get a NaiveDiagonal block, curvature_blocks.py:111-133).
extern "C" int ds_kfac_factors(ds_ctx* c, const double* x, int64_t batch, double* const* a_out, const int64_t* a_sizes,
                               double* const* g_out, const int64_t* g_sizes, int n_layers, double* const* env_abs,
                               double* const* env_phase, const int64_t* env_sizes, int n_env, void* stream) {
    DS_REQUIRE(c && c->params_set, "parameters have not been set (ds_set_params)");
    DS_REQUIRE(a_out && a_sizes && g_out && g_sizes && env_abs && env_phase && env_sizes, "null argument");
    DS_REQUIRE(batch >= 0, "negative batch");
    Guard g(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    const DsDims& d = c->sys.d;
    const int L = d.L, H = d.H, P = d.P, N = d.N;
    const int Lpair = ds_pair_levels(d) - 1;          // tagged pair layers
    const int Kin = d.use_last ? 3 * H + 2 * P : H;   // inputs of the orbital projections
    DS_REQUIRE(n_layers == L + Lpair + 2, "expected %d tagged layers, got %d", L + Lpair + 2, n_layers);
    DS_REQUIRE(n_env == 4, "expected 4 envelope leaves, got %d", n_env);
    if (int rc = prepare_grad(c, st)) return rc;
    auto need = [&](double** p, size_t n) -> int { if (!*p) { if (int rc = dev_alloc(c, p, n)) return rc; } return 0; };
    const size_t sig_mult = (d.env_type == 0) ? 1 : (d.env_type == 1 ? 3 : 9);
    for (int l = 0; l < L; ++l) {
        const int C = (l == 0) ? d.C0 : H, K = (l == 0) ? d.K0 : d.K1;
        const size_t ldx = 2 * C + K + 2;
        if (int rc = need(&c->fA1[l], ldx * ldx)) return rc;
        if (int rc = need(&c->fG1[l], (size_t)H * H)) return rc;
        DS_CUDA_CHECK(cudaMemsetAsync(c->fA1[l], 0, ldx * ldx * sizeof(double), st));
        DS_CUDA_CHECK(cudaMemsetAsync(c->fG1[l], 0, (size_t)H * H * sizeof(double), st));
    }
    for (int l = 0; l < Lpair; ++l) {
        if (int rc = need(&c->fAp[l], 32 * 32)) return rc;
        if (int rc = need(&c->fAps[l], 32)) return rc;
        if (int rc = need(&c->fGp[l], 32 * 32)) return rc;
        DS_CUDA_CHECK(cudaMemsetAsync(c->fAp[l], 0, 32 * 32 * sizeof(double), st));
        DS_CUDA_CHECK(cudaMemsetAsync(c->fAps[l], 0, 32 * sizeof(double), st));
        DS_CUDA_CHECK(cudaMemsetAsync(c->fGp[l], 0, 32 * 32 * sizeof(double), st));
    }
    for (int s = 0; s < 2; ++s) {
        const size_t np2 = 2 * (size_t)c->npar[s];
        if (int rc = need(&c->fAo[s], (size_t)(Kin + 2) * (Kin + 2))) return rc;
        if (int rc = need(&c->fGo[s], std::max<size_t>(np2 * np2, 1))) return rc;
        DS_CUDA_CHECK(cudaMemsetAsync(c->fAo[s], 0, (size_t)(Kin + 2) * (Kin + 2) * sizeof(double), st));
        DS_CUDA_CHECK(cudaMemsetAsync(c->fGo[s], 0, np2 * np2 * sizeof(double), st));
        for (int ps = 0; ps < 2; ++ps) {
            const size_t npi = std::max<size_t>((size_t)d.A * c->npar[s], 1);
            if (int rc = need(&c->fenv_pi[ps][s], npi)) return rc;
            if (int rc = need(&c->fenv_sigma[ps][s], sig_mult * npi)) return rc;
            DS_CUDA_CHECK(cudaMemsetAsync(c->fenv_pi[ps][s], 0, (size_t)d.A * c->npar[s] * sizeof(double), st));
            DS_CUDA_CHECK(cudaMemsetAsync(c->fenv_sigma[ps][s], 0, sig_mult * d.A * c->npar[s] * sizeof(double), st));
        }
    }
    if (batch > 0) {
        DS_REQUIRE(x, "null argument");
        c->fact_on = true;
        int Wc = 0;
        int rc = plan_chunk(c, batch, false, &Wc, true);
        if (!rc && (size_t)Wc > c->fones_n) {
            const size_t n = (size_t)Wc;
            double* buf = nullptr;
            rc = dev_alloc(c, &buf, 2 * n);
            if (!rc) {
                std::vector<double> h(2 * n, 0.0);
                for (size_t q = 0; q < n; ++q) h[q] = 1.0;
                cudaError_t e = cudaMemcpyAsync(buf, h.data(), 2 * n * sizeof(double), cudaMemcpyHostToDevice, st);
                if (e == cudaSuccess) e = cudaStreamSynchronize(st);
                if (e != cudaSuccess) { ds_set_error("unit cotangent upload failed: %s", cudaGetErrorString(e)); rc = DS_ERR_CUDA; }
                c->fones = buf; c->fones_n = n;
            }
        }
        const int n3 = 3 * N;
        static const double dummy = 0.0;
        for (long long w0 = 0; w0 < batch && !rc; w0 += Wc) {
            int wc = (int)std::min<long long>(Wc, batch - w0);
            rc = run_chunk(c, x + w0 * n3, wc, false, nullptr, nullptr, nullptr, nullptr, nullptr, st, &dummy, &dummy);
        }
        c->fact_on = false;
        if (rc) return rc;
    }
    // pack into the reference's layouts
    int li = 0;
    for (int l = 0; l < L; ++l, ++li) {
        const int C = (l == 0) ? d.C0 : H, K = (l == 0) ? d.K0 : d.K1;
        const int nin = 2 * C + K + 1, ldx = nin + 1;
        DS_REQUIRE(a_sizes[li] == (int64_t)nin * nin && g_sizes[li] == (int64_t)H * H, "factor %d has the wrong size", li);
        if (int rc = ds_launch_copy2d(c->fA1[l], ldx, a_out[li], nin, nin, nin, st)) return rc;
        DS_CUDA_CHECK(cudaMemcpyAsync(g_out[li], c->fG1[l], (size_t)H * H * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    for (int l = 0; l < Lpair; ++l, ++li) {
        const int pin = (l == 0) ? d.F : P;
        DS_REQUIRE(a_sizes[li] == (int64_t)(pin + 1) * (pin + 1) && g_sizes[li] == (int64_t)P * P, "factor %d has the wrong size", li);
        if (int rc = ds_launch_pair_fact_pack(c->fAp[l], c->fAps[l], (double)batch * N * N, pin, a_out[li], st)) return rc;
        if (int rc = ds_launch_copy2d(c->fGp[l], 32, g_out[li], P, P, P, st)) return rc;
    }
    for (int s = 0; s < 2; ++s, ++li) {
        const int64_t np2 = 2 * (int64_t)c->npar[s];
        DS_REQUIRE(a_sizes[li] == (int64_t)(Kin + 1) * (Kin + 1) && g_sizes[li] == np2 * np2, "factor %d has the wrong size", li);
        if (int rc = ds_launch_copy2d(c->fAo[s], Kin + 2, a_out[li], Kin + 1, Kin + 1, Kin + 1, st)) return rc;
        if (int rc = ds_launch_deinterleave2(c->fGo[s], g_out[li], c->npar[s], st)) return rc;
    }
    for (int s = 0; s < 2; ++s) {
        const int64_t npi = (int64_t)d.A * c->npar[s];
        DS_REQUIRE(env_sizes[2 * s] == npi && env_sizes[2 * s + 1] == (int64_t)sig_mult * npi, "envelope leaf %d has the wrong size", 2 * s);
        DS_CUDA_CHECK(cudaMemcpyAsync(env_abs[2 * s], c->fenv_pi[0][s], npi * sizeof(double), cudaMemcpyDeviceToDevice, st));
        DS_CUDA_CHECK(cudaMemcpyAsync(env_abs[2 * s + 1], c->fenv_sigma[0][s], sig_mult * npi * sizeof(double), cudaMemcpyDeviceToDevice, st));
        DS_CUDA_CHECK(cudaMemcpyAsync(env_phase[2 * s], c->fenv_pi[1][s], npi * sizeof(double), cudaMemcpyDeviceToDevice, st));
        DS_CUDA_CHECK(cudaMemcpyAsync(env_phase[2 * s + 1], c->fenv_sigma[1][s], sig_mult * npi * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

// log|psi|, phase and their gradients with respect to the electron coordinates (what
// jax.value_and_grad(slog_network, argnums=1) gives importance_update, qmc.py:101-118); the gradients come out of
// the forward-Laplacian sweep (first-derivative half: d log psi / d x_d = sum_k w_k sum_s tr(X dM_d)).
extern "C" int ds_logpsi_grad_x(ds_ctx* c, const double* x, int64_t batch, double* log_abs, double* phase,
                                double* grad_abs, double* grad_phase, void* stream) {
    DS_REQUIRE(c, "null context");
    DS_REQUIRE(grad_abs || grad_phase, "no gradient output requested");
    Guard g(c->device);
    c->gx_abs = grad_abs; c->gx_phase = grad_phase;
    int rc = run_batched(c, x, batch, true, log_abs, phase, nullptr, nullptr, nullptr, (cudaStream_t)stream);
    c->gx_abs = nullptr; c->gx_phase = nullptr;
    return rc;
}

extern "C" int64_t ds_orbitals_size(const ds_ctx* c) {
    if (!c) return -1;
    return (int64_t)mats_per_walker(c);
}

extern "C" int ds_orbitals(ds_ctx* c, const double* x, int64_t batch, double* out, void* stream) {
    DS_REQUIRE(c && out, "null argument");
    Guard g(c->device);
    return run_batched(c, x, batch, false, nullptr, nullptr, nullptr, nullptr, out, (cudaStream_t)stream);
}

extern "C" int ds_ewald(ds_ctx* c, const double* x, int64_t batch, double* ee, double* ei, void* stream) {
    DS_REQUIRE(c, "null context");
    if (batch == 0) return 0;
    DS_REQUIRE(x, "null argument");
    Guard g(c->device);
    int rc = ds_launch_ewald(c->ew, x, batch, ee, ei, nullptr, (cudaStream_t)stream);
    if (!rc) c->launches++;
    return rc;
}

// Plane-wave sums of estimator.py (make_structure_factor :42-85, make_complex_polarization :15-40): q_dev [nq][3],
// out_dev [batch][nq] complex (re, im).  mode 0: sum_i exp(i q.x_i); mode 1: exp(i sum_i q.x_i).
extern "C" int ds_rho_q(ds_ctx* c, const double* x, int64_t batch, const double* q, int nq, int mode, double* out,
                        void* stream) {
    DS_REQUIRE(c, "null context");
    DS_REQUIRE(batch >= 0 && nq >= 0, "negative size");
    DS_REQUIRE(mode == 0 || mode == 1, "unknown mode %d", mode);
    if (batch == 0 || nq == 0) return 0;
    DS_REQUIRE(x && q && out, "null argument");
    Guard g(c->device);
    if (int rc = ds_launch_rho_q(x, batch, c->sys.d.N, q, nq, mode, out, (cudaStream_t)stream)) return rc;
    c->launches++;
    return 0;
}

extern "C" double ds_ewald_ii(const ds_ctx* c) { return c ? c->ew.ii_total : 0.0; }

extern "C" int ds_local_energy(ds_ctx* c, const double* x, int64_t batch, int mode, int partition_number,
                               double* ke_re, double* ke_im, double* ewald, void* stream) {
    DS_REQUIRE(c, "null context");
    DS_REQUIRE(mode >= DS_LAP_FOR && mode <= DS_LAP_PARTITION, "Unrecognized laplacian evaluation mode.");
    if (mode == DS_LAP_PARTITION)
        DS_REQUIRE(partition_number >= 1 && (3 * c->sys.d.N) % partition_number == 0,
                   "partition_number (%d) must divide 3*N_elec (%d)", partition_number, 3 * c->sys.d.N);
    DS_REQUIRE(batch >= 0, "negative batch");
    if (batch == 0) return 0;
    DS_REQUIRE(ke_re && ke_im && ewald, "null output pointer");
    Guard g(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (c->prof_on) {
        if (!c->tot_a) { DS_CUDA_CHECK(cudaEventCreate(&c->tot_a)); DS_CUDA_CHECK(cudaEventCreate(&c->tot_b)); }
        DS_CUDA_CHECK(cudaEventRecord(c->tot_a, st));
    }
    int rc = run_batched(c, x, batch, true, nullptr, nullptr, ke_re, ke_im, nullptr, st);
    if (rc) return rc;
    rc = ds_launch_ewald(c->ew, x, batch, nullptr, nullptr, ewald, st);
    if (rc) return rc;
    c->launches++;
    if (c->prof_on) {
        DS_CUDA_CHECK(cudaEventRecord(c->tot_b, st));
        DS_CUDA_CHECK(cudaEventSynchronize(c->tot_b));
        float ms = 0.f;
        DS_CUDA_CHECK(cudaEventElapsedTime(&ms, c->tot_a, c->tot_b));
        c->tot_ms += ms;
    }
    return 0;
}

static int mcmc_impl(ds_ctx* c, double* x, int64_t batch, int steps, double width, uint64_t seed, const double* xi,
                     const double* u, uint8_t* accept, double* n_accept, void* stream, bool one_electron);

extern "C" int ds_mcmc_step(ds_ctx* c, double* x, int64_t batch, int steps, double width, uint64_t seed,
                            const double* xi, const double* u, uint8_t* accept, double* n_accept, void* stream) {
    return mcmc_impl(c, x, batch, steps, width, seed, xi, u, accept, n_accept, stream, false);
}

// qmc.mh_one_electron_update (qmc.py:227-287) driven as make_mcmc_step does (qmc.py:355-358): steps * N
// single-electron moves, move i displaces electron i % N; xi has shape (steps*N, batch, 3), u and the accept
// masks (steps*N, batch).
extern "C" int ds_mcmc_step_one_electron(ds_ctx* c, double* x, int64_t batch, int steps, double width, uint64_t seed,
                                         const double* xi, const double* u, uint8_t* accept, double* n_accept,
                                         void* stream) {
    return mcmc_impl(c, x, batch, steps, width, seed, xi, u, accept, n_accept, stream, true);
}

static int mcmc_impl(ds_ctx* c, double* x, int64_t batch, int steps, double width, uint64_t seed, const double* xi,
                     const double* u, uint8_t* accept, double* n_accept, void* stream, bool one_electron) {
    DS_REQUIRE(c && n_accept, "null argument");
    DS_REQUIRE(steps >= 0, "negative number of MCMC steps");
    DS_REQUIRE(x || batch == 0, "null walker pointer");
    Guard g(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int n3 = 3 * c->sys.d.N;
    if (int rc = ensure(c, c->mc_x2, (size_t)batch * n3)) return rc;
    if (int rc = ensure(c, c->mc_lp, (size_t)batch)) return rc;
    if (int rc = ensure(c, c->mc_lp2, (size_t)batch)) return rc;
    DS_CUDA_CHECK(cudaMemsetAsync(n_accept, 0, sizeof(double), st));
    if (batch == 0) return 0;
    // logprob = 2 log|psi|   (qmc.py:357)
    if (int rc = run_batched(c, x, batch, false, c->mc_lp.p, nullptr, nullptr, nullptr, nullptr, st)) return rc;
    if (int rc = ds_launch_scale(c->mc_lp.p, c->mc_lp.p, 2.0, batch, st)) return rc;
    c->launches++;
    const int n_el = c->sys.d.N;
    const long long nsteps = one_electron ? (long long)steps * n_el : steps;
    const size_t xi_stride = one_electron ? (size_t)batch * 3 : (size_t)batch * n3;
    for (long long s = 0; s < nsteps; ++s) {
        if (int rc = ds_launch_propose(c->sys.sim, x, c->mc_x2.p, batch, n3, width,
                                       xi ? xi + (size_t)s * xi_stride : nullptr, seed, (unsigned long long)s, st,
                                       one_electron ? (int)(s % n_el) : -1)) return rc;
        if (int rc = run_batched(c, c->mc_x2.p, batch, false, c->mc_lp2.p, nullptr, nullptr, nullptr, nullptr, st)) return rc;
        if (int rc = ds_launch_scale(c->mc_lp2.p, c->mc_lp2.p, 2.0, batch, st)) return rc;
        if (int rc = ds_launch_accept(x, c->mc_x2.p, c->mc_lp.p, c->mc_lp2.p, batch, n3,
                                      u ? u + (size_t)s * batch : nullptr, seed, (unsigned long long)s,
                                      accept ? accept + (size_t)s * batch : nullptr, n_accept, st)) return rc;
        c->launches += 3;
    }
    return 0;
}

extern "C" int ds_energy_stats(ds_ctx* c, const double* ke_re, const double* ke_im, const double* ew,
                               int64_t batch, double* out6, void* stream) {
    DS_REQUIRE(c && out6, "null argument");
    Guard g(c->device);
    int rc = ds_launch_stats(ke_re, ke_im, ew, batch, out6, (cudaStream_t)stream);
    if (!rc) c->launches++;
    return rc;
}

// ---- host-buffer forms ----------------------------------------------------
extern "C" int ds_logpsi_host(ds_ctx* c, const double* x, int64_t batch, double* log_abs, double* phase) {
    DS_REQUIRE(c, "null context");
    if (batch == 0) return 0;
    DS_REQUIRE(x, "null argument");
    Guard g(c->device);
    const size_t n3 = 3 * (size_t)c->sys.d.N;
    if (int rc = ensure(c, c->host_stage, (size_t)batch * (n3 + 2))) return rc;
    double* dx = c->host_stage.p;
    double* dl = dx + (size_t)batch * n3;
    double* dp = dl + batch;
    DS_CUDA_CHECK(cudaMemcpyAsync(dx, x, (size_t)batch * n3 * sizeof(double), cudaMemcpyHostToDevice, 0));
    if (int rc = ds_logpsi(c, dx, batch, dl, dp, nullptr)) return rc;
    if (log_abs) DS_CUDA_CHECK(cudaMemcpyAsync(log_abs, dl, batch * sizeof(double), cudaMemcpyDeviceToHost, 0));
    if (phase) DS_CUDA_CHECK(cudaMemcpyAsync(phase, dp, batch * sizeof(double), cudaMemcpyDeviceToHost, 0));
    DS_CUDA_CHECK(cudaStreamSynchronize(0));
    return 0;
}

extern "C" int ds_local_energy_host(ds_ctx* c, const double* x, int64_t batch, int mode, int partition_number,
                                    double* ke_re, double* ke_im, double* ewald) {
    DS_REQUIRE(c, "null context");
    if (batch == 0) return ds_local_energy(c, x, 0, mode, partition_number, ke_re, ke_im, ewald, nullptr);
    DS_REQUIRE(x && ke_re && ke_im && ewald, "null argument");
    Guard g(c->device);
    const size_t n3 = 3 * (size_t)c->sys.d.N;
    if (int rc = ensure(c, c->host_stage, (size_t)batch * (n3 + 3))) return rc;
    double* dx = c->host_stage.p;
    double* d0 = dx + (size_t)batch * n3;
    DS_CUDA_CHECK(cudaMemcpyAsync(dx, x, (size_t)batch * n3 * sizeof(double), cudaMemcpyHostToDevice, 0));
    if (int rc = ds_local_energy(c, dx, batch, mode, partition_number, d0, d0 + batch, d0 + 2 * batch, nullptr)) return rc;
    DS_CUDA_CHECK(cudaMemcpyAsync(ke_re, d0, batch * sizeof(double), cudaMemcpyDeviceToHost, 0));
    DS_CUDA_CHECK(cudaMemcpyAsync(ke_im, d0 + batch, batch * sizeof(double), cudaMemcpyDeviceToHost, 0));
    DS_CUDA_CHECK(cudaMemcpyAsync(ewald, d0 + 2 * batch, batch * sizeof(double), cudaMemcpyDeviceToHost, 0));
    DS_CUDA_CHECK(cudaStreamSynchronize(0));
    return 0;
}

extern "C" int ds_mcmc_step_host(ds_ctx* c, double* x, int64_t batch, int steps, double width, uint64_t seed,
                                 const double* xi, const double* u, uint8_t* accept, double* n_accept) {
    DS_REQUIRE(c && x && n_accept, "null argument");
    Guard g(c->device);
    const size_t n3 = 3 * (size_t)c->sys.d.N;
    size_t need = (size_t)batch * n3 + 8;
    if (xi) need += (size_t)steps * batch * n3;
    if (u) need += (size_t)steps * batch;
    need += ((size_t)steps * batch + 7) / 8 + 8;
    if (int rc = ensure(c, c->host_stage, need)) return rc;
    double* dx = c->host_stage.p;
    double* dn = dx + (size_t)batch * n3;
    double* p = dn + 8;
    double *dxi = nullptr, *du = nullptr;
    if (xi) { dxi = p; p += (size_t)steps * batch * n3; }
    if (u) { du = p; p += (size_t)steps * batch; }
    uint8_t* dacc = accept ? reinterpret_cast<uint8_t*>(p) : nullptr;
    DS_CUDA_CHECK(cudaMemcpyAsync(dx, x, (size_t)batch * n3 * sizeof(double), cudaMemcpyHostToDevice, 0));
    if (xi) DS_CUDA_CHECK(cudaMemcpyAsync(dxi, xi, (size_t)steps * batch * n3 * sizeof(double), cudaMemcpyHostToDevice, 0));
    if (u) DS_CUDA_CHECK(cudaMemcpyAsync(du, u, (size_t)steps * batch * sizeof(double), cudaMemcpyHostToDevice, 0));
    if (int rc = ds_mcmc_step(c, dx, batch, steps, width, seed, dxi, du, dacc, dn, nullptr)) return rc;
    DS_CUDA_CHECK(cudaMemcpyAsync(x, dx, (size_t)batch * n3 * sizeof(double), cudaMemcpyDeviceToHost, 0));
    DS_CUDA_CHECK(cudaMemcpyAsync(n_accept, dn, sizeof(double), cudaMemcpyDeviceToHost, 0));
    if (accept) DS_CUDA_CHECK(cudaMemcpyAsync(accept, dacc, (size_t)steps * batch, cudaMemcpyDeviceToHost, 0));
    DS_CUDA_CHECK(cudaStreamSynchronize(0));
    return 0;
}

// ---- instrumentation ------------------------------------------------------
extern "C" int64_t ds_launch_count(const ds_ctx* c) { return c ? c->launches : -1; }

extern "C" int ds_profile_enable(ds_ctx* c, int on) {
    DS_REQUIRE(c, "null context");
    c->prof_on = on != 0;
    return 0;
}

extern "C" int ds_profile_reset(ds_ctx* c) {
    DS_REQUIRE(c, "null context");
    Guard g(c->device);
    for (auto& ev : c->prof) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    c->prof.clear();
    c->tot_ms = 0.0;
    return 0;
}

extern "C" int ds_profile_get(ds_ctx* c, double* jac_ms, int64_t* jac_launches, double* jac_flops, double* total_ms) {
    DS_REQUIRE(c, "null context");
    Guard g(c->device);
    double ms = 0.0, fl = 0.0;
    for (auto& ev : c->prof) {
        DS_CUDA_CHECK(cudaEventSynchronize(ev.b));
        float t = 0.f;
        DS_CUDA_CHECK(cudaEventElapsedTime(&t, ev.a, ev.b));
        ms += t; fl += ev.flops;
    }
    if (jac_ms) *jac_ms = ms;
    if (jac_launches) *jac_launches = (int64_t)c->prof.size();
    if (jac_flops) *jac_flops = fl;
    if (total_ms) *total_ms = c->tot_ms;
    return 0;
}

extern "C" int ds_workspace_info(ds_ctx* c, int64_t* chunk_walkers, int64_t* workspace_bytes) {
    DS_REQUIRE(c, "null context");
    if (chunk_walkers) *chunk_walkers = c->last_chunk;
    if (workspace_bytes) *workspace_bytes = (int64_t)(c->ws.cap * sizeof(double));
    return 0;
}

extern "C" int ds_debug_set_int(ds_ctx* c, const char* key, int value) {
    DS_REQUIRE(c && key, "null argument");
    if (!strcmp(key, "stop_layer")) { c->dbg_stop_layer = value; return 0; }
    if (!strcmp(key, "i8")) { c->use_i8 = value != 0; return 0; }
    if (!strcmp(key, "l0_kernel")) { c->use_l0_kernel = value != 0; return 0; }
    if (!strcmp(key, "slice_means")) { c->use_slice_means = value != 0; return 0; }
    if (!strcmp(key, "fused_digits")) { c->use_fused_digits = value != 0; return 0; }
    ds_set_error("unknown debug key %s", key);
    return -1;
}

extern "C" int64_t ds_debug_buffer(ds_ctx* c, const char* name, double* dst, int64_t max_doubles) {
    if (!c || !name) return -1;
    Guard g(c->device);
    if (!strcmp(name, "oz_prof")) {          // role clocks of oz_gemm_kernel (DS_OZ_OPT bit 32), 8 modes x 8 counters; read clears
        if (dst && max_doubles >= 64) {
            unsigned long long raw[64];
            if (ds_oz_prof_read(raw, 1)) return -2;
            double v[64];
            for (int i = 0; i < 64; ++i) v[i] = (double)raw[i];
            if (cudaMemcpy(dst, v, sizeof(v), cudaMemcpyHostToDevice) != cudaSuccess) return -2;
        }
        return 64;
    }
    for (const Region& r : c->last_regions) {
        if (!strcmp(r.name, name)) {
            if (!r.p) return 0;
            int64_t n = std::min<int64_t>((int64_t)r.n, max_doubles);
            if (dst && n > 0) {
                if (cudaMemcpy(dst, r.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice) != cudaSuccess) return -2;
            }
            return (int64_t)r.n;
        }
    }
    ds_set_error("no workspace region named %s", name);
    return -1;
}

extern "C" int ds_dgemm_probe(int device, const double* a, const double* b, double* cc, int64_t m, int n, int k,
                              void* stream) {
    Guard g(device);
    GemmParams p{};
    p.A = a; p.lda = k; p.B = b; p.ldb = n; p.M = m; p.N = n; p.K = k; p.C = cc; p.ldc = n;
    return ds_launch_gemm(p, GEMM_PLAIN, false, (cudaStream_t)stream);
}

// Stand-alone run of the tcgen05 int8-slice GEMM (C = A.B, row-major fp64 in / out): digits of A
// and of B^T are formed on the device, then the GEMM is launched `reps` times; *gemm_ms = average
// device time of one GEMM launch, *slice_ms = time of the digit kernel over A.
extern "C" int ds_ozaki_dgemm_probe(int device, const double* a, const double* b, double* cc, int64_t m, int n, int k,
                                    int reps, double* gemm_ms, double* slice_ms, void* stream) {
    Guard g(device);
    cudaStream_t st = (cudaStream_t)stream;
    DS_REQUIRE(m > 0 && n > 0 && k > 0 && reps >= 1, "bad probe sizes");
    signed char *Ad = nullptr, *Wd = nullptr;
    double *sa = nullptr, *sb = nullptr, *bt = nullptr;
    const bool mn = getenv("DS_OZ_PROBE_MN") && atoi(getenv("DS_OZ_PROBE_MN")) != 0;   // A as row-contiguous digits (MN-major operand)
    const long long Rp = (m + 63) / 64 * 64;
    DS_CUDA_CHECK(cudaMalloc((void**)&Ad, (size_t)Rp * OZ_S * k));
    DS_CUDA_CHECK(cudaMalloc((void**)&Wd, (size_t)n * OZ_S * k));
    DS_CUDA_CHECK(cudaMalloc((void**)&sa, (size_t)m * sizeof(double)));
    DS_CUDA_CHECK(cudaMalloc((void**)&sb, (size_t)n * sizeof(double)));
    DS_CUDA_CHECK(cudaMalloc((void**)&bt, (size_t)n * k * sizeof(double)));
    cudaEvent_t e0, e1, e2, e3;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); cudaEventCreate(&e3);
    int rc = ds_launch_transpose(b, k, n, bt, st);
    if (!rc) rc = ds_launch_slice_rows(bt, k, n, k, Wd, sb, st);
    cudaEventRecord(e0, st);
    if (mn) { cudaMemsetAsync(Ad, 0, (size_t)Rp * OZ_S * k, st); if (!rc) rc = ds_launch_slice_rows_mn(a, k, m, k, Rp, Ad, sa, st); }
    else if (!rc) rc = ds_launch_slice_rows(a, k, m, k, Ad, sa, st);
    cudaEventRecord(e1, st);
    OzParams p{};
    p.Ad = Ad; p.sa = sa; p.rpg = m; p.gstride = m; p.goff = 0; p.n_groups = 1;
    p.Wd = Wd; p.sb = sb; p.N = n; p.K = k; p.C = cc; p.ldc = n;
    p.bmn = mn ? 1 : 0; p.Rp_in = Rp;
    const int pmode = mn ? OZ_PLAIN + 100 : OZ_PLAIN;
    if (const char* dv = getenv("DS_OZ_DBG")) p.dbg = atoi(dv);
    if (!rc) rc = ds_launch_oz_gemm(p, pmode, false, st);        // warm-up (also configures the kernel)
    cudaEventRecord(e2, st);
    for (int i = 0; i < reps && !rc; ++i) rc = ds_launch_oz_gemm(p, pmode, false, st);
    cudaEventRecord(e3, st);
    cudaError_t ce = cudaStreamSynchronize(st);
    float t01 = 0.f, t23 = 0.f;
    cudaEventElapsedTime(&t01, e0, e1);
    cudaEventElapsedTime(&t23, e2, e3);
    if (slice_ms) *slice_ms = t01;
    if (gemm_ms) *gemm_ms = t23 / reps;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3);
    cudaFree(Ad); cudaFree(Wd); cudaFree(sa); cudaFree(sb); cudaFree(bt);
    if (rc) return rc;
    if (ce != cudaSuccess) { ds_set_error("ozaki probe failed: %s", cudaGetErrorString(ce)); return -2; }
    return 0;
}

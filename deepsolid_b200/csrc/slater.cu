// Orbitals -> Slater matrices -> log-determinants and the determinant part of the
// Laplacian (network.py:335-337, 375-427, 449-458, 536-557; hamiltonian.py:45-70).
//
//   M[k][i,o]   = (h_i . W_orb)[k,o] * E_i[k,o],   E_i[k,o] = env_i[k,o] * exp(i k_o.x_i)
//   psi         = sum_k prod_s det M_s[k]
//   lap psi/psi = sum_k w_k { sum_s [ tr(X lapM) - sum_d tr((X dM_d)^2) ] + sum_d (sum_s tr(X dM_d))^2 },
//                 X = M^-1, w_k = D_k / psi
#include "kernels.cuh"
#include <stdlib.h>

namespace {

// ---------------------------------------------------------------------------
// E table: envelope * Bloch phase and its own-electron jet.  grid = Wc*N, threads over p.
// ---------------------------------------------------------------------------
template <bool JETS>
__global__ void __launch_bounds__(256) etab_kernel(const DsSys sys, const SlaterBufs sb, int npar_max) {
    const DsDims& dm = sys.d;
    const long long e = blockIdx.x;
    const int N = dm.N, A = dm.A, D = dm.D;
    const int w = (int)(e / N), i = (int)(e % N);
    const int s = (i < dm.n_up) ? 0 : 1;
    const int ns = s ? dm.n_dn : dm.n_up;
    const int npar = ns * D;
    const double* x = sb.X + (long long)w * 3 * N + 3 * i;
    const double x0 = x[0], x1 = x[1], x2 = x[2];
    const double* rae = sb.RAE + e * A * 5;
    const double* pi_ = sb.env_pi[s];
    const double* sg_ = sb.env_sigma[s];
    const double* kl = sb.klist[s];
    double* out = sb.ETAB + e * 5LL * npar_max * 2;
    for (int p = threadIdx.x; p < npar; p += blockDim.x) {
        double ev = 0.0, eg0 = 0.0, eg1 = 0.0, eg2 = 0.0, el = 0.0;
        for (int a = 0; a < A; ++a) {
            const double r = rae[a * 5], sig = sg_[a * npar + p], pw = pi_[a * npar + p];
            const double asig = fabs(sig);
            const double ex = exp(-fabs(r * sig)) * pw;
            ev += ex;
            if (JETS) {
                const double g0 = rae[a * 5 + 1], g1 = rae[a * 5 + 2], g2 = rae[a * 5 + 3], l = rae[a * 5 + 4];
                const double sr = ds_sign(r);      // r >= 0; keeps d|r sigma|/dr = sign(r sigma) sigma exact
                const double d1 = -asig * sr * ex;
                eg0 += d1 * g0; eg1 += d1 * g1; eg2 += d1 * g2;
                el += d1 * l + sig * sig * ex * (g0 * g0 + g1 * g1 + g2 * g2);
            }
        }
        const int o = p % ns;
        const double k0 = kl[o * 3], k1 = kl[o * 3 + 1], k2 = kl[o * 3 + 2];
        double sn, cs;
        sincos(k0 * x0 + k1 * x1 + k2 * x2, &sn, &cs);
        // value
        out[2 * p] = ev * cs;
        out[2 * p + 1] = ev * sn;
        if (JETS) {
            // d/dc (env * ph) = env_g * ph + env * (i k_c) ph ;  i*ph = (-sn, cs)
            const double kk[3] = {k0, k1, k2};
            const double eg[3] = {eg0, eg1, eg2};
            double lre = el * cs, lim = el * sn;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double* oc = out + (long long)(1 + c) * npar_max * 2;
                oc[2 * p] = eg[c] * cs - ev * kk[c] * sn;
                oc[2 * p + 1] = eg[c] * sn + ev * kk[c] * cs;
                lre += 2.0 * eg[c] * kk[c] * (-sn);
                lim += 2.0 * eg[c] * kk[c] * cs;
            }
            const double k2n = k0 * k0 + k1 * k1 + k2 * k2;
            lre -= ev * k2n * cs;
            lim -= ev * k2n * sn;
            double* ol = out + 4LL * npar_max * 2;
            ol[2 * p] = lre;
            ol[2 * p + 1] = lim;
        }
    }
}

// ---------------------------------------------------------------------------
// Assemble the Slater matrices from the raw orbital-layer outputs.  grid = Wc*N.
// ---------------------------------------------------------------------------
template <bool JETS>
__global__ void __launch_bounds__(256) assemble_kernel(const DsSys sys, const SlaterBufs sb, int npar_max) {
    const DsDims& dm = sys.d;
    const long long e = blockIdx.x;
    const int N = dm.N, D = dm.D, NDp = dm.NDp;
    const long long w = e / N;
    const int i = (int)(e % N);
    const int s = (i < dm.n_up) ? 0 : 1;
    const int ns = s ? dm.n_dn : dm.n_up;
    const int is = s ? i - dm.n_up : i;
    const int npar = ns * D;
    const cplx* E = reinterpret_cast<const cplx*>(sb.ETAB) + e * 5LL * npar_max;
    const cplx* yv = reinterpret_cast<const cplx*>(sb.YV) + e * (long long)npar_max;
    const cplx* yl = reinterpret_cast<const cplx*>(sb.YL) + e * (long long)npar_max;
    const cplx* yo = reinterpret_cast<const cplx*>(sb.YOWN) + e * 3LL * npar_max;
    cplx* mat = reinterpret_cast<cplx*>(sb.MAT[s]);
    cplx* lapm = reinterpret_cast<cplx*>(sb.LAPM[s]);
    cplx* da = reinterpret_cast<cplx*>(sb.DA[s]);
    for (int p = threadIdx.x; p < npar; p += blockDim.x) {
        const int k = p / ns, o = p - k * ns;
        const cplx O = yv[p];
        const cplx E0 = E[p];
        const long long mi = ((w * D + k) * ns + is) * ns + o;
        mat[mi] = cmul(O, E0);
        if (JETS) {
            cplx l = cmul(yl[p], E0);
            cfma(l, O, E[4LL * npar_max + p]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const cplx Ec = E[(long long)(1 + c) * npar_max + p];
                const cplx oj = yo[(long long)c * npar_max + p];
                l.re += 2.0 * (oj.re * Ec.re - oj.im * Ec.im);
                l.im += 2.0 * (oj.re * Ec.im + oj.im * Ec.re);
                const long long di = (((w * D + k) * NDp + 3 * i + c) * ns + is) * (long long)ns + o;
                cplx cur = da[di];
                cfma(cur, O, Ec);
                da[di] = cur;
            }
            lapm[mi] = l;
        }
    }
}

// ---------------------------------------------------------------------------
// Determinant kernel: one CTA per (walker, spin, determinant).
//   LAP=false: LU with partial pivoting -> log|det|, phase.
//   LAP=true : in-place Gauss-Jordan inverse, then tr(X lapM), tr(X dM_d), tr((X dM_d)^2).
// ---------------------------------------------------------------------------
constexpr int DET_THREADS = 256;

__device__ __forceinline__ double cabs1(cplx a) { return fabs(a.re) + fabs(a.im); }

template <bool LAP>
__global__ void __launch_bounds__(DET_THREADS) det_kernel(const DsSys sys, const SlaterBufs sb, int G) {
    const DsDims& dm = sys.d;
    const int D = dm.D, NDp = dm.NDp, ND = dm.ND;
    const int k = blockIdx.x % D;
    const int s = (blockIdx.x / D) % 2;
    const long long w = blockIdx.x / (2 * D);
    const int n = s ? dm.n_dn : dm.n_up;
    const int np = ((n + 2) / 3) * 3;               // padded to the 3x3 register block
    const int tid = threadIdx.x;

    extern __shared__ __align__(16) double smraw[];
    cplx* Xs = reinterpret_cast<cplx*>(smraw);      // [np][np] rows = orbital (after inversion), cols = electron
    cplx* colk = Xs + np * np;                      // [np]
    cplx* dAs = colk + np;                          // [G][np][np]  rows = electron, cols = orbital
    cplx* Ys = dAs + (LAP ? G * np * np : 0);       // [G][np][np]
    __shared__ int piv[128];
    __shared__ int s_p;
    __shared__ double s_red[2 * 16 + 4];

    const cplx* mat = reinterpret_cast<const cplx*>(sb.MAT[s]) + (w * D + k) * (long long)n * n;
    for (int t = tid; t < np * np; t += DET_THREADS) {
        int r = t / np, c = t - r * np;
        Xs[t] = (r < n && c < n) ? mat[r * n + c] : cplx{0.0, 0.0};
    }
    __syncthreads();

    double logabs = 0.0;
    cplx phase{1.0, 0.0};
    for (int kk = 0; kk < n; ++kk) {
        // pivot search in column kk (warp 0)
        if (tid < 32) {
            double best = -1.0;
            int bi = kk;
            for (int r = kk + tid; r < n; r += 32) {
                double v = cabs1(Xs[r * np + kk]);
                if (v > best) { best = v; bi = r; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, best, off);
                int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (tid == 0) { s_p = bi; piv[kk] = bi; }
        }
        __syncthreads();
        const int p = s_p;
        if (p != kk) {
            for (int c = tid; c < n; c += DET_THREADS) {
                cplx a = Xs[kk * np + c];
                Xs[kk * np + c] = Xs[p * np + c];
                Xs[p * np + c] = a;
            }
            phase = cplx{-phase.re, -phase.im};
        }
        __syncthreads();
        const cplx pv = Xs[kk * np + kk];
        {
            double a = hypot(pv.re, pv.im);
            logabs += log(a);
            phase = cmul(phase, cplx{pv.re / a, pv.im / a});
        }
        const cplx ipv = cinv(pv);
        // save column kk, then scale the pivot row
        for (int r = tid; r < n; r += DET_THREADS) colk[r] = Xs[r * np + kk];
        __syncthreads();
        if (LAP) {
            for (int c = tid; c < n; c += DET_THREADS) {
                cplx a = (c == kk) ? cplx{1.0, 0.0} : Xs[kk * np + c];
                Xs[kk * np + c] = cmul(a, ipv);
            }
            __syncthreads();
            for (int t = tid; t < n * n; t += DET_THREADS) {
                int r = t / n, c = t - r * n;
                if (r == kk) continue;
                cplx f = colk[r];
                cplx a = (c == kk) ? cplx{0.0, 0.0} : Xs[r * np + c];
                cplx b = Xs[kk * np + c];
                a.re -= f.re * b.re - f.im * b.im;
                a.im -= f.re * b.im + f.im * b.re;
                Xs[r * np + c] = a;
            }
        } else {
            const int m = n - kk - 1;
            for (int t = tid; t < m * m; t += DET_THREADS) {
                int r = kk + 1 + t / m, c = kk + 1 + t % m;
                cplx f = cmul(colk[r], ipv);
                cplx b = Xs[kk * np + c];
                cplx a = Xs[r * np + c];
                a.re -= f.re * b.re - f.im * b.im;
                a.im -= f.re * b.im + f.im * b.re;
                Xs[r * np + c] = a;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        double* ld = sb.LOGDET + ((w * 2 + s) * D + k) * 3;
        ld[0] = logabs; ld[1] = phase.re; ld[2] = phase.im;
    }
    if (!LAP) return;

    // undo the row pivoting: swap columns in reverse order
    for (int kk = n - 1; kk >= 0; --kk) {
        const int p = piv[kk];
        if (p != kk) {
            for (int r = tid; r < n; r += DET_THREADS) {
                cplx a = Xs[r * np + kk];
                Xs[r * np + kk] = Xs[r * np + p];
                Xs[r * np + p] = a;
            }
        }
        __syncthreads();
    }
    // now Xs[o][i] = (M^-1)[o,i]
    if (sb.XINV[s] != nullptr) {                     // parameter-gradient path: only the inverse is wanted
        cplx* xo = reinterpret_cast<cplx*>(sb.XINV[s]) + (w * D + k) * (long long)n * n;
        for (int t = tid; t < n * n; t += DET_THREADS) {
            int o = t / n, i = t - o * n;
            xo[t] = Xs[o * np + i];
        }
        return;
    }

    const int lane = tid & 31, warp = tid >> 5;
    auto block_sum2 = [&](double a, double b, double& oa, double& ob) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, off);
            b += __shfl_xor_sync(0xffffffffu, b, off);
        }
        __syncthreads();
        if (lane == 0) { s_red[2 * warp] = a; s_red[2 * warp + 1] = b; }
        __syncthreads();
        oa = 0.0; ob = 0.0;
        for (int q = 0; q < DET_THREADS / 32; ++q) { oa += s_red[2 * q]; ob += s_red[2 * q + 1]; }
    };

    // tr(X lapM) = sum_{i,o} X[o,i] lapM[i,o]
    {
        const cplx* lm = reinterpret_cast<const cplx*>(sb.LAPM[s]) + (w * D + k) * (long long)n * n;
        cplx acc{0.0, 0.0};
        for (int t = tid; t < n * n; t += DET_THREADS) {
            int i = t / n, o = t - i * n;
            cfma(acc, Xs[o * np + i], lm[t]);
        }
        double ra, rb;
        block_sum2(acc.re, acc.im, ra, rb);
        if (tid == 0) {
            double* tl = sb.TRLAP + ((w * 2 + s) * D + k) * 2;
            tl[0] = ra; tl[1] = rb;
        }
    }

    const cplx* da = reinterpret_cast<const cplx*>(sb.DA[s]) + (w * D + k) * (long long)NDp * n * n;
    double* tau_out = sb.TAU + ((w * 2 + s) * D + k) * (long long)NDp * 2;
    const int nb = np / 3;
    cplx trsq{0.0, 0.0};                             // accumulated by every thread redundantly after block sums
    for (int d0 = 0; d0 < ND; d0 += G) {
        const int g_cnt = min(G, ND - d0);
        __syncthreads();
        // stage dM_d (zero padded)
        for (int t = tid; t < g_cnt * np * np; t += DET_THREADS) {
            int g = t / (np * np), rem = t - g * np * np;
            int i = rem / np, o = rem - i * np;
            dAs[t] = (i < n && o < n) ? da[((long long)(d0 + g) * n + i) * n + o] : cplx{0.0, 0.0};
        }
        __syncthreads();
        // Y_g = X . dM_g with 3x3 register blocks
        for (int t = tid; t < g_cnt * nb * nb; t += DET_THREADS) {
            int g = t / (nb * nb), rem = t - g * nb * nb;
            int rb = (rem / nb) * 3, cb = (rem % nb) * 3;
            cplx y[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) y[a][b] = cplx{0.0, 0.0};
            const cplx* xa = Xs + rb * np;
            const cplx* db = dAs + g * np * np + cb;
            for (int i = 0; i < n; ++i) {
                cplx xv[3], dv[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) xv[a] = xa[a * np + i];
#pragma unroll
                for (int b = 0; b < 3; ++b) dv[b] = db[i * np + b];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) cfma(y[a][b], xv[a], dv[b]);
            }
            cplx* yo = Ys + g * np * np;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) yo[(rb + a) * np + cb + b] = y[a][b];
        }
        __syncthreads();
        // traces: warp `g` round-robin over directions
        for (int g = warp; g < g_cnt; g += DET_THREADS / 32) {
            const cplx* y = Ys + g * np * np;
            cplx tr{0.0, 0.0}, sq{0.0, 0.0};
            for (int t = lane; t < n * n; t += 32) {
                int a = t / n, b = t - a * n;
                cplx yab = y[a * np + b];
                if (a == b) { tr.re += yab.re; tr.im += yab.im; }
                cfma(sq, yab, y[b * np + a]);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                tr.re += __shfl_xor_sync(0xffffffffu, tr.re, off);
                tr.im += __shfl_xor_sync(0xffffffffu, tr.im, off);
                sq.re += __shfl_xor_sync(0xffffffffu, sq.re, off);
                sq.im += __shfl_xor_sync(0xffffffffu, sq.im, off);
            }
            if (lane == 0) {
                tau_out[2 * (d0 + g)] = tr.re;
                tau_out[2 * (d0 + g) + 1] = tr.im;
                trsq.re += sq.re; trsq.im += sq.im;      // lane 0 of each warp holds a partial
            }
        }
    }
    {
        double ra, rb;
        block_sum2(lane == 0 ? trsq.re : 0.0, lane == 0 ? trsq.im : 0.0, ra, rb);
        if (tid == 0) {
            double* ts = sb.TRSQ + ((w * 2 + s) * D + k) * 2;
            ts[0] = ra; ts[1] = rb;
        }
    }
}

// ---------------------------------------------------------------------------
// Determinant kernel of the Laplacian sweep, second version: one CTA per (walker, spin, determinant).
//   * in-place Gauss-Jordan inverse X = M^-1 in shared memory (as det_kernel<true>)
//   * directions are processed in groups of G: the G derivative matrices dM_d are staged with cp.async
//     (no register round trip), every thread owns one 3x3 block position (rb, cb) of TWO directions
//     (X operands loaded once for both: 9 LDS.128 per 72 DFMA), the products Y_d = X dM_d are written
//     IN PLACE over the staged operands, and  sum_d tr(Y_d^2) = sum_{a,b} Y[a,b] Y[b,a]  is accumulated
//     per thread from its register block and the transposed block read back from shared memory; only
//     tr(Y_d) is reduced per direction (diagonal blocks).
// blockDim.x >= (G/2) * nb^2 work items (nb = np/3), a multiple of 32, at most 512.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_s(void* smem_dst, const void* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(512, 1) det_lap_kernel(const DsSys sys, const SlaterBufs sb, int G) {
    const DsDims& dm = sys.d;
    const int D = dm.D, NDp = dm.NDp, ND = dm.ND;
    const int k = blockIdx.x % D;
    const int s = (blockIdx.x / D) % 2;
    const long long w = blockIdx.x / (2 * D);
    const int n = s ? dm.n_dn : dm.n_up;
    const int np = ((n + 2) / 3) * 3;               // padded to the 3x3 register block
    const int nb = np / 3;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;

    extern __shared__ __align__(16) double smraw[];
    cplx* Xs = reinterpret_cast<cplx*>(smraw);      // [np][np] rows = orbital (after inversion), cols = electron
    cplx* colk = Xs + np * np;                      // [np]
    cplx* buf = colk + np;                          // [G][np][np]  dM_d (rows = electron, cols = orbital), then Y_d
    cplx* s_tau = buf + G * np * np;                // [G][nb] partial traces of the diagonal blocks
    __shared__ int piv[128];
    __shared__ int s_p;
    __shared__ double s_red[2 * 16 + 4];

    const cplx* mat = reinterpret_cast<const cplx*>(sb.MAT[s]) + (w * D + k) * (long long)n * n;
    for (int t = tid; t < np * np; t += nthr) {
        int r = t / np, c = t - r * np;
        Xs[t] = (r < n && c < n) ? mat[r * n + c] : cplx{0.0, 0.0};
    }
    if (np != n)                                     // padding of the staged operands stays zero for the whole kernel
        for (int t = tid; t < G * np * np; t += nthr) buf[t] = cplx{0.0, 0.0};
    __syncthreads();

    // first group of derivative matrices: in flight during the inversion
    const cplx* da = reinterpret_cast<const cplx*>(sb.DA[s]) + (w * D + k) * (long long)NDp * n * n;
    auto stage = [&](int d0, int g_cnt) {
        if (np == n) {
            const int tot = g_cnt * n * n;
            const cplx* src = da + (long long)d0 * n * n;
            for (int t = tid; t < tot; t += nthr) cp_async16_s(buf + t, src + t);
        } else {
            const int rows = g_cnt * n;              // one warp per matrix row
            for (int rr = warp; rr < rows; rr += nwarp) {
                const int g = rr / n, i = rr - g * n;
                const cplx* src = da + ((long long)(d0 + g) * n + i) * n;
                cplx* dst = buf + (g * np + i) * np;
                for (int o = lane; o < n; o += 32) cp_async16_s(dst + o, src + o);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    stage(0, min(G, ND));

    double logabs = 0.0;
    cplx phase{1.0, 0.0};
    for (int kk = 0; kk < n; ++kk) {
        if (tid < 32) {
            double best = -1.0;
            int bi = kk;
            for (int r = kk + tid; r < n; r += 32) {
                double v = cabs1(Xs[r * np + kk]);
                if (v > best) { best = v; bi = r; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, best, off);
                int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (tid == 0) { s_p = bi; piv[kk] = bi; }
        }
        __syncthreads();
        const int p = s_p;
        if (p != kk) {
            for (int c = tid; c < n; c += nthr) {
                cplx a = Xs[kk * np + c];
                Xs[kk * np + c] = Xs[p * np + c];
                Xs[p * np + c] = a;
            }
            phase = cplx{-phase.re, -phase.im};
        }
        __syncthreads();
        const cplx pv = Xs[kk * np + kk];
        {
            double a = hypot(pv.re, pv.im);
            logabs += log(a);
            phase = cmul(phase, cplx{pv.re / a, pv.im / a});
        }
        const cplx ipv = cinv(pv);
        for (int r = tid; r < n; r += nthr) colk[r] = Xs[r * np + kk];
        __syncthreads();
        for (int c = tid; c < n; c += nthr) {
            cplx a = (c == kk) ? cplx{1.0, 0.0} : Xs[kk * np + c];
            Xs[kk * np + c] = cmul(a, ipv);
        }
        __syncthreads();
        for (int t = tid; t < n * n; t += nthr) {
            int r = t / n, c = t - r * n;
            if (r == kk) continue;
            cplx f = colk[r];
            cplx a = (c == kk) ? cplx{0.0, 0.0} : Xs[r * np + c];
            cplx b = Xs[kk * np + c];
            a.re -= f.re * b.re - f.im * b.im;
            a.im -= f.re * b.im + f.im * b.re;
            Xs[r * np + c] = a;
        }
        __syncthreads();
    }
    if (tid == 0) {
        double* ld = sb.LOGDET + ((w * 2 + s) * D + k) * 3;
        ld[0] = logabs; ld[1] = phase.re; ld[2] = phase.im;
    }
    // undo the row pivoting: swap columns in reverse order
    for (int kk = n - 1; kk >= 0; --kk) {
        const int p = piv[kk];
        if (p != kk) {
            for (int r = tid; r < n; r += nthr) {
                cplx a = Xs[r * np + kk];
                Xs[r * np + kk] = Xs[r * np + p];
                Xs[r * np + p] = a;
            }
        }
        __syncthreads();
    }
    // now Xs[o][i] = (M^-1)[o,i]

    auto block_sum2 = [&](double a, double b, double& oa, double& ob) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, off);
            b += __shfl_xor_sync(0xffffffffu, b, off);
        }
        __syncthreads();
        if (lane == 0) { s_red[2 * warp] = a; s_red[2 * warp + 1] = b; }
        __syncthreads();
        oa = 0.0; ob = 0.0;
        for (int q = 0; q < nwarp; ++q) { oa += s_red[2 * q]; ob += s_red[2 * q + 1]; }
    };

    // tr(X lapM) = sum_{i,o} X[o,i] lapM[i,o]
    {
        const cplx* lm = reinterpret_cast<const cplx*>(sb.LAPM[s]) + (w * D + k) * (long long)n * n;
        cplx acc{0.0, 0.0};
        for (int t = tid; t < n * n; t += nthr) {
            int i = t / n, o = t - i * n;
            cfma(acc, Xs[o * np + i], lm[t]);
        }
        double ra, rb;
        block_sum2(acc.re, acc.im, ra, rb);
        if (tid == 0) {
            double* tl = sb.TRLAP + ((w * 2 + s) * D + k) * 2;
            tl[0] = ra; tl[1] = rb;
        }
    }

    // ---- direction groups ---------------------------------------------------
    double* tau_out = sb.TAU + ((w * 2 + s) * D + k) * (long long)NDp * 2;
    const int nb2 = nb * nb;
    const int n_items = (G >> 1) * nb2;
    const bool active = tid < n_items;
    int pi = 0, rb = 0, cb = 0;
    if (active) { pi = tid / nb2; int rem = tid - pi * nb2; rb = (rem / nb) * 3; cb = (rem - (rem / nb) * nb) * 3; }
    const int g0 = 2 * pi;
    const cplx* xa = Xs + rb * np;
    cplx* b0 = buf + (g0 * np) * np + cb;            // dM_{g0}[i][cb..]
    cplx* b1 = b0 + np * np;                         // dM_{g0+1}
    cplx sq{0.0, 0.0};
    for (int d0 = 0; d0 < ND; d0 += G) {
        const int g_cnt = min(G, ND - d0);
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        cplx y0[3][3], y1[3][3];
        if (active) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) { y0[a][b] = cplx{0.0, 0.0}; y1[a][b] = cplx{0.0, 0.0}; }
#pragma unroll 3
            for (int i = 0; i < n; ++i) {
                cplx xv[3], dv0[3], dv1[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) xv[a] = xa[a * np + i];
#pragma unroll
                for (int b = 0; b < 3; ++b) { dv0[b] = b0[i * np + b]; dv1[b] = b1[i * np + b]; }
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) { cfma(y0[a][b], xv[a], dv0[b]); cfma(y1[a][b], xv[a], dv1[b]); }
            }
        }
        __syncthreads();                             // every operand of this group has been read
        if (active) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    b0[(rb + a) * np + b] = y0[a][b];
                    b1[(rb + a) * np + b] = y1[a][b];
                }
            if (rb == cb) {
                s_tau[g0 * nb + rb / 3] = cplx{y0[0][0].re + y0[1][1].re + y0[2][2].re, y0[0][0].im + y0[1][1].im + y0[2][2].im};
                s_tau[(g0 + 1) * nb + rb / 3] = cplx{y1[0][0].re + y1[1][1].re + y1[2][2].re, y1[0][0].im + y1[1][1].im + y1[2][2].im};
            }
        }
        __syncthreads();
        if (active) {
            // sum_{a,b in block} Y[rb+a, cb+b] Y[cb+b, rb+a]; the transposed block sits at rows cb.., columns rb..
            const cplx* t0 = buf + (g0 * np + cb) * np + rb;
            const cplx* t1 = t0 + np * np;
            cplx q0{0.0, 0.0}, q1{0.0, 0.0};
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    cfma(q0, y0[a][b], t0[b * np + a]);
                    cfma(q1, y1[a][b], t1[b * np + a]);
                }
            if (g0 < g_cnt) { sq.re += q0.re; sq.im += q0.im; }
            if (g0 + 1 < g_cnt) { sq.re += q1.re; sq.im += q1.im; }
        }
        if (tid < g_cnt) {
            cplx tr{0.0, 0.0};
            for (int j = 0; j < nb; ++j) { cplx v = s_tau[tid * nb + j]; tr.re += v.re; tr.im += v.im; }
            tau_out[2 * (d0 + tid)] = tr.re;
            tau_out[2 * (d0 + tid) + 1] = tr.im;
        }
        if (d0 + G < ND) {
            __syncthreads();                         // transposed blocks and s_tau consumed: the buffer may be refilled
            stage(d0 + G, min(G, ND - d0 - G));
        }
    }
    {
        double ra, rb2;
        block_sum2(sq.re, sq.im, ra, rb2);
        if (tid == 0) {
            double* ts = sb.TRSQ + ((w * 2 + s) * D + k) * 2;
            ts[0] = ra; ts[1] = rb2;
        }
    }
}

// ---------------------------------------------------------------------------
// Combine determinants (logdet_matmul, network.py:395-427) and the kinetic energy.
// One warp per walker.
// ---------------------------------------------------------------------------
template <bool LAP>
__global__ void __launch_bounds__(128) combine_kernel(const DsSys sys, const SlaterBufs sb, int Wc, double* log_abs,
                                                      double* phase, double* ke_re, double* ke_im) {
    const DsDims& dm = sys.d;
    const int D = dm.D, NDp = dm.NDp, ND = dm.ND;
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= Wc) return;
    const double* ld = sb.LOGDET + w * 2 * D * 3;
    // log-sum-exp over determinants (every lane redundantly; D is small)
    double mx = -INFINITY;
    for (int k = 0; k < D; ++k) {
        double l = ld[k * 3] + ld[(D + k) * 3];
        mx = fmax(mx, l);
    }
    cplx tot{0.0, 0.0};
    for (int k = 0; k < D; ++k) {
        double l = ld[k * 3] + ld[(D + k) * 3];
        cplx ph = cmul(cplx{ld[k * 3 + 1], ld[k * 3 + 2]}, cplx{ld[(D + k) * 3 + 1], ld[(D + k) * 3 + 2]});
        double m = exp(l - mx);
        tot.re += ph.re * m; tot.im += ph.im * m;
    }
    if (lane == 0) {
        if (log_abs) log_abs[w] = log(hypot(tot.re, tot.im)) + mx;
        if (phase) phase[w] = atan2(tot.im, tot.re);
    }
    if (!LAP) return;
    const cplx itot = cinv(tot);
    cplx ke{0.0, 0.0};
    for (int k = 0; k < D; ++k) {
        double l = ld[k * 3] + ld[(D + k) * 3];
        cplx ph = cmul(cplx{ld[k * 3 + 1], ld[k * 3 + 2]}, cplx{ld[(D + k) * 3 + 1], ld[(D + k) * 3 + 2]});
        double m = exp(l - mx);
        cplx wk = cmul(cplx{ph.re * m, ph.im * m}, itot);
        const double* t0 = sb.TAU + ((w * 2 + 0) * D + k) * (long long)NDp * 2;
        const double* t1 = sb.TAU + ((w * 2 + 1) * D + k) * (long long)NDp * 2;
        cplx sq{0.0, 0.0};
        for (int d = lane; d < ND; d += 32) {
            cplx t{t0[2 * d] + t1[2 * d], t0[2 * d + 1] + t1[2 * d + 1]};
            cfma(sq, t, t);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sq.re += __shfl_xor_sync(0xffffffffu, sq.re, off);
            sq.im += __shfl_xor_sync(0xffffffffu, sq.im, off);
        }
        const double* tl0 = sb.TRLAP + ((w * 2 + 0) * D + k) * 2;
        const double* tl1 = sb.TRLAP + ((w * 2 + 1) * D + k) * 2;
        const double* ts0 = sb.TRSQ + ((w * 2 + 0) * D + k) * 2;
        const double* ts1 = sb.TRSQ + ((w * 2 + 1) * D + k) * 2;
        cplx per{tl0[0] + tl1[0] - ts0[0] - ts1[0] + sq.re, tl0[1] + tl1[1] - ts0[1] - ts1[1] + sq.im};
        cfma(ke, wk, per);
    }
    if (lane == 0) {
        ke_re[w] = -0.5 * ke.re;
        ke_im[w] = -0.5 * ke.im;
    }
}

}  // namespace

static size_t g_det_smem[2] = {0, 0};     // largest dynamic shared memory configured for det_kernel<false/true>

int ds_launch_etab(const DsSys& sys, const SlaterBufs& sb, int Wc, int npar_max, bool jets, cudaStream_t stream) {
    dim3 grid((unsigned)((long long)Wc * sys.d.N));
    if (jets) etab_kernel<true><<<grid, 256, 0, stream>>>(sys, sb, npar_max);
    else etab_kernel<false><<<grid, 256, 0, stream>>>(sys, sb, npar_max);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_orb_assemble(const DsSys& sys, const SlaterBufs& sb, int Wc, int npar_max, bool jets,
                           cudaStream_t stream) {
    dim3 grid((unsigned)((long long)Wc * sys.d.N));
    if (jets) assemble_kernel<true><<<grid, 256, 0, stream>>>(sys, sb, npar_max);
    else assemble_kernel<false><<<grid, 256, 0, stream>>>(sys, sb, npar_max);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_det(const DsSys& sys, const SlaterBufs& sb, int Wc, bool lap, cudaStream_t stream) {
    const int nmax = sys.d.n_up > sys.d.n_dn ? sys.d.n_up : sys.d.n_dn;
    DS_REQUIRE(nmax <= 128, "determinants larger than 128x128 are not supported (n_s=%d)", nmax);
    const int np = ((nmax + 2) / 3) * 3;
    const int nb = np / 3;
    dim3 grid((unsigned)((long long)Wc * 2 * sys.d.D));
    static const bool use_v1 = getenv("DS_DET_V1") && atoi(getenv("DS_DET_V1")) != 0;
    if (lap && !use_v1) {
        // det_lap_kernel: G directions per group, two per thread
        const int nb2 = nb * nb;
        int G = 2 * (256 / nb2 > 1 ? 256 / nb2 : 1);
        auto smem_of = [&](int g) { return (size_t)(np * np + np + g * np * np + g * nb) * sizeof(cplx); };
        while (G > 2 && smem_of(G) > 100 * 1024) G -= 2;
        const int nd_even = (sys.d.ND + 1) & ~1;
        if (G > nd_even) G = nd_even;
        const int items = (G / 2) * nb2;
        int threads = ((items > G ? items : G) + 31) & ~31;
        const size_t smem = smem_of(G);
        if (threads <= 512 && smem <= 226 * 1024) {
            static size_t cfg = 0;
            if (smem > cfg) {
                DS_CUDA_CHECK(cudaFuncSetAttribute(det_lap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                cfg = smem;
            }
            det_lap_kernel<<<grid, threads, smem, stream>>>(sys, sb, G);
            DS_CUDA_CHECK(cudaGetLastError());
            return 0;
        }
    }
    int G = 1;
    if (lap) {
        G = (DET_THREADS + nb * nb - 1) / (nb * nb);
        if (G < 1) G = 1;
        // keep the CTA under ~100 KB so two fit per SM when matrices are small
        while (G > 1 && (size_t)(np * np + np + 2 * G * np * np) * sizeof(cplx) > 100 * 1024) --G;
    }
    size_t smem = (size_t)(np * np + np + (lap ? 2 * G * np * np : 0)) * sizeof(cplx);
    DS_REQUIRE(smem <= 226 * 1024, "determinant kernel needs %zu bytes of shared memory", smem);
    if (smem > g_det_smem[lap ? 1 : 0]) {
        if (lap) DS_CUDA_CHECK(cudaFuncSetAttribute(det_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else DS_CUDA_CHECK(cudaFuncSetAttribute(det_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g_det_smem[lap ? 1 : 0] = smem;
    }
    if (lap) det_kernel<true><<<grid, DET_THREADS, smem, stream>>>(sys, sb, G);
    else det_kernel<false><<<grid, DET_THREADS, smem, stream>>>(sys, sb, G);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_det_inverse(const DsSys& sys, const SlaterBufs& sb, int Wc, cudaStream_t stream) {
    const int nmax = sys.d.n_up > sys.d.n_dn ? sys.d.n_up : sys.d.n_dn;
    DS_REQUIRE(nmax <= 128, "determinants larger than 128x128 are not supported (n_s=%d)", nmax);
    DS_REQUIRE(sb.XINV[0] && sb.XINV[1], "det_inverse: output buffers missing");
    const int np = ((nmax + 2) / 3) * 3;
    const size_t smem = (size_t)(np * np + np) * sizeof(cplx);
    DS_REQUIRE(smem <= 226 * 1024, "determinant kernel needs %zu bytes of shared memory", smem);
    if (smem > g_det_smem[1]) {
        DS_CUDA_CHECK(cudaFuncSetAttribute(det_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g_det_smem[1] = smem;
    }
    dim3 grid((unsigned)((long long)Wc * 2 * sys.d.D));
    det_kernel<true><<<grid, DET_THREADS, smem, stream>>>(sys, sb, 0);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_combine(const DsSys& sys, const SlaterBufs& sb, int Wc, bool lap, double* log_abs, double* phase,
                      double* ke_re, double* ke_im, cudaStream_t stream) {
    dim3 grid((unsigned)((Wc + 3) / 4));
    if (lap) combine_kernel<true><<<grid, 128, 0, stream>>>(sys, sb, Wc, log_abs, phase, ke_re, ke_im);
    else combine_kernel<false><<<grid, 128, 0, stream>>>(sys, sb, Wc, log_abs, phase, ke_re, ke_im);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

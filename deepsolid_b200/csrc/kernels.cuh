// Kernel launchers of the deepsolid_b200 library (definitions in the .cu files).
#pragma once
#include "ds_common.cuh"
#include "gemm_f64.cuh"

#define DS_MAX_LAYERS 4

// features.cu ---------------------------------------------------------------
struct FeatParams {
    const double* X;                 // [Wc, 3N]
    double *A0V, *A0L, *A0J;         // layer-0 operand matrices, ld = K0
    double* AV[DS_MAX_LAYERS];       // value rows of layer l >= 1 (ld K1); pair-mean columns written here
    double* AL[DS_MAX_LAYERS];       // Laplacian rows
    double* AJ[DS_MAX_LAYERS];       // Jacobian rows: the 2P pair-mean columns of row r live at AJ[l][r * ldj[l] + joff[l] ..]
    int ldj[DS_MAX_LAYERS], joff[DS_MAX_LAYERS];   // (ld K1, offset H inside the layer's operand; ld 2P, offset 0 when compact)
    double* RAE;                     // [(w*N+i)*A + a][DS_RAE_STRIDE] jets of the electron-atom distance and relative vector
    const double* Wp[DS_MAX_LAYERS]; // pair-stream weights [in x P]
    const double* bp[DS_MAX_LAYERS]; // pair-stream biases [P]
};
int ds_launch_features(const DsSys& sys, const FeatParams& fp, int Wc, bool jets, cudaStream_t stream);

// stream.cu -----------------------------------------------------------------
// spin-channel means of the one-electron stream: rows of the shared-mean operand
// GIN[w, d, s*C + c] = mean_{i in s} src[(w,i,d), c]; d = NDp -> value, NDp+1 -> Laplacian.
int ds_launch_means(const DsDims& dm, int Wc, int C, const double* AJ, int ldj, const double* AV,
                    const double* AL, int ldv, double* GIN, int ldgin, bool jets, cudaStream_t stream,
                    bool skip_jacobian_rows = false);

// layer-0 Jacobian rows (no GEMM): J1 = (1-T^2)(A0J.B + G), S = sum_d (A0J.B + G)^2
int ds_launch_l0_jac(const DsDims& dm, int Wc, const double* A0J, const double* B, const double* G, int ldg,
                     const double* T, int ldt, double* S, double* OJ, int ldc, cudaStream_t stream);

// slater.cu -----------------------------------------------------------------
struct SlaterBufs {
    const double* X;        // [Wc,3N]
    const double* RAE;      // electron-atom distance jets
    double* ETAB;           // complex [(w*N+i)][5][npar_max]
    const double* YV;       // complex raw orbital values   [(w*N+i)][npar_max]
    const double* YL;       // complex raw orbital Laplacians
    const double* YOWN;     // complex raw own-gradient     [(w*N+i)][3][npar_max]
    double* MAT[2];         // complex orbital matrices per spin [(w*D+k)][n_s][n_s]
    double* LAPM[2];        // complex Laplacian of the matrices
    double* DA[2];          // complex derivative matrices  [(w*D+k)][NDp][n_s][n_s]
    double* LOGDET;         // [w][2][D][3] : log|det|, cos(arg), sin(arg)
    double* TAU;            // complex [w][2][D][NDp]  tr(X dA_d)
    double* TRSQ;           // complex [w][2][D]       sum_d tr((X dA_d)^2)
    double* TRLAP;          // complex [w][2][D]       tr(X lapA)
    const double* env_pi[2];
    const double* env_sigma[2];     // [A][n_s*D]
    const double* klist[2];         // [n_s][3]
    double* XINV[2];        // complex inverse matrices [(w*D+k)][o][i] (parameter-gradient path only, else null)
};
int ds_launch_etab(const DsSys& sys, const SlaterBufs& sb, int Wc, int npar_max, bool jets, cudaStream_t stream);
// use_last_layer: add the shared spin-mean contribution GO (per spin, [Wc*rows x ldgo], rows = NDg with jets else 1) to YV / YL / YOWN / DA
int ds_launch_orb_mean_addend(const DsSys& sys, const SlaterBufs& sb, int Wc, int npar_max, const double* GO0, const double* GO1,
                              int ldgo, bool jets, cudaStream_t stream);
int ds_launch_orb_assemble(const DsSys& sys, const SlaterBufs& sb, int Wc, int npar_max, bool jets, cudaStream_t stream);
int ds_launch_det(const DsSys& sys, const SlaterBufs& sb, int Wc, bool lap, cudaStream_t stream);
int ds_launch_combine(const DsSys& sys, const SlaterBufs& sb, int Wc, bool lap, double* log_abs, double* phase,
                      double* ke_re, double* ke_im, cudaStream_t stream, double* gx_abs = nullptr,
                      double* gx_phase = nullptr);
// LOGDET and the inverse matrices XINV of every (walker, spin, determinant)
int ds_launch_det_inverse(const DsSys& sys, const SlaterBufs& sb, int Wc, cudaStream_t stream);

// grad.cu: reverse sweep of (log|psi|, phase) w.r.t. the parameters --------------------------------
struct GradBufs {
    // cotangents of the outputs, per walker (dloss = sum_w a_w dlog|psi_w| + b_w dphase_w)
    const double* cot_abs; const double* cot_phase;
    // alternatively the cotangent of the orbital matrices themselves (eval_mats, pretraining): complex (re, im)
    // interleaved in the layout of ds_orbitals, dloss = sum cot_re dRe(M) + cot_im dIm(M); null otherwise
    const double* cot_mats; long long cot_mats_stride;
    // orbital layer
    double* GY[2];              // [Wc*n_s][2 npar_s] cotangent of the raw orbital outputs, columns (re, im) interleaved
    double* g_pi[2];            // [A][npar_s]  accumulated
    double* g_sigma[2];
    // one-electron stream
    const double* T;            // tanh values of the layer [Wc*N][H]
    const double* GH;           // cotangent of the layer output [Wc*N][H]
    double* GZ;                 // cotangent of the pre-activation [Wc*N][H]
    double* GZS;                // per-walker sums of GZ [Wc][H]
    double* g_bias;             // [H] accumulated
    const double* GA; int lda;  // GZ . B_am^T [Wc*N][K]
    const double* GG; int ldgg; // GZS . B_g^T [Wc][2C]
    double* GHin;               // cotangent of the layer input (own columns) [Wc*N][C]
    double* GPM;                // pair-mean cotangents of the layer [Wc*N][2P]
    // pair stream
    const double* GPMl[DS_MAX_LAYERS];          // cotangents of the pair means of level l (null for l = 0)
    double* g_Wp[DS_MAX_LAYERS]; double* g_bp[DS_MAX_LAYERS];   // accumulated
    // Kronecker-factor statistics of the pair layers (ds_kfac_factors), accumulated: raw [32][32] Gram matrices of the
    // layer inputs (fact_A, with their column sums fact_As [32]) and of the pre-activation cotangents (fact_G)
    double* fact_A[DS_MAX_LAYERS]; double* fact_As[DS_MAX_LAYERS]; double* fact_G[DS_MAX_LAYERS];
};
int ds_launch_orb_grad(const DsSys& sys, const SlaterBufs& sb, const GradBufs& gb, int Wc, int npar_max, cudaStream_t stream);
int ds_launch_gz(const DsDims& dm, const GradBufs& gb, int Wc, bool residual, cudaStream_t stream);
int ds_launch_hin(const DsDims& dm, const GradBufs& gb, int Wc, int C, int K, bool residual, bool want_pm, cudaStream_t stream);
// fact = 0: weight / bias gradients; 1: Gram matrices of the layer inputs (forward only); 2: of the cotangents
int ds_launch_pair_grad(const DsSys& sys, const FeatParams& fp, const GradBufs& gb, int Wc, cudaStream_t stream, int fact = 0);
// rows [own C | spin means 2C | pair means K-C | 1 | 0] of a one-electron layer's input (network.py:305-332), ldx = 2C+K+2
int ds_launch_layer_input_rows(const double* Ain, int K, const double* GINV, int C, int N, long long rows, double* out,
                               cudaStream_t stream);
// rows of one spin channel of src [Wc*N][lds] (first `cols` columns) followed by (1, 0): out [Wc*ns][cols+2]
int ds_launch_spin_rows(const double* src, int lds, int cols, int N, int off_s, int ns, int Wc, double* out, cudaStream_t stream);
// [(pin+1) x (pin+1)] Gram matrix of (x, 1) from the raw [32][32] block, the column sums and the row count
int ds_launch_pair_fact_pack(const double* A32, const double* As, double count, int pin, double* out, cudaStream_t stream);
// dst[q(r)][q(c)] = src[r][c], q(2p+im) = im*np + p : (re, im)-interleaved -> (re block | im block) on both indices
int ds_launch_deinterleave2(const double* src, double* dst, int np, cudaStream_t stream);
// dst[r, c] (+)= src[r, c] for a [rows x cols] block (leading dimensions lds, ldd); deinterleave: see grad.cu
int ds_launch_copy2d(const double* src, int lds, double* dst, int ldd, int rows, int cols, cudaStream_t stream);
// dst[g][c] = sum over the rows_per_group rows of group g of src[.][c]
int ds_launch_group_rowsum(const double* src, int lds, int rows_per_group, int n_groups, int cols, double* dst, cudaStream_t stream);
int ds_launch_colsum_add(const double* src, int lds, int rows, int cols, double* dst, cudaStream_t stream);
int ds_launch_deinterleave(const double* src, double* dst, int rows, int np, cudaStream_t stream);

// ewald.cu ------------------------------------------------------------------
struct EwaldDev {
    int n_elec, n_atoms, dist_kind, n_g;
    double lat[9], inv[9], alpha;
    const double* atoms;        // [n_atoms*3]
    const double* charges;
    const double* mi_shifts;    // [27*3]
    const double* disp;         // [27*3]
    const double* gpoints;      // [n_g*3]
    const double* gweight;
    const double* ion_re;
    const double* ion_im;
    double ee_const, ei_const, ii_total;
};
int ds_launch_ewald(const EwaldDev& ew, const double* X, long long batch, double* ee, double* ei,
                    double* total_or_null, cudaStream_t stream);

// mcmc.cu -------------------------------------------------------------------
int ds_launch_propose(const DsLattice& sim, const double* x, double* x2, long long batch, int n3, double width,
                      const double* xi_or_null, unsigned long long seed, unsigned long long step,
                      cudaStream_t stream, int only_electron = -1);
int ds_launch_accept(double* x, const double* x2, double* lp, const double* lp2, long long batch, int n3,
                     const double* u_or_null, unsigned long long seed, unsigned long long step,
                     unsigned char* mask_or_null, double* n_accept, cudaStream_t stream);
int ds_launch_scale(double* dst, const double* src, double a, long long n, cudaStream_t stream);
// estimator.py plane-wave sums: out [batch][nq] complex (re, im); mode 0 = sum_i e^{iq.x_i}, 1 = e^{i sum_i q.x_i}
int ds_launch_rho_q(const double* X, long long batch, int n_elec, const double* Q, int nq, int mode, double* out,
                    cudaStream_t stream);
struct ds_ctx;
double* ds_ctx_scratch8(ds_ctx* c);          // 8 doubles of device scratch owned by the context
void ds_ctx_count_launches(ds_ctx* c, int n);
int ds_launch_stats(const double* ke_re, const double* ke_im, const double* ew, long long n, double* out6,
                    cudaStream_t stream);

// fp64-accurate GEMM on the 5th-generation tensor cores (tcgen05.mma kind::i8, TMEM
// accumulators, TMA-fed shared memory).
//
// tcgen05 has no f64 kind, so the Jacobian-sweep GEMMs  Z[r,n] = sum_k A[r,k] W[k,n]  are
// evaluated with error-free integer slices (Ozaki scheme I):
//     A[r,k] ~ sa[r] * sum_s dA_s[r,k] 256^-s ,   W[k,n] ~ sb[n] * sum_t dW_t[k,n] 256^-t ,
// dA_s, dW_t balanced base-256 digits (int8), S = 6 digits each (46 bits below the row /
// column maximum).  Digit products accumulate EXACTLY in int32 (|sum| < 2^26), and the
// diagonals g = s+t < 6 are recombined in fp64 in the epilogue:
//     Z[r,n] = sa[r] sb[n] sum_g 256^-g C_g[r,n],   C_g = sum_{s+t=g} dA_s . dW_t      (21 int8 GEMMs).
//
// The product is formed TRANSPOSED: the UMMA M side (128 TMEM lanes) holds 128 output channels
// n, the N side holds 64 rows r of A.  All six digit matrices of a K block are staged in one
// shared-memory stage, slice-major, so that ONE instruction  W_t x [A_0 .. A_{5-t}]  (N up to
// 256) adds W_t.A_s into the TMEM column block of diagonal s+t: 8 tcgen05.mma per K=32 step
// instead of 21.  TMEM: 6 diagonals x 64 columns of int32.  Epilogue thread = output channel,
// loop over columns = rows of A (consecutive directions d of one electron), so the
// sum over d of zJ^2, the shared-mean addend G[d,:], tanh factors and stores are all coalesced
// across lanes and sequential in a thread.
#pragma once
#include "ds_common.cuh"

#define OZ_S 6            // digits per operand
#define OZ_BK 64          // K bytes per pipeline stage (SWIZZLE_64B rows)
#define OZ_TM 128         // output channels per tile (UMMA M)
#define OZ_TN 64          // rows of A per tile (UMMA N per digit)

enum OzMode {
    OZ_PLAIN = 0,   // C[r, n] = Z
    OZ_JAC = 1,     // one-electron stream Jacobian rows (see gemm_f64.cuh GEMM_JAC)
    OZ_ORBJ = 2,    // orbital-layer Jacobian rows (see gemm_f64.cuh GEMM_ORBJ)
    OZ_VALUE = 3,   // one-electron stream value rows:     h' = res(tanh(Z + G_val + b)), tanh values kept in Tout
    OZ_LAP = 4,     // one-electron stream Laplacian rows: l' = res((1-t^2)(Z + G_lap) - 2t(1-t^2) S)
    OZ_JACD = 5     // OZ_JAC whose output IS the next GEMM's operand: the epilogue forms the int8 digits and row scales
                    // of [J' | pair-mean columns] itself (row maximum over both 128-channel CTAs of a cluster pair through
                    // distributed shared memory), reads the residual rows from the INPUT digits, and writes per-8-row
                    // partial sums of zJ^2 (fixed order: deterministic) -- no fp64 Jacobian ever reaches HBM
};

struct OzParams {
    // digits of A: int8 [rows][OZ_S][K] (row pitch OZ_S*K bytes), scales sa[rows]
    const signed char* Ad; const double* sa;
    // A rows are addressed as (group w, row-in-group q): physical row = w*gstride + goff + q, q < rpg.
    // Plain matrices: n_groups = 1, rpg = M.
    long long rpg, gstride, goff; int n_groups;
    // digits of W^T: int8 [N][OZ_S][K], scales sb[N]
    const signed char* Wd; const double* sb;
    int N, K;
    // outputs / epilogue operands (same meaning as GemmParams)
    double* C; int ldc;
    const double* G; int ldg;
    int n_elec, NDp, NDg;
    const double* T; int ldt;
    double* Tout;            // OZ_VALUE: tanh values written here (same leading dimension ldt)
    const double* colbias;   // OZ_VALUE: layer bias [N]
    double* S;
    const double* R; int ldr;
    const double* etab; int npar_max;
    int n_s, off_s, n_det;
    int n_orb, n_rows_mat, row0;   // orbitals per determinant (matrix columns), matrix rows, row of this channel's first electron
    double* DA; double* YOWN;
    int dbg;                 // probe only (DS_OZ_DBG in the stand-alone probe, DS_OZ_OPT on the production path): 1 = epilogue
                             // skips TMEM reads/stores, 2 = no MMA issue, 4 = no TMA; with a -DDS_OZ_PROF build: 32 = role
                             // clocks (scripts/oz_roles.py), 256 / 512 (+1024, 2048, 4096) = synthetic phase B (fp64 chain
                             // only / the loads and stores only, without the shared-mean / residual loads / stores), 8192 =
                             // no residual prefetch
    // Blocked row-contiguous digit layout [Rp/64 blocks][OZ_S][K][64 rows] (bytes; Rp = rows rounded up to 64): what
    // OZ_JACD writes -- a thread's 16 rows of one slice are one 128-bit store, a 64-row tile is one contiguous region --
    // and what the next GEMM reads as an MN-major operand.  bmn != 0: Ad is in this layout (Rp_in rows).
    int bmn; long long Rp_in;
    // OZ_JACD outputs: blocked digits [Rp_out/64][OZ_S][Kout][64] + scales of the rows [own N channels | npm pair-mean columns]
    signed char* Dout; double* sa_out; int Kout; long long Rp_out;
    const double* PM; int npm;   // fp64 pair-mean Jacobian rows [rows][npm] of the NEXT layer (null / 0: own columns only)
    double* SP;                  // [rows / 8][ldt] partial sums of zJ^2 over aligned groups of 8 rows
};

// digits + scales of `rows` rows of a row-major fp64 matrix (leading dimension lda, K columns)
int ds_launch_slice_rows(const double* A, int lda, long long rows, int K, signed char* Ad, double* sa,
                         cudaStream_t stream);
// the same digits for the Jacobian rows (w, i, d) of a layer, fused with the spin-channel means over i of
// the first C columns:  GIN[(w*NDg + d)*ldgin + s*C + c]  (rows d < NDp only)
int ds_launch_slice_means(const double* A, int lda, int K, int C, int n_walkers, int n_up, int n_elec, int NDp, int NDg,
                          signed char* Ad, double* sa, double* GIN, int ldgin, cudaStream_t stream);
// Bt[N][K] = B[K][N]^T
int ds_launch_transpose(const double* B, int K, int N, double* Bt, cudaStream_t stream);
int ds_launch_oz_gemm(const OzParams& p, int mode, bool residual, cudaStream_t stream);
int ds_oz_prof_read(unsigned long long* out, int reset);
// S[e][n] = sum_b SP[e * (NDp/8) + b][n], fixed order
int ds_launch_sp_reduce(const double* SP, int ld, long long n_elec_rows, int blocks_per_electron, int H, double* S,
                        cudaStream_t stream);
// spin-channel means of Jacobian rows given as row-contiguous digits [OZ_S][K][Rp]:
// GIN[(w*NDg + d)*ldgin + s*C + c] = mean_{i in s} A[(w,i,d), c], c < C
int ds_launch_means_digits(const signed char* Ad, const double* sa, int K, long long Rp, int C, int n_walkers, int n_up,
                           int n_elec, int NDp, int NDg, double* GIN, int ldgin, cudaStream_t stream);
// fp64 rows -> row-contiguous digits (probe / tests)
int ds_launch_slice_rows_mn(const double* A, int lda, long long rows, int K, long long Rp, signed char* Ad, double* sa,
                            cudaStream_t stream);

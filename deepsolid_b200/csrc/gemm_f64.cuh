// fp64 tensor-core GEMM (mma.sync m8n8k4 f64 -> DMMA on sm_100a) with the fused
// epilogues of the forward-Laplacian sweep.  tcgen05 has no f64 kind, so the fp64
// parity path runs on the legacy DMMA pipe; operands are staged with cp.async into
// padded (bank-conflict-free) shared-memory tiles, 3 stages deep.
#pragma once
#include "ds_common.cuh"

enum GemmMode {
    GEMM_PLAIN = 0,   // C = A.B (+ colbias)
    GEMM_VALUE = 1,   // one-electron stream, value rows:    h' = res(tanh(A.B + G_val + b))
    GEMM_JAC = 2,     // Jacobian rows (w,i,d):               J' = res((1-t^2)(A.B + G_d)), S += (A.B+G_d)^2
    GEMM_LAP = 3,     // Laplacian rows:                      l' = res((1-t^2)(A.B+G_lap) - 2t(1-t^2) S)
    GEMM_ORBJ = 4,    // orbital layer Jacobian rows: complexify, scale by envelope*phase, scatter to
                      // per-determinant matrices; raw own-electron rows kept for the product rule
    GEMM_TN = 5       // C (+)= A^T . B with A stored [K x M] row-major (weight gradients: reduction over rows);
                      // the row mapping (rpg, gstride, goff) applies to the K index of A only
};

struct GemmParams {
    const double* A; int lda;
    const double* B; int ldb;
    long long M; int N; int K;
    // logical row r -> physical A row: (r / rpg) * gstride + goff + r % rpg   (rpg<=0: identity)
    long long rpg, gstride, goff;
    double* C; int ldc;          // output (logical->physical row mapping as for A when cmap != 0)
    int cmap;
    int no_amap;                 // PLAIN: the row mapping applies to the rows of C only (A rows are compact)
    int accumulate;              // PLAIN / TN: C += A.B instead of C = A.B
    const double* colbias;       // [N] or null
    // one-electron stream epilogues
    const double* G;             // GOUT [Wc*NDg x ldg]
    int ldg;
    int n_elec, NDp, NDg;
    double* T;                   // tanh values [rows_e x ldt]
    int ldt;
    double* S;                   // sum_d zJ^2  [rows_e x ldt]
    const double* R; int ldr;    // residual input (same row indexing as C)
    // orbital epilogue
    const double* etab;          // complex E value table [(w*N+i) * (5*npar_max) + p] (re,im)
    int npar_max;                // stride unit of etab / yown
    int n_s, off_s, n_det;
    int n_orb, n_rows_mat, row0;   // orbitals per determinant (matrix columns), matrix rows, row of this channel's first electron
    double* DA;                  // complex [((w*D+k)*NDp+d)*n_s*n_s + i_s*n_s + o]
    double* YOWN;                // raw complex own rows [((w*N+i)*3+c)*npar_max + p]
};

int ds_launch_gemm(const GemmParams& p, int mode, bool residual, cudaStream_t stream);

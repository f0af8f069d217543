// Spin-channel means of the one-electron stream (network.py:322-330).
//
// construct_symmetric_features concatenates [h_i, mean_up h, mean_dn h, ...]; the two
// mean blocks are identical for every electron of a walker, so their contribution to
// the next layer (and to every Jacobian direction d) is computed ONCE per walker:
//   GOUT[w,d,:] = GIN[w,d,:] . W[C:3C]   with   GIN[w,d,s*C+c] = mean_{i in s} J[(w,i,d), c].
// Rows d = NDp / NDp+1 of GIN carry the value / Laplacian means.
#include "kernels.cuh"
#include <stdlib.h>

namespace {

__global__ void __launch_bounds__(256) means_kernel(DsDims dm, int C, const double* __restrict__ AJ, int ldj,
                                                    const double* __restrict__ AV, const double* __restrict__ AL,
                                                    int ldv, double* __restrict__ GIN, int ldgin, int tile0) {
    const int w = blockIdx.x;
    const int d0 = (tile0 + blockIdx.y) * 8;
    const int N = dm.N, NDp = dm.NDp, NDg = dm.NDg;
    const double inv_up = 1.0 / dm.n_up, inv_dn = 1.0 / dm.n_dn;
    for (int idx = threadIdx.x; idx < 8 * C; idx += blockDim.x) {
        const int dl = idx / C, c = idx - dl * C, d = d0 + dl;
        if (d >= NDg) continue;
        double su = 0.0, sd = 0.0;
        if (d < NDp) {
            const double* src = AJ + ((long long)w * N * NDp + d) * ldj + c;
            const long long stride = (long long)NDp * ldj;
            for (int i = 0; i < dm.n_up; ++i) su += src[i * stride];
            for (int i = dm.n_up; i < N; ++i) sd += src[i * stride];
        } else if (d == NDp || d == NDp + 1) {
            const double* src = (d == NDp ? AV : AL);
            if (src != nullptr) {
                src += (long long)w * N * ldv + c;
                for (int i = 0; i < dm.n_up; ++i) su += src[(long long)i * ldv];
                for (int i = dm.n_up; i < N; ++i) sd += src[(long long)i * ldv];
            }
        }
        double* dst = GIN + ((long long)w * NDg + d) * ldgin;
        dst[c] = su * inv_up;
        dst[C + c] = sd * inv_dn;
    }
}

// ---------------------------------------------------------------------------
// Layer-0 Jacobian rows without a GEMM.  The layer-0 operand rows have only K0 = 4A+8 columns, and the
// 4A own-feature columns of row (i, d) vanish unless d is a coordinate of electron i, so the
// "GEMM" is 8 (or K0) FMAs per output and the kernel is a pure HBM stream of the output rows:
//   z = A0J[(e,d),:] . B[:, n] + G[w, d, n];  S[e,n] = sum_d z^2;  J1[(e,d), n] = (1 - T[e,n]^2) z.
// One CTA per electron e = (w, i); thread = channel n, its weight column in registers (KT = K0) or
// read through L1 (KT = 0, any K0); the electron's operand rows are staged in shared memory.
// ---------------------------------------------------------------------------
template <int KT>
__global__ void __launch_bounds__(256) l0_jac_kernel(DsDims dm, const double* __restrict__ A0J,
                                                     const double* __restrict__ B, const double* __restrict__ G, int ldg,
                                                     const double* __restrict__ T, int ldt, double* __restrict__ S,
                                                     double* __restrict__ OJ, int ldc) {
    extern __shared__ __align__(16) double l0_sm[];          // [NDp][K0]
    const int K0 = dm.K0, C0 = dm.C0, H = dm.H, NDp = dm.NDp;
    const long long e = blockIdx.x;
    const int w = (int)(e / dm.N), i = (int)(e - (long long)w * dm.N);
    {
        const double2* src = reinterpret_cast<const double2*>(A0J + e * (long long)NDp * K0);
        double2* dst = reinterpret_cast<double2*>(l0_sm);
        for (int t = threadIdx.x; t < NDp * K0 / 2; t += blockDim.x) dst[t] = src[t];
    }
    __syncthreads();
    const double* g0 = G + (long long)w * dm.NDg * ldg;
    double* o0 = OJ + e * (long long)NDp * ldc;
    for (int n = threadIdx.x; n < H; n += blockDim.x) {
        double wc[KT > 0 ? KT : 1];
        if (KT > 0) {
#pragma unroll
            for (int k = 0; k < KT; ++k) wc[k] = B[(long long)k * H + n];
        }
        const double t = T[e * (long long)ldt + n];
        const double d1 = 1.0 - t * t;
        double sacc = 0.0;
        double zn[4];                                        // shared-mean rows of the next group, in flight
#pragma unroll
        for (int u = 0; u < 4; ++u) zn[u] = g0[(long long)u * ldg + n];
        for (int db = 0; db < NDp; db += 4) {                // NDp is a multiple of 8
            double z[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) z[u] = zn[u];
            if (db + 4 < NDp) {
#pragma unroll
                for (int u = 0; u < 4; ++u) zn[u] = g0[(long long)(db + 4 + u) * ldg + n];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int d = db + u;
                const double* a = l0_sm + d * K0;
                const bool own = (d / 3) == i;               // warp-uniform
                if (KT > 0) {
                    if (own) {
#pragma unroll
                        for (int k = 0; k < KT - 8; ++k) z[u] = fma(a[k], wc[k], z[u]);
                    }
#pragma unroll
                    for (int k = KT - 8; k < KT; ++k) z[u] = fma(a[k], wc[k], z[u]);
                } else {
                    for (int k = own ? 0 : C0; k < K0; ++k) z[u] = fma(a[k], B[(long long)k * H + n], z[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                sacc = fma(z[u], z[u], sacc);
                o0[(long long)(db + u) * ldc + n] = d1 * z[u];
            }
        }
        S[e * (long long)ldt + n] = sacc;
    }
}

// Two channels (n, n + H/2) per thread for the common K0 = 16 case: the operand row of a direction is read from
// shared memory once for both, which halves the per-output instruction count of this instruction-bound stream.
__global__ void __launch_bounds__(128) l0_jac2_kernel(DsDims dm, const double* __restrict__ A0J,
                                                      const double* __restrict__ B, const double* __restrict__ G, int ldg,
                                                      const double* __restrict__ T, int ldt, double* __restrict__ S,
                                                      double* __restrict__ OJ, int ldc) {
    constexpr int KT = 16;
    extern __shared__ __align__(16) double l0_sm[];          // [NDp][16]
    const int H = dm.H, NDp = dm.NDp, H2 = dm.H >> 1, n_up = dm.n_up;
    const long long e = blockIdx.x;
    const int w = (int)(e / dm.N), i = (int)(e - (long long)w * dm.N);
    {
        const double2* src = reinterpret_cast<const double2*>(A0J + e * (long long)NDp * KT);
        double2* dst = reinterpret_cast<double2*>(l0_sm);
        for (int t = threadIdx.x; t < NDp * KT / 2; t += blockDim.x) dst[t] = src[t];
    }
    __syncthreads();
    const double* g0 = G + (long long)w * dm.NDg * ldg;
    double* o0 = OJ + e * (long long)NDp * ldc;
    for (int n = threadIdx.x; n < H2; n += blockDim.x) {
        const int n1 = n + H2;
        // pair-mean weights (every direction) stay in registers; the own-feature weights (3 of 3N directions) come from L1
        double wa[8], wb[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { wa[k] = B[(long long)(k + KT - 8) * H + n]; wb[k] = B[(long long)(k + KT - 8) * H + n1]; }
        const double ta = T[e * (long long)ldt + n], tb = T[e * (long long)ldt + n1];
        const double da = 1.0 - ta * ta, db_ = 1.0 - tb * tb;
        double sa = 0.0, sb = 0.0;
        double zna[4], znb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { zna[u] = g0[(long long)u * ldg + n]; znb[u] = g0[(long long)u * ldg + n1]; }
        for (int d0 = 0; d0 < NDp; d0 += 4) {
            double za[4], zb[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { za[u] = zna[u]; zb[u] = znb[u]; }
            if (d0 + 4 < NDp) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    zna[u] = g0[(long long)(d0 + 4 + u) * ldg + n];
                    znb[u] = g0[(long long)(d0 + 4 + u) * ldg + n1];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int d = d0 + u;
                const double* a = l0_sm + d * KT;
                if ((d / 3) == i) {                          // warp-uniform: own-feature columns only for own directions
#pragma unroll
                    for (int k = 0; k < KT - 8; ++k) {
                        const double av = a[k];
                        za[u] = fma(av, __ldg(B + (long long)k * H + n), za[u]);
                        zb[u] = fma(av, __ldg(B + (long long)k * H + n1), zb[u]);
                    }
                }
                // pair-mean columns: a direction of another electron j moves only the mean over j's own spin channel
                // (4 of the 8 columns are zero); own directions move both (minus the sums over all partners)
                if ((d / 3) == i) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) { const double av = a[KT - 8 + k]; za[u] = fma(av, wa[k], za[u]); zb[u] = fma(av, wb[k], zb[u]); }
                } else if (d / 3 < n_up) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) { const double av = a[KT - 8 + k]; za[u] = fma(av, wa[k], za[u]); zb[u] = fma(av, wb[k], zb[u]); }
                } else {
#pragma unroll
                    for (int k = 4; k < 8; ++k) { const double av = a[KT - 8 + k]; za[u] = fma(av, wa[k], za[u]); zb[u] = fma(av, wb[k], zb[u]); }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                sa = fma(za[u], za[u], sa);
                sb = fma(zb[u], zb[u], sb);
                __stcs(o0 + (long long)(d0 + u) * ldc + n, da * za[u]);      // streamed: next touched by the digit pass
                __stcs(o0 + (long long)(d0 + u) * ldc + n1, db_ * zb[u]);
            }
        }
        S[e * (long long)ldt + n] = sa;
        S[e * (long long)ldt + n1] = sb;
    }
}

}  // namespace

int ds_launch_l0_jac(const DsDims& dm, int Wc, const double* A0J, const double* B, const double* G, int ldg,
                     const double* T, int ldt, double* S, double* OJ, int ldc, cudaStream_t stream) {
    const size_t smem = (size_t)dm.NDp * dm.K0 * sizeof(double);
    DS_REQUIRE(smem <= 200 * 1024, "layer-0 Jacobian kernel: operand rows of one electron need %zu bytes of shared memory", smem);
    int threads = (dm.H + 31) & ~31;
    if (threads > 256) threads = 256;
    dim3 grid((unsigned)((long long)Wc * dm.N));
    static size_t cfg[2] = {0, 0};
    static const bool one_ch = getenv("DS_L0_ONE") && atoi(getenv("DS_L0_ONE")) != 0;
    if (dm.K0 == 16 && dm.C0 == 8 && (dm.H & 63) == 0 && !one_ch) {
        static size_t cfg2 = 0;
        if (smem > 48 * 1024 && smem > cfg2) {
            DS_CUDA_CHECK(cudaFuncSetAttribute(l0_jac2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cfg2 = smem;
        }
        int t2 = dm.H / 2;
        if (t2 > 128) t2 = 128;
        l0_jac2_kernel<<<grid, t2, smem, stream>>>(dm, A0J, B, G, ldg, T, ldt, S, OJ, ldc);
    } else if (dm.K0 == 16) {
        if (smem > 48 * 1024 && smem > cfg[0]) {
            DS_CUDA_CHECK(cudaFuncSetAttribute(l0_jac_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cfg[0] = smem;
        }
        l0_jac_kernel<16><<<grid, threads, smem, stream>>>(dm, A0J, B, G, ldg, T, ldt, S, OJ, ldc);
    } else {
        if (smem > 48 * 1024 && smem > cfg[1]) {
            DS_CUDA_CHECK(cudaFuncSetAttribute(l0_jac_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cfg[1] = smem;
        }
        l0_jac_kernel<0><<<grid, threads, smem, stream>>>(dm, A0J, B, G, ldg, T, ldt, S, OJ, ldc);
    }
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_means(const DsDims& dm, int Wc, int C, const double* AJ, int ldj, const double* AV,
                    const double* AL, int ldv, double* GIN, int ldgin, bool jets, cudaStream_t stream,
                    bool skip_jacobian_rows) {
    int ntiles = dm.NDg / 8;
    int tile0 = 0;
    if (!jets || skip_jacobian_rows) { tile0 = dm.NDp / 8; ntiles = 1; }     // only the value (/ Laplacian) rows
    dim3 grid(Wc, ntiles);
    means_kernel<<<grid, 256, 0, stream>>>(dm, C, AJ, ldj, AV, jets ? AL : nullptr, ldv, GIN, ldgin, tile0);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

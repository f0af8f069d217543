// Spin-channel means of the one-electron stream (network.py:322-330).
//
// construct_symmetric_features concatenates [h_i, mean_up h, mean_dn h, ...]; the two
// mean blocks are identical for every electron of a walker, so their contribution to
// the next layer (and to every Jacobian direction d) is computed ONCE per walker:
//   GOUT[w,d,:] = GIN[w,d,:] . W[C:3C]   with   GIN[w,d,s*C+c] = mean_{i in s} J[(w,i,d), c].
// Rows d = NDp / NDp+1 of GIN carry the value / Laplacian means.
#include "kernels.cuh"

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

}  // namespace

int ds_launch_means(const DsDims& dm, int Wc, int C, const double* AJ, int ldj, const double* AV,
                    const double* AL, int ldv, double* GIN, int ldgin, bool jets, cudaStream_t stream) {
    int ntiles = dm.NDg / 8;
    int tile0 = 0;
    if (!jets) { tile0 = dm.NDp / 8; ntiles = 1; }
    dim3 grid(Wc, ntiles);
    means_kernel<<<grid, 256, 0, stream>>>(dm, C, AJ, ldj, AV, jets ? AL : nullptr, ldv, GIN, ldgin, tile0);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

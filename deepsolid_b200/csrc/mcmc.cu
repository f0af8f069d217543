// Metropolis proposal / accept kernels (qmc.py:153-224, distance.py:144-163) and the
// energy statistics reduction (train.py:74-80).
#include "kernels.cuh"

namespace {

// Philox4x32-10 (Salmon et al. 2011), counter-based: one call per (step, element).
__device__ __forceinline__ void philox4x32_10(unsigned c[4], unsigned k0, unsigned k1) {
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
        unsigned hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
        unsigned n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += W0; k1 += W1;
    }
}
// two uniforms in (0,1) with 53 / 32 random bits
__device__ __forceinline__ void philox_uniform2(unsigned long long seed, unsigned long long stream_id,
                                                unsigned long long idx, double& u1, double& u2) {
    unsigned c[4] = {(unsigned)idx, (unsigned)(idx >> 32), (unsigned)stream_id, (unsigned)(stream_id >> 32)};
    philox4x32_10(c, (unsigned)seed, (unsigned)(seed >> 32));
    unsigned long long a = ((unsigned long long)c[0] << 21) ^ (unsigned long long)(c[1] >> 11);   // 53 bits
    u1 = ((double)a + 0.5) * (1.0 / 9007199254740992.0);
    u2 = ((double)c[2] + 0.5) * (1.0 / 4294967296.0);
}

// only_electron < 0: all-electron move (mh_update, qmc.py:192); otherwise only that electron of every walker is
// displaced (mh_one_electron_update, qmc.py:270-273; noise xi then has shape (batch, 3)) while every electron
// is re-wrapped, as distance.enforce_pbc does to the whole configuration.
__global__ void __launch_bounds__(256) propose_kernel(const DsLattice sim, const double* __restrict__ x,
                                                      double* __restrict__ x2, long long n_elec_total, int n3,
                                                      double width, const double* __restrict__ xi,
                                                      unsigned long long seed, unsigned long long step, int only_electron) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // electron index over the batch
    if (t >= n_elec_total) return;
    const int n_el = n3 / 3;
    const long long b = t / n_el;
    const bool moved = only_electron < 0 || (int)(t - b * n_el) == only_electron;
    double p[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const long long idx = 3 * t + c;
        const long long nidx = only_electron < 0 ? idx : 3 * b + c;           // index into the noise array
        double z = 0.0;
        if (moved) {
            if (xi) {
                z = xi[nidx];
            } else {
                double u1, u2;
                philox_uniform2(seed, 2 * step, (unsigned long long)nidx, u1, u2);
                z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
            }
        }
        p[c] = moved ? __dadd_rn(x[idx], __dmul_rn(width, z)) : x[idx];      // x1 + stddev * N(0,1), no FMA contraction
    }
    double o[3];
    ds_wrap(sim, p, o);                                     // distance.enforce_pbc: divmod(frac, 1)
    x2[3 * t] = o[0]; x2[3 * t + 1] = o[1]; x2[3 * t + 2] = o[2];
    (void)n3;
}

// one warp per walker: cond = (lp2 - lp1) > log(u); select; count.
__global__ void __launch_bounds__(256) accept_kernel(double* __restrict__ x, const double* __restrict__ x2,
                                                     double* __restrict__ lp, const double* __restrict__ lp2,
                                                     long long batch, int n3, const double* __restrict__ u,
                                                     unsigned long long seed, unsigned long long step,
                                                     unsigned char* __restrict__ mask, double* __restrict__ n_accept) {
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    if (b < batch) {
        int cond = 0;
        if (lane == 0) {
            double uu;
            if (u) uu = u[b];
            else { double u2; philox_uniform2(seed, 2 * step + 1, (unsigned long long)b, uu, u2); }
            const double l1 = lp[b], l2 = lp2[b];
            cond = (l2 - l1) > log(uu);
            if (cond) { lp[b] = l2; atomicAdd(&cnt, 1); }
            if (mask) mask[b] = (unsigned char)cond;
        }
        cond = __shfl_sync(0xffffffffu, cond, 0);
        if (cond)
            for (int t = lane; t < n3; t += 32) x[b * n3 + t] = x2[b * n3 + t];
    }
    __syncthreads();
    if (threadIdx.x == 0 && cnt > 0) atomicAdd(n_accept, (double)cnt);
}

__global__ void scale_kernel(double* __restrict__ dst, const double* __restrict__ src, double a, long long n) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[t] = a * src[t];
}

// one block, fixed summation order: the statistics are bit-reproducible from run to run (the reference is deterministic)
__global__ void __launch_bounds__(1024) stats_kernel(const double* __restrict__ ke_re, const double* __restrict__ ke_im,
                                                     const double* __restrict__ ew, long long n, double* __restrict__ out6) {
    double v[5] = {0, 0, 0, 0, 0};
    for (long long t = threadIdx.x; t < n; t += blockDim.x) {
        double re = ke_re[t] + ew[t], im = ke_im[t];
        v[0] += re; v[1] += im; v[2] += re * re + im * im; v[3] += ke_re[t]; v[4] += ew[t];
    }
    __shared__ double red[5][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
        if (lane == 0) red[q][warp] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0.0;
        for (int w = 0; w < 32; ++w) s += red[threadIdx.x][w];
        out6[threadIdx.x] = s;
    }
    if (threadIdx.x == 5) out6[5] = (double)n;
}

}  // namespace

int ds_launch_propose(const DsLattice& sim, const double* x, double* x2, long long batch, int n3, double width,
                      const double* xi_or_null, unsigned long long seed, unsigned long long step,
                      cudaStream_t stream, int only_electron) {
    long long ne = batch * (n3 / 3);
    if (ne <= 0) return 0;
    propose_kernel<<<(unsigned)((ne + 255) / 256), 256, 0, stream>>>(sim, x, x2, ne, n3, width, xi_or_null, seed, step,
                                                                  only_electron);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_accept(double* x, const double* x2, double* lp, const double* lp2, long long batch, int n3,
                     const double* u_or_null, unsigned long long seed, unsigned long long step,
                     unsigned char* mask_or_null, double* n_accept, cudaStream_t stream) {
    if (batch <= 0) return 0;
    accept_kernel<<<(unsigned)((batch + 7) / 8), 256, 0, stream>>>(x, x2, lp, lp2, batch, n3, u_or_null, seed, step,
                                                                  mask_or_null, n_accept);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_scale(double* dst, const double* src, double a, long long n, cudaStream_t stream) {
    if (n <= 0) return 0;
    scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(dst, src, a, n);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_stats(const double* ke_re, const double* ke_im, const double* ew, long long n, double* out6,
                    cudaStream_t stream) {
    DS_CUDA_CHECK(cudaMemsetAsync(out6, 0, 6 * sizeof(double), stream));
    if (n <= 0) return 0;
    stats_kernel<<<1, 1024, 0, stream>>>(ke_re, ke_im, ew, n, out6);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

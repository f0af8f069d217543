// Per-walker Ewald energy (ewaldsum.py:138-191) with the minimal-image conventions of
// distance.MinimalImageDistance (distance.py:70-141).  One CTA per walker; fp64 erfc
// and sincos; the G table streams from L2.
#include "kernels.cuh"

namespace {

constexpr int EW_THREADS = 256;

__device__ __forceinline__ void min_image(const EwaldDev& ew, const double* __restrict__ shifts, double d[3]) {
    if (ew.dist_kind == 0) {                 // distance.py:110-128
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double L = ew.lat[k * 3 + k];
            double r = d[k] + 0.5 * L;
            d[k] = (r - floor(r / L) * L) - 0.5 * L;
        }
    } else if (ew.dist_kind == 1) {          // distance.py:91-108
        double f[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double s = d[0] * ew.inv[0 * 3 + k] + d[1] * ew.inv[1 * 3 + k] + d[2] * ew.inv[2 * 3 + k] + 0.5;
            f[k] = (s - floor(s)) - 0.5;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) d[k] = f[0] * ew.lat[0 * 3 + k] + f[1] * ew.lat[1 * 3 + k] + f[2] * ew.lat[2 * 3 + k];
    } else {                                 // distance.py:70-89 (first minimum wins, like argmin)
        double best = INFINITY;
        int bi = 0;
        for (int t = 0; t < 27; ++t) {
            double a = d[0] + shifts[3 * t], b = d[1] + shifts[3 * t + 1], c = d[2] + shifts[3 * t + 2];
            double r = sqrt(a * a + b * b + c * c);
            if (r < best) { best = r; bi = t; }
        }
        d[0] += shifts[3 * bi]; d[1] += shifts[3 * bi + 1]; d[2] += shifts[3 * bi + 2];
    }
}

__device__ __forceinline__ double real_sum(const EwaldDev& ew, const double* __restrict__ disp, const double d[3]) {
    double acc = 0.0;
    for (int t = 0; t < 27; ++t) {
        double a = d[0] + disp[3 * t], b = d[1] + disp[3 * t + 1], c = d[2] + disp[3 * t + 2];
        double r = sqrt(a * a + b * b + c * c);
        acc += erfc(ew.alpha * r) / r;
    }
    return acc;
}

__global__ void __launch_bounds__(EW_THREADS) ewald_kernel(const EwaldDev ew, const double* __restrict__ X,
                                                           double* __restrict__ ee_out, double* __restrict__ ei_out,
                                                           double* __restrict__ tot_out) {
    extern __shared__ double sm[];
    const int N = ew.n_elec, A = ew.n_atoms;
    double* x = sm;                 // [3N]
    double* shifts = x + 3 * N;     // [81]
    double* disp = shifts + 81;     // [81]
    __shared__ double red[3][EW_THREADS / 32];
    const long long b = blockIdx.x;
    const int tid = threadIdx.x;
    for (int t = tid; t < 3 * N; t += EW_THREADS) x[t] = X[b * 3 * N + t];
    for (int t = tid; t < 81; t += EW_THREADS) { shifts[t] = ew.mi_shifts[t]; disp[t] = ew.disp[t]; }
    __syncthreads();

    double ee = 0.0, ei = 0.0;
    // electron-ion real space
    for (int t = tid; t < N * A; t += EW_THREADS) {
        int i = t / A, a = t - i * A;
        double d[3] = {x[3 * i] - ew.atoms[3 * a], x[3 * i + 1] - ew.atoms[3 * a + 1], x[3 * i + 2] - ew.atoms[3 * a + 2]};
        min_image(ew, shifts, d);
        ei -= ew.charges[a] * real_sum(ew, disp, d);
    }
    // electron-electron real space, i < j
    const int npair = N * (N - 1) / 2;
    for (int t = tid; t < npair; t += EW_THREADS) {
        // unrank t -> (i, j), i < j, row-major over the strict upper triangle
        int i = (int)((2.0 * N - 1.0 - sqrt((2.0 * N - 1.0) * (2.0 * N - 1.0) - 8.0 * t)) * 0.5);
        while (i > 0 && (long long)i * (2 * N - i - 1) / 2 > t) --i;
        while ((long long)(i + 1) * (2 * N - i - 2) / 2 <= t) ++i;
        int j = t - (int)((long long)i * (2 * N - i - 1) / 2) + i + 1;
        double d[3] = {x[3 * i] - x[3 * j], x[3 * i + 1] - x[3 * j + 1], x[3 * i + 2] - x[3 * j + 2]};
        min_image(ew, shifts, d);
        ee += real_sum(ew, disp, d);
    }
    // reciprocal space
    for (int g = tid; g < ew.n_g; g += EW_THREADS) {
        const double g0 = ew.gpoints[3 * g], g1 = ew.gpoints[3 * g + 1], g2 = ew.gpoints[3 * g + 2];
        double sc = 0.0, ss = 0.0;
        for (int i = 0; i < N; ++i) {
            double s, c;
            sincos(g0 * x[3 * i] + g1 * x[3 * i + 1] + g2 * x[3 * i + 2], &s, &c);
            sc += c; ss += s;
        }
        const double wg = ew.gweight[g];
        ee += wg * (ss * ss + sc * sc);
        ei += 2.0 * wg * (-ew.ion_re[g] * sc - ew.ion_im[g] * ss);
    }
    // block reduction
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ee += __shfl_xor_sync(0xffffffffu, ee, off);
        ei += __shfl_xor_sync(0xffffffffu, ei, off);
    }
    if (lane == 0) { red[0][warp] = ee; red[1][warp] = ei; }
    __syncthreads();
    if (tid == 0) {
        double see = 0.0, sei = 0.0;
        for (int q = 0; q < EW_THREADS / 32; ++q) { see += red[0][q]; sei += red[1][q]; }
        see += ew.ee_const; sei += ew.ei_const;
        if (ee_out) ee_out[b] = see;
        if (ei_out) ei_out[b] = sei;
        if (tot_out) tot_out[b] = see + sei + ew.ii_total;
    }
}

}  // namespace

int ds_launch_ewald(const EwaldDev& ew, const double* X, long long batch, double* ee, double* ei,
                    double* total_or_null, cudaStream_t stream) {
    if (batch <= 0) return 0;
    size_t smem = (size_t)(3 * ew.n_elec + 162) * sizeof(double);
    ewald_kernel<<<(unsigned)batch, EW_THREADS, smem, stream>>>(ew, X, ee, ei, total_or_null);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}


// ---------------------------------------------------------------------------
// Observables of estimator.py: plane-wave sums over the electrons of each walker.
//   mode 0 (structure factor, estimator.py:68-73):  out[b][k] = sum_i exp(i q_k . x_bi)
//   mode 1 (complex polarisation, estimator.py:27-31): out[b][k] = exp(i sum_i q_k . x_bi)
// One thread per (walker, q-vector); positions are read through L1 (every q of a walker re-reads the same 3N doubles).
// ---------------------------------------------------------------------------
namespace {
__global__ void rho_q_kernel(const double* __restrict__ X, long long batch, int n_elec, const double* __restrict__ Q, int nq,
                             int mode, double* __restrict__ out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= batch * nq) return;
    const long long b = t / nq;
    const int k = (int)(t - b * nq);
    const double q0 = Q[3 * k], q1 = Q[3 * k + 1], q2 = Q[3 * k + 2];
    const double* x = X + b * 3 * n_elec;
    double re = 0.0, im = 0.0, tot = 0.0;
    for (int i = 0; i < n_elec; ++i) {
        const double d = q0 * x[3 * i] + q1 * x[3 * i + 1] + q2 * x[3 * i + 2];
        if (mode == 0) {
            double sn, cs;
            sincos(d, &sn, &cs);
            re += cs; im += sn;
        } else {
            tot += d;
        }
    }
    if (mode != 0) sincos(tot, &im, &re);
    out[2 * t] = re; out[2 * t + 1] = im;
}
}  // namespace

int ds_launch_rho_q(const double* X, long long batch, int n_elec, const double* Q, int nq, int mode, double* out,
                    cudaStream_t stream) {
    const long long n = batch * nq;
    if (n <= 0) return 0;
    rho_q_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(X, batch, n_elec, Q, nq, mode, out);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Are the fp64 FMA pipe and the fp64 tensor path (DMMA, mma.sync m8n8k4) separate execution resources on B200?
// Three kernels of the same shape: every warp issues DFMAs, every warp issues DMMAs, half of the warps each.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes scripts/micro/fp64_pipes.cu && ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode 0: DFMA only, 1: DMMA only, 2: even warps DFMA / odd warps DMMA
__global__ void __launch_bounds__(256) pipes(int mode, int iters, double* out) {
    const int warp = threadIdx.x >> 5;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = i;
    const bool fma_warp = mode == 0 || (mode == 2 && (warp & 1) == 0);
    if (fma_warp) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);          // 16 independent DFMA per iteration
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) dmma884(c[i], c[i + 1], a, b);  // 8 independent DMMA per iteration
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 1.2345) out[0] = s;
}

int main() {
    double* out; cudaMalloc(&out, 8);
    const int iters = 20000, blocks = 148 * 2;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep)
    for (int mode = 0; mode < 3; ++mode) {
        pipes<<<blocks, 256>>>(mode, 100, out);
        cudaEventRecord(e0);
        pipes<<<blocks, 256>>>(mode, iters, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double warps = (double)blocks * 8;
        const double fma_warps = mode == 0 ? warps : mode == 1 ? 0 : warps / 2, mma_warps = warps - fma_warps;
        const double flop = fma_warps * iters * 16 * 32 * 2.0 + mma_warps * iters * 8 * 512.0;
        printf("mode %d (%s): %.3f ms  %.2f TFLOP/s fp64  (DFMA warps %.0f, DMMA warps %.0f)\n", mode,
               mode == 0 ? "DFMA only" : mode == 1 ? "DMMA only" : "half / half", ms, flop / ms * 1e-9, fma_warps, mma_warps);
    }
    return 0;
}

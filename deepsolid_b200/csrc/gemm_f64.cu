// fp64 DMMA GEMM with fused forward-Laplacian epilogues.  See gemm_f64.cuh.
#include "gemm_f64.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 3;
constexpr int AS_LD = BK + 4;     // 20 doubles: conflict-free 64-bit fragment loads
constexpr int BS_LD = BN + 4;     // 132 doubles
constexpr int A_STAGE = BM * AS_LD;
constexpr int B_STAGE = BK * BS_LD;
constexpr int SMEM_BYTES = STAGES * (A_STAGE + B_STAGE) * (int)sizeof(double);
constexpr int NTHREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ long long map_row(long long r, long long rpg, long long gstride, long long goff) {
    if (rpg <= 0) return r;
    long long g = r / rpg;
    return g * gstride + goff + (r - g * rpg);
}

template <int MODE, bool RES>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_f64_kernel(const GemmParams p) {
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + STAGES * A_STAGE;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const long long row0 = (long long)blockIdx.x * BM;
    const int col0 = blockIdx.y * BN;

    // ---- global->shared assignment ---------------------------------------
    constexpr bool TA = (MODE == GEMM_TN);
    const double* a_src[4];
    bool a_ok[4];
    int a_dst[4], a_col[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int idx = tid + q * NTHREADS;
        if (TA) {       // A^T tile [BK][BM+4]: 16 k-rows of 128 output rows m, two m per 16-byte copy
            int kr = idx >> 6, ch = idx & 63;
            long long m = row0 + ch * 2;
            a_ok[q] = m < p.M;
            a_src[q] = p.A + (a_ok[q] ? m : 0);
            a_dst[q] = kr * BS_LD + ch * 2;
            a_col[q] = kr;
        } else {
            int r = idx >> 3, ch = idx & 7;
            long long gr = row0 + r;
            a_ok[q] = gr < p.M;
            long long pr = (a_ok[q] && !p.no_amap) ? map_row(gr, p.rpg, p.gstride, p.goff) : (a_ok[q] ? gr : 0);
            a_src[q] = p.A + pr * (long long)p.lda + ch * 2;
            a_dst[q] = r * AS_LD + ch * 2;
            a_col[q] = ch * 2;
        }
    }
    const double* b_src[4];
    bool b_ok[4];
    int b_dst[4], b_kr[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        int idx = tid + q * NTHREADS;
        int kr = idx >> 6, ch = idx & 63;
        int col = col0 + ch * 2;
        b_ok[q] = col < p.N;
        b_src[q] = p.B + (long long)kr * p.ldb + (b_ok[q] ? col : 0);
        b_dst[q] = kr * BS_LD + ch * 2;
        b_kr[q] = kr;
    }
    auto load_stage = [&](int stage, int k0) {
        double* as = As + stage * A_STAGE;
        double* bs = Bs + stage * B_STAGE;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (TA) {
                const int k = k0 + a_col[q];
                const bool ok = a_ok[q] && k < p.K;
                const long long pk = ok ? map_row(k, p.rpg, p.gstride, p.goff) : 0;
                cp_async16(as + a_dst[q], a_src[q] + pk * (long long)p.lda, ok);
            } else {
                cp_async16(as + a_dst[q], a_src[q] + k0, a_ok[q] && (k0 + a_col[q] < p.K));
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            cp_async16(bs + b_dst[q], b_src[q] + (long long)k0 * p.ldb, b_ok[q] && (k0 + b_kr[q] < p.K));
    };

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // split-K (GEMM_TN only: tiny outputs, reductions over hundreds of thousands of rows): gridDim.z CTAs share one
    // output tile, each reduces a contiguous range of K blocks and adds its partial tile atomically
    const int nk_all = (p.K + BK - 1) / BK;
    int kb0 = 0, nk = nk_all;
    if (TA && gridDim.z > 1) {
        const int per = (nk_all + gridDim.z - 1) / gridDim.z;
        kb0 = blockIdx.z * per;
        nk = nk_all - kb0 < per ? nk_all - kb0 : per;
        if (nk <= 0) return;
    }
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, (kb0 + s) * BK);
        cp_async_commit();
    }
    const int a_frag = TA ? (lane & 3) * BS_LD + wm * 64 + (lane >> 2) : (wm * 64 + (lane >> 2)) * AS_LD + (lane & 3);
    const int b_frag = (lane & 3) * BS_LD + wn * 32 + (lane >> 2);
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nx = kt + STAGES - 1;
            if (nx < nk) load_stage(nx % STAGES, (kb0 + nx) * BK);
            cp_async_commit();
        }
        const double* as = As + (kt % STAGES) * A_STAGE + a_frag;
        const double* bs = Bs + (kt % STAGES) * B_STAGE + b_frag;
#pragma unroll
        for (int kk = 0; kk < BK; kk += 4) {
            double a[8], b[4];
#pragma unroll
            for (int mt = 0; mt < 8; ++mt) a[mt] = TA ? as[kk * BS_LD + mt * 8] : as[mt * 8 * AS_LD + kk];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) b[nt] = bs[kk * BS_LD + nt * 8];
#pragma unroll
            for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue ---------------------------------------------------------
    const double rs2 = 0.70710678118654752440;
    const int ncol_base = col0 + wn * 32 + (lane & 3) * 2;

    if (MODE == GEMM_PLAIN || MODE == GEMM_TN) {
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
            long long r = row0 + wm * 64 + mt * 8 + (lane >> 2);
            if (r >= p.M) continue;
            long long cr = (p.cmap && MODE == GEMM_PLAIN) ? map_row(r, p.rpg, p.gstride, p.goff) : r;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                int n = ncol_base + nt * 8;
                if (n >= p.N) continue;
                double v0 = acc[mt][nt][0], v1 = acc[mt][nt][1];
                if (p.colbias) { v0 += p.colbias[n]; v1 += p.colbias[n + 1]; }
                double2* dst = reinterpret_cast<double2*>(p.C + cr * (long long)p.ldc + n);
                if (TA && gridDim.z > 1) {
                    atomicAdd(&dst->x, v0);
                    atomicAdd(&dst->y, v1);
                    continue;
                }
                if (p.accumulate) { const double2 o = *dst; v0 += o.x; v1 += o.y; }
                *dst = make_double2(v0, v1);
            }
        }
    } else if (MODE == GEMM_VALUE || MODE == GEMM_LAP) {
        // rows r = w*n_elec + i
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
            long long r = row0 + wm * 64 + mt * 8 + (lane >> 2);
            if (r >= p.M) continue;
            long long w = r / p.n_elec;
            const double* g = p.G + (w * p.NDg + p.NDp + (MODE == GEMM_VALUE ? 0 : 1)) * (long long)p.ldg;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                int n = ncol_base + nt * 8;
                if (n >= p.N) continue;
                double2 gg = *reinterpret_cast<const double2*>(g + n);
                double z0 = acc[mt][nt][0] + gg.x, z1 = acc[mt][nt][1] + gg.y;
                double o0, o1;
                if (MODE == GEMM_VALUE) {
                    z0 += p.colbias[n]; z1 += p.colbias[n + 1];
                    o0 = tanh(z0); o1 = tanh(z1);
                    *reinterpret_cast<double2*>(p.T + r * (long long)p.ldt + n) = make_double2(o0, o1);
                } else {
                    double2 t = *reinterpret_cast<const double2*>(p.T + r * (long long)p.ldt + n);
                    double2 s = *reinterpret_cast<const double2*>(p.S + r * (long long)p.ldt + n);
                    double d0 = 1.0 - t.x * t.x, d1 = 1.0 - t.y * t.y;
                    o0 = d0 * z0 - 2.0 * t.x * d0 * s.x;
                    o1 = d1 * z1 - 2.0 * t.y * d1 * s.y;
                }
                if (RES) {
                    double2 h = *reinterpret_cast<const double2*>(p.R + r * (long long)p.ldr + n);
                    o0 = (h.x + o0) * rs2; o1 = (h.y + o1) * rs2;
                }
                *reinterpret_cast<double2*>(p.C + r * (long long)p.ldc + n) = make_double2(o0, o1);
            }
        }
    } else if (MODE == GEMM_JAC) {
        // rows r = (w*n_elec + i)*NDp + d ; every aligned group of 8 rows shares (w,i)
        long long cur_e = -1;
        double sacc[4][2], d1v[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) sacc[nt][0] = sacc[nt][1] = d1v[nt][0] = d1v[nt][1] = 0.0;
        auto flush = [&]() {
            if (cur_e < 0) return;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    double v = sacc[nt][j];
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    int n = ncol_base + nt * 8 + j;
                    if ((lane >> 2) == 0 && n < p.N) atomicAdd(p.S + cur_e * (long long)p.ldt + n, v);
                    sacc[nt][j] = 0.0;
                }
            }
        };
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
            long long rbase = row0 + wm * 64 + mt * 8;       // warp-uniform
            if (rbase >= p.M) break;
            long long e = rbase / p.NDp;
            if (e != cur_e) {
                flush();
                cur_e = e;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    int n = ncol_base + nt * 8;
                    if (n < p.N) {
                        double2 t = *reinterpret_cast<const double2*>(p.T + e * (long long)p.ldt + n);
                        d1v[nt][0] = 1.0 - t.x * t.x; d1v[nt][1] = 1.0 - t.y * t.y;
                    }
                }
            }
            long long r = rbase + (lane >> 2);
            int d = (int)(r - e * p.NDp);
            long long w = e / p.n_elec;
            const double* g = p.G + (w * p.NDg + d) * (long long)p.ldg;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                int n = ncol_base + nt * 8;
                if (n >= p.N) continue;
                double2 gg = *reinterpret_cast<const double2*>(g + n);
                double z0 = acc[mt][nt][0] + gg.x, z1 = acc[mt][nt][1] + gg.y;
                sacc[nt][0] = fma(z0, z0, sacc[nt][0]);
                sacc[nt][1] = fma(z1, z1, sacc[nt][1]);
                double o0 = d1v[nt][0] * z0, o1 = d1v[nt][1] * z1;
                if (RES) {
                    double2 h = *reinterpret_cast<const double2*>(p.R + r * (long long)p.ldr + n);
                    o0 = (h.x + o0) * rs2; o1 = (h.y + o1) * rs2;
                }
                *reinterpret_cast<double2*>(p.C + r * (long long)p.ldc + n) = make_double2(o0, o1);
            }
        }
        flush();
    } else if (MODE == GEMM_ORBJ) {
        // logical rows r = (w*n_s + i_s)*NDp + d ; columns n = 2p (re), 2p+1 (im)
        const long long rpe = (long long)p.n_s * p.NDp;
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
            long long r = row0 + wm * 64 + mt * 8 + (lane >> 2);
            if (r >= p.M) continue;
            long long w = r / rpe;
            int rem = (int)(r - w * rpe);
            int is = rem / p.NDp, d = rem - is * p.NDp;
            int ND = 3 * p.n_elec;
            if (d >= ND) continue;
            long long e = w * p.n_elec + p.off_s + is;
            const double* et = p.etab + e * 5LL * p.npar_max * 2;
            bool own = (d / 3) == (p.off_s + is);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                int n = ncol_base + nt * 8;
                if (n >= p.N) continue;
                int pp = n >> 1;
                int k = pp / p.n_orb, o = pp - k * p.n_orb;
                double vr = acc[mt][nt][0], vi = acc[mt][nt][1];
                double2 E = *reinterpret_cast<const double2*>(et + 2 * pp);
                double outr = vr * E.x - vi * E.y, outi = vr * E.y + vi * E.x;
                long long di = (((w * p.n_det + k) * p.NDp + d) * p.n_rows_mat + p.row0 + is) * (long long)p.n_orb + o;
                *reinterpret_cast<double2*>(p.DA + 2 * di) = make_double2(outr, outi);
                if (own) {
                    int c = d - 3 * (d / 3);
                    *reinterpret_cast<double2*>(p.YOWN + 2 * ((e * 3 + c) * (long long)p.npar_max + pp)) =
                        make_double2(vr, vi);
                }
            }
        }
    }
}

template <int MODE, bool RES>
int launch(const GemmParams& p, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        DS_CUDA_CHECK(cudaFuncSetAttribute(gemm_f64_kernel<MODE, RES>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)((p.N + BN - 1) / BN));
    if (MODE == GEMM_TN && p.accumulate && !p.colbias) {
        // enough CTAs for two waves of the machine, at least 8 K blocks each
        const long long tiles = (long long)grid.x * grid.y;
        const int nkb = (p.K + BK - 1) / BK;
        long long z = (2 * 148 + tiles - 1) / tiles;
        if (z > nkb / 8) z = nkb / 8;
        if (z < 1) z = 1;
        grid.z = (unsigned)z;
    }
    gemm_f64_kernel<MODE, RES><<<grid, NTHREADS, SMEM_BYTES, stream>>>(p);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace

int ds_launch_gemm(const GemmParams& p, int mode, bool residual, cudaStream_t stream) {
    if (p.M <= 0 || p.N <= 0 || p.K <= 0) return 0;
    DS_REQUIRE((mode == GEMM_TN || (p.K % 2) == 0) && (p.N % 2) == 0 && (p.lda % 2) == 0 && (p.ldb % 2) == 0,
               "gemm: K, N, lda, ldb must be even (got K=%d N=%d lda=%d ldb=%d)", p.K, p.N, p.lda, p.ldb);
    switch (mode) {
        case GEMM_PLAIN: return launch<GEMM_PLAIN, false>(p, stream);
        case GEMM_VALUE: return residual ? launch<GEMM_VALUE, true>(p, stream) : launch<GEMM_VALUE, false>(p, stream);
        case GEMM_JAC: return residual ? launch<GEMM_JAC, true>(p, stream) : launch<GEMM_JAC, false>(p, stream);
        case GEMM_LAP: return residual ? launch<GEMM_LAP, true>(p, stream) : launch<GEMM_LAP, false>(p, stream);
        case GEMM_ORBJ: return launch<GEMM_ORBJ, false>(p, stream);
        case GEMM_TN:
            DS_REQUIRE((p.M % 2) == 0, "gemm TN: M must be even (M=%lld)", p.M);
            return launch<GEMM_TN, false>(p, stream);
    }
    ds_set_error("gemm: unknown mode %d", mode);
    return -1;
}

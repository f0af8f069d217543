// tcgen05 / TMEM / TMA implementation of the sliced-integer fp64 GEMM.  See ozaki.cuh.
#include "ozaki.cuh"
#include <cstdlib>

#include <cuda.h>
#include <cudaTypedefs.h>

namespace {

// Compiled in only with -DDS_OZ_PROF (the clocks cost the epilogue registers): DS_EXTRA_NVCC_FLAGS=-DDS_OZ_PROF python -m deepsolid_b200.build -f
#ifdef DS_OZ_PROF
constexpr bool OZ_PROF = true;
#else
constexpr bool OZ_PROF = false;
#endif

constexpr int EPI_WARPS = 16;                         // four per TMEM lane quarter
constexpr int EPI_COLS = 16;                          // columns (rows of A) per epilogue warp
constexpr int OZ_THREADS = 64 + EPI_WARPS * 32;       // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..: epilogue
constexpr int STAGES = 3;
constexpr int W_SLICE = OZ_TM * OZ_BK;                // 8192 B
constexpr int A_SLICE = OZ_TN * OZ_BK;                // 4096 B
constexpr int W_STAGE = OZ_S * W_SLICE;               // 49152 B
constexpr int A_STAGE = OZ_S * A_SLICE;               // 24576 B
constexpr int STAGE_BYTES = W_STAGE + A_STAGE;        // 73728 B
constexpr int SMEM_DYN = STAGES * STAGE_BYTES + 1024; // + alignment slack
constexpr int TMEM_COLS = 512;                        // 6 x 64 used (power of two required)
// Row tile TN = 64: one accumulator set (6 x 64 columns).  TN = 32: TWO sets (2 x 6 x 32 columns), so the TMEM read of
// a tile (64 B/clk, ~1.6 us per 196 KB) overlaps the MMAs of the next one; the epilogue warps split into two sets of
// eight that alternate tiles.
template <int TN> struct OzCfg {
    static constexpr int NBUF = (TN == 32) ? 2 : 1;
    static constexpr int SET_WARPS = EPI_WARPS / NBUF;
    static constexpr int A_SLICE_T = TN * OZ_BK;
    static constexpr int A_STAGE_T = OZ_S * A_SLICE_T;
    static constexpr int STAGE_T = W_STAGE + A_STAGE_T;
    static constexpr int SMEM_T = STAGES * STAGE_T + 1024;
    static_assert(TN / (SET_WARPS / 4) == EPI_COLS, "epilogue warps x EPI_COLS must tile the rows of a tile");
};

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error traps (kernel fails) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, int8 x int8 -> int32
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
// 16 consecutive int32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, int (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ---- thread-block cluster / distributed shared memory (OZ_JACD: the two 128-channel CTAs of a row tile) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t caddr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(caddr), "r"(v) : "memory");
}
// arrive on an mbarrier of another CTA; release at cluster scope orders this thread's earlier remote stores before it
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cbar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cbar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
// asynchronous store into another CTA's shared memory that completes 4 transaction bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_u32(uint32_t caddr, uint32_t v, uint32_t cbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(caddr), "r"(v), "r"(cbar)
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// int32 -> fp64 without the conversion unit: the integer sits in the low mantissa word of 2^52 + 2^31 + v
__device__ __forceinline__ double i32_to_f64(int v) {
    return __hiloint2double(0x43300000, (int)((unsigned)v ^ 0x80000000u)) - 4503601774854144.0;
}
// transposed warp reductions (max of unsigned): v[r] over the 32 lanes for 16 (4) rows in 16 (6) shuffles.
// Result for row r(lane) = 8 b4 + 4 b3 + 2 b2 + b1 (resp. 2 b4 + b3) of the lane index is returned in every lane.
__device__ __forceinline__ unsigned warp_rowmax16(const unsigned (&v)[16], int lane) {
    unsigned a[8], b[4], c[2];
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const unsigned keep = h16 ? v[i + 8] : v[i], give = h16 ? v[i] : v[i + 8];
        a[i] = max(keep, __shfl_xor_sync(0xffffffffu, give, 16));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const unsigned keep = h8 ? a[i + 4] : a[i], give = h8 ? a[i] : a[i + 4];
        b[i] = max(keep, __shfl_xor_sync(0xffffffffu, give, 8));
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const unsigned keep = h4 ? b[i + 2] : b[i], give = h4 ? b[i] : b[i + 2];
        c[i] = max(keep, __shfl_xor_sync(0xffffffffu, give, 4));
    }
    const unsigned keep = h2 ? c[1] : c[0], give = h2 ? c[0] : c[1];
    unsigned m = max(keep, __shfl_xor_sync(0xffffffffu, give, 2));
    return max(m, __shfl_xor_sync(0xffffffffu, m, 1));
}
__device__ __forceinline__ unsigned warp_rowmax4(const unsigned (&v)[4], int lane) {
    unsigned a[2];
    const bool h16 = lane & 16, h8 = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const unsigned keep = h16 ? v[i + 2] : v[i], give = h16 ? v[i] : v[i + 2];
        a[i] = max(keep, __shfl_xor_sync(0xffffffffu, give, 16));
    }
    const unsigned keep = h8 ? a[1] : a[0], give = h8 ? a[0] : a[1];
    unsigned m = max(keep, __shfl_xor_sync(0xffffffffu, give, 8));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 4));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
    return max(m, __shfl_xor_sync(0xffffffffu, m, 1));
}
// six balanced base-256 digits of x at row exponent e (see slice_row_warp), stored at out[s * pitch], s = 0..5
__device__ __forceinline__ void store_digits(double x, double f, bool bad, signed char* __restrict__ out, long long pitch) {
    const double MAGIC = 6755399441055744.0;          // 2^52 + 2^51
    const double t = bad ? MAGIC : fma(x, f, MAGIC);
    unsigned l = (unsigned)__double2loint(t), h = (unsigned)__double2hiint(t);
    const unsigned l2 = l + 0x80808080u;
    h += 0x80u + (l2 < l ? 1u : 0u);
    l = l2 ^ 0x80808080u;
    h ^= 0x80u;
    out[0] = (signed char)(h >> 8);
    out[pitch] = (signed char)h;
    out[2 * pitch] = (signed char)(l >> 24);
    out[3 * pitch] = (signed char)(l >> 16);
    out[4 * pitch] = (signed char)(l >> 8);
    out[5 * pitch] = (signed char)l;
}

// K-major, SWIZZLE_64B shared-memory matrix descriptor (rows of 64 bytes, 8-row groups 512 B apart)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address
    d |= (uint64_t)1 << 16;                     // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(512 >> 4) << 32;            // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                     // descriptor version of sm_100
    d |= (uint64_t)4 << 61;                     // SWIZZLE_64B
    return d;
}
// MN-major, SWIZZLE_64B descriptor of the B operand when the digits are stored row-contiguous ([slice][k][row], the
// layout the fused-digit epilogue writes): per k one 64-byte line of 64 consecutive rows, 8 k = one 512 B swizzle atom
// (stride byte offset), the next 64 rows = the next digit slice, OZ_BK * 64 B further (leading byte offset).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((OZ_BK * 64) >> 4) << 16;   // leading byte offset: between 64-row blocks (= digit slices)
    d |= (uint64_t)(512 >> 4) << 32;            // stride byte offset: between groups of 8 k
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                     // SWIZZLE_64B
    return d;
}
// kind::i8 instruction descriptor: D = s32, A = B = signed int8, M = 128, N = n; A K-major, B K-major or MN-major
__host__ __device__ constexpr uint32_t make_idesc(int n, bool b_mn = false) {
    return (2u << 4) | (1u << 7) | (1u << 10) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(OZ_TM >> 4) << 24);
}

// ---------------------------------------------------------------------------
// digits of fp64 rows: one warp per row.
//   e      = exponent of the row maximum + 1 (taken from the IEEE exponent field: integer max of the
//            high words, no fp64 compares),  sa[r] = 2^(e-6)
//   q[k]   = rn(A[r,k] * 2^(46-e))           (|q| < 2^46; formed with the 2^52+2^51 magic add)
//   digits = balanced base-256 digits of q: the bytes of q + 0x8080808080 with their top bit flipped
//            (adding 128 to every byte position propagates exactly the balanced carries); the top
//            digit is the signed remainder byte.
// 8 consecutive k per lane; a byte transpose (PRMT) turns them into one 8-byte word per slice, so a
// warp stores 256 contiguous bytes per slice.
// ---------------------------------------------------------------------------
// digits + scale of ONE row held by a warp as v[it][j] = A[r, it*256 + lane*8 + j] (zero beyond K)
__device__ __forceinline__ void slice_row_warp(const double (&v)[2][8], int K, int lane, signed char* __restrict__ out,
                                               double* __restrict__ sa_r) {
    unsigned mxh = 0u;                                // max over the row of the high words of |A|
#pragma unroll
    for (int it = 0; it < 2; ++it)
#pragma unroll
        for (int j = 0; j < 8; ++j) mxh = max(mxh, (unsigned)__double2hiint(v[it][j]) & 0x7fffffffu);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mxh = max(mxh, __shfl_xor_sync(0xffffffffu, mxh, off));
    const bool bad = mxh >= 0x7ff00000u;              // inf or nan somewhere in the row
    int e = (int)(mxh >> 20) - 1022;                  // row max < 2^e (subnormal rows: e = -1022)
    if (e < -900) e = -900;
    // 2^(46-e) as an IEEE double (46-e in [-978, 946]: always a normal number)
    const double f = __hiloint2double((1023 + 8 * OZ_S - 2 - e) << 20, 0);
    if (lane == 0) *sa_r = bad ? __longlong_as_double(0x7ff8000000000000LL) : __hiloint2double((1023 + e - 6) << 20, 0);
    const double MAGIC = 6755399441055744.0;          // 2^52 + 2^51
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int c0 = it * 256 + lane * 8;
        if (c0 >= K) continue;
        unsigned lo[8], hi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double t = bad ? MAGIC : fma(v[it][j], f, MAGIC);     // low 48 bits = q (two's complement)
            unsigned l = (unsigned)__double2loint(t), h = (unsigned)__double2hiint(t);
            const unsigned l2 = l + 0x80808080u;
            h += 0x80u + (l2 < l ? 1u : 0u);
            lo[j] = l2 ^ 0x80808080u;                 // digits 2..5 (bytes 3..0)
            hi[j] = h ^ 0x80u;                        // digit 1 (byte 0), digit 0 = signed byte 1
        }
        unsigned w[OZ_S][2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const unsigned *L = lo + 4 * hf, *H = hi + 4 * hf;
            const unsigned t01 = __byte_perm(L[0], L[1], 0x5140), t23 = __byte_perm(L[2], L[3], 0x5140);   // bytes 0,1
            const unsigned u01 = __byte_perm(L[0], L[1], 0x7362), u23 = __byte_perm(L[2], L[3], 0x7362);   // bytes 2,3
            const unsigned h01 = __byte_perm(H[0], H[1], 0x5140), h23 = __byte_perm(H[2], H[3], 0x5140);
            w[5][hf] = __byte_perm(t01, t23, 0x5410);
            w[4][hf] = __byte_perm(t01, t23, 0x7632);
            w[3][hf] = __byte_perm(u01, u23, 0x5410);
            w[2][hf] = __byte_perm(u01, u23, 0x7632);
            w[1][hf] = __byte_perm(h01, h23, 0x5410);
            w[0][hf] = __byte_perm(h01, h23, 0x7632);
        }
#pragma unroll
        for (int s = 0; s < OZ_S; ++s)
            *reinterpret_cast<uint2*>(out + (long long)s * K + c0) = make_uint2(w[s][0], w[s][1]);
    }
}

__device__ __forceinline__ void load_row_warp(double (&v)[2][8], const double* __restrict__ a, int K, int lane) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int c0 = it * 256 + lane * 8;
        if (c0 < K) {
            const double2* p = reinterpret_cast<const double2*>(a + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double2 t = p[j];
                v[it][2 * j] = t.x; v[it][2 * j + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[it][j] = 0.0;
        }
    }
}

__global__ void __launch_bounds__(256) slice_rows_kernel(const double* __restrict__ A, int lda, long long rows, int K,
                                                         signed char* __restrict__ Ad, double* __restrict__ sa) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    double v[2][8];
    load_row_warp(v, A + r * (long long)lda, K, lane);
    slice_row_warp(v, K, lane, Ad + r * (long long)OZ_S * K, sa + r);
}

// Digits of the Jacobian rows of one-electron-stream layer l >= 1 AND the spin-channel means of the same
// rows in ONE pass over the fp64 Jacobian: a warp owns (walker w, direction d) and walks the electrons
// i = 0..N-1 (rows (w*N + i)*NDp + d), so the column sums over the electrons of a spin channel stay in
// registers:  GIN[w, d, s*C + c] = mean_{i in s} J[(w,i,d), c],  c < C  (the own columns; network.py:322-330).
__global__ void __launch_bounds__(256, 2) slice_means_kernel(const double* __restrict__ A, int lda, int K, int C,
                                                          int n_walkers, int n_up, int n_elec, int NDp, int NDg,
                                                          signed char* __restrict__ Ad, double* __restrict__ sa,
                                                          double* __restrict__ GIN, int ldgin) {
    const long long wd = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wd >= (long long)n_walkers * NDp) return;
    const int lane = threadIdx.x & 31;
    const int w = (int)(wd / NDp), d = (int)(wd - (long long)w * NDp);
    double sum[2][8];
    double* gout = GIN + ((long long)w * NDg + d) * ldgin;
    double v[2][8], vn[2][8];
    long long r = ((long long)w * n_elec) * NDp + d;
    load_row_warp(vn, A + r * (long long)lda, K, lane);
    for (int s = 0; s < 2; ++s) {
        const int ibeg = s ? n_up : 0, iend = s ? n_elec : n_up;
#pragma unroll
        for (int it = 0; it < 2; ++it)
#pragma unroll
            for (int j = 0; j < 8; ++j) sum[it][j] = 0.0;
        for (int i = ibeg; i < iend; ++i, r += NDp) {
#pragma unroll
            for (int it = 0; it < 2; ++it)
#pragma unroll
                for (int j = 0; j < 8; ++j) v[it][j] = vn[it][j];
            if (i + 1 < n_elec) load_row_warp(vn, A + (r + NDp) * (long long)lda, K, lane);   // next row in flight
#pragma unroll
            for (int it = 0; it < 2; ++it)
#pragma unroll
                for (int j = 0; j < 8; ++j) sum[it][j] += v[it][j];
            slice_row_warp(v, K, lane, Ad + r * (long long)OZ_S * K, sa + r);
        }
        const double inv = 1.0 / (double)(iend - ibeg);
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int c0 = it * 256 + lane * 8;
            if (c0 < C) {                             // C is a multiple of 8: a lane's 8 columns are all own or all pair-mean
                double2* o = reinterpret_cast<double2*>(gout + s * C + c0);
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = make_double2(sum[it][2 * j] * inv, sum[it][2 * j + 1] * inv);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// epilogue blocks: 8 consecutive rows of A (one electron) for one output channel.  FULL: the whole
// 128-channel block is inside N (no per-lane predicates, straight-line code the scheduler can
// interleave across the 8 columns); all loads are issued before the first use.
// ---------------------------------------------------------------------------
// sa[r] * sb[n] for power-of-two scales, formed on the integer pipe (the epilogue waits on the fp64 pipe): both are
// 2^k (slice_row_warp) or NaN for a row / column holding inf or nan, so the product is an exponent-field addition.
// sch = high word of sa[r]; sbo = high word of sb[n] 2^-16 minus the exponent bias; badc = column flagged.
// A product below the normal range (an all-zero padding row times a small column scale) becomes 0, as fp64 would.
__device__ __forceinline__ double pow2_scale(int sch, int sbo, bool badc) {
    int h = max(sch + sbo, 0);
    if (badc || sch >= 0x7ff00000) h = 0x7ff80000;
    return __hiloint2double(h, 0);
}

template <bool RES, bool FULL>
__device__ __forceinline__ double jac_block8(const double* zz8, const double* __restrict__ gp, long long ldg,
                                             const double* __restrict__ rp, long long ldr, double* __restrict__ cp,
                                             long long ldc, const double* __restrict__ sap, int sbo, bool badc, double d1,
                                             bool nv, double sacc) {
    const double rs2 = 0.70710678118654752440;
    double gv[8], rv[8];
    int sch[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        sch[jj] = __ldg(reinterpret_cast<const int*>(sap + jj) + 1);   // same address in every lane: one broadcast transaction
        gv[jj] = (FULL || nv) ? gp[jj * ldg] : 0.0;
        rv[jj] = (RES && (FULL || nv)) ? rp[jj * ldr] : 0.0;
    }
    double s1 = 0.0;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const double zj = fma(zz8[jj], pow2_scale(sch[jj], sbo, badc), gv[jj]);
        if (jj & 1) s1 = fma(zj, zj, s1); else sacc = fma(zj, zj, sacc);
        double o = d1 * zj;
        if (RES) o = (rv[jj] + o) * rs2;
        if (FULL || nv) cp[jj * ldc] = o;
    }
    return sacc + s1;
}

// One aligned group of 8 rows (directions d .. d+7 of one electron) of the orbital Jacobian for one channel
// n = 2 pp + (re|im):  dM = (y_re + i y_im)(Ex + i Ey), the partner component comes from the neighbouring lane.
// Lean on purpose (the epilogue is issue-bound): the column scale and the sign of the cross term are folded into
// two per-block factors, rows beyond ND are predicated off (no per-row branches), and the three own-coordinate
// rows (raw orbital derivatives for the Laplacian assembly) are handled by a separate, rarely taken loop.
template <bool FULL>
__device__ __forceinline__ void orbj_block8(const double* zz8, const double* __restrict__ sap, double sbn, double Ex,
                                            double Ey, int im, double* __restrict__ dp, long long ns2,
                                            double* __restrict__ yp, long long ystride, int d, int ND, int c0, bool nv) {
    double sc[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) sc[jj] = __ldg(sap + jj);      // powers of two (or NaN for a bad row)
    const double sbp = __shfl_xor_sync(0xffffffffu, sbn, 1);     // column scale of the partner component
    const double ax = sbn * Ex, ay = im ? sbp * Ey : -(sbp * Ey);
    const int nrows = ND - d;                                   // rows jj < nrows are real directions
    const bool ok = FULL || nv;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const double zs = zz8[jj] * sc[jj];                     // exact (power-of-two scale)
        const double zp = __shfl_xor_sync(0xffffffffu, zs, 1);
        // (vr + i vi)(Ex + i Ey): re = vr Ex - vi Ey, im = vi Ex + vr Ey
        const double o = fma(zp, ay, zs * ax);
        if (ok && jj < nrows) dp[jj * ns2] = o;
    }
    if (c0 > -8 && c0 < 3) {                                    // this block holds own coordinates of the electron
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const unsigned c = (unsigned)(c0 + jj);
            if (ok && c < 3u && jj < nrows) yp[c * ystride] = zz8[jj] * sc[jj] * sbn;
        }
    }
}

// ---------------------------------------------------------------------------
// the GEMM
// ---------------------------------------------------------------------------
// dbg & 32: per-role wait / phase clocks, summed over the CTAs, per MODE (ds_oz_prof_read):
//   [0] kernel clocks of the MMA thread  [1] MMA waits for drained accumulators  [2] MMA waits for a full smem stage
//   [3] producer waits for a free stage  [4] epilogue warp 2 waits for finished accumulators  [5] its phase A (TMEM read)
//   [6] its phase B (epilogue math, loads, stores)  [7] tiles of the CTA
__device__ unsigned long long g_oz_prof[8][8];
template <int MODE, bool RES, int TN, int ND, bool BMN>
// 18 warps: five share one SM sub-partition (16384 registers), so at most 96 registers per thread
__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_gemm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmA, const OzParams p,
               const int tiles_per_group, const int n_cb, const long long n_tiles) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    using Cfg = OzCfg<TN>;
    constexpr int NBUF = Cfg::NBUF, SET_WARPS = Cfg::SET_WARPS, A_SLICE_T = Cfg::A_SLICE_T, STAGE_T = Cfg::STAGE_T;
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_s;
    // OZ_JACD: row maxima (IEEE high words of |x|) of a tile's 64 rows -- per epilogue warp, of the pair-mean columns,
    // and the other CTA's (written by it through DSMEM, two slots alternating with the tile parity)
    constexpr bool JD = (MODE == OZ_JACD);
    __shared__ unsigned xown[JD ? 2 : 1][4][4][16], xpm[JD ? 2 : 1][4][16], xrem[JD ? 2 : 1][64];
    __shared__ __align__(8) uint64_t xbar[JD ? 2 : 1][4];

    // (the shuffle tells the compiler that `warp` is warp-uniform: role branches become uniform branches and the epilogue keeps
    //  its memory descriptors and loop state in uniform registers instead of re-materialising them per access)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int nkb = p.K / OZ_BK;
    constexpr int ndiag = ND;                            // 6; 5 is an accuracy / speed experiment (DS_OZ_DIAGS=5)

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], SET_WARPS); }
        if (JD)
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 4; ++j) mbar_init(&xbar[i][j], 1);      // one local arrive.expect_tx(64) per use; the other CTA's 16 stores complete the bytes
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (JD) cluster_sync_all();          // the peer's barriers exist before anything is sent to them
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            long long cnt = 0, prof_a = 0;
            // (tile counts fit 32 bits, checked by the launcher: unsigned divisions, not the 64-bit software routine)
            for (unsigned tile = blockIdx.x; tile < (unsigned)n_tiles; tile += gridDim.x) {
                const int cb = (int)(tile % (unsigned)n_cb);
                const unsigned rt = tile / (unsigned)n_cb;
                const int grp = (int)(rt / (unsigned)tiles_per_group);
                const int q0 = (int)(rt % (unsigned)tiles_per_group) * TN;
                for (int kb = 0; kb < nkb; ++kb, ++cnt) {
                    const int st = (int)(cnt % STAGES);
                    const uint32_t ph = (uint32_t)((cnt / STAGES) & 1);
                    const long long tp0 = (OZ_PROF && (p.dbg & 32)) ? clock64() : 0;
                    mbar_wait(&empty_bar[st], ph ^ 1u);
                    if (OZ_PROF && (p.dbg & 32)) prof_a += clock64() - tp0;
                    unsigned char* sW = smem + st * STAGE_T;
                    if (p.dbg & 4) { mbar_arrive(&full_bar[st]); continue; }
                    mbar_expect_tx(&full_bar[st], STAGE_T);
                    tma_load_4d(sW, &tmW, &full_bar[st], kb * OZ_BK, cb * OZ_TM, 0, 0);
                    // blocked digits: the window of a group starts at the 64-row block that holds its first row
                    if (BMN) tma_load_4d(sW + W_STAGE, &tmA, &full_bar[st], 0, kb * OZ_BK, 0, (int)((((long long)grp * p.gstride + p.goff) >> 6) + q0 / TN));
                    else tma_load_4d(sW + W_STAGE, &tmA, &full_bar[st], kb * OZ_BK, q0, 0, grp);
                }
            }
            if (OZ_PROF && (p.dbg & 32)) atomicAdd(&g_oz_prof[MODE][3], (unsigned long long)prof_a);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            long long cnt = 0, prof_t = 0, prof_f = 0;
            const long long prof_0 = (OZ_PROF && (p.dbg & 32)) ? clock64() : 0;
            uint32_t it = 0;
            for (unsigned tile = blockIdx.x; tile < (unsigned)n_tiles; tile += gridDim.x, ++it) {
                const uint32_t buf = it % NBUF, use = it / NBUF;
                const long long tm0 = (OZ_PROF && (p.dbg & 32)) ? clock64() : 0;
                mbar_wait(&tmem_empty_bar[buf], (use & 1u) ^ 1u);
                if (OZ_PROF && (p.dbg & 32)) prof_t += clock64() - tm0;
                tc_fence_after();
                const uint32_t tacc = tmem_base + buf * (OZ_S * TN);
                for (int kb = 0; kb < nkb; ++kb, ++cnt) {
                    const int st = (int)(cnt % STAGES);
                    const uint32_t ph = (uint32_t)((cnt / STAGES) & 1);
                    const long long tf0 = (OZ_PROF && (p.dbg & 32)) ? clock64() : 0;
                    mbar_wait(&full_bar[st], ph);
                    if (OZ_PROF && (p.dbg & 32)) prof_f += clock64() - tf0;
                    tc_fence_after();
                    const uint32_t sW = smem_u32(smem + st * STAGE_T);
                    const uint32_t sA = sW + W_STAGE;
                    if (!(p.dbg & 2))
#pragma unroll
                    for (int ks = 0; ks < OZ_BK / 32; ++ks) {
#pragma unroll
                        for (int t = 0; t < OZ_S; ++t) {
                            // W_t x [A_0 .. A_{5-t}] -> diagonals t .. 5 (TMEM column blocks of 64)
                            const uint64_t wdesc = make_desc(sW + t * W_SLICE + ks * 32);
                            const int nsl = ndiag - t;
                            if (nsl <= 0) continue;
                            constexpr int MAXSL = 256 / TN;                  // slices of A one instruction can span (N <= 256)
                            const int n1 = (nsl > MAXSL ? MAXSL : nsl) * TN;
                            const uint32_t acc = (kb > 0 || ks > 0 || t > 0) ? 1u : 0u;
                            // K-major rows: 32 k = 32 B along a row; MN-major lines: 32 k = 32 lines of 64 B
                            const uint32_t koff = BMN ? ks * 32 * 64 : ks * 32;
                            umma_i8(tacc + t * TN, wdesc, BMN ? make_desc_mn(sA + koff) : make_desc(sA + koff), make_idesc(n1, BMN), acc);
                            if (nsl > MAXSL)
                                umma_i8(tacc + (t + MAXSL) * TN, wdesc,
                                        BMN ? make_desc_mn(sA + MAXSL * A_SLICE_T + koff) : make_desc(sA + MAXSL * A_SLICE_T + koff),
                                        make_idesc((nsl - MAXSL) * TN, BMN), acc);
                        }
                    }
                    umma_commit(&empty_bar[st]);          // frees the smem stage when the MMAs have read it
                }
                umma_commit(&tmem_full_bar[buf]);         // accumulators complete
            }
            if (OZ_PROF && (p.dbg & 32)) {
                atomicAdd(&g_oz_prof[MODE][0], (unsigned long long)(clock64() - prof_0));
                atomicAdd(&g_oz_prof[MODE][1], (unsigned long long)prof_t);
                atomicAdd(&g_oz_prof[MODE][2], (unsigned long long)prof_f);
                atomicAdd(&g_oz_prof[MODE][7], (unsigned long long)it);
            }
        }
    } else {
        // ===================== epilogue: thread = output channel =====================
        // EPI_WARPS warps: four per TMEM lane quarter, each owning EPI_COLS of the tile's 64 columns
        // (= rows of A).  Phase A (holds TMEM): read the six diagonals, pack them EXACTLY into two
        // int64 per output
        //   hi = c0 2^16 + c1 2^8 + c2,  lo = c3 2^16 + c4 2^8 + c5   (|c_g| < 2^26),
        // round hi + lo 2^-24 ONCE to fp64 (the only rounding of the whole product) and release the
        // accumulators so the next tile's MMAs start.  Phase B: scales, fused epilogue math, stores,
        // in blocks of 8 columns: rows of A come in aligned groups of 8 that share one electron
        // (NDp and rows-per-group are multiples of 8), so a block needs no per-column bookkeeping.
        const int q = warp & 3;                           // TMEM lane quarter this warp may access
        const int set = (warp - 2) / SET_WARPS;           // accumulator set (and tile parity) of this warp
        const int cg = ((warp - 2) % SET_WARPS) >> 2;     // which EPI_COLS columns
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + set * (OZ_S * TN) + cg * EPI_COLS;
        uint32_t it = 0;
        long long prof_w = 0, prof_pa = 0, prof_pb = 0, te2_prev = 0;
        const double MAGIC_HL = 6755399441055744.0 + 402653184.0;     // (2^52 + 2^51)(1 + 2^-24), exact
        int cur_cb = -1;
        double sbn = 0.0;
        int sbo = 0;
        bool badc = false;
        for (unsigned tile = blockIdx.x + (unsigned)set * gridDim.x; tile < (unsigned)n_tiles; tile += (unsigned)NBUF * gridDim.x, ++it) {
            const int cb = (int)(tile % (unsigned)n_cb);
            const unsigned rt = tile / (unsigned)n_cb;
            const long long grp = rt / (unsigned)tiles_per_group;
            const long long q0 = (long long)(rt % (unsigned)tiles_per_group) * TN + cg * EPI_COLS;      // first row (in group) of this warp
            const int n = cb * OZ_TM + q * 32 + lane;
            const bool nv = n < p.N;
            const bool full = (cb + 1) * OZ_TM <= p.N;                             // warp-uniform
            if (cb != cur_cb) {                                                    // a CTA normally keeps its cb
                cur_cb = cb;
                sbn = nv ? p.sb[n] * (1.0 / 65536.0) : 0.0;
                sbo = __double2hiint(sbn) - 0x3ff00000;                            // (sb[n] is a power of two or NaN)
                badc = __double2hiint(sbn) >= 0x7ff00000;
            }
            // physical row of column 0; with blocked digits (BMN) a group's window starts at the 64-row block holding
            // its first row, gskip rows early
            const long long gstart = grp * p.gstride + p.goff;
            const int gskip = BMN ? (int)(gstart & 63) : 0;
            const long long prow0 = gstart - gskip + q0;
            const long long left = p.rpg + gskip - q0;                             // (modes with gskip > 0 test validity per 8-row block)
            const int nvalid = left < 0 ? 0 : (left < EPI_COLS ? (int)left : EPI_COLS);   // warp-uniform
            const double* sap = p.sa + prow0;                                      // scales of this warp's rows
            if (MODE == OZ_JAC && RES && nv && nvalid == EPI_COLS && !(OZ_PROF && (p.dbg & 8192))) {
                // pull the residual rows towards L2 while the MMAs of this tile run
                const double* pr = p.R + prow0 * (long long)p.ldr + n;
#pragma unroll
                for (int j = 0; j < EPI_COLS; ++j) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pr));
                    pr += p.ldr;
                }
            }

            double zz[EPI_COLS];
            const bool prof = (OZ_PROF && (p.dbg & 32)) && warp == 2 && lane == 0;
            const long long te0 = prof ? clock64() : 0;
            if (prof && te2_prev) prof_pb += te0 - te2_prev;       // phase B of the previous tile (incl. the prologue of this one)
            mbar_wait(&tmem_full_bar[set], it & 1u);
            const long long te1 = prof ? clock64() : 0;
            tc_fence_after();
            if (!(MODE == OZ_PLAIN && (p.dbg & 1))) {
#pragma unroll
                for (int c0 = 0; c0 < EPI_COLS; c0 += 8) {
                    int v[OZ_S][8];
#pragma unroll
                    for (int g = 0; g < OZ_S; ++g) {
                        if (g < ndiag) tmem_ld8(lane_addr + g * TN + c0, v[g]);
                        else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[g][j] = 0;
                        }
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const long long hi = (long long)v[0][j] * 65536 + ((long long)v[1][j] * 256 + (long long)v[2][j]);
                        const long long lo = (long long)v[3][j] * 65536 + ((long long)v[4][j] * 256 + (long long)v[5][j]);
                        // (MAGIC + lo) 2^-24 + (MAGIC + hi) - MAGIC (1 + 2^-24): the subtraction of the constant
                        // 2^52 + 2^51 + 2^28 + 2^27 from MAGIC + hi is exact, so the fma rounds hi + lo 2^-24 ONCE
                        // (two fp64 instructions per output instead of three: the fp64 pipe is what this epilogue waits on)
                        const double dh = __longlong_as_double(hi + 0x4338000000000000LL) - MAGIC_HL;
                        zz[c0 + j] = fma(__longlong_as_double(lo + 0x4338000000000000LL), 1.0 / 16777216.0, dh);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[set]);
            if (MODE == OZ_PLAIN && (p.dbg & 1)) continue;
            const long long te2 = prof ? clock64() : 0;
            if (prof) { prof_w += te1 - te0; prof_pa += te2 - te1; te2_prev = te2; }
            if (OZ_PROF && (p.dbg & (256 | 512)) && (MODE == OZ_JAC || MODE == OZ_ORBJ)) {
                // synthetic phase B (contention probes; results are garbage): 256 = register-only fp64 chain of the real
                // epilogue's length (8 ops per output), 512 = the real epilogue's global loads and stores without the math
                if (p.dbg & 256) {
#pragma unroll 1
                    for (int r = 0; r < 8; ++r)
#pragma unroll
                        for (int j = 0; j < EPI_COLS; ++j) zz[j] = fma(zz[j], 1.0000001, 0.5);
                    if (zz[3] == 1.2345) p.C[n] = zz[3] + zz[7] + zz[11] + zz[0] + zz[15];
                } else if (MODE == OZ_JAC && nv && nvalid == EPI_COLS) {
                    const double* gp = p.G + (prow0 & 1023) * p.ldg + n;
                    const double* rp = p.R + prow0 * (long long)p.ldr + n;
                    double* cp = p.C + prow0 * (long long)p.ldc + n;
                    const bool useg = !(p.dbg & 2048), user = !(p.dbg & 1024), st = !(p.dbg & 4096);
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        double gv[8], rv[8];
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            gv[jj] = useg ? gp[(8 * b + jj) * (long long)p.ldg] : 1.0;
                            rv[jj] = user ? rp[(8 * b + jj) * (long long)p.ldr] : 2.0;
                        }
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const double o = gv[jj] + rv[jj] + zz[8 * b + jj];
                            if (st || o == 1.2345) cp[(8 * b + jj) * (long long)p.ldc] = o;
                        }
                    }
                }
                continue;
            }

            if (MODE == OZ_JAC) {
                // physical row = e * NDp + d  (e = walker * n_elec + electron); rows < 2^31 (checked by the launcher)
                unsigned e = (unsigned)prow0 / (unsigned)p.NDp;
                int d = (int)((unsigned)prow0 - e * (unsigned)p.NDp);
                unsigned w = e / (unsigned)p.n_elec;
                int ie = (int)(e - w * (unsigned)p.n_elec);
#pragma unroll
                for (int b = 0; b < EPI_COLS / 8; ++b) {
                    if (8 * b < nvalid) {                                   // warp-uniform; blocks are all-or-nothing
                        const double t = nv ? p.T[(long long)e * p.ldt + n] : 0.0;
                        const double d1 = 1.0 - t * t;
                        const double* gp = p.G + ((long long)w * p.NDg + d) * p.ldg + n;
                        const double* rp = RES ? p.R + (prow0 + 8 * b) * (long long)p.ldr + n : nullptr;
                        double* cp = p.C + (prow0 + 8 * b) * (long long)p.ldc + n;
                        double sacc;
                        if (full) sacc = jac_block8<RES, true>(zz + 8 * b, gp, p.ldg, rp, p.ldr, cp, p.ldc, sap + 8 * b, sbo, badc, d1, true, 0.0);
                        else sacc = jac_block8<RES, false>(zz + 8 * b, gp, p.ldg, rp, p.ldr, cp, p.ldc, sap + 8 * b, sbo, badc, d1, nv, 0.0);
                        // partial sum of zJ^2 over this aligned group of 8 directions; summed per electron in a fixed
                        // order by sp_reduce_kernel (deterministic, unlike an atomicAdd into S)
                        if (nv) p.SP[((prow0 + 8 * b) >> 3) * (long long)p.ldt + n] = sacc;
                    }
                    d += 8;
                    if (d >= p.NDp) { d -= p.NDp; ++e; if (++ie == p.n_elec) { ie = 0; ++w; } }
                }
            } else if (MODE == OZ_JACD) {
                // Both channel blocks are full (N == 2 OZ_TM, checked by the launcher): no channel predicates.
                // Digits are written ROW-CONTIGUOUS, [slice][k][row] with row pitch Rp: this thread's 16 rows of one
                // slice are 16 consecutive bytes (one 128-bit store / load), and the next GEMM reads them as an
                // MN-major operand.  prow0 is a multiple of 16.
                // B1: the Jacobian rows of the layer output, kept in registers (zz[j] <- J'[(e,d_j), n]).
                const double rs2 = 0.70710678118654752440;
                const double MAGICJ = 6755399441055744.0;      // 2^52 + 2^51
                {
                    unsigned e = (unsigned)prow0 / (unsigned)p.NDp;
                    int d = (int)((unsigned)prow0 - e * (unsigned)p.NDp);
                    unsigned w = e / (unsigned)p.n_elec;
                    int ie = (int)(e - w * (unsigned)p.n_elec);
                    // residual rows = own columns of the INPUT operand: from its digits when those are row-contiguous
                    // (6 x 128-bit loads, L2-resident: the TMA producer fetched this tile microseconds ago), else from
                    // the fp64 rows the layer-0 kernel wrote
                    uint4 rd[OZ_S];
                    if (RES && BMN && nvalid > 0) {
#pragma unroll
                        for (int t = 0; t < OZ_S; ++t)
                            rd[t] = *reinterpret_cast<const uint4*>(p.Ad + ((((prow0 >> 6) * OZ_S + t) * p.K + n) << 6) + (prow0 & 63));
                    }
#pragma unroll
                    for (int b = 0; b < EPI_COLS / 8; ++b) {
                        if (8 * b < nvalid) {                               // warp-uniform; aligned groups of 8 rows share an electron
                            const double t = p.T[(long long)e * p.ldt + n];
                            const double d1 = 1.0 - t * t;
                            const double* gp = p.G + ((long long)w * p.NDg + d) * p.ldg + n;
                            double gv[8], sc[8], rv[8];
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) {
                                sc[jj] = __ldg(sap + 8 * b + jj);
                                gv[jj] = gp[(long long)jj * p.ldg];
                                if (RES && !BMN) rv[jj] = p.R[(prow0 + 8 * b + jj) * (long long)p.ldr + n];
                            }
                            if (RES && BMN) {
                                // rows 8b .. 8b+7 = words 2b, 2b+1 of every slice: 4 x 4 byte transposes give the
                                // biased 48-bit integer (hi16 : lo32) of each row; undo the bias, rebuild the double
#pragma unroll
                                for (int hw = 0; hw < 2; ++hw) {
                                    const int wi = 2 * b + hw;
                                    unsigned W[OZ_S];
#pragma unroll
                                    for (int t2 = 0; t2 < OZ_S; ++t2)
                                        W[t2] = wi == 0 ? rd[t2].x : wi == 1 ? rd[t2].y : wi == 2 ? rd[t2].z : rd[t2].w;
                                    const unsigned a54 = __byte_perm(W[5], W[4], 0x5140), b54 = __byte_perm(W[5], W[4], 0x7362);
                                    const unsigned a32 = __byte_perm(W[3], W[2], 0x5140), b32 = __byte_perm(W[3], W[2], 0x7362);
                                    const unsigned a10 = __byte_perm(W[1], W[0], 0x5140), b10 = __byte_perm(W[1], W[0], 0x7362);
                                    unsigned lo[4], hi[4];
                                    lo[0] = __byte_perm(a54, a32, 0x5410); lo[1] = __byte_perm(a54, a32, 0x7632);
                                    lo[2] = __byte_perm(b54, b32, 0x5410); lo[3] = __byte_perm(b54, b32, 0x7632);
                                    hi[0] = __byte_perm(a10, 0u, 0x4410); hi[1] = __byte_perm(a10, 0u, 0x4432);
                                    hi[2] = __byte_perm(b10, 0u, 0x4410); hi[3] = __byte_perm(b10, 0u, 0x4432);
#pragma unroll
                                    for (int r4 = 0; r4 < 4; ++r4) {
                                        const unsigned l2 = lo[r4] ^ 0x80808080u;
                                        const int hh = (int)(short)(hi[r4] ^ 0x80u);           // digit 0 is the signed top byte
                                        const unsigned l = l2 - 0x80808080u;
                                        const int hq = hh - 0x80 - (l2 < 0x80808080u ? 1 : 0);
                                        // (2^-40 / sqrt 2: the residual's 1/sqrt2 is folded into the row scale)
                                        rv[4 * hw + r4] = (__hiloint2double(0x43380000 + hq, (int)l) - MAGICJ) * (sc[4 * hw + r4] * (9.094947017729282e-13 * 0.70710678118654752440));
                                    }
                                }
                            }
                            double s0 = 0.0, s1 = 0.0;
                            const double d1s = RES ? d1 * rs2 : d1;            // (r + d1 z) / sqrt2 = r / sqrt2 + (d1 / sqrt2) z
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) {
                                const double zj = fma(zz[8 * b + jj], sc[jj] * sbn, gv[jj]);
                                if (jj & 1) s1 = fma(zj, zj, s1); else s0 = fma(zj, zj, s0);
                                if (RES) zz[8 * b + jj] = fma(d1s, zj, BMN ? rv[jj] : rv[jj] * rs2);
                                else zz[8 * b + jj] = d1s * zj;
                            }
                            p.SP[((prow0 + 8 * b) >> 3) * (long long)p.ldt + n] = s0 + s1;
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) zz[8 * b + jj] = 0.0;
                        }
                        d += 8;
                        if (d >= p.NDp) { d -= p.NDp; ++e; if (++ie == p.n_elec) { ie = 0; ++w; } }
                    }
                }
                // B2: row maximum over the 2 x 128 channels and the pair-mean columns.
                const uint32_t crank = cluster_ctarank();
                const int slot = (int)(it & 1u);
                const int pmc = (int)crank * 32 + lane;                      // pair-mean column of this lane (this CTA's half)
                const bool pmv_ok = p.PM != nullptr && pmc < p.npm;
                double pmv[4];
                {
                    unsigned hw[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) hw[j] = (unsigned)__double2hiint(zz[j]) & 0x7fffffffu;
                    const unsigned m = warp_rowmax16(hw, lane);
                    if ((lane & 1) == 0) xown[slot][cg][q][((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1)] = m;
                    unsigned hp[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int j = 4 * q + r;                             // this warp digitises pair-mean rows 4q .. 4q+3 of the 16
                        pmv[r] = (pmv_ok && j < nvalid) ? p.PM[(prow0 + j) * (long long)p.npm + pmc] : 0.0;
                        hp[r] = (unsigned)__double2hiint(pmv[r]) & 0x7fffffffu;
                    }
                    const unsigned mp = warp_rowmax4(hp, lane);
                    if ((lane & 7) == 0) xpm[slot][cg][4 * q + ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1)] = mp;
                }
                named_bar_sync(1 + cg, 128);                                 // the four warps of this 16-row group
                unsigned mine = 0u;
                if (lane < 16) {
                    mine = max(max(xown[slot][cg][0][lane], xown[slot][cg][1][lane]), max(xown[slot][cg][2][lane], xown[slot][cg][3][lane]));
                    mine = max(mine, xpm[slot][cg][lane]);
                    if (q == 0) {
                        // one warp of the group sends the 16 maxima to the other CTA: asynchronous DSMEM stores that
                        // complete transaction bytes on ITS barrier (no fences, no L1 invalidation on either side)
                        if (lane == 0) mbar_expect_tx(&xbar[slot][cg], 64);  // the 64 bytes the other CTA sends to us
                        st_async_u32(map_to_cta(smem_u32(&xrem[slot][cg * 16 + lane]), crank ^ 1u), mine,
                                     map_to_cta(smem_u32(&xbar[slot][cg]), crank ^ 1u));
                    }
                }
                mbar_wait(&xbar[slot][cg], (it >> 1) & 1u);
                if (lane < 16) mine = max(mine, xrem[slot][cg * 16 + lane]);
                // B3: digits of the own channel (16 rows x 6 slices -> six 128-bit stores) and of this CTA's share of
                // the pair-mean columns (4 rows x 6 slices -> six 32-bit stores), row scales.
                {
                    // lane j < 16 turns the maximum of row j into the high word of 2^(46-e) once (0 = inf / nan in the
                    // row); every lane then fetches the 16 words with one shuffle each
                    unsigned fhi = 0u;
                    if (lane < 16) {
                        const bool badr = mine >= 0x7ff00000u;
                        int ex = (int)(mine >> 20) - 1022;
                        if (ex < -900) ex = -900;
                        fhi = badr ? 0u : (unsigned)((1023 + 8 * OZ_S - 2 - ex) << 20);
                        if (crank == 0 && q == 0 && lane < nvalid)
                            p.sa_out[prow0 + lane] = badr ? __longlong_as_double(0x7ff8000000000000LL) : __hiloint2double((1023 + ex - 6) << 20, 0);
                    }
                    unsigned dl[16], dh[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const unsigned fw = __shfl_sync(0xffffffffu, fhi, j);
                        const bool bad = fw == 0u;
                        const double f = __hiloint2double((int)fw, 0);
                        const double tq = bad ? MAGICJ : fma(zz[j], f, MAGICJ);
                        unsigned l = (unsigned)__double2loint(tq), h = (unsigned)__double2hiint(tq);
                        const unsigned l2 = l + 0x80808080u;
                        h += 0x80u + (l2 < l ? 1u : 0u);
                        dl[j] = l2 ^ 0x80808080u;
                        dh[j] = h ^ 0x80u;
                        if ((j >> 2) == q) {                                 // pair-mean rows of this warp: same scale
                            const double tp = bad ? MAGICJ : fma(pmv[j & 3], f, MAGICJ);
                            unsigned pl = (unsigned)__double2loint(tp), ph = (unsigned)__double2hiint(tp);
                            const unsigned pl2 = pl + 0x80808080u;
                            ph += 0x80u + (pl2 < pl ? 1u : 0u);
                            pmv[j & 3] = __hiloint2double((int)(ph ^ 0x80u), (int)(pl2 ^ 0x80808080u));   // (digits 0,1 | digits 2..5)
                        }
                    }
                    // byte transposes: word g of slice s = digit s of rows 4g .. 4g+3
                    // blocked layout [64-row block][slice][k][64 rows]: a tile is one contiguous 6 Kout 64-byte region
                    signed char* ob = p.Dout + ((((prow0 >> 6) * OZ_S) * p.Kout + n) << 6) + (prow0 & 63);
                    const long long spitch = (long long)p.Kout << 6;
                    unsigned sw[OZ_S][4];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const unsigned *L = dl + 4 * g, *Hh = dh + 4 * g;
                        const unsigned t01 = __byte_perm(L[0], L[1], 0x5140), t23 = __byte_perm(L[2], L[3], 0x5140);
                        const unsigned u01 = __byte_perm(L[0], L[1], 0x7362), u23 = __byte_perm(L[2], L[3], 0x7362);
                        const unsigned h01 = __byte_perm(Hh[0], Hh[1], 0x5140), h23 = __byte_perm(Hh[2], Hh[3], 0x5140);
                        sw[5][g] = __byte_perm(t01, t23, 0x5410); sw[4][g] = __byte_perm(t01, t23, 0x7632);
                        sw[3][g] = __byte_perm(u01, u23, 0x5410); sw[2][g] = __byte_perm(u01, u23, 0x7632);
                        sw[1][g] = __byte_perm(h01, h23, 0x5410); sw[0][g] = __byte_perm(h01, h23, 0x7632);
                    }
#pragma unroll
                    for (int t2 = 0; t2 < OZ_S; ++t2)
                        *reinterpret_cast<uint4*>(ob + t2 * spitch) = make_uint4(sw[t2][0], sw[t2][1], sw[t2][2], sw[t2][3]);
                    if (pmv_ok) {
                        unsigned L[4], Hh[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) { L[r] = (unsigned)__double2loint(pmv[r]); Hh[r] = (unsigned)__double2hiint(pmv[r]); }
                        const unsigned t01 = __byte_perm(L[0], L[1], 0x5140), t23 = __byte_perm(L[2], L[3], 0x5140);
                        const unsigned u01 = __byte_perm(L[0], L[1], 0x7362), u23 = __byte_perm(L[2], L[3], 0x7362);
                        const unsigned h01 = __byte_perm(Hh[0], Hh[1], 0x5140), h23 = __byte_perm(Hh[2], Hh[3], 0x5140);
                        signed char* op = p.Dout + ((((prow0 >> 6) * OZ_S) * p.Kout + p.N + pmc) << 6) + (prow0 & 63) + 4 * q;
                        *reinterpret_cast<unsigned*>(op + 5 * spitch) = __byte_perm(t01, t23, 0x5410);
                        *reinterpret_cast<unsigned*>(op + 4 * spitch) = __byte_perm(t01, t23, 0x7632);
                        *reinterpret_cast<unsigned*>(op + 3 * spitch) = __byte_perm(u01, u23, 0x5410);
                        *reinterpret_cast<unsigned*>(op + 2 * spitch) = __byte_perm(u01, u23, 0x7632);
                        *reinterpret_cast<unsigned*>(op + 1 * spitch) = __byte_perm(h01, h23, 0x5410);
                        *reinterpret_cast<unsigned*>(op + 0 * spitch) = __byte_perm(h01, h23, 0x7632);
                    }
                }
            } else if (MODE == OZ_VALUE || MODE == OZ_LAP) {
                // rows are electrons e = prow0 + j (value or Laplacian row of each); walker = e / n_elec
                unsigned w = (unsigned)prow0 / (unsigned)p.n_elec;
                int ie = (int)((unsigned)prow0 - w * (unsigned)p.n_elec);
                const double rs2v = 0.70710678118654752440;
                const double bias = (MODE == OZ_VALUE && nv) ? p.colbias[n] : 0.0;
                const int grow = p.NDp + (MODE == OZ_VALUE ? 0 : 1);
#pragma unroll
                for (int j = 0; j < EPI_COLS; ++j) {
                    if (j < nvalid && nv) {
                        const long long e = prow0 + j;
                        const double z = fma(zz[j], __ldg(sap + j) * sbn, p.G[((long long)w * p.NDg + grow) * p.ldg + n]);
                        double o;
                        if (MODE == OZ_VALUE) {
                            o = tanh(z + bias);
                            p.Tout[e * p.ldt + n] = o;
                        } else {
                            const double t = p.T[e * p.ldt + n], sv = p.S[e * p.ldt + n];
                            const double d1 = 1.0 - t * t;
                            o = d1 * z - 2.0 * t * d1 * sv;
                        }
                        if (RES) o = (p.R[e * (long long)p.ldr + n] + o) * rs2v;
                        p.C[e * (long long)p.ldc + n] = o;
                    }
                    if (++ie == p.n_elec) { ie = 0; ++w; }
                }
            } else if (MODE == OZ_PLAIN) {
                double* cptr = p.C + prow0 * (long long)p.ldc + n;
#pragma unroll
                for (int j = 0; j < EPI_COLS; ++j) {
                    if (nv && j < nvalid) *cptr = zz[j] * (__ldg(sap + j) * sbn);
                    cptr += p.ldc;
                }
            } else {
                // OZ_ORBJ: group = walker, row in group = is*NDp + d, channel n = 2*pp + (re|im), pp = k*n_s + o.
                const int pp = n >> 1, im = n & 1;
                const int kdet = pp / p.n_orb, oo = pp - kdet * p.n_orb;
                const int ND = 3 * p.n_elec;
                const long long ns2 = 2LL * p.n_rows_mat * p.n_orb;
                // gskip (a multiple of 8): with blocked digits the window of a group starts at a 64-row block boundary; the
                // rows before the group's first row belong to the group before and are skipped block by block
                const long long qg = q0 - gskip;                           // first row of this warp inside the group proper
                const long long rpg_real = p.rpg;
                double* dab = p.DA + 2 * ((((long long)grp * p.n_det + kdet) * p.NDp) * p.n_rows_mat * p.n_orb + oo) + im;
#pragma unroll
                for (int b = 0; b < EPI_COLS / 8; ++b) {
                    const long long qb = qg + 8 * b;
                    if (qb >= 0 && qb < rpg_real) {                         // warp-uniform
                        const int is = (int)((unsigned)qb / (unsigned)p.NDp);
                        const int d = (int)qb - is * p.NDp;
                        const long long e = grp * p.n_elec + p.off_s + is;
                        double Ex = 0.0, Ey = 0.0;
                        if (nv) {
                            const double2 E = *reinterpret_cast<const double2*>(p.etab + e * 10LL * p.npar_max + 2 * pp);
                            Ex = E.x; Ey = E.y;
                        }
                        double* dp = dab + d * ns2 + 2LL * (p.row0 + is) * p.n_orb;
                        double* yp = p.YOWN + 2 * (e * 3LL * p.npar_max + pp) + im;
                        const int c0 = d - 3 * (p.off_s + is);              // own-coordinate index of column 0
                        if (full) orbj_block8<true>(zz + 8 * b, sap + 8 * b, sbn, Ex, Ey, im, dp, ns2, yp, 2LL * p.npar_max, d, ND, c0, true);
                        else orbj_block8<false>(zz + 8 * b, sap + 8 * b, sbn, Ex, Ey, im, dp, ns2, yp, 2LL * p.npar_max, d, ND, c0, nv);
                    }
                }
            }
        }
        if ((OZ_PROF && (p.dbg & 32)) && warp == 2 && lane == 0) {
            if (te2_prev) prof_pb += clock64() - te2_prev;
            atomicAdd(&g_oz_prof[MODE][4], (unsigned long long)prof_w);
            atomicAdd(&g_oz_prof[MODE][5], (unsigned long long)prof_pa);
            atomicAdd(&g_oz_prof[MODE][6], (unsigned long long)prof_pb);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (JD) cluster_sync_all();          // neither CTA leaves while the other may still write into its shared memory
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ B, int K, int N, double* __restrict__ Bt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // index into Bt [N][K]
    if (i >= (long long)K * N) return;
    const int n = (int)(i / K), k = (int)(i - (long long)n * K);
    Bt[i] = B[(long long)k * N + n];
}

// S[e][n] = sum over the NDp/8 aligned 8-row groups of electron e of the partial sums written by OZ_JACD (fixed order)
__global__ void __launch_bounds__(256) sp_reduce_kernel(const double* __restrict__ SP, int ld, int nblk, int H,
                                                        double* __restrict__ S) {
    const long long e = blockIdx.x;
    for (int n = threadIdx.x; n < H; n += blockDim.x) {
        const double* src = SP + e * (long long)nblk * ld + n;
        double acc = 0.0;
        for (int b = 0; b < nblk; ++b) acc += src[(long long)b * ld];
        S[e * (long long)ld + n] = acc;
    }
}

// decode 4 rows (one 32-bit word per digit slice, byte r = row r) of row-contiguous digits into the exact integers
// q_r = x_r 2^(46-e) as doubles (the inverse of the digit formation in slice_row_warp / OZ_JACD)
__device__ __forceinline__ void decode4(const unsigned (&W)[OZ_S], double (&q)[4]) {
    const double MAGIC = 6755399441055744.0;
    const unsigned a54 = __byte_perm(W[5], W[4], 0x5140), b54 = __byte_perm(W[5], W[4], 0x7362);
    const unsigned a32 = __byte_perm(W[3], W[2], 0x5140), b32 = __byte_perm(W[3], W[2], 0x7362);
    const unsigned a10 = __byte_perm(W[1], W[0], 0x5140), b10 = __byte_perm(W[1], W[0], 0x7362);
    unsigned lo[4], hi[4];
    lo[0] = __byte_perm(a54, a32, 0x5410); lo[1] = __byte_perm(a54, a32, 0x7632);
    lo[2] = __byte_perm(b54, b32, 0x5410); lo[3] = __byte_perm(b54, b32, 0x7632);
    hi[0] = __byte_perm(a10, 0u, 0x4410); hi[1] = __byte_perm(a10, 0u, 0x4432);
    hi[2] = __byte_perm(b10, 0u, 0x4410); hi[3] = __byte_perm(b10, 0u, 0x4432);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const unsigned l2 = lo[r] ^ 0x80808080u;
        const int hh = (int)(short)(hi[r] ^ 0x80u);
        const unsigned l = l2 - 0x80808080u;
        const int hq = hh - 0x80 - (l2 < 0x80808080u ? 1 : 0);
        q[r] = __hiloint2double(0x43380000 + hq, (int)l) - MAGIC;
    }
}

// Spin-channel means of Jacobian rows that exist only as blocked row-contiguous digits ([block][slice][k][64]): a warp owns
// (walker w, channel k), a lane 8 consecutive directions d (one 64-bit load per slice and electron, the lanes of a
// warp read NDp contiguous bytes); GIN[(w*NDg + d)*ldgin + s*C + k] = mean_{i in s} sa[r] 2^-40 q[r, k].
__global__ void __launch_bounds__(256) means_digits_kernel(const signed char* __restrict__ Ad, const double* __restrict__ sa,
                                                           int K, long long Rp, int C, int n_up, int n_elec, int NDp,
                                                           int NDg, double* __restrict__ GIN, int ldgin) {
    const int w = blockIdx.y;
    const int G8 = NDp / 8;
    // work item = (channel k, group g of 8 consecutive directions): consecutive threads walk the directions of one
    // channel (contiguous bytes), then the next channel
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= C * G8) return;
    const int k = item / G8, g = item - k * G8;
    const long long r0 = (long long)w * n_elec * NDp + 8 * g;
    (void)Rp;
    for (int s = 0; s < 2; ++s) {
        const int ibeg = s ? n_up : 0, iend = s ? n_elec : n_up;
        double sum[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sum[j] = 0.0;
#pragma unroll 2
        for (int i = ibeg; i < iend; ++i) {
            const long long r = r0 + (long long)i * NDp;
            uint2 dg[OZ_S];
#pragma unroll
            for (int t = 0; t < OZ_S; ++t)
                dg[t] = *reinterpret_cast<const uint2*>(Ad + ((((r >> 6) * OZ_S + t) * K + k) << 6) + (r & 63));
            const double4 sA = *reinterpret_cast<const double4*>(sa + r), sB = *reinterpret_cast<const double4*>(sa + r + 4);
            unsigned Wx[OZ_S], Wy[OZ_S];
#pragma unroll
            for (int t = 0; t < OZ_S; ++t) { Wx[t] = dg[t].x; Wy[t] = dg[t].y; }
            double q0[4], q1[4];
            decode4(Wx, q0);
            decode4(Wy, q1);
            sum[0] = fma(q0[0], sA.x, sum[0]); sum[1] = fma(q0[1], sA.y, sum[1]);
            sum[2] = fma(q0[2], sA.z, sum[2]); sum[3] = fma(q0[3], sA.w, sum[3]);
            sum[4] = fma(q1[0], sB.x, sum[4]); sum[5] = fma(q1[1], sB.y, sum[5]);
            sum[6] = fma(q1[2], sB.z, sum[6]); sum[7] = fma(q1[3], sB.w, sum[7]);
        }
        const double inv = 9.094947017729282e-13 / (double)(iend - ibeg);      // 2^-40 / n_s
#pragma unroll
        for (int j = 0; j < 8; ++j) GIN[((long long)w * NDg + 8 * g + j) * ldgin + s * C + k] = sum[j] * inv;
    }
}

// fp64 rows -> row-contiguous digits (probe / tests only: byte-scattered stores)
__global__ void __launch_bounds__(256) slice_rows_mn_kernel(const double* __restrict__ A, int lda, long long rows, int K,
                                                            long long Rp, signed char* __restrict__ Ad, double* __restrict__ sa) {
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    unsigned mxh = 0u;
    for (int k = lane; k < K; k += 32) mxh = max(mxh, (unsigned)__double2hiint(A[r * lda + k]) & 0x7fffffffu);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mxh = max(mxh, __shfl_xor_sync(0xffffffffu, mxh, off));
    int e = (int)(mxh >> 20) - 1022;
    if (e < -900) e = -900;
    const double f = __hiloint2double((1023 + 8 * OZ_S - 2 - e) << 20, 0);
    if (lane == 0) sa[r] = __hiloint2double((1023 + e - 6) << 20, 0);
    const double MAGIC = 6755399441055744.0;
    for (int k = lane; k < K; k += 32) {
        const double t = fma(A[r * lda + k], f, MAGIC);
        unsigned l = (unsigned)__double2loint(t), h = (unsigned)__double2hiint(t);
        const unsigned l2 = l + 0x80808080u;
        h += 0x80u + (l2 < l ? 1u : 0u);
        l = l2 ^ 0x80808080u; h ^= 0x80u;
        const unsigned dgt[OZ_S] = {h >> 8, h, l >> 24, l >> 16, l >> 8, l};
#pragma unroll
        for (int t2 = 0; t2 < OZ_S; ++t2) Ad[((((r >> 6) * OZ_S + t2) * K + k) << 6) + (r & 63)] = (signed char)dgt[t2];
    }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

// 4-D map over int8 digits [group][row][slice][k]: dims (k, row, slice, group)
int make_map(CUtensorMap* tm, const signed char* base, int K, long long rows, long long group_stride_rows, int n_groups,
             int box_rows) {
    auto enc = get_encode();
    if (!enc) { ds_set_error("cuTensorMapEncodeTiled is not available from the driver"); return -2; }
    cuuint64_t dims[4] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)OZ_S, (cuuint64_t)n_groups};
    cuuint64_t strides[3] = {(cuuint64_t)OZ_S * K, (cuuint64_t)K, (cuuint64_t)group_stride_rows * OZ_S * K};
    cuuint32_t box[4] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows, (cuuint32_t)OZ_S, 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<signed char*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ds_set_error("cuTensorMapEncodeTiled failed with code %d (K=%d rows=%lld groups=%d)", (int)r, K, rows, n_groups);
        return -2;
    }
    return 0;
}

// 4-D map over blocked row-contiguous digits [64-row block][slice][k][64 rows]: dims (row in block, k, slice, block)
int make_map_mn(CUtensorMap* tm, const signed char* base, int K, long long Rp) {
    auto enc = get_encode();
    if (!enc) { ds_set_error("cuTensorMapEncodeTiled is not available from the driver"); return -2; }
    cuuint64_t dims[4] = {(cuuint64_t)OZ_TN, (cuuint64_t)K, (cuuint64_t)OZ_S, (cuuint64_t)(Rp / OZ_TN)};
    cuuint64_t strides[3] = {(cuuint64_t)OZ_TN, (cuuint64_t)K * OZ_TN, (cuuint64_t)OZ_S * K * OZ_TN};
    cuuint32_t box[4] = {(cuuint32_t)OZ_TN, (cuuint32_t)OZ_BK, (cuuint32_t)OZ_S, 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<signed char*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        ds_set_error("cuTensorMapEncodeTiled (blocked digits) failed with code %d (K=%d Rp=%lld)", (int)r, K, Rp);
        return -2;
    }
    return 0;
}

template <int MODE, bool RES, int TN, int ND, bool BMN>
int launch_tn(const OzParams& p, cudaStream_t stream) {
    static bool configured = false;
    static int n_sm = 0;
    constexpr int SMEM_T = OzCfg<TN>::SMEM_T;
    if (!configured) {
        DS_CUDA_CHECK(cudaFuncSetAttribute(oz_gemm_kernel<MODE, RES, TN, ND, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_T));
        int dev = 0;
        DS_CUDA_CHECK(cudaGetDevice(&dev));
        DS_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    CUtensorMap tmW, tmA;
    if (int rc = make_map(&tmW, p.Wd, p.K, p.N, p.N, 1, OZ_TM)) return rc;
    if (BMN) {
        if (int rc = make_map_mn(&tmA, p.Ad, p.K, p.Rp_in)) return rc;
    } else if (int rc = make_map(&tmA, p.Ad + p.goff * (long long)OZ_S * p.K, p.K, p.rpg, p.gstride, p.n_groups, TN)) return rc;
    // blocked digits: a group's window starts up to 56 rows before its first row (64-row block boundary)
    // (only groups that can start inside a block need it: a single group at row 0 -- OZ_JACD, whose extra tile would
    //  write digits past the operand -- starts on a block boundary)
    const bool windows = BMN && (p.n_groups > 1 || (p.goff & 63) != 0);
    const int tpg = (int)((p.rpg + (windows ? 56 : 0) + TN - 1) / TN);
    const int n_cb = (p.N + OZ_TM - 1) / OZ_TM;
    const long long n_tiles = (long long)tpg * p.n_groups * n_cb;
    DS_REQUIRE(n_tiles + 2LL * n_sm < (1LL << 31), "oz_gemm: %lld tiles in one launch (the kernel indexes tiles with 32 bits)", n_tiles);
    int grid = (int)(n_tiles < n_sm ? n_tiles : n_sm);
    if (MODE == OZ_JACD) {
        // CTA pairs (2j, 2j+1) = the two channel blocks of one row tile: launched as clusters of two so that the
        // row maxima can cross through distributed shared memory (n_cb == 2: tile % 2 = blockIdx.x % 2 when the grid is even)
        grid &= ~1;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(OZ_THREADS);
        cfg.dynamicSmemBytes = SMEM_T;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        DS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, oz_gemm_kernel<MODE, RES, TN, ND, BMN>, tmW, tmA, p, tpg, n_cb, n_tiles));
        return 0;
    }
    oz_gemm_kernel<MODE, RES, TN, ND, BMN><<<grid, OZ_THREADS, SMEM_T, stream>>>(tmW, tmA, p, tpg, n_cb, n_tiles);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// 64-row tiles, one accumulator set of 6 diagonals.  Two measured-and-rejected variants are no longer built:
// 32-row tiles with double-buffered TMEM (profiles/r1_probe_tn32.log: hides the TMEM read but MMA-only 1.00 ms vs
// 0.75 ms and W re-read per 32 rows, 1.42 ms vs 1.14 ms in total) and 5 diagonals / 15 slice products
// (profiles/r2_diag5.log: +6 % local energies/s but max |dE_L| 2e-8 .. 2e-7 Ha against the fp64 path, over the 1e-8 budget).
template <int MODE, bool RES>
int launch(const OzParams& p, cudaStream_t stream) {
    return launch_tn<MODE, RES, 64, 6, false>(p, stream);
}

}  // namespace

// debug: read (and optionally clear) the dbg & 32 role clocks, 8 modes x 8 counters
int ds_oz_prof_read(unsigned long long* out, int reset) {
    DS_CUDA_CHECK(cudaDeviceSynchronize());
    DS_CUDA_CHECK(cudaMemcpyFromSymbol(out, g_oz_prof, sizeof(unsigned long long) * 64));
    if (reset) {
        static const unsigned long long zeros[64] = {};
        DS_CUDA_CHECK(cudaMemcpyToSymbol(g_oz_prof, zeros, sizeof(zeros)));
    }
    return 0;
}

int ds_launch_slice_rows(const double* A, int lda, long long rows, int K, signed char* Ad, double* sa,
                         cudaStream_t stream) {
    if (rows <= 0) return 0;
    DS_REQUIRE(K % 8 == 0 && K <= 512 && lda % 2 == 0, "slice_rows: K must be a multiple of 8 and <= 512 (K=%d lda=%d)", K, lda);
    const int wpb = 8;
    slice_rows_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, stream>>>(A, lda, rows, K, Ad, sa);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_slice_means(const double* A, int lda, int K, int C, int n_walkers, int n_up, int n_elec, int NDp, int NDg,
                          signed char* Ad, double* sa, double* GIN, int ldgin, cudaStream_t stream) {
    if (n_walkers <= 0) return 0;
    DS_REQUIRE(K % 8 == 0 && K <= 512 && lda % 2 == 0 && C % 8 == 0 && C <= K && ldgin % 2 == 0,
               "slice_means: K, C must be multiples of 8, C <= K <= 512 (K=%d C=%d lda=%d)", K, C, lda);
    const long long warps = (long long)n_walkers * NDp;
    const int wpb = 8;
    slice_means_kernel<<<(unsigned)((warps + wpb - 1) / wpb), wpb * 32, 0, stream>>>(A, lda, K, C, n_walkers, n_up, n_elec,
                                                                                     NDp, NDg, Ad, sa, GIN, ldgin);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_sp_reduce(const double* SP, int ld, long long n_elec_rows, int blocks_per_electron, int H, double* S,
                        cudaStream_t stream) {
    if (n_elec_rows <= 0) return 0;
    sp_reduce_kernel<<<(unsigned)n_elec_rows, 256, 0, stream>>>(SP, ld, blocks_per_electron, H, S);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_means_digits(const signed char* Ad, const double* sa, int K, long long Rp, int C, int n_walkers, int n_up,
                           int n_elec, int NDp, int NDg, double* GIN, int ldgin, cudaStream_t stream) {
    if (n_walkers <= 0) return 0;
    DS_REQUIRE(NDp % 8 == 0 && C <= K && Rp % 64 == 0, "means_digits: NDp must be a multiple of 8 (NDp=%d C=%d K=%d)", NDp, C, K);
    dim3 grid((unsigned)((C * (NDp / 8) + 255) / 256), (unsigned)n_walkers);
    means_digits_kernel<<<grid, 256, 0, stream>>>(Ad, sa, K, Rp, C, n_up, n_elec, NDp, NDg, GIN, ldgin);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_slice_rows_mn(const double* A, int lda, long long rows, int K, long long Rp, signed char* Ad, double* sa,
                            cudaStream_t stream) {
    if (rows <= 0) return 0;
    slice_rows_mn_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(A, lda, rows, K, Rp, Ad, sa);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_transpose(const double* B, int K, int N, double* Bt, cudaStream_t stream) {
    const long long tot = (long long)K * N;
    if (tot <= 0) return 0;
    transpose_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(B, K, N, Bt);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_oz_gemm(const OzParams& p_in, int mode, bool residual, cudaStream_t stream) {
    if (p_in.rpg <= 0 || p_in.n_groups <= 0 || p_in.N <= 0) return 0;
    // DS_OZ_OPT: probe bits ORed into dbg (see OzParams::dbg; 32 = role clocks, needs a -DDS_OZ_PROF build)
    static const int opt_bits = getenv("DS_OZ_OPT") ? atoi(getenv("DS_OZ_OPT")) : 0;
    OzParams p = p_in;
    p.dbg |= opt_bits;
    DS_REQUIRE(p.K % OZ_BK == 0 && p.K >= OZ_BK, "oz_gemm: K must be a multiple of %d (K=%d)", OZ_BK, p.K);
    DS_REQUIRE((reinterpret_cast<uintptr_t>(p.Ad) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.Wd) & 15) == 0,
               "oz_gemm: digit buffers must be 16-byte aligned");
    if (p.bmn) DS_REQUIRE(p.Rp_in % 64 == 0 && p.Rp_in >= 64, "oz_gemm: row pitch of row-contiguous digits must be a multiple of 64");
    if (mode == OZ_VALUE || mode == OZ_LAP)
        DS_REQUIRE(p.n_groups == 1 && p.rpg < (1LL << 31), "oz_gemm: value / Laplacian rows come as one group");
    if (mode == OZ_JACD) {
        DS_REQUIRE(p.N == 2 * OZ_TM && p.n_groups == 1 && p.goff == 0, "oz_gemm: the fused-digit mode needs exactly two channel blocks (N = %d)", p.N);
        DS_REQUIRE(p.Dout && p.sa_out && p.SP && p.Kout >= p.N + p.npm && p.npm <= 64 && p.npm >= 0 && p.Rp_out % 64 == 0 &&
                       p.Rp_out >= p.rpg && (reinterpret_cast<uintptr_t>(p.Dout) & 15) == 0,
                   "oz_gemm: bad digit output (Kout=%d npm=%d Rp=%lld)", p.Kout, p.npm, p.Rp_out);
        DS_REQUIRE(!residual || p.K >= p.N, "oz_gemm: the residual rows are the first N columns of the operand");
        DS_REQUIRE(!residual || p.bmn || p.R, "oz_gemm: residual rows need fp64 rows or row-contiguous input digits");
    }
    if (mode == OZ_JAC || mode == OZ_ORBJ || mode == OZ_JACD) {
        DS_REQUIRE(p.NDp % 8 == 0 && p.rpg % 8 == 0 && p.goff % 8 == 0 && p.gstride % 8 == 0,
                   "oz_gemm: Jacobian rows must come in aligned groups of 8 (NDp=%d rpg=%lld)", p.NDp, p.rpg);
        DS_REQUIRE(p.rpg * (mode == OZ_ORBJ ? p.n_groups : 1) < (1LL << 31), "oz_gemm: too many Jacobian rows in one launch");
    }
    switch (mode) {
        case OZ_PLAIN: return launch<OZ_PLAIN, false>(p, stream);
        case OZ_JAC: return residual ? launch<OZ_JAC, true>(p, stream) : launch<OZ_JAC, false>(p, stream);
        case OZ_ORBJ: return p.bmn ? launch_tn<OZ_ORBJ, false, 64, 6, true>(p, stream) : launch<OZ_ORBJ, false>(p, stream);
        case OZ_VALUE: return residual ? launch<OZ_VALUE, true>(p, stream) : launch<OZ_VALUE, false>(p, stream);
        case OZ_LAP: return residual ? launch<OZ_LAP, true>(p, stream) : launch<OZ_LAP, false>(p, stream);
        case OZ_JACD:
            if (p.bmn) return residual ? launch_tn<OZ_JACD, true, 64, 6, true>(p, stream) : launch_tn<OZ_JACD, false, 64, 6, true>(p, stream);
            return residual ? launch_tn<OZ_JACD, true, 64, 6, false>(p, stream) : launch_tn<OZ_JACD, false, 64, 6, false>(p, stream);
        case OZ_PLAIN + 100: return launch_tn<OZ_PLAIN, false, 64, 6, true>(p, stream);      // probe: plain product from row-contiguous digits
    }
    ds_set_error("oz_gemm: unknown mode %d", mode);
    return -1;
}

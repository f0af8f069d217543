// Periodic input features (network.py:249-302), the two-electron stream
// (network.py:525-528) and its spin-channel means (network.py:323,328), with analytic
// value / gradient / Laplacian jets when JETS is set.
//
// h_two[j,i] depends on x_j - x_i only (the pair stream never mixes), so a pair's jet
// w.r.t. r = x_j - x_i gives d/dx_j = +grad, d/dx_i = -grad, full Laplacian = 2 lap_r.
//
// Grid: one CTA per (walker w, electron i).  A warp owns a pair (j,i); lane = hidden
// channel of the pair stream.  Outputs are written straight into the operand
// matrices of the one-electron-stream GEMMs (own + pair-mean columns).
#include "kernels.cuh"
#include <stdlib.h>

namespace {

constexpr int FEAT_THREADS = 256;

template <int FT>
__device__ __forceinline__ Jet pick4(const Jet (&f)[FT], int q) {    // f[q] without a dynamically indexed register array
    Jet r = f[0];
#pragma unroll
    for (int t = 1; t < FT; ++t)
        if (q == t) r = f[t];
    return r;
}

// FT = features per pair: 4 (nu_distance) or 7 (tri_distance)
template <bool JETS, int FT>
__global__ void __launch_bounds__(FEAT_THREADS, 2) features_pair_kernel(const DsSys sys, const FeatParams fp) {
    const DsDims& dm = sys.d;
    // L counts the two-electron feature LEVELS here (dm.L, or dm.L + 1 with use_last_layer): level l feeds layer l,
    // the last level the orbital projection; pair layers 0 .. L-2
    const int N = dm.N, A = dm.A, P = dm.P, L = ds_pair_levels(dm), C0 = dm.C0, K0 = dm.K0, K1 = dm.K1, H = dm.H;
    const int NDp = dm.NDp, ND = dm.ND;
    constexpr int F = FT;
    const long long e = blockIdx.x;             // w*N + i
    const int w = (int)(e / N), i = (int)(e % N);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    extern __shared__ __align__(16) double sm[];
    // per-warp broadcast scratch [32 channels][v g0 g1 g2 l pad]: the pair layers read the jets of input channel c
    // from here (three 128-bit broadcast loads) instead of ten 32-bit shuffles per channel
    double* cs = sm + warp * (32 * 6);
    double* sx = sm + 8 * (32 * 6);             // [N*3] wrapped (simulation cell) positions
    double* sums = sx + 3 * N;                  // [2 spins][L levels][5 comps][32 lanes]
    double* wsm = sums + 2 * L * 5 * 32;        // pair weights: level l>=1: [P_in x P] + [P] bias
    const double* x = fp.X + (long long)w * 3 * N;

    for (int t = tid; t < N; t += blockDim.x) {
        double xi[3] = {x[3 * t], x[3 * t + 1], x[3 * t + 2]}, o[3];
        ds_wrap(sys.sim, xi, o);
        sx[3 * t] = o[0]; sx[3 * t + 1] = o[1]; sx[3 * t + 2] = o[2];
    }
    for (int t = tid; t < 2 * L * 5 * 32; t += blockDim.x) sums[t] = 0.0;
    {   // stage pair-stream weights
        int off = 0;
        for (int l = 0; l < L - 1; ++l) {
            int pin = (l == 0) ? F : P;
            for (int t = tid; t < pin * P; t += blockDim.x) wsm[off + t] = fp.Wp[l][t];
            for (int t = tid; t < P; t += blockDim.x) wsm[off + pin * P + t] = fp.bp[l][t];
            off += pin * P + P;
        }
    }
    __syncthreads();

    // ---- electron-atom features of electron i (primitive cell) ---------------
    if (tid < A) {
        double xi[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]}, px[3], d[3];
        ds_wrap(sys.prim, xi, px);
        for (int k = 0; k < 3; ++k) d[k] = px[k] - sys.atoms[3 * tid + k];
        Jet f[FT];
        if (FT == 7) ds_tri_distance<JETS>(sys.prim, d, f);
        else ds_nu_distance<JETS>(sys.prim, d, f);
#pragma unroll
        for (int q = 0; q < FT; ++q) {
            fp.A0V[e * K0 + tid * F + q] = f[q].v;
            if (JETS) {
                fp.A0L[e * K0 + tid * F + q] = f[q].l;
                long long r = e * NDp + 3 * i;
                fp.A0J[(r + 0) * K0 + tid * F + q] = f[q].g0;
                fp.A0J[(r + 1) * K0 + tid * F + q] = f[q].g1;
                fp.A0J[(r + 2) * K0 + tid * F + q] = f[q].g2;
            }
        }
        double* ra = fp.RAE + (e * A + tid) * DS_RAE_STRIDE;
#pragma unroll
        for (int q = 0; q < 4; ++q) {               // distance, then the relative vector (its first 3 components)
            ra[5 * q] = f[q].v; ra[5 * q + 1] = f[q].g0; ra[5 * q + 2] = f[q].g1; ra[5 * q + 3] = f[q].g2; ra[5 * q + 4] = f[q].l;
        }
    }

    // ---- pairs (j, i) ---------------------------------------------------------
    const double inv_up = 1.0 / dm.n_up, inv_dn = 1.0 / dm.n_dn;
    // one spin channel of partners j at a time: the per-lane sums of a channel stay in 5 registers per level
    for (int sj = 0; sj < 2; ++sj) {
    const int jbeg = sj ? dm.n_up : 0, jend = sj ? N : dm.n_up;
    const double invn = sj ? inv_dn : inv_up;
    double acc[DS_MAX_LAYERS][5];
#pragma unroll
    for (int l = 0; l < DS_MAX_LAYERS; ++l)
#pragma unroll
        for (int c = 0; c < 5; ++c) acc[l][c] = 0.0;

    for (int j = jbeg + warp; j < jend; j += nwarps) {
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = sx[3 * j + k] - sx[3 * i + k] + (j == i ? 1.0 : 0.0);
        Jet f[FT];
        if (FT == 7) ds_tri_distance<JETS>(sys.sim, d, f);
        else ds_nu_distance<JETS>(sys.sim, d, f);
        if (j == i) {
#pragma unroll
            for (int q = 0; q < FT; ++q) f[q] = jet_const(0.0);
        }
        Jet cur = jet_const(0.0);
        if (lane < F) cur = pick4(f, lane);
        const long long rj = e * NDp + 3 * j;      // Jacobian rows of directions (j, c)
        int woff = 0;
#pragma unroll
        for (int l = 0; l < DS_MAX_LAYERS; ++l) {
            if (l >= L) break;
            const int Pl = (l == 0) ? F : P;
            // accumulate spin-channel sums of level l
            if (lane < Pl) {
                acc[l][0] += cur.v;
                if (JETS) { acc[l][1] += cur.g0; acc[l][2] += cur.g1; acc[l][3] += cur.g2; acc[l][4] += cur.l; }
            }
            // Jacobian rows of directions (j,c), j != i: +grad / n_{s(j)} in the spin-of-j half
            if (JETS && j != i) {
                if (l == 0) {
                    if (lane < K0) {
                        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
                        if (lane >= C0) {
                            int s = (lane - C0) / F, q = (lane - C0) - s * F;
                            Jet fq = pick4(f, q);
                            if (s == sj) { v0 = fq.g0 * invn; v1 = fq.g1 * invn; v2 = fq.g2 * invn; }
                        }
                        fp.A0J[(rj + 0) * K0 + lane] = v0;
                        fp.A0J[(rj + 1) * K0 + lane] = v1;
                        fp.A0J[(rj + 2) * K0 + lane] = v2;
                    }
                } else if (lane < P) {
                    double* aj = fp.AJ[l];
                    const long long ldj = fp.ldj[l];
                    long long c_on = fp.joff[l] + sj * P + lane, c_off = fp.joff[l] + (1 - sj) * P + lane;
                    aj[(rj + 0) * ldj + c_on] = cur.g0 * invn; aj[(rj + 0) * ldj + c_off] = 0.0;
                    aj[(rj + 1) * ldj + c_on] = cur.g1 * invn; aj[(rj + 1) * ldj + c_off] = 0.0;
                    aj[(rj + 2) * ldj + c_on] = cur.g2 * invn; aj[(rj + 2) * ldj + c_off] = 0.0;
                }
            }
            if (l == L - 1) break;
            // pair layer l: z = cur . W + b ; tanh ; residual when shapes agree (l >= 1)
            const double* W = wsm + woff;
            const double* bb = W + Pl * P;
            Jet z = jet_const(lane < P ? bb[lane] : 0.0);
            if (JETS) {
                __syncwarp();
                *reinterpret_cast<double2*>(cs + lane * 6) = make_double2(cur.v, cur.g0);
                *reinterpret_cast<double2*>(cs + lane * 6 + 2) = make_double2(cur.g1, cur.g2);
                cs[lane * 6 + 4] = cur.l;
                __syncwarp();
#pragma unroll 4
                for (int c = 0; c < Pl; ++c) {
                    const double wv = (lane < P) ? W[c * P + lane] : 0.0;
                    const double2 a = *reinterpret_cast<const double2*>(cs + c * 6);
                    const double2 b = *reinterpret_cast<const double2*>(cs + c * 6 + 2);
                    const double cl = cs[c * 6 + 4];
                    z.v = fma(a.x, wv, z.v);
                    z.g0 = fma(a.y, wv, z.g0); z.g1 = fma(b.x, wv, z.g1); z.g2 = fma(b.y, wv, z.g2);
                    z.l = fma(cl, wv, z.l);
                }
            } else {
                for (int c = 0; c < Pl; ++c) {
                    const double wv = (lane < P) ? W[c * P + lane] : 0.0;
                    z.v = fma(__shfl_sync(0xffffffffu, cur.v, c), wv, z.v);
                }
            }
            Jet t;
            if (JETS) t = jet_tanh(z);
            else t = jet_const(tanh(z.v));
            if (l >= 1) {
                const double rs2 = 0.70710678118654752440;
                t = jet_scale(jet_add(cur, t), rs2);
            }
            cur = (lane < P) ? t : jet_const(0.0);
            woff += Pl * P + P;
        }
    }
    // cross-warp reduction of the sums of this spin channel, one warp after the other: a fixed summation order, so
    // that the features (and everything downstream) are bit-reproducible from run to run
    for (int wp = 0; wp < nwarps; ++wp) {
        if (warp == wp) {
#pragma unroll
            for (int l = 0; l < DS_MAX_LAYERS; ++l) {
                if (l >= L) break;
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    if (!JETS && c > 0) break;
                    sums[((sj * L + l) * 5 + c) * 32 + lane] += acc[l][c];
                }
            }
        }
        __syncthreads();
    }
    }   // spin channel
    __syncthreads();

    // ---- finalize: value / Laplacian rows, own-direction Jacobian rows, padding rows
#pragma unroll
    for (int l = 0; l < DS_MAX_LAYERS; ++l) {
        if (l >= L) break;
        const int Pl = (l == 0) ? F : P;
        for (int t = tid; t < 64; t += blockDim.x) {
            const int lane_c = t & 31, s = t >> 5;
            if (lane_c >= Pl) continue;
            const double* sv = sums + ((s * L + l) * 5) * 32 + lane_c;
            const double invn = s ? inv_dn : inv_up;
            const long long r = e * NDp + 3 * i;
            if (l == 0) {
                long long col = C0 + s * F + lane_c;
                fp.A0V[e * K0 + col] = sv[0] * invn;
                if (JETS) {
                    fp.A0L[e * K0 + col] = 2.0 * sv[4 * 32] * invn;
                    fp.A0J[(r + 0) * K0 + col] = -sv[1 * 32] * invn;
                    fp.A0J[(r + 1) * K0 + col] = -sv[2 * 32] * invn;
                    fp.A0J[(r + 2) * K0 + col] = -sv[3 * 32] * invn;
                }
            } else {
                long long col = H + s * P + lane_c;
                fp.AV[l][e * K1 + col] = sv[0] * invn;
                if (JETS) {
                    fp.AL[l][e * K1 + col] = 2.0 * sv[4 * 32] * invn;
                    const long long ldj = fp.ldj[l], cj = fp.joff[l] + s * P + lane_c;
                    fp.AJ[l][(r + 0) * ldj + cj] = -sv[1 * 32] * invn;
                    fp.AJ[l][(r + 1) * ldj + cj] = -sv[2 * 32] * invn;
                    fp.AJ[l][(r + 2) * ldj + cj] = -sv[3 * 32] * invn;
                }
            }
        }
    }
    if (JETS) {
        // zero rows of the padding directions d in [ND, NDp)
        const int npad = NDp - ND;
        for (int t = tid; t < npad * K0; t += blockDim.x)
            fp.A0J[(e * NDp + ND + t / K0) * K0 + t % K0] = 0.0;
#pragma unroll
        for (int l = 1; l < DS_MAX_LAYERS; ++l) {
            if (l >= L) break;
            for (int t = tid; t < npad * 2 * P; t += blockDim.x)
                fp.AJ[l][(e * NDp + ND + t / (2 * P)) * fp.ldj[l] + fp.joff[l] + t % (2 * P)] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------
// Value-only variant (log psi of a Metropolis proposal, the forward of the gradient pass): LANE = pair.
// A warp owns (walker w, electron i, spin channel s) and its lanes the partners j of that channel, so the distance
// features are computed once per pair (not once per lane), the pair-layer products are plain register FMAs against
// weights broadcast from shared memory (no shuffles), and the sum over partners is one shared-memory transpose per
// level.  P <= 32 channels live in registers (z[], cur[]).
// ---------------------------------------------------------------------------
constexpr int FV_WARPS = 4;
constexpr int FV_NB = 4;          // output channels per step of the rolled channel loop (independent FMA chains)

// shared-memory doubles of the staged pair weights: per layer W^T [32][pin_pad] + b [32]
__host__ __device__ inline int fv_pin_pad(int pin) { return (pin + 1) & ~1; }

template <int FT>
__global__ void __launch_bounds__(FV_WARPS * 32, 4) features_value_kernel(const DsSys sys, const FeatParams fp, long long n_items) {
    const DsDims& dm = sys.d;
    // L counts the two-electron feature LEVELS here (dm.L, or dm.L + 1 with use_last_layer): level l feeds layer l,
    // the last level the orbital projection; pair layers 0 .. L-2
    const int N = dm.N, A = dm.A, P = dm.P, L = ds_pair_levels(dm), C0 = dm.C0, K0 = dm.K0, K1 = dm.K1, H = dm.H;
    constexpr int F = FT;
    constexpr int FP = (FT + 1) & ~1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double rs2 = 0.70710678118654752440;

    extern __shared__ __align__(16) double fv_sm[];
    // layer 0: Wt [32][FP], b [32]; layers l >= 1: Wt [32][32], b [32]   (output channels >= P zero)
    double* w0 = fv_sm;
    double* w1 = w0 + 32 * FP + 32;
    constexpr int W1_STRIDE = 32 * 32 + 32;
    for (int t = tid; t < 32 * FP; t += blockDim.x) {
        const int co = t / FP, ci = t - co * FP;
        w0[t] = (co < P && ci < F && L > 1) ? fp.Wp[0][ci * P + co] : 0.0;
    }
    for (int t = tid; t < 32; t += blockDim.x) w0[32 * FP + t] = (t < P && L > 1) ? fp.bp[0][t] : 0.0;
    for (int l = 1; l < L - 1; ++l) {
        double* wl = w1 + (l - 1) * W1_STRIDE;
        for (int t = tid; t < 32 * 32; t += blockDim.x) {
            const int co = t >> 5, ci = t & 31;
            wl[t] = (co < P && ci < P) ? fp.Wp[l][ci * P + co] : 0.0;
        }
        for (int t = tid; t < 32; t += blockDim.x) wl[32 * 32 + t] = (t < P) ? fp.bp[l][t] : 0.0;
    }
    double* tr = w1 + (L > 2 ? L - 2 : 0) * W1_STRIDE + warp * (32 * 33);      // per-warp scratch [32 lanes][33]
    double* my = tr + lane * 33;
    __syncthreads();

    for (long long item = (long long)blockIdx.x * FV_WARPS + warp; item < n_items; item += (long long)gridDim.x * FV_WARPS) {
        const int s = (int)(item & 1);
        const long long e = item >> 1;                   // w * N + i
        const int w = (int)(e / N), i = (int)(e - (long long)w * N);
        const int jbeg = s ? dm.n_up : 0, jend = s ? N : dm.n_up;
        const double invn = (jend > jbeg) ? 1.0 / (double)(jend - jbeg) : 0.0;
        const double* x = fp.X + (long long)w * 3 * N;
        double xi3[3] = {x[3 * i], x[3 * i + 1], x[3 * i + 2]}, wi[3];
        ds_wrap(sys.sim, xi3, wi);

        // electron-atom features of electron i (primitive cell): done once, by the spin-up item of the electron
        if (s == 0 && lane < A) {
            double px[3], d[3];
            ds_wrap(sys.prim, xi3, px);
            for (int q = 0; q < 3; ++q) d[q] = px[q] - sys.atoms[3 * lane + q];
            Jet f[FT];
            if (FT == 7) ds_tri_distance<false>(sys.prim, d, f);
            else ds_nu_distance<false>(sys.prim, d, f);
#pragma unroll
            for (int q = 0; q < FT; ++q) fp.A0V[e * K0 + lane * F + q] = f[q].v;
            double* ra = fp.RAE + (e * A + lane) * DS_RAE_STRIDE;
#pragma unroll
            for (int q = 0; q < 4; ++q) { ra[5 * q] = f[q].v; ra[5 * q + 1] = 0.0; ra[5 * q + 2] = 0.0; ra[5 * q + 3] = 0.0; ra[5 * q + 4] = 0.0; }
        }

        double colsum[DS_MAX_LAYERS];                    // lane c: sum over partners of channel c, per level
#pragma unroll
        for (int l = 0; l < DS_MAX_LAYERS; ++l) colsum[l] = 0.0;

        for (int j0 = jbeg; j0 < jend; j0 += 32) {
            const int j = j0 + lane;
            const bool valid = j < jend;
            const double vmask = valid ? 1.0 : 0.0;
            double f[FP];
            {
                double xj3[3] = {0.0, 0.0, 0.0}, wj[3], d[3];
                if (valid) { xj3[0] = x[3 * j]; xj3[1] = x[3 * j + 1]; xj3[2] = x[3 * j + 2]; }
                ds_wrap(sys.sim, xj3, wj);
                for (int q = 0; q < 3; ++q) d[q] = wj[q] - wi[q] + ((valid && j == i) ? 1.0 : 0.0);
                Jet fj[FT];
                if (FT == 7) ds_tri_distance<false>(sys.sim, d, fj);
                else ds_nu_distance<false>(sys.sim, d, fj);
#pragma unroll
                for (int q = 0; q < FT; ++q) f[q] = (valid && j != i) ? fj[q].v : 0.0;
                if (FP > FT) f[FP - 1] = 0.0;
            }
            // level 0: sum over partners of the F features
#pragma unroll
            for (int q = 0; q < FT; ++q) my[q] = f[q];
            __syncwarp();
            double bsum = 0.0;                           // this block's column sum at the current level
            if (lane < F) {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int r = 0; r < 32; r += 4) {
                    s0 += tr[r * 33 + lane]; s1 += tr[(r + 1) * 33 + lane]; s2 += tr[(r + 2) * 33 + lane]; s3 += tr[(r + 3) * 33 + lane];
                }
                bsum = (s0 + s1) + (s2 + s3);
                colsum[0] += bsum;
            }
            __syncwarp();
            if (L <= 1) continue;

            // pair layer 0: t = tanh(f . W0 + b0), raw values to the scratch row
            {
                const double* bb = w0 + 32 * FP;
#pragma unroll 1
                for (int n = 0; n < 32; n += FV_NB) {
                    double z[FV_NB];
#pragma unroll
                    for (int u = 0; u < FV_NB; ++u) z[u] = bb[n + u];
#pragma unroll
                    for (int c = 0; c < FP; c += 2) {
#pragma unroll
                        for (int u = 0; u < FV_NB; ++u) {
                            const double2 w2 = *reinterpret_cast<const double2*>(w0 + (n + u) * FP + c);
                            z[u] = fma(f[c], w2.x, z[u]); z[u] = fma(f[c + 1], w2.y, z[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < FV_NB; ++u) my[n + u] = tanh(z[u]) * vmask;    // lanes past the channel end contribute 0
                }
            }
            double cur[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) cur[c] = my[c];
            __syncwarp();
            {
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int r = 0; r < 32; r += 4) {
                    s0 += tr[r * 33 + lane]; s1 += tr[(r + 1) * 33 + lane]; s2 += tr[(r + 2) * 33 + lane]; s3 += tr[(r + 3) * 33 + lane];
                }
                bsum = (s0 + s1) + (s2 + s3);
            }
            __syncwarp();
            colsum[1] += bsum;

            // pair layers l >= 1 (square, residual): cur <- (cur + tanh(cur . Wl + bl)) / sqrt 2
#pragma unroll 1
            for (int l = 1; l < L - 1; ++l) {
                const double* wl = w1 + (l - 1) * W1_STRIDE;
                const double* bb = wl + 32 * 32;
#pragma unroll 1
                for (int n = 0; n < 32; n += FV_NB) {
                    double z[FV_NB];
#pragma unroll
                    for (int u = 0; u < FV_NB; ++u) z[u] = bb[n + u];
#pragma unroll
                    for (int c = 0; c < 32; c += 2) {
#pragma unroll
                        for (int u = 0; u < FV_NB; ++u) {
                            const double2 w2 = *reinterpret_cast<const double2*>(wl + (n + u) * 32 + c);
                            z[u] = fma(cur[c], w2.x, z[u]); z[u] = fma(cur[c + 1], w2.y, z[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < FV_NB; ++u) my[n + u] = tanh(z[u]) * vmask;
                }
                __syncwarp();
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
                for (int r = 0; r < 32; r += 4) {
                    s0 += tr[r * 33 + lane]; s1 += tr[(r + 1) * 33 + lane]; s2 += tr[(r + 2) * 33 + lane]; s3 += tr[(r + 3) * 33 + lane];
                }
                bsum = (bsum + ((s0 + s1) + (s2 + s3))) * rs2;       // column sum of the residual update
#pragma unroll
                for (int c = 0; c < 32; ++c) cur[c] = (cur[c] + my[c]) * rs2;
                __syncwarp();
                if (l + 1 < DS_MAX_LAYERS) {
                    // colsum[] is indexed by a runtime level here: spelled out to keep it in registers
#pragma unroll
                    for (int q = 2; q < DS_MAX_LAYERS; ++q) if (q == l + 1) colsum[q] += bsum;
                }
            }
        }
        // means of this spin channel -> pair-mean columns of the one-electron operands
        if (lane < F) fp.A0V[e * K0 + C0 + s * F + lane] = colsum[0] * invn;
#pragma unroll
        for (int l = 1; l < DS_MAX_LAYERS; ++l) {
            if (l >= L) break;
            if (lane < P) fp.AV[l][e * K1 + H + s * P + lane] = colsum[l] * invn;
        }
    }
}

}  // namespace

size_t ds_features_smem(const DsDims& d) {
    const int Lv = ds_pair_levels(d);
    size_t n = 8 * (32 * 6) + 3 * d.N + 2 * Lv * 5 * 32;
    for (int l = 0; l < Lv - 1; ++l) n += ((l == 0) ? d.F : d.P) * d.P + d.P;
    return n * sizeof(double);
}

int ds_launch_features(const DsSys& sys, const FeatParams& fp, int Wc, bool jets, cudaStream_t stream) {
    size_t smem = ds_features_smem(sys.d);
    dim3 grid((unsigned)((long long)Wc * sys.d.N));
    const bool tri = sys.d.F == 7;
    if (smem > 48 * 1024) {
        DS_CUDA_CHECK(cudaFuncSetAttribute(features_pair_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DS_CUDA_CHECK(cudaFuncSetAttribute(features_pair_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DS_CUDA_CHECK(cudaFuncSetAttribute(features_pair_kernel<true, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DS_CUDA_CHECK(cudaFuncSetAttribute(features_pair_kernel<false, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    // warps per CTA: a warp owns one partner j at a time, one spin channel after the other; 4 or 8 warps
    // (whole warps per SM sub-partition, 128 registers each), whichever wastes fewer warp-rounds
    int best_w = 8;
    double best_eff = -1.0;
    for (int w = 8; w >= 4; w -= 4) {
        const int rounds = (sys.d.n_up + w - 1) / w + (sys.d.n_dn + w - 1) / w;
        const double eff = (double)sys.d.N / ((double)rounds * w);
        if (eff > best_eff + 1e-9) { best_eff = eff; best_w = w; }
    }
    const int threads = jets ? 32 * best_w : FEAT_THREADS;
    static const bool old_value = getenv("DS_FEAT_OLD") && atoi(getenv("DS_FEAT_OLD")) != 0;
    if (!jets && sys.d.P <= 32 && !old_value) {
        // lane-per-pair value kernel
        const int fpad = fv_pin_pad(sys.d.F);
        const int Lv = ds_pair_levels(sys.d);
        size_t n = (size_t)FV_WARPS * 32 * 33 + 32 * fpad + 32 + (size_t)(Lv > 2 ? Lv - 2 : 0) * (32 * 32 + 32);
        const size_t sm2 = n * sizeof(double);
        const long long n_items = 2LL * Wc * sys.d.N;
        long long blocks = (n_items + FV_WARPS - 1) / FV_WARPS;
        if (blocks > 148 * 8) blocks = 148 * 8;
        if (tri) {
            DS_CUDA_CHECK(cudaFuncSetAttribute(features_value_kernel<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
            features_value_kernel<7><<<(unsigned)blocks, FV_WARPS * 32, sm2, stream>>>(sys, fp, n_items);
        } else {
            DS_CUDA_CHECK(cudaFuncSetAttribute(features_value_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
            features_value_kernel<4><<<(unsigned)blocks, FV_WARPS * 32, sm2, stream>>>(sys, fp, n_items);
        }
        DS_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
    if (jets) {
        if (tri) features_pair_kernel<true, 7><<<grid, threads, smem, stream>>>(sys, fp);
        else features_pair_kernel<true, 4><<<grid, threads, smem, stream>>>(sys, fp);
    } else {
        if (tri) features_pair_kernel<false, 7><<<grid, threads, smem, stream>>>(sys, fp);
        else features_pair_kernel<false, 4><<<grid, threads, smem, stream>>>(sys, fp);
    }
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

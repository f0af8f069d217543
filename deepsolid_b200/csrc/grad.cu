// Reverse sweep of (log|psi|, phase) with respect to the network parameters: the backward half of
// DeepSolid's energy-gradient estimator (train.py:91-142: tangents_dot = mean(Re(clip_diff * conj(d log psi))),
// i.e. a vector-Jacobian product of batch_network with per-walker cotangents a_w = Re clip_diff_w / B on
// log|psi_w| and b_w = Im clip_diff_w / B on the phase).  The weight-gradient contractions themselves are fp64
// DMMA GEMMs (gemm_f64.cu, GEMM_TN / GEMM_PLAIN); this file holds the element-wise and per-pair kernels.
//
//   psi = sum_k D_k,  D_k = prod_s det M_s^k,  d log psi = sum_k w_k sum_s tr(X_s^k dM_s^k),  w_k = D_k / psi
//   dloss = Re[ conj(c) d log psi ],  c = a + i b      =>  cotangent of M_s^k[i,o]:  Gm = conj(c) w_k X_s^k[o,i]
//   (dloss = Re sum Gm dM),   M[i,o] = Y[i,(k,o)] E_i[(k,o)],  Y = h.W_re + i h.W_im,  E = env * exp(i k_o.x_i).
#include "kernels.cuh"

namespace {

// ---------------------------------------------------------------------------
// Orbital layer: cotangents of the raw orbital outputs and the envelope parameter gradients.
// grid = Wc*N (one CTA per electron), threads over p = k*n_s + o.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) orb_grad_kernel(const DsSys sys, const SlaterBufs sb, const GradBufs gb, int npar_max) {
    const DsDims& dm = sys.d;
    const long long e = blockIdx.x;
    const int N = dm.N, A = dm.A, D = dm.D;
    const long long w = e / N;
    const int i = (int)(e % N);
    const int s = (i < dm.n_up) ? 0 : 1;
    const int ne_s = s ? dm.n_dn : dm.n_up;         // electrons of this spin channel (rows of GY)
    const int is_loc = s ? i - dm.n_up : i;
    const int ns = ds_norb(dm, s);                  // orbitals per determinant = matrix columns
    const int blk = dm.full_det ? 0 : s;
    const int nrow = ds_blk_n(dm, blk);
    const int is = dm.full_det ? i : is_loc;        // row of this electron in its matrix
    const int npar = ns * D;
    __shared__ cplx wk_s[64];                       // conj(c) * w_k per determinant
    const bool direct = gb.cot_mats != nullptr;     // cotangent of the matrices given by the caller (eval_mats)
    if (threadIdx.x == 0 && !direct) {
        const double* ld = sb.LOGDET + w * 2 * D * 3;
        double mx = -INFINITY;
        for (int k = 0; k < D; ++k) mx = fmax(mx, ld[k * 3] + ld[(D + k) * 3]);
        cplx tot{0.0, 0.0};
        for (int k = 0; k < D; ++k) {
            const double l = ld[k * 3] + ld[(D + k) * 3];
            const cplx ph = cmul(cplx{ld[k * 3 + 1], ld[k * 3 + 2]}, cplx{ld[(D + k) * 3 + 1], ld[(D + k) * 3 + 2]});
            const double m = exp(l - mx);
            wk_s[k] = cplx{ph.re * m, ph.im * m};
            tot.re += ph.re * m; tot.im += ph.im * m;
        }
        const cplx itot = cinv(tot);
        const cplx cbar{gb.cot_abs[w], -gb.cot_phase[w]};
        for (int k = 0; k < D; ++k) wk_s[k] = cmul(cbar, cmul(wk_s[k], itot));
    }
    __syncthreads();
    const double* x = sb.X + w * 3 * N + 3 * i;
    const double x0 = x[0], x1 = x[1], x2 = x[2];
    const double* rae = sb.RAE + e * A * DS_RAE_STRIDE;
    const int env_type = dm.env_type;
    const double* pi_ = sb.env_pi[s];
    const double* sg_ = sb.env_sigma[s];
    const double* kl = sb.klist[s];
    const cplx* E = reinterpret_cast<const cplx*>(sb.ETAB) + e * 5LL * npar_max;
    const cplx* yv = reinterpret_cast<const cplx*>(sb.YV) + e * (long long)npar_max;
    const cplx* xinv = reinterpret_cast<const cplx*>(sb.XINV[blk]);
    double* gy = gb.GY[s] + (w * ne_s + is_loc) * 2LL * npar;
    for (int p = threadIdx.x; p < npar; p += blockDim.x) {
        const int k = p / ns, o = p - k * ns;
        cplx Gm;
        if (direct) {
            // ds_orbitals layout: per walker [spin0: D n0 n0][spin1: D n1 n1] complex, element (k, i_s, o)
            const long long soff = (s && !dm.full_det) ? (long long)D * dm.n_up * dm.n_up : 0;
            const double* cm = gb.cot_mats + w * gb.cot_mats_stride + 2 * (soff + ((long long)k * nrow + is) * ns + o);
            Gm = cplx{cm[0], -cm[1]};               // dloss = Re(conj(cot) dM)
        } else {
            const cplx X = xinv[((w * D + k) * nrow + o) * (long long)nrow + is];
            Gm = cmul(wk_s[k], X);
        }
        const cplx g = cmul(Gm, E[p]);              // cotangent of Y: dloss = Re(g dY) = g.re dYr - g.im dYi
        gy[2 * p] = g.re;
        gy[2 * p + 1] = -g.im;
        double sn, cs;
        sincos(kl[o * 3] * x0 + kl[o * 3 + 1] * x1 + kl[o * 3 + 2] * x2, &sn, &cs);
        const cplx gyp = cmul(cmul(Gm, yv[p]), cplx{cs, sn});
        const double genv = gyp.re;                  // cotangent of the (real) envelope value
        for (int a = 0; a < A; ++a) {
            const double* ra = rae + a * DS_RAE_STRIDE;
            const double pw = pi_[a * npar + p];
            if (env_type == 0) {
                const double r = ra[0], sig = sg_[a * npar + p];
                const double ex = exp(-fabs(r * sig));
                atomicAdd(gb.g_pi[s] + a * npar + p, genv * ex);
                atomicAdd(gb.g_sigma[s] + a * npar + p, -genv * pw * ex * r * ds_sign(r * sig));
            } else {
                double S[9];
                if (env_type == 1) {
#pragma unroll
                    for (int m = 0; m < 3; ++m) S[m * 3 + m] = sg_[((long long)a * 3 + m) * npar + p];
                } else {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
#pragma unroll
                        for (int m = 0; m < 3; ++m) S[k * 3 + m] = sg_[(((long long)k * 3 + m) * A + a) * npar + p];
                }
                Jet rel[3];
#pragma unroll
                for (int m = 0; m < 3; ++m) rel[m] = jet_const(ra[5 * (1 + m)]);
                double yv[3], rr;
                const Jet en = ds_aniso_env(rel, S, env_type == 1, yv, &rr);
                atomicAdd(gb.g_pi[s] + a * npar + p, genv * en.v);
                // d exp(-r) / d S[k][m] = -exp(-r) y_m rel_k / r
                const double c = -genv * pw * en.v / rr;
                if (env_type == 1) {
#pragma unroll
                    for (int m = 0; m < 3; ++m) atomicAdd(gb.g_sigma[s] + ((long long)a * 3 + m) * npar + p, c * yv[m] * rel[m].v);
                } else {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
#pragma unroll
                        for (int m = 0; m < 3; ++m)
                            atomicAdd(gb.g_sigma[s] + (((long long)k * 3 + m) * A + a) * npar + p, c * yv[m] * rel[k].v);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// One-electron stream: cotangent of the pre-activations, their per-walker sums and the bias gradient.
// grid = Wc, thread = channel.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gz_kernel(DsDims dm, GradBufs gb, int residual) {
    const int w = blockIdx.x, N = dm.N, H = dm.H;
    const double rs2 = 0.70710678118654752440;
    for (int c = threadIdx.x; c < H; c += blockDim.x) {
        double sum = 0.0;
        for (int i = 0; i < N; ++i) {
            const long long o = ((long long)w * N + i) * H + c;
            const double t = gb.T[o];
            double g = gb.GH[o];
            if (residual) g *= rs2;
            const double gz = g * (1.0 - t * t);
            gb.GZ[o] = gz;
            sum += gz;
        }
        gb.GZS[(long long)w * H + c] = sum;
        atomicAdd(gb.g_bias + c, sum);
    }
}

// cotangent of the layer input: own columns of GZ.B_am^T, the share of the spin-channel means, the residual path;
// the pair-mean columns are kept for the pair-stream sweep.  grid = Wc*N.
__global__ void __launch_bounds__(256) hin_kernel(DsDims dm, GradBufs gb, int C, int residual, int want_pm) {
    const long long e = blockIdx.x;
    const int N = dm.N, H = dm.H, P = dm.P;
    const long long w = e / N;
    const int i = (int)(e % N);
    const int s = (i < dm.n_up) ? 0 : 1;
    const double invn = 1.0 / (double)(s ? dm.n_dn : dm.n_up);
    const double rs2 = 0.70710678118654752440;
    const double* ga = gb.GA + e * (long long)gb.lda;
    const double* gg = gb.GG + w * (long long)gb.ldgg + s * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double v = ga[c] + gg[c] * invn;
        if (residual) v += rs2 * gb.GH[e * H + c];
        gb.GHin[e * C + c] = v;
    }
    if (want_pm)
        for (int q = threadIdx.x; q < 2 * P; q += blockDim.x) gb.GPM[e * 2 * P + q] = ga[C + q];
}

// ---------------------------------------------------------------------------
// Pair stream: forward values are recomputed per pair (they are never stored), then the reverse sweep.
// grid = Wc*N (CTA = electron i, the second index of h_two[j,i]); warp = partner j; lane = channel.
// Level 0 = the 4 input features, level l+1 = output of pair layer l.  GPM_l[e_i][s(j)*P + c] / n_s(j) is the
// cotangent reaching h_two^l[j,i] through the pair means of one-electron layer l.
// ---------------------------------------------------------------------------
constexpr int PG_THREADS = 256;

// NPL = number of P x P pair layers (= L - 2): sizes the per-lane weight-gradient accumulators
// FACT = 0: parameter gradients.  FACT = 1 / 2: Kronecker-factor statistics of the tagged pair layers
// (curvature_blocks.py:262-281 through RepeatedDenseBlock, curvature_tags_and_blocks.py:142-156): sums over all pairs of
// x x^T of the layer inputs (1, forward only) or of gz gz^T of the pre-activation cotangents (2).
template <int NPL, int FACT>
__global__ void __launch_bounds__(PG_THREADS, (NPL <= 1 && FACT != 2) ? 2 : 1) pair_grad_kernel(const DsSys sys, const FeatParams fp,
                                                                                                const GradBufs gb, long long n_e) {
    const DsDims& dm = sys.d;
    // L counts the pair-stream LEVELS (dm.L, or dm.L + 1 with use_last_layer: the last level feeds the orbital projection)
    const int N = dm.N, P = dm.P, L = ds_pair_levels(dm), F = dm.F;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const double rs2 = 0.70710678118654752440;

    extern __shared__ double sm[];
    double* sx = sm;                             // [3N] wrapped positions
    double* wsm = sx + 3 * N;                    // per pair layer: W [Pin x P], b [P], W^T [P x Pin]
    double* red = wsm;                           // reduction scratch is carved after the weights (set below)
    int woff[DS_MAX_LAYERS];
    {
        int off = 0;
        for (int l = 0; l < L - 1; ++l) {
            const int pin = (l == 0) ? F : P;
            woff[l] = off;
            for (int t = tid; t < pin * P; t += blockDim.x) {
                const double v = fp.Wp[l][t];
                wsm[off + t] = v;
                const int ci = t / P, co = t - ci * P;
                wsm[off + pin * P + P + co * pin + ci] = v;        // transposed copy
            }
            for (int t = tid; t < P; t += blockDim.x) wsm[off + pin * P + t] = fp.bp[l][t];
            off += 2 * pin * P + P;
        }
        red = wsm + off;                         // [nwarps][P+1] scratch rows
    }
    __syncthreads();

    // per-lane gradient accumulators: column `lane` of every pair-layer weight and bias
    constexpr int NACC = (FACT == 2) ? NPL + 1 : (NPL > 0 ? NPL : 1);
    double gW0[7], gW[NACC][32], gB[DS_MAX_LAYERS];
#pragma unroll
    for (int q = 0; q < 7; ++q) gW0[q] = 0.0;
#pragma unroll
    for (int l = 0; l < NACC; ++l)
#pragma unroll
        for (int c = 0; c < 32; ++c) gW[l][c] = 0.0;
#pragma unroll
    for (int l = 0; l < DS_MAX_LAYERS; ++l) gB[l] = 0.0;

    const double inv_up = 1.0 / dm.n_up, inv_dn = 1.0 / dm.n_dn;
    // persistent CTAs: every CTA walks over many electrons and adds its parameter-gradient partials once at the end
    // (one atomic per parameter and CTA instead of per electron)
    int w_cur = -1;
    for (long long e = blockIdx.x; e < n_e; e += gridDim.x) {
    const int w = (int)(e / N), i = (int)(e % N);
    if (w != w_cur) {                            // CTA-uniform
        __syncthreads();
        const double* x = fp.X + (long long)w * 3 * N;
        for (int t = tid; t < N; t += blockDim.x) {
            double xi[3] = {x[3 * t], x[3 * t + 1], x[3 * t + 2]}, o[3];
            ds_wrap(sys.sim, xi, o);
            sx[3 * t] = o[0]; sx[3 * t + 1] = o[1]; sx[3 * t + 2] = o[2];
        }
        __syncthreads();
        w_cur = w;
    }
    for (int j = warp; j < N; j += nwarps) {
        const int sj = (j < dm.n_up) ? 0 : 1;
        const double invn = sj ? inv_dn : inv_up;
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = sx[3 * j + k] - sx[3 * i + k] + (j == i ? 1.0 : 0.0);
        Jet f[7];
        ds_distance<false>(dm.dist_type, sys.sim, d, f);
        double cur[DS_MAX_LAYERS], tv[DS_MAX_LAYERS];       // level values and tanh outputs of this lane's channel
        cur[0] = 0.0;
        if (j != i && lane < F)
            cur[0] = (lane == 0) ? f[0].v : (lane == 1) ? f[1].v : (lane == 2) ? f[2].v : (lane == 3) ? f[3].v
                   : (lane == 4) ? f[4].v : (lane == 5) ? f[5].v : f[6].v;
#pragma unroll
        for (int l = 0; l < DS_MAX_LAYERS - 1; ++l) {
            if (l >= L - 1) break;
            const int pin = (l == 0) ? F : P;
            const double* W = wsm + woff[l];
            double z = (lane < P) ? W[pin * P + lane] : 0.0;
            for (int c = 0; c < pin; ++c) {
                const double cv = __shfl_sync(0xffffffffu, cur[l], c);
                if (lane < P) z = fma(cv, W[c * P + lane], z);
            }
            const double t = tanh(z);
            tv[l + 1] = t;
            double nv = (l >= 1) ? (cur[l] + t) * rs2 : t;
            cur[l + 1] = (lane < P) ? nv : 0.0;
        }
        if (FACT == 1) {                                    // Gram matrices of the inputs of every pair layer
#pragma unroll
            for (int c = 0; c < 7; ++c) gW0[c] = fma(__shfl_sync(0xffffffffu, cur[0], c), cur[0], gW0[c]);
            gB[0] += cur[0];
#pragma unroll
            for (int l = 1; l <= NPL; ++l) {
                if (l >= L - 1) break;
#pragma unroll
                for (int c = 0; c < 32; ++c)
                    gW[l - 1 < NACC ? l - 1 : 0][c] = fma(__shfl_sync(0xffffffffu, cur[l], c), cur[l], gW[l - 1 < NACC ? l - 1 : 0][c]);
                gB[l] += cur[l];
            }
            continue;
        }
        // reverse: cotangents of the levels
        double g[DS_MAX_LAYERS];
#pragma unroll
        for (int l = 0; l < DS_MAX_LAYERS; ++l) g[l] = 0.0;
        if (lane < P) {
#pragma unroll
            for (int lv = 1; lv < DS_MAX_LAYERS; ++lv)
                if (lv < L && gb.GPMl[lv]) g[lv] = gb.GPMl[lv][e * 2 * P + sj * P + lane] * invn;
        }
#pragma unroll
        for (int lv = DS_MAX_LAYERS - 1; lv >= 1; --lv) {
            if (lv > L - 1) continue;
            const int l = lv - 1;                            // pair layer that produced level lv
            const int pin = (l == 0) ? F : P;
            const bool res = (l >= 1);
            const double gt = res ? g[lv] * rs2 : g[lv];
            const double gz = (lane < P) ? gt * (1.0 - tv[lv] * tv[lv]) : 0.0;
            gB[l] += gz;
            if (FACT == 2) {
                if (l < NACC) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) gW[l < NACC ? l : 0][c] = fma(__shfl_sync(0xffffffffu, gz, c), gz, gW[l < NACC ? l : 0][c]);
                }
            } else if (l == 0) {
#pragma unroll
                for (int c = 0; c < 7; ++c) gW0[c] = fma(__shfl_sync(0xffffffffu, cur[0], c), gz, gW0[c]);      // cur[0] = 0 beyond F
            } else if (l - 1 < NPL) {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const double cv = __shfl_sync(0xffffffffu, cur[l], c);
                    if (c < P) gW[l - 1 < NACC ? l - 1 : 0][c] = fma(cv, gz, gW[l - 1 < NACC ? l - 1 : 0][c]);
                }
            }
            if (l >= 1 && l - 1 < NPL) {
                // cotangent of level l (it has parameters upstream): W.gz over the output channels + residual path
                const double* WT = wsm + woff[l] + pin * P + P;      // [P_out][pin]
                double acc = res ? g[lv] * rs2 : 0.0;
                for (int c = 0; c < P; ++c) {
                    const double gzc = __shfl_sync(0xffffffffu, gz, c);
                    if (lane < pin) acc = fma(WT[c * pin + lane], gzc, acc);
                }
                g[l] += (lane < pin) ? acc : 0.0;
            }
        }
    }

    }   // electrons of this CTA

    // cross-warp reduction, then one atomic per parameter and CTA
    double* row = red + warp * 33;
    auto reduce_and_add = [&](double v, double* dst, bool valid) {
        __syncthreads();
        row[lane] = v;
        __syncthreads();
        if (warp == 0) {
            double s = 0.0;
            for (int q = 0; q < nwarps; ++q) s += red[q * 33 + lane];
            if (valid) atomicAdd(dst, s);
        }
    };
    if (FACT == 1) {
        if (L > 1) {
#pragma unroll
            for (int c = 0; c < 7; ++c)
                if (c < F) reduce_and_add(gW0[c], gb.fact_A[0] + c * 32 + lane, true);
            reduce_and_add(gB[0], gb.fact_As[0] + lane, true);
        }
#pragma unroll
        for (int l = 1; l <= NPL; ++l) {
            if (l >= L - 1) break;
#pragma unroll
            for (int c = 0; c < 32; ++c) reduce_and_add(gW[l - 1 < NACC ? l - 1 : 0][c], gb.fact_A[l] + c * 32 + lane, true);
            reduce_and_add(gB[l], gb.fact_As[l] + lane, true);
        }
        return;
    }
    if (FACT == 2) {
#pragma unroll
        for (int l = 0; l < NACC; ++l) {
            if (l >= L - 1) break;
#pragma unroll
            for (int c = 0; c < 32; ++c) reduce_and_add(gW[l][c], gb.fact_G[l] + c * 32 + lane, true);
        }
        return;
    }
    if (L > 1) {
#pragma unroll
        for (int c = 0; c < 7; ++c)
            if (c < F) reduce_and_add(gW0[c], gb.g_Wp[0] + c * P + lane, lane < P);
        reduce_and_add(gB[0], gb.g_bp[0] + lane, lane < P);
    }
#pragma unroll
    for (int l = 1; l <= NPL; ++l) {
        if (l >= L - 1) break;
#pragma unroll
        for (int c = 0; c < 32; ++c)
            if (c < P) reduce_and_add(gW[l - 1][c], gb.g_Wp[l] + c * P + lane, lane < P);
        reduce_and_add(gB[l], gb.g_bp[l] + lane, lane < P);
    }
}

// dst[c] += sum_r src[r, c]
__global__ void colsum_add_kernel(const double* __restrict__ src, int lds, int rows, int cols, double* __restrict__ dst) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += src[(long long)r * lds + c];
    dst[c] += s;
}

__global__ void copy2d_kernel(const double* __restrict__ src, int lds, double* __restrict__ dst, int ldd, int rows, int cols) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)rows * cols) return;
    const int r = (int)(t / cols), c = (int)(t - (long long)r * cols);
    dst[(long long)r * ldd + c] = src[(long long)r * lds + c];
}

// src [rows][2 np] with columns (re, im) interleaved  ->  dst [rows][2 np] = (re block | im block)   (network.py:543-545)
__global__ void deinterleave_kernel(const double* __restrict__ src, double* __restrict__ dst, int rows, int np) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)rows * 2 * np) return;
    const int r = (int)(t / (2 * np)), c = (int)(t - (long long)r * 2 * np);
    const int p = c >> 1, im = c & 1;
    dst[(long long)r * 2 * np + im * np + p] = src[t];
}

__global__ void layer_input_rows_kernel(const double* __restrict__ Ain, int K, const double* __restrict__ GINV, int C, int N,
                                        long long rows, double* __restrict__ out) {
    const int ldx = 2 * C + K + 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * ldx) return;
    const long long e = t / ldx;
    const int c = (int)(t - e * ldx);
    double v;
    if (c < C) v = Ain[e * K + c];
    else if (c < 3 * C) v = GINV[(e / N) * 2 * C + (c - C)];
    else if (c < 2 * C + K) v = Ain[e * K + (c - 2 * C)];
    else v = (c == 2 * C + K) ? 1.0 : 0.0;
    out[t] = v;
}

__global__ void spin_rows_kernel(const double* __restrict__ src, int lds, int cols, int N, int off_s, int ns, long long rows,
                                 double* __restrict__ out) {
    const int ldx = cols + 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rows * ldx) return;
    const long long r = t / ldx;
    const int c = (int)(t - r * ldx);
    const long long w = r / ns;
    const long long e = w * N + off_s + (r - w * ns);
    out[t] = (c < cols) ? src[e * lds + c] : (c == cols ? 1.0 : 0.0);
}

__global__ void pair_fact_pack_kernel(const double* __restrict__ A32, const double* __restrict__ As, double count, int pin,
                                      double* __restrict__ out) {
    const int n = pin + 1;
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
        const int r = t / n, c = t - r * n;
        double v;
        if (r < pin && c < pin) v = A32[r * 32 + c];
        else if (r < pin) v = As[r];
        else if (c < pin) v = As[c];
        else v = count;
        out[t] = v;
    }
}

__global__ void deinterleave2_kernel(const double* __restrict__ src, double* __restrict__ dst, int np) {
    const long long n2 = 2LL * np;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n2 * n2) return;
    const int r = (int)(t / n2), c = (int)(t - (long long)r * n2);
    const long long qr = (long long)(r & 1) * np + (r >> 1), qc = (long long)(c & 1) * np + (c >> 1);
    dst[qr * n2 + qc] = src[t];
}

}  // namespace

int ds_launch_layer_input_rows(const double* Ain, int K, const double* GINV, int C, int N, long long rows, double* out,
                               cudaStream_t stream) {
    const long long n = rows * (2 * C + K + 2);
    if (n <= 0) return 0;
    layer_input_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(Ain, K, GINV, C, N, rows, out);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_spin_rows(const double* src, int lds, int cols, int N, int off_s, int ns, int Wc, double* out, cudaStream_t stream) {
    const long long rows = (long long)Wc * ns, n = rows * (cols + 2);
    if (n <= 0) return 0;
    spin_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, lds, cols, N, off_s, ns, rows, out);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_pair_fact_pack(const double* A32, const double* As, double count, int pin, double* out, cudaStream_t stream) {
    pair_fact_pack_kernel<<<1, 256, 0, stream>>>(A32, As, count, pin, out);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_deinterleave2(const double* src, double* dst, int np, cudaStream_t stream) {
    const long long n = 4LL * np * np;
    if (n <= 0) return 0;
    deinterleave2_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, dst, np);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_orb_grad(const DsSys& sys, const SlaterBufs& sb, const GradBufs& gb, int Wc, int npar_max, cudaStream_t stream) {
    DS_REQUIRE(sys.d.D <= 64, "orbital gradient kernel supports at most 64 determinants");
    orb_grad_kernel<<<(unsigned)((long long)Wc * sys.d.N), 256, 0, stream>>>(sys, sb, gb, npar_max);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_gz(const DsDims& dm, const GradBufs& gb, int Wc, bool residual, cudaStream_t stream) {
    gz_kernel<<<Wc, 256, 0, stream>>>(dm, gb, residual ? 1 : 0);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_hin(const DsDims& dm, const GradBufs& gb, int Wc, int C, int K, bool residual, bool want_pm, cudaStream_t stream) {
    (void)K;
    hin_kernel<<<(unsigned)((long long)Wc * dm.N), 256, 0, stream>>>(dm, gb, C, residual ? 1 : 0, want_pm ? 1 : 0);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_pair_grad(const DsSys& sys, const FeatParams& fp, const GradBufs& gb, int Wc, cudaStream_t stream, int fact) {
    const DsDims& d = sys.d;
    const int Lv = ds_pair_levels(d);
    if (Lv < 2) return 0;
    DS_REQUIRE(d.P <= 32, "pair-stream gradient kernel needs hidden_two <= 32");
    size_t n = 3 * d.N;
    for (int l = 0; l < Lv - 1; ++l) n += 2 * ((l == 0) ? d.F : d.P) * d.P + d.P;
    n += (PG_THREADS / 32) * 33;
    const size_t smem = n * sizeof(double);
    const long long n_e = (long long)Wc * d.N;
    const int npl = Lv - 2;
    const int grid = (int)(n_e < 4 * 148 ? n_e : 4 * 148);
    DS_REQUIRE(fact == 0 || npl <= 2, "pair-layer factor statistics support at most 3 pair layers");
#define DS_PG_LAUNCH2(NPL_, FACT_)                                                                                      \
    do {                                                                                                                \
        if (smem > 48 * 1024)                                                                                           \
            DS_CUDA_CHECK(cudaFuncSetAttribute(pair_grad_kernel<NPL_, FACT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        pair_grad_kernel<NPL_, FACT_><<<grid, PG_THREADS, smem, stream>>>(sys, fp, gb, n_e);                              \
    } while (0)
#define DS_PG_LAUNCH(NPL_)                                                                                              \
    do {                                                                                                                \
        if (fact == 0) DS_PG_LAUNCH2(NPL_, 0);                                                                          \
        else if (fact == 1) DS_PG_LAUNCH2(NPL_, 1);                                                                     \
        else DS_PG_LAUNCH2(NPL_, 2);                                                                                    \
    } while (0)
    if (npl <= 0) DS_PG_LAUNCH(0);
    else if (npl == 1) DS_PG_LAUNCH(1);
    else DS_PG_LAUNCH(2);
#undef DS_PG_LAUNCH
#undef DS_PG_LAUNCH2
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// dst[g][c] = sum_{r < rows_per_group} src[(g * rows_per_group + r) * lds + c]
__global__ void __launch_bounds__(256) group_rowsum_kernel(const double* __restrict__ src, int lds, int rows_per_group, int cols,
                                                           double* __restrict__ dst, int ldd) {
    const long long g = blockIdx.x;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < rows_per_group; ++r) acc += src[(g * rows_per_group + r) * (long long)lds + c];
        dst[g * (long long)ldd + c] = acc;
    }
}

int ds_launch_group_rowsum(const double* src, int lds, int rows_per_group, int n_groups, int cols, double* dst, cudaStream_t stream) {
    if (n_groups <= 0 || cols <= 0) return 0;
    group_rowsum_kernel<<<(unsigned)n_groups, 256, 0, stream>>>(src, lds, rows_per_group, cols, dst, cols);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_colsum_add(const double* src, int lds, int rows, int cols, double* dst, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return 0;
    colsum_add_kernel<<<(cols + 127) / 128, 128, 0, stream>>>(src, lds, rows, cols, dst);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_copy2d(const double* src, int lds, double* dst, int ldd, int rows, int cols, cudaStream_t stream) {
    const long long tot = (long long)rows * cols;
    if (tot <= 0) return 0;
    copy2d_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(src, lds, dst, ldd, rows, cols);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_deinterleave(const double* src, double* dst, int rows, int np, cudaStream_t stream) {
    const long long tot = (long long)rows * 2 * np;
    if (tot <= 0) return 0;
    deinterleave_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(src, dst, rows, np);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

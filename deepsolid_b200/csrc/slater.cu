// Orbitals -> Slater matrices -> log-determinants and the determinant part of the
// Laplacian (network.py:335-337, 375-427, 449-458, 536-557; hamiltonian.py:45-70).
//
//   M[k][i,o]   = (h_i . W_orb)[k,o] * E_i[k,o],   E_i[k,o] = env_i[k,o] * exp(i k_o.x_i)
//   psi         = sum_k prod_s det M_s[k]
//   lap psi/psi = sum_k w_k { sum_s [ tr(X lapM) - sum_d tr((X dM_d)^2) ] + sum_d (sum_s tr(X dM_d))^2 },
//                 X = M^-1, w_k = D_k / psi
#include "kernels.cuh"
#include <stdlib.h>

namespace {

// ---------------------------------------------------------------------------
// E table: envelope * Bloch phase and its own-electron jet.  grid = Wc*N, threads over p.
// ---------------------------------------------------------------------------
template <bool JETS>
__global__ void __launch_bounds__(256) etab_kernel(const DsSys sys, const SlaterBufs sb, int npar_max) {
    const DsDims& dm = sys.d;
    const long long e = blockIdx.x;
    const int N = dm.N, A = dm.A, D = dm.D;
    const int w = (int)(e / N), i = (int)(e % N);
    const int s = (i < dm.n_up) ? 0 : 1;
    const int ns = ds_norb(dm, s);              // orbitals of this spin channel
    const int npar = ns * D;
    const double* x = sb.X + (long long)w * 3 * N + 3 * i;
    const double x0 = x[0], x1 = x[1], x2 = x[2];
    const double* rae = sb.RAE + e * A * DS_RAE_STRIDE;
    const double* pi_ = sb.env_pi[s];
    const int env_type = dm.env_type;
    const double* sg_ = sb.env_sigma[s];
    const double* kl = sb.klist[s];
    double* out = sb.ETAB + e * 5LL * npar_max * 2;
    for (int p = threadIdx.x; p < npar; p += blockDim.x) {
        double ev = 0.0, eg0 = 0.0, eg1 = 0.0, eg2 = 0.0, el = 0.0;
        for (int a = 0; a < A; ++a) {
            const double* ra = rae + a * DS_RAE_STRIDE;
            const double pw = pi_[a * npar + p];
            if (env_type == 0) {
                const double r = ra[0], sig = sg_[a * npar + p];
                const double asig = fabs(sig);
                const double ex = exp(-fabs(r * sig)) * pw;
                ev += ex;
                if (JETS) {
                    const double g0 = ra[1], g1 = ra[2], g2 = ra[3], l = ra[4];
                    const double sr = ds_sign(r);      // r >= 0; keeps d|r sigma|/dr = sign(r sigma) sigma exact
                    const double d1 = -asig * sr * ex;
                    eg0 += d1 * g0; eg1 += d1 * g1; eg2 += d1 * g2;
                    el += d1 * l + sig * sig * ex * (g0 * g0 + g1 * g1 + g2 * g2);
                }
            } else {
                // diagonal: sigma [A][3][npar]; full: sigma [3][3][A][npar]   (network.py:146-152)
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
                for (int m = 0; m < 3; ++m) {
                    const double* rj = ra + 5 * (1 + m);
                    rel[m] = Jet{rj[0], JETS ? rj[1] : 0.0, JETS ? rj[2] : 0.0, JETS ? rj[3] : 0.0, JETS ? rj[4] : 0.0};
                }
                double yv[3], rr;
                const Jet en = ds_aniso_env(rel, S, env_type == 1, yv, &rr);
                ev += pw * en.v;
                if (JETS) { eg0 += pw * en.g0; eg1 += pw * en.g1; eg2 += pw * en.g2; el += pw * en.l; }
            }
        }
        const int o = p % ns;
        const double k0 = kl[o * 3], k1 = kl[o * 3 + 1], k2 = kl[o * 3 + 2];
        double sn, cs;
        sincos(k0 * x0 + k1 * x1 + k2 * x2, &sn, &cs);
        // value
        out[2 * p] = ev * cs;
        out[2 * p + 1] = ev * sn;
        if (JETS) {
            // d/dc (env * ph) = env_g * ph + env * (i k_c) ph ;  i*ph = (-sn, cs)
            const double kk[3] = {k0, k1, k2};
            const double eg[3] = {eg0, eg1, eg2};
            double lre = el * cs, lim = el * sn;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double* oc = out + (long long)(1 + c) * npar_max * 2;
                oc[2 * p] = eg[c] * cs - ev * kk[c] * sn;
                oc[2 * p + 1] = eg[c] * sn + ev * kk[c] * cs;
                lre += 2.0 * eg[c] * kk[c] * (-sn);
                lim += 2.0 * eg[c] * kk[c] * cs;
            }
            const double k2n = k0 * k0 + k1 * k1 + k2 * k2;
            lre -= ev * k2n * cs;
            lim -= ev * k2n * sn;
            double* ol = out + 4LL * npar_max * 2;
            ol[2 * p] = lre;
            ol[2 * p + 1] = lim;
        }
    }
}

// ---------------------------------------------------------------------------
// Assemble the Slater matrices from the raw orbital-layer outputs.  grid = Wc*N.
// ---------------------------------------------------------------------------
template <bool JETS>
__global__ void __launch_bounds__(256) assemble_kernel(const DsSys sys, const SlaterBufs sb, int npar_max) {
    const DsDims& dm = sys.d;
    const long long e = blockIdx.x;
    const int N = dm.N, D = dm.D, NDp = dm.NDp;
    const long long w = e / N;
    const int i = (int)(e % N);
    const int s = (i < dm.n_up) ? 0 : 1;
    const int ns = ds_norb(dm, s);                              // orbitals = columns of the matrix
    const int blk = dm.full_det ? 0 : s;
    const int nrow = ds_blk_n(dm, blk);                         // rows of the matrix
    const int is = dm.full_det ? i : (s ? i - dm.n_up : i);     // row of this electron
    const int npar = ns * D;
    const cplx* E = reinterpret_cast<const cplx*>(sb.ETAB) + e * 5LL * npar_max;
    const cplx* yv = reinterpret_cast<const cplx*>(sb.YV) + e * (long long)npar_max;
    const cplx* yl = reinterpret_cast<const cplx*>(sb.YL) + e * (long long)npar_max;
    const cplx* yo = reinterpret_cast<const cplx*>(sb.YOWN) + e * 3LL * npar_max;
    cplx* mat = reinterpret_cast<cplx*>(sb.MAT[blk]);
    cplx* lapm = reinterpret_cast<cplx*>(sb.LAPM[blk]);
    cplx* da = reinterpret_cast<cplx*>(sb.DA[blk]);
    for (int p = threadIdx.x; p < npar; p += blockDim.x) {
        const int k = p / ns, o = p - k * ns;
        const cplx O = yv[p];
        const cplx E0 = E[p];
        const long long mi = ((w * D + k) * nrow + is) * ns + o;
        mat[mi] = cmul(O, E0);
        if (JETS) {
            cplx l = cmul(yl[p], E0);
            cfma(l, O, E[4LL * npar_max + p]);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const cplx Ec = E[(long long)(1 + c) * npar_max + p];
                const cplx oj = yo[(long long)c * npar_max + p];
                l.re += 2.0 * (oj.re * Ec.re - oj.im * Ec.im);
                l.im += 2.0 * (oj.re * Ec.im + oj.im * Ec.re);
                const long long di = (((w * D + k) * NDp + 3 * i + c) * nrow + is) * (long long)ns + o;
                cplx cur = da[di];
                cfma(cur, O, Ec);
                da[di] = cur;
            }
            lapm[mi] = l;
        }
    }
}

// ---------------------------------------------------------------------------
// use_last_layer (network.py:528-533): the orbital projection also takes the spin-channel means of the last layer.
// Their product with the mean rows of the weights, GO[w, row, (p, re|im)], is the same for every electron of a walker
// and is added here to the raw outputs the per-electron GEMMs produced: value rows YV (row NDp of GO, or row 0 when
// only values exist), Laplacian rows YL (row NDp+1), the own-gradient rows YOWN (rows 3i+c), and -- times the
// envelope-phase factor E -- every Jacobian row d of the derivative matrices DA.  grid = Wc*N, before assemble.
// ---------------------------------------------------------------------------
template <bool JETS>
__global__ void __launch_bounds__(256) orb_mean_addend_kernel(const DsSys sys, const SlaterBufs sb, int npar_max,
                                                              const double* __restrict__ GO0, const double* __restrict__ GO1,
                                                              int ldgo) {
    const DsDims& dm = sys.d;
    const long long e = blockIdx.x;
    const int N = dm.N, D = dm.D, NDp = dm.NDp, ND = dm.ND;
    const long long w = e / N;
    const int i = (int)(e % N);
    const int s = (i < dm.n_up) ? 0 : 1;
    const int ns = ds_norb(dm, s);
    const int blk = dm.full_det ? 0 : s;
    const int nrow = ds_blk_n(dm, blk);
    const int is = dm.full_det ? i : (s ? i - dm.n_up : i);
    const int npar = ns * D;
    const int rows = JETS ? dm.NDg : 1, vrow = JETS ? NDp : 0;
    const cplx* go = reinterpret_cast<const cplx*>((s ? GO1 : GO0) + w * (long long)rows * ldgo);
    const long long ldc = ldgo / 2;
    const cplx* E = reinterpret_cast<const cplx*>(sb.ETAB) + e * 5LL * npar_max;
    cplx* yv = reinterpret_cast<cplx*>(const_cast<double*>(sb.YV)) + e * (long long)npar_max;
    for (int p = threadIdx.x; p < npar; p += blockDim.x) {
        const cplx g = go[vrow * ldc + p];
        yv[p].re += g.re; yv[p].im += g.im;
    }
    if (!JETS) return;
    cplx* yl = reinterpret_cast<cplx*>(const_cast<double*>(sb.YL)) + e * (long long)npar_max;
    cplx* yo = reinterpret_cast<cplx*>(const_cast<double*>(sb.YOWN)) + e * 3LL * npar_max;
    for (int p = threadIdx.x; p < npar; p += blockDim.x) {
        const cplx gl = go[(NDp + 1) * ldc + p];
        yl[p].re += gl.re; yl[p].im += gl.im;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const cplx gc = go[(3 * i + c) * ldc + p];
            yo[(long long)c * npar_max + p].re += gc.re; yo[(long long)c * npar_max + p].im += gc.im;
        }
    }
    cplx* da = reinterpret_cast<cplx*>(sb.DA[blk]);
    for (int idx = threadIdx.x; idx < ND * npar; idx += blockDim.x) {
        const int d = idx / npar, p = idx - d * npar;
        const int k = p / ns, o = p - k * ns;
        const long long di = (((w * D + k) * NDp + d) * nrow + is) * (long long)ns + o;
        cplx cur = da[di];
        cfma(cur, go[d * ldc + p], E[p]);
        da[di] = cur;
    }
}

// ---------------------------------------------------------------------------
// Determinant kernel: one CTA per (walker, spin, determinant).
//   LAP=false: LU with partial pivoting -> log|det|, phase.
//   LAP=true : in-place Gauss-Jordan inverse, then tr(X lapM), tr(X dM_d), tr((X dM_d)^2).
// ---------------------------------------------------------------------------
constexpr int DET_THREADS = 256;

__device__ __forceinline__ double cabs1(cplx a) { return fabs(a.re) + fabs(a.im); }

template <bool LAP>
__global__ void __launch_bounds__(DET_THREADS) det_kernel(const DsSys sys, const SlaterBufs sb, int G) {
    const DsDims& dm = sys.d;
    const int D = dm.D, NDp = dm.NDp, ND = dm.ND;
    const int k = blockIdx.x % D;
    const int nblk = ds_nblk(dm);
    const int s = (blockIdx.x / D) % nblk;          // matrix block: spin channel, or the single full determinant
    const long long w = blockIdx.x / (nblk * D);
    const int n = ds_blk_n(dm, s);
    if (dm.full_det) {                               // the second slot of the per-spin results stays neutral
        const long long slot1 = (w * 2 + 1) * D + k;
        if (threadIdx.x == 0) {
            double* ld1 = sb.LOGDET + slot1 * 3;
            ld1[0] = 0.0; ld1[1] = 1.0; ld1[2] = 0.0;
            if (sb.TRSQ) { sb.TRSQ[slot1 * 2] = 0.0; sb.TRSQ[slot1 * 2 + 1] = 0.0; }
            if (sb.TRLAP) { sb.TRLAP[slot1 * 2] = 0.0; sb.TRLAP[slot1 * 2 + 1] = 0.0; }
        }
        if (sb.TAU)
            for (int t = threadIdx.x; t < 2 * dm.NDp; t += blockDim.x) sb.TAU[slot1 * 2 * dm.NDp + t] = 0.0;
    }
    const int np = ((n + 2) / 3) * 3;               // padded to the 3x3 register block
    const int tid = threadIdx.x;

    extern __shared__ __align__(16) double smraw[];
    cplx* Xs = reinterpret_cast<cplx*>(smraw);      // [np][np] rows = orbital (after inversion), cols = electron
    cplx* colk = Xs + np * np;                      // [np]
    cplx* dAs = colk + np;                          // [G][np][np]  rows = electron, cols = orbital
    cplx* Ys = dAs + (LAP ? G * np * np : 0);       // [G][np][np]
    __shared__ int piv[128];
    __shared__ int s_p;
    __shared__ double s_red[2 * 16 + 4];

    const cplx* mat = reinterpret_cast<const cplx*>(sb.MAT[s]) + (w * D + k) * (long long)n * n;
    for (int t = tid; t < np * np; t += DET_THREADS) {
        int r = t / np, c = t - r * np;
        Xs[t] = (r < n && c < n) ? mat[r * n + c] : cplx{0.0, 0.0};
    }
    __syncthreads();

    double logabs = 0.0;
    cplx phase{1.0, 0.0};
    for (int kk = 0; kk < n; ++kk) {
        // pivot search in column kk (warp 0)
        if (tid < 32) {
            double best = -1.0;
            int bi = kk;
            for (int r = kk + tid; r < n; r += 32) {
                double v = cabs1(Xs[r * np + kk]);
                if (v > best) { best = v; bi = r; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, best, off);
                int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (tid == 0) { s_p = bi; piv[kk] = bi; }
        }
        __syncthreads();
        const int p = s_p;
        if (p != kk) {
            for (int c = tid; c < n; c += DET_THREADS) {
                cplx a = Xs[kk * np + c];
                Xs[kk * np + c] = Xs[p * np + c];
                Xs[p * np + c] = a;
            }
            phase = cplx{-phase.re, -phase.im};
        }
        __syncthreads();
        const cplx pv = Xs[kk * np + kk];
        {
            double a = hypot(pv.re, pv.im);
            logabs += log(a);
            phase = cmul(phase, cplx{pv.re / a, pv.im / a});
        }
        const cplx ipv = cinv(pv);
        // save column kk, then scale the pivot row
        for (int r = tid; r < n; r += DET_THREADS) colk[r] = Xs[r * np + kk];
        __syncthreads();
        if (LAP) {
            for (int c = tid; c < n; c += DET_THREADS) {
                cplx a = (c == kk) ? cplx{1.0, 0.0} : Xs[kk * np + c];
                Xs[kk * np + c] = cmul(a, ipv);
            }
            __syncthreads();
            for (int t = tid; t < n * n; t += DET_THREADS) {
                int r = t / n, c = t - r * n;
                if (r == kk) continue;
                cplx f = colk[r];
                cplx a = (c == kk) ? cplx{0.0, 0.0} : Xs[r * np + c];
                cplx b = Xs[kk * np + c];
                a.re -= f.re * b.re - f.im * b.im;
                a.im -= f.re * b.im + f.im * b.re;
                Xs[r * np + c] = a;
            }
        } else {
            const int m = n - kk - 1;
            for (int t = tid; t < m * m; t += DET_THREADS) {
                int r = kk + 1 + t / m, c = kk + 1 + t % m;
                cplx f = cmul(colk[r], ipv);
                cplx b = Xs[kk * np + c];
                cplx a = Xs[r * np + c];
                a.re -= f.re * b.re - f.im * b.im;
                a.im -= f.re * b.im + f.im * b.re;
                Xs[r * np + c] = a;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        double* ld = sb.LOGDET + ((w * 2 + s) * D + k) * 3;
        ld[0] = logabs; ld[1] = phase.re; ld[2] = phase.im;
    }
    if (!LAP) return;

    // undo the row pivoting: swap columns in reverse order
    for (int kk = n - 1; kk >= 0; --kk) {
        const int p = piv[kk];
        if (p != kk) {
            for (int r = tid; r < n; r += DET_THREADS) {
                cplx a = Xs[r * np + kk];
                Xs[r * np + kk] = Xs[r * np + p];
                Xs[r * np + p] = a;
            }
        }
        __syncthreads();
    }
    // now Xs[o][i] = (M^-1)[o,i]
    if (sb.XINV[s] != nullptr) {                     // parameter-gradient path: only the inverse is wanted
        cplx* xo = reinterpret_cast<cplx*>(sb.XINV[s]) + (w * D + k) * (long long)n * n;
        for (int t = tid; t < n * n; t += DET_THREADS) {
            int o = t / n, i = t - o * n;
            xo[t] = Xs[o * np + i];
        }
        return;
    }

    const int lane = tid & 31, warp = tid >> 5;
    auto block_sum2 = [&](double a, double b, double& oa, double& ob) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, off);
            b += __shfl_xor_sync(0xffffffffu, b, off);
        }
        __syncthreads();
        if (lane == 0) { s_red[2 * warp] = a; s_red[2 * warp + 1] = b; }
        __syncthreads();
        oa = 0.0; ob = 0.0;
        for (int q = 0; q < DET_THREADS / 32; ++q) { oa += s_red[2 * q]; ob += s_red[2 * q + 1]; }
    };

    // tr(X lapM) = sum_{i,o} X[o,i] lapM[i,o]
    {
        const cplx* lm = reinterpret_cast<const cplx*>(sb.LAPM[s]) + (w * D + k) * (long long)n * n;
        cplx acc{0.0, 0.0};
        for (int t = tid; t < n * n; t += DET_THREADS) {
            int i = t / n, o = t - i * n;
            cfma(acc, Xs[o * np + i], lm[t]);
        }
        double ra, rb;
        block_sum2(acc.re, acc.im, ra, rb);
        if (tid == 0) {
            double* tl = sb.TRLAP + ((w * 2 + s) * D + k) * 2;
            tl[0] = ra; tl[1] = rb;
        }
    }

    const cplx* da = reinterpret_cast<const cplx*>(sb.DA[s]) + (w * D + k) * (long long)NDp * n * n;
    double* tau_out = sb.TAU + ((w * 2 + s) * D + k) * (long long)NDp * 2;
    const int nb = np / 3;
    cplx trsq{0.0, 0.0};                             // accumulated by every thread redundantly after block sums
    for (int d0 = 0; d0 < ND; d0 += G) {
        const int g_cnt = min(G, ND - d0);
        __syncthreads();
        // stage dM_d (zero padded)
        for (int t = tid; t < g_cnt * np * np; t += DET_THREADS) {
            int g = t / (np * np), rem = t - g * np * np;
            int i = rem / np, o = rem - i * np;
            dAs[t] = (i < n && o < n) ? da[((long long)(d0 + g) * n + i) * n + o] : cplx{0.0, 0.0};
        }
        __syncthreads();
        // Y_g = X . dM_g with 3x3 register blocks
        for (int t = tid; t < g_cnt * nb * nb; t += DET_THREADS) {
            int g = t / (nb * nb), rem = t - g * nb * nb;
            int rb = (rem / nb) * 3, cb = (rem % nb) * 3;
            cplx y[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) y[a][b] = cplx{0.0, 0.0};
            const cplx* xa = Xs + rb * np;
            const cplx* db = dAs + g * np * np + cb;
            for (int i = 0; i < n; ++i) {
                cplx xv[3], dv[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) xv[a] = xa[a * np + i];
#pragma unroll
                for (int b = 0; b < 3; ++b) dv[b] = db[i * np + b];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) cfma(y[a][b], xv[a], dv[b]);
            }
            cplx* yo = Ys + g * np * np;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) yo[(rb + a) * np + cb + b] = y[a][b];
        }
        __syncthreads();
        // traces: warp `g` round-robin over directions
        for (int g = warp; g < g_cnt; g += DET_THREADS / 32) {
            const cplx* y = Ys + g * np * np;
            cplx tr{0.0, 0.0}, sq{0.0, 0.0};
            for (int t = lane; t < n * n; t += 32) {
                int a = t / n, b = t - a * n;
                cplx yab = y[a * np + b];
                if (a == b) { tr.re += yab.re; tr.im += yab.im; }
                cfma(sq, yab, y[b * np + a]);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                tr.re += __shfl_xor_sync(0xffffffffu, tr.re, off);
                tr.im += __shfl_xor_sync(0xffffffffu, tr.im, off);
                sq.re += __shfl_xor_sync(0xffffffffu, sq.re, off);
                sq.im += __shfl_xor_sync(0xffffffffu, sq.im, off);
            }
            if (lane == 0) {
                tau_out[2 * (d0 + g)] = tr.re;
                tau_out[2 * (d0 + g) + 1] = tr.im;
                trsq.re += sq.re; trsq.im += sq.im;      // lane 0 of each warp holds a partial
            }
        }
    }
    {
        double ra, rb;
        block_sum2(lane == 0 ? trsq.re : 0.0, lane == 0 ? trsq.im : 0.0, ra, rb);
        if (tid == 0) {
            double* ts = sb.TRSQ + ((w * 2 + s) * D + k) * 2;
            ts[0] = ra; ts[1] = rb;
        }
    }
}

// ---------------------------------------------------------------------------
// Determinant kernel of the Laplacian sweep, second version: one CTA per (walker, spin, determinant).
//   * in-place Gauss-Jordan inverse X = M^-1 in shared memory (as det_kernel<true>)
//   * directions are processed in groups of G: the G derivative matrices dM_d are staged with cp.async
//     (no register round trip), every thread owns one 3x3 block position (rb, cb) of TWO directions
//     (X operands loaded once for both: 9 LDS.128 per 72 DFMA), the products Y_d = X dM_d are written
//     IN PLACE over the staged operands, and  sum_d tr(Y_d^2) = sum_{a,b} Y[a,b] Y[b,a]  is accumulated
//     per thread from its register block and the transposed block read back from shared memory; only
//     tr(Y_d) is reduced per direction (diagonal blocks).
// blockDim.x >= (G/2) * nb^2 work items (nb = np/3), a multiple of 32, at most 512.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16_s(void* smem_dst, const void* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}

__global__ void __launch_bounds__(512, 1) det_lap_kernel(const DsSys sys, const SlaterBufs sb, int G) {
    const DsDims& dm = sys.d;
    const int D = dm.D, NDp = dm.NDp, ND = dm.ND;
    const int k = blockIdx.x % D;
    const int nblk = ds_nblk(dm);
    const int s = (blockIdx.x / D) % nblk;          // matrix block: spin channel, or the single full determinant
    const long long w = blockIdx.x / (nblk * D);
    const int n = ds_blk_n(dm, s);
    if (dm.full_det) {                               // the second slot of the per-spin results stays neutral
        const long long slot1 = (w * 2 + 1) * D + k;
        if (threadIdx.x == 0) {
            double* ld1 = sb.LOGDET + slot1 * 3;
            ld1[0] = 0.0; ld1[1] = 1.0; ld1[2] = 0.0;
            if (sb.TRSQ) { sb.TRSQ[slot1 * 2] = 0.0; sb.TRSQ[slot1 * 2 + 1] = 0.0; }
            if (sb.TRLAP) { sb.TRLAP[slot1 * 2] = 0.0; sb.TRLAP[slot1 * 2 + 1] = 0.0; }
        }
        if (sb.TAU)
            for (int t = threadIdx.x; t < 2 * dm.NDp; t += blockDim.x) sb.TAU[slot1 * 2 * dm.NDp + t] = 0.0;
    }
    const int np = ((n + 2) / 3) * 3;               // padded to the 3x3 register block
    const int nb = np / 3;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;

    extern __shared__ __align__(16) double smraw[];
    cplx* Xs = reinterpret_cast<cplx*>(smraw);      // [np][np] rows = orbital (after inversion), cols = electron
    cplx* colk = Xs + np * np;                      // [np]
    cplx* buf = colk + np;                          // [G][np][np]  dM_d (rows = electron, cols = orbital), then Y_d
    cplx* s_tau = buf + G * np * np;                // [G][nb] partial traces of the diagonal blocks
    __shared__ int piv[128];
    __shared__ int s_p;
    __shared__ double s_red[2 * 16 + 4];

    const cplx* mat = reinterpret_cast<const cplx*>(sb.MAT[s]) + (w * D + k) * (long long)n * n;
    for (int t = tid; t < np * np; t += nthr) {
        int r = t / np, c = t - r * np;
        Xs[t] = (r < n && c < n) ? mat[r * n + c] : cplx{0.0, 0.0};
    }
    if (np != n)                                     // padding of the staged operands stays zero for the whole kernel
        for (int t = tid; t < G * np * np; t += nthr) buf[t] = cplx{0.0, 0.0};
    __syncthreads();

    // first group of derivative matrices: in flight during the inversion
    const cplx* da = reinterpret_cast<const cplx*>(sb.DA[s]) + (w * D + k) * (long long)NDp * n * n;
    auto stage = [&](int d0, int g_cnt) {
        if (np == n) {
            const int tot = g_cnt * n * n;
            const cplx* src = da + (long long)d0 * n * n;
            for (int t = tid; t < tot; t += nthr) cp_async16_s(buf + t, src + t);
        } else {
            const int rows = g_cnt * n;              // one warp per matrix row
            for (int rr = warp; rr < rows; rr += nwarp) {
                const int g = rr / n, i = rr - g * n;
                const cplx* src = da + ((long long)(d0 + g) * n + i) * n;
                cplx* dst = buf + (g * np + i) * np;
                for (int o = lane; o < n; o += 32) cp_async16_s(dst + o, src + o);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    stage(0, min(G, ND));

    double logabs = 0.0;
    cplx phase{1.0, 0.0};
    for (int kk = 0; kk < n; ++kk) {
        if (tid < 32) {
            double best = -1.0;
            int bi = kk;
            for (int r = kk + tid; r < n; r += 32) {
                double v = cabs1(Xs[r * np + kk]);
                if (v > best) { best = v; bi = r; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, best, off);
                int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (tid == 0) { s_p = bi; piv[kk] = bi; }
        }
        __syncthreads();
        const int p = s_p;
        if (p != kk) {
            for (int c = tid; c < n; c += nthr) {
                cplx a = Xs[kk * np + c];
                Xs[kk * np + c] = Xs[p * np + c];
                Xs[p * np + c] = a;
            }
            phase = cplx{-phase.re, -phase.im};
        }
        __syncthreads();
        const cplx pv = Xs[kk * np + kk];
        {
            double a = hypot(pv.re, pv.im);
            logabs += log(a);
            phase = cmul(phase, cplx{pv.re / a, pv.im / a});
        }
        const cplx ipv = cinv(pv);
        for (int r = tid; r < n; r += nthr) colk[r] = Xs[r * np + kk];
        __syncthreads();
        for (int c = tid; c < n; c += nthr) {
            cplx a = (c == kk) ? cplx{1.0, 0.0} : Xs[kk * np + c];
            Xs[kk * np + c] = cmul(a, ipv);
        }
        __syncthreads();
        for (int t = tid; t < n * n; t += nthr) {
            int r = t / n, c = t - r * n;
            if (r == kk) continue;
            cplx f = colk[r];
            cplx a = (c == kk) ? cplx{0.0, 0.0} : Xs[r * np + c];
            cplx b = Xs[kk * np + c];
            a.re -= f.re * b.re - f.im * b.im;
            a.im -= f.re * b.im + f.im * b.re;
            Xs[r * np + c] = a;
        }
        __syncthreads();
    }
    if (tid == 0) {
        double* ld = sb.LOGDET + ((w * 2 + s) * D + k) * 3;
        ld[0] = logabs; ld[1] = phase.re; ld[2] = phase.im;
    }
    // undo the row pivoting: swap columns in reverse order
    for (int kk = n - 1; kk >= 0; --kk) {
        const int p = piv[kk];
        if (p != kk) {
            for (int r = tid; r < n; r += nthr) {
                cplx a = Xs[r * np + kk];
                Xs[r * np + kk] = Xs[r * np + p];
                Xs[r * np + p] = a;
            }
        }
        __syncthreads();
    }
    // now Xs[o][i] = (M^-1)[o,i]

    auto block_sum2 = [&](double a, double b, double& oa, double& ob) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, off);
            b += __shfl_xor_sync(0xffffffffu, b, off);
        }
        __syncthreads();
        if (lane == 0) { s_red[2 * warp] = a; s_red[2 * warp + 1] = b; }
        __syncthreads();
        oa = 0.0; ob = 0.0;
        for (int q = 0; q < nwarp; ++q) { oa += s_red[2 * q]; ob += s_red[2 * q + 1]; }
    };

    // tr(X lapM) = sum_{i,o} X[o,i] lapM[i,o]
    {
        const cplx* lm = reinterpret_cast<const cplx*>(sb.LAPM[s]) + (w * D + k) * (long long)n * n;
        cplx acc{0.0, 0.0};
        for (int t = tid; t < n * n; t += nthr) {
            int i = t / n, o = t - i * n;
            cfma(acc, Xs[o * np + i], lm[t]);
        }
        double ra, rb;
        block_sum2(acc.re, acc.im, ra, rb);
        if (tid == 0) {
            double* tl = sb.TRLAP + ((w * 2 + s) * D + k) * 2;
            tl[0] = ra; tl[1] = rb;
        }
    }

    // ---- direction groups ---------------------------------------------------
    double* tau_out = sb.TAU + ((w * 2 + s) * D + k) * (long long)NDp * 2;
    const int nb2 = nb * nb;
    const int n_items = (G >> 1) * nb2;
    const bool active = tid < n_items;
    int pi = 0, rb = 0, cb = 0;
    if (active) { pi = tid / nb2; int rem = tid - pi * nb2; rb = (rem / nb) * 3; cb = (rem - (rem / nb) * nb) * 3; }
    const int g0 = 2 * pi;
    const cplx* xa = Xs + rb * np;
    cplx* b0 = buf + (g0 * np) * np + cb;            // dM_{g0}[i][cb..]
    cplx* b1 = b0 + np * np;                         // dM_{g0+1}
    cplx sq{0.0, 0.0};
    for (int d0 = 0; d0 < ND; d0 += G) {
        const int g_cnt = min(G, ND - d0);
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        cplx y0[3][3], y1[3][3];
        if (active) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) { y0[a][b] = cplx{0.0, 0.0}; y1[a][b] = cplx{0.0, 0.0}; }
#pragma unroll 3
            for (int i = 0; i < n; ++i) {
                cplx xv[3], dv0[3], dv1[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) xv[a] = xa[a * np + i];
#pragma unroll
                for (int b = 0; b < 3; ++b) { dv0[b] = b0[i * np + b]; dv1[b] = b1[i * np + b]; }
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) { cfma(y0[a][b], xv[a], dv0[b]); cfma(y1[a][b], xv[a], dv1[b]); }
            }
        }
        __syncthreads();                             // every operand of this group has been read
        if (active) {
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    b0[(rb + a) * np + b] = y0[a][b];
                    b1[(rb + a) * np + b] = y1[a][b];
                }
            if (rb == cb) {
                s_tau[g0 * nb + rb / 3] = cplx{y0[0][0].re + y0[1][1].re + y0[2][2].re, y0[0][0].im + y0[1][1].im + y0[2][2].im};
                s_tau[(g0 + 1) * nb + rb / 3] = cplx{y1[0][0].re + y1[1][1].re + y1[2][2].re, y1[0][0].im + y1[1][1].im + y1[2][2].im};
            }
        }
        __syncthreads();
        if (active) {
            // sum_{a,b in block} Y[rb+a, cb+b] Y[cb+b, rb+a]; the transposed block sits at rows cb.., columns rb..
            const cplx* t0 = buf + (g0 * np + cb) * np + rb;
            const cplx* t1 = t0 + np * np;
            cplx q0{0.0, 0.0}, q1{0.0, 0.0};
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) {
                    cfma(q0, y0[a][b], t0[b * np + a]);
                    cfma(q1, y1[a][b], t1[b * np + a]);
                }
            if (g0 < g_cnt) { sq.re += q0.re; sq.im += q0.im; }
            if (g0 + 1 < g_cnt) { sq.re += q1.re; sq.im += q1.im; }
        }
        if (tid < g_cnt) {
            cplx tr{0.0, 0.0};
            for (int j = 0; j < nb; ++j) { cplx v = s_tau[tid * nb + j]; tr.re += v.re; tr.im += v.im; }
            tau_out[2 * (d0 + tid)] = tr.re;
            tau_out[2 * (d0 + tid) + 1] = tr.im;
        }
        if (d0 + G < ND) {
            __syncthreads();                         // transposed blocks and s_tau consumed: the buffer may be refilled
            stage(d0 + G, min(G, ND - d0 - G));
        }
    }
    {
        double ra, rb2;
        block_sum2(sq.re, sq.im, ra, rb2);
        if (tid == 0) {
            double* ts = sb.TRSQ + ((w * 2 + s) * D + k) * 2;
            ts[0] = ra; ts[1] = rb2;
        }
    }
}

// ---------------------------------------------------------------------------
// Determinant kernel of the Laplacian sweep, third version: the products Y_d = X dM_d on the fp64 tensor
// path (mma.sync m8n8k4 f64 -> DMMA), one CTA per (walker, spin, determinant).
//   * in-place Gauss-Jordan inverse X = M^-1 (complex, shared memory), then its real embedding
//         Xe = [ Xr  -Xi ; Xi  Xr ]   (2n x 2n, padded to MT*8, leading dimension = 4 mod 16: conflict-free A fragments)
//   * G directions per group are staged de-interleaved with 8-byte cp.async into
//         Be = [ Re dM_g ; Im dM_g ]  (rows k = (re|im, electron i), columns (g, orbital o), ld = 4 mod 16)
//     so that ONE real GEMM  Ye = Xe . Be  yields [ Re Y ; Im Y ] of all G directions; warp w owns the
//     n-tiles {w, w+8} and all MT m-tiles (7 A-fragment + 2 B-fragment loads per 14 DMMAs at n = 27).
//   * Y is written in place over Be; tr(Y_d) per direction and sum_d tr(Y_d^2) are read back from there.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8_s(void* smem_dst, const void* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

template <int MT>
__global__ void __launch_bounds__(256, (MT <= 8) ? 2 : 1) det_dmma_kernel(const DsSys sys, const SlaterBufs sb, int G, int LDB, int alias) {
    const DsDims& dm = sys.d;
    const int D = dm.D, NDp = dm.NDp, ND = dm.ND;
    const int k = blockIdx.x % D;
    const int nblk = ds_nblk(dm);
    const int s = (blockIdx.x / D) % nblk;          // matrix block: spin channel, or the single full determinant
    const long long w = blockIdx.x / (nblk * D);
    const int n = ds_blk_n(dm, s);
    if (dm.full_det) {                               // the second slot of the per-spin results stays neutral
        const long long slot1 = (w * 2 + 1) * D + k;
        if (threadIdx.x == 0) {
            double* ld1 = sb.LOGDET + slot1 * 3;
            ld1[0] = 0.0; ld1[1] = 1.0; ld1[2] = 0.0;
            if (sb.TRSQ) { sb.TRSQ[slot1 * 2] = 0.0; sb.TRSQ[slot1 * 2 + 1] = 0.0; }
            if (sb.TRLAP) { sb.TRLAP[slot1 * 2] = 0.0; sb.TRLAP[slot1 * 2 + 1] = 0.0; }
        }
        if (sb.TAU)
            for (int t = threadIdx.x; t < 2 * dm.NDp; t += blockDim.x) sb.TAU[slot1 * 2 * dm.NDp + t] = 0.0;
    }
    const int np = n;                               // no padding of the complex inverse
    constexpr int MP = MT * 8;                      // rows of Xe / Be (>= 2n)
    constexpr int LDX = MP + 4;                     // = 4 or 12 mod 16
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;

    extern __shared__ __align__(16) double smraw[];
    double* Xe = smraw;                             // [MP][LDX]
    double* Be = Xe + MP * LDX;                     // [MP][LDB]
    cplx* Xs = reinterpret_cast<cplx*>(alias ? Be : Be + MP * LDB);   // [n][n] complex inverse (aliases Be when tight)
    cplx* colk = Xs + n * n;                        // [n]
    __shared__ int piv[128];
    __shared__ int s_p;
    __shared__ double s_red[2 * 16 + 4];

    const cplx* mat = reinterpret_cast<const cplx*>(sb.MAT[s]) + (w * D + k) * (long long)n * n;
    for (int t = tid; t < n * n; t += nthr) Xs[t] = mat[t];
    if (!alias)
        for (int t = tid; t < MP * LDB; t += nthr) Be[t] = 0.0;
    __syncthreads();

    const double* da = sb.DA[s] + (w * D + k) * (long long)NDp * n * n * 2;     // complex (re, im) interleaved
    // every thread owns up to JMAX fixed matrix elements (i, o) = idx / n, idx % n, idx = tid + j * nthr, of every
    // direction: their operand-buffer offsets (straight and transposed) are computed once, not per group
    // (matrices beyond 32 x 32 run one CTA per SM: enough registers to keep every element's offsets resident too)
    constexpr int NMAXT = MT * 4;                    // largest n this instantiation serves
    constexpr int JMAX = (MT <= 8) ? 4 : (NMAXT * NMAXT + 255) / 256;
    constexpr int JTRI = (MT <= 8) ? 4 : (NMAXT * (NMAXT + 1) / 2 + 255) / 256;
    const int nn = n * n;
    int off_ab[JMAX];
#pragma unroll
    for (int j = 0; j < JMAX; ++j) {
        const int idx = tid + j * nthr;
        const int a = idx / n, b = idx - a * n;
        off_ab[j] = a * LDB + b;
    }
    // trace elements: sum_{a,b} Y[a,b] Y[b,a] = sum_a Y[a,a]^2 + 2 sum_{a<b} Y[a,b] Y[b,a]: every thread owns up to
    // JMAX fixed pairs (a <= b) of the upper triangle (linear index idx = tid + j * nthr, row-major over the triangle)
    const int n_tri = n * (n + 1) / 2;
    int tr_ab[JTRI], tr_ba[JTRI];
    double tr_w[JTRI];
#pragma unroll
    for (int j = 0; j < JTRI; ++j) {
        const int idx = tid + j * nthr;
        tr_ab[j] = -1; tr_ba[j] = 0; tr_w[j] = 0.0;
        if (idx < n_tri) {
            // row a holds n - a entries (b = a .. n-1); first index of row a: a n - a (a - 1) / 2
            int a = (int)((2.0 * n + 1.0 - sqrt((2.0 * n + 1.0) * (2.0 * n + 1.0) - 8.0 * idx)) * 0.5);
            while (a > 0 && a * n - a * (a - 1) / 2 > idx) --a;
            while ((a + 1) * n - (a + 1) * a / 2 <= idx) ++a;
            const int b = a + (idx - (a * n - a * (a - 1) / 2));
            tr_ab[j] = a * LDB + b;
            tr_ba[j] = b * LDB + a;
            tr_w[j] = (a == b) ? 1.0 : 2.0;
        }
    }
    auto stage = [&](int d0, int g_cnt) {
        for (int g = 0; g < g_cnt; ++g) {
            const double* src = da + (long long)(d0 + g) * nn * 2;
            double* dst = Be + g * n;
#pragma unroll
            for (int j = 0; j < JMAX; ++j) {
                const int idx = tid + j * nthr;
                if (idx < nn) {
                    cp_async8_s(dst + off_ab[j], src + 2 * idx);
                    cp_async8_s(dst + n * LDB + off_ab[j], src + 2 * idx + 1);
                }
            }
            for (int idx = tid + JMAX * nthr; idx < nn; idx += nthr) {      // matrices larger than JMAX * nthr elements
                const int a = idx / n, b = idx - a * n;
                cp_async8_s(dst + a * LDB + b, src + 2 * idx);
                cp_async8_s(dst + (n + a) * LDB + b, src + 2 * idx + 1);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
    if (!alias) stage(0, min(G, ND));

    double logabs = 0.0;
    cplx phase{1.0, 0.0};
    for (int kk = 0; kk < n; ++kk) {
        if (tid < 32) {
            double best = -1.0;
            int bi = kk;
            for (int r = kk + tid; r < n; r += 32) {
                double v = cabs1(Xs[r * np + kk]);
                if (v > best) { best = v; bi = r; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, best, off);
                int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (tid == 0) { s_p = bi; piv[kk] = bi; }
        }
        __syncthreads();
        const int p = s_p;
        if (p != kk) {
            for (int c = tid; c < n; c += nthr) {
                cplx a = Xs[kk * np + c];
                Xs[kk * np + c] = Xs[p * np + c];
                Xs[p * np + c] = a;
            }
            phase = cplx{-phase.re, -phase.im};
        }
        __syncthreads();
        const cplx pv = Xs[kk * np + kk];
        {
            double a = hypot(pv.re, pv.im);
            logabs += log(a);
            phase = cmul(phase, cplx{pv.re / a, pv.im / a});
        }
        const cplx ipv = cinv(pv);
        for (int r = tid; r < n; r += nthr) colk[r] = Xs[r * np + kk];
        __syncthreads();
        for (int c = tid; c < n; c += nthr) {
            cplx a = (c == kk) ? cplx{1.0, 0.0} : Xs[kk * np + c];
            Xs[kk * np + c] = cmul(a, ipv);
        }
        __syncthreads();
        for (int t = tid; t < n * n; t += nthr) {
            int r = t / n, c = t - r * n;
            if (r == kk) continue;
            cplx f = colk[r];
            cplx a = (c == kk) ? cplx{0.0, 0.0} : Xs[r * np + c];
            cplx b = Xs[kk * np + c];
            a.re -= f.re * b.re - f.im * b.im;
            a.im -= f.re * b.im + f.im * b.re;
            Xs[r * np + c] = a;
        }
        __syncthreads();
    }
    if (tid == 0) {
        double* ld = sb.LOGDET + ((w * 2 + s) * D + k) * 3;
        ld[0] = logabs; ld[1] = phase.re; ld[2] = phase.im;
    }
    for (int kk = n - 1; kk >= 0; --kk) {            // undo the row pivoting: swap columns in reverse order
        const int p = piv[kk];
        if (p != kk) {
            for (int r = tid; r < n; r += nthr) {
                cplx a = Xs[r * np + kk];
                Xs[r * np + kk] = Xs[r * np + p];
                Xs[r * np + p] = a;
            }
        }
        __syncthreads();
    }
    // now Xs[o][i] = (M^-1)[o,i]

    auto block_sum2 = [&](double a, double b, double& oa, double& ob) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, off);
            b += __shfl_xor_sync(0xffffffffu, b, off);
        }
        __syncthreads();
        if (lane == 0) { s_red[2 * warp] = a; s_red[2 * warp + 1] = b; }
        __syncthreads();
        oa = 0.0; ob = 0.0;
        for (int q = 0; q < nwarp; ++q) { oa += s_red[2 * q]; ob += s_red[2 * q + 1]; }
    };
    {   // tr(X lapM) = sum_{i,o} X[o,i] lapM[i,o]
        const cplx* lm = reinterpret_cast<const cplx*>(sb.LAPM[s]) + (w * D + k) * (long long)n * n;
        cplx acc{0.0, 0.0};
        for (int t = tid; t < n * n; t += nthr) {
            int i = t / n, o = t - i * n;
            cfma(acc, Xs[o * np + i], lm[t]);
        }
        double ra, rb;
        block_sum2(acc.re, acc.im, ra, rb);
        if (tid == 0) {
            double* tl = sb.TRLAP + ((w * 2 + s) * D + k) * 2;
            tl[0] = ra; tl[1] = rb;
        }
    }
    // real embedding of the inverse
    for (int t = tid; t < MP * LDX; t += nthr) {
        const int r = t / LDX, c = t - r * LDX;
        double v = 0.0;
        if (r < 2 * n && c < 2 * n) {
            const int o = (r < n) ? r : r - n, i = (c < n) ? c : c - n;
            const cplx x = Xs[o * np + i];
            v = ((r < n) == (c < n)) ? x.re : ((r < n) ? -x.im : x.im);
        }
        Xe[t] = v;
    }
    __syncthreads();
    if (alias) {                                     // Xs is dead: its storage becomes the operand buffer
        for (int t = tid; t < MP * LDB; t += nthr) Be[t] = 0.0;
        __syncthreads();
        stage(0, min(G, ND));
    }

    // ---- direction groups ---------------------------------------------------
    double* tau_out = sb.TAU + ((w * 2 + s) * D + k) * (long long)NDp * 2;
    const int ntiles = (G * n + 7) >> 3;             // n-tiles of 8 columns
    const int gid = lane >> 2, tig = lane & 3;
    const double* xa = Xe + gid * LDX + tig;         // A fragment: row gid, column tig of the 8x4 tile
    double sq_re = 0.0, sq_im = 0.0;
    for (int d0 = 0; d0 < ND; d0 += G) {
        const int g_cnt = min(G, ND - d0);
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        __syncthreads();
        double acc[MT][2][2];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[mt][j][0] = acc[mt][j][1] = 0.0;
        const int nt0 = warp, nt1 = warp + nwarp;
        const bool has0 = nt0 < ntiles, has1 = nt1 < ntiles;
        // (two copies of the k loop: a predicated-off DMMA still occupies its issue slot and the pipe)
        if (has0 && has1) {
            const double* xb0 = Be + tig * LDB + nt0 * 8 + gid;       // B fragment: row tig, column gid of the 4x8 tile
            const double* xb1 = Be + tig * LDB + nt1 * 8 + gid;
#pragma unroll 2
            for (int k0 = 0; k0 < MP; k0 += 4) {
                double a[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) a[mt] = xa[mt * 8 * LDX + k0];
                const double b0 = xb0[k0 * LDB];
                const double b1 = xb1[k0 * LDB];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    dmma884(acc[mt][0][0], acc[mt][0][1], a[mt], b0);
                    dmma884(acc[mt][1][0], acc[mt][1][1], a[mt], b1);
                }
            }
        } else if (has0) {
            const double* xb0 = Be + tig * LDB + nt0 * 8 + gid;
#pragma unroll 2
            for (int k0 = 0; k0 < MP; k0 += 4) {
                double a[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) a[mt] = xa[mt * 8 * LDX + k0];
                const double b0 = xb0[k0 * LDB];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) dmma884(acc[mt][0][0], acc[mt][0][1], a[mt], b0);
            }
        }
        __syncthreads();                             // every operand of this group has been read
        if (has0) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                double* y0 = Be + (mt * 8 + gid) * LDB + nt0 * 8 + 2 * tig;
                *reinterpret_cast<double2*>(y0) = make_double2(acc[mt][0][0], acc[mt][0][1]);
                if (has1) {
                    double* y1 = Be + (mt * 8 + gid) * LDB + nt1 * 8 + 2 * tig;
                    *reinterpret_cast<double2*>(y1) = make_double2(acc[mt][1][0], acc[mt][1][1]);
                }
            }
        }
        __syncthreads();
        // sum_{a,b} Y[a,b] Y[b,a] of the valid directions (complex), the thread's fixed pairs (a <= b)
        for (int g = 0; g < g_cnt; ++g) {
            const double* col = Be + g * n;
            const double* coli = col + n * LDB;
#pragma unroll
            for (int j = 0; j < JTRI; ++j) {
                if (tr_ab[j] >= 0) {
                    const double yr = col[tr_ab[j]] * tr_w[j], yi = coli[tr_ab[j]] * tr_w[j];
                    const double zr = col[tr_ba[j]], zi = coli[tr_ba[j]];
                    sq_re = fma(yr, zr, sq_re); sq_re = fma(-yi, zi, sq_re);
                    sq_im = fma(yr, zi, sq_im); sq_im = fma(yi, zr, sq_im);
                }
            }
            for (int idx = tid + JTRI * nthr; idx < n_tri; idx += nthr) {       // triangles larger than JTRI * nthr pairs
                int a = (int)((2.0 * n + 1.0 - sqrt((2.0 * n + 1.0) * (2.0 * n + 1.0) - 8.0 * idx)) * 0.5);
                while (a > 0 && a * n - a * (a - 1) / 2 > idx) --a;
                while ((a + 1) * n - (a + 1) * a / 2 <= idx) ++a;
                const int b = a + (idx - (a * n - a * (a - 1) / 2));
                const double wgt = (a == b) ? 1.0 : 2.0;
                const double yr = col[a * LDB + b] * wgt, yi = coli[a * LDB + b] * wgt;
                const double zr = col[b * LDB + a], zi = coli[b * LDB + a];
                sq_re = fma(yr, zr, sq_re); sq_re = fma(-yi, zi, sq_re);
                sq_im = fma(yr, zi, sq_im); sq_im = fma(yi, zr, sq_im);
            }
        }
        // tr(Y_g): warp g
        for (int g = warp; g < g_cnt; g += nwarp) {
            double tr = 0.0, ti = 0.0;
            for (int o = lane; o < n; o += 32) { tr += Be[o * LDB + g * n + o]; ti += Be[(n + o) * LDB + g * n + o]; }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                tr += __shfl_xor_sync(0xffffffffu, tr, off);
                ti += __shfl_xor_sync(0xffffffffu, ti, off);
            }
            if (lane == 0) { tau_out[2 * (d0 + g)] = tr; tau_out[2 * (d0 + g) + 1] = ti; }
        }
        if (d0 + G < ND) {
            __syncthreads();                         // Y consumed: the buffer may be refilled
            stage(d0 + G, min(G, ND - d0 - G));
        }
    }
    {
        double ra, rb2;
        block_sum2(sq_re, sq_im, ra, rb2);
        if (tid == 0) {
            double* ts = sb.TRSQ + ((w * 2 + s) * D + k) * 2;
            ts[0] = ra; ts[1] = rb2;
        }
    }
}

template <int MT>
int launch_det_dmma(const DsSys& sys, const SlaterBufs& sb, int Wc, int nmax, cudaStream_t stream, bool* done) {
    *done = false;
    constexpr int MP = MT * 8, LDX = MP + 4;
    // directions per group: at most 16 n-tiles (two per warp), operand buffer small enough for two CTAs per SM
    int G = (16 * 8) / nmax;
    if (G < 1) return 0;
    if (G > sys.d.ND) G = sys.d.ND;
    auto ldb_of = [&](int g) { int npad = (g * nmax + 7) & ~7; return npad + ((4 - (npad & 15)) & 15); };    // = 4 mod 16
    auto smem_of = [&](int g, bool alias) {
        size_t xs = (size_t)(nmax * nmax + nmax) * sizeof(cplx);
        size_t be = (size_t)MP * ldb_of(g) * sizeof(double);
        return (size_t)MP * LDX * sizeof(double) + (alias ? (be > xs ? be : xs) : be + xs);
    };
    while (G > 1 && smem_of(G, false) > 112 * 1024) --G;
    bool alias = false;
    if (smem_of(G, false) > 112 * 1024) alias = true;
    const size_t smem = smem_of(G, alias);
    if (smem > 226 * 1024) return 0;
    static size_t cfg = 0;
    if (smem > cfg) {
        DS_CUDA_CHECK(cudaFuncSetAttribute(det_dmma_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cfg = smem;
    }
    dim3 grid((unsigned)((long long)Wc * ds_nblk(sys.d) * sys.d.D));
    // (measured and rejected, profiles/r2_det_variants.log: a start-up stagger of the second CTA of every SM, and a
    //  192-thread configuration with three CTAs per SM)
    det_dmma_kernel<MT><<<grid, 256, smem, stream>>>(sys, sb, G, ldb_of(G), alias ? 1 : 0);
    DS_CUDA_CHECK(cudaGetLastError());
    *done = true;
    return 0;
}

// ---------------------------------------------------------------------------
// Value-only log-determinants (the Metropolis inner loop): ONE WARP per matrix, lane r holds row r in registers,
// complex LU with partial pivoting done with warp shuffles and no shared memory / block barriers.  Pivoting is
// implicit (rows never move: the pivot row of step k is the not-yet-used lane with the largest |a[k]|, the same
// choice as the row-swapping LU of det_kernel), the permutation parity is recovered from the pivot order.
// NMAX >= n is the compile-time row length (8, 16 or 32).
// ---------------------------------------------------------------------------
template <int NMAX>
__global__ void __launch_bounds__(128) det_warp_kernel(const DsSys sys, const SlaterBufs sb, long long n_mats) {
    const DsDims& dm = sys.d;
    const int D = dm.D;
    const int lane = threadIdx.x & 31;
    const long long m = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (m >= n_mats) return;
    const int nblk = ds_nblk(dm);
    const int k = (int)(m % D);
    const int s = (int)((m / D) % nblk);
    const long long w = m / ((long long)nblk * D);
    const int n = ds_blk_n(dm, s);
    const cplx* mat = reinterpret_cast<const cplx*>(sb.MAT[s]) + (w * D + k) * (long long)n * n;
    double ar[NMAX], ai[NMAX];
#pragma unroll
    for (int c = 0; c < NMAX; ++c) {
        cplx v{0.0, 0.0};
        if (lane < n && c < n) v = mat[lane * n + c];
        ar[c] = v.re; ai[c] = v.im;
    }
    // det = prod of the pivots, kept as a complex mantissa P (rescaled by a power of two every step) and an exponent:
    // log|det| = log|P| + esum ln 2, phase = P / |P|: no per-step log / hypot / division by the modulus
    cplx P{1.0, 0.0};
    int esum = 0;
    bool done = lane >= n;                       // rows already used as pivots (and the padding lanes)
    int my_step = -1;                            // step at which this row was the pivot
    // fully unrolled over the elimination steps: every register-array index is a compile-time constant
#pragma unroll(NMAX)
    for (int kk = 0; kk < NMAX; ++kk) {
        if (kk < n) {                            // warp-uniform
            const double pr = ar[kk], pi = ai[kk];
            double best = done ? -1.0 : fabs(pr) + fabs(pi);
            int bi = lane;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            const int p = bi;                    // warp-uniform pivot row
            const double pvr = __shfl_sync(0xffffffffu, pr, p), pvi = __shfl_sync(0xffffffffu, pi, p);
            {
                P = cmul(P, cplx{pvr, pvi});
                int e;
                (void)frexp(fmax(fabs(P.re), fabs(P.im)), &e);
                if (e > -900 && e < 900) {       // zero / inf / nan products keep their value (and propagate)
                    P.re = ldexp(P.re, -e); P.im = ldexp(P.im, -e);
                    esum += e;
                }
            }
            if (lane == p) { done = true; my_step = kk; }
            const cplx ipv = cinv(cplx{pvr, pvi});
            const cplx f = done ? cplx{0.0, 0.0} : cmul(cplx{pr, pi}, ipv);      // multiplier of the active rows
#pragma unroll(NMAX)
            for (int c = kk + 1; c < NMAX; ++c) {
                const double br = __shfl_sync(0xffffffffu, ar[c], p), bim = __shfl_sync(0xffffffffu, ai[c], p);
                ar[c] -= f.re * br - f.im * bim;
                ai[c] -= f.re * bim + f.im * br;
            }
        }
    }
    const double pa = hypot(P.re, P.im);
    const double logabs = log(pa) + (double)esum * 0.69314718055994530942;
    cplx phase{P.re / pa, P.im / pa};
    // parity of the permutation step -> row: sign = (-1)^(n - #cycles)
    unsigned visited = 0u;
    int cycles = 0;
    for (int r0 = 0; r0 < n; ++r0) {
        if (visited & (1u << r0)) continue;
        ++cycles;
        int r = r0;
        while (!(visited & (1u << r))) {
            visited |= 1u << r;
            r = __shfl_sync(0xffffffffu, my_step, r);     // row r was the pivot of step my_step[r]: next element of the cycle
        }
    }
    if (((n - cycles) & 1) != 0) phase = cplx{-phase.re, -phase.im};
    if (lane == 0) {
        double* ld = sb.LOGDET + ((w * 2 + s) * D + k) * 3;
        ld[0] = logabs; ld[1] = phase.re; ld[2] = phase.im;
        if (dm.full_det) {
            double* ld1 = sb.LOGDET + ((w * 2 + 1) * D + k) * 3;
            ld1[0] = 0.0; ld1[1] = 1.0; ld1[2] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------
// Combine determinants (logdet_matmul, network.py:395-427) and the kinetic energy.
// One warp per walker.
// ---------------------------------------------------------------------------
template <bool LAP>
__global__ void __launch_bounds__(128) combine_kernel(const DsSys sys, const SlaterBufs sb, int Wc, double* log_abs,
                                                      double* phase, double* ke_re, double* ke_im, double* gx_abs,
                                                      double* gx_phase) {
    const DsDims& dm = sys.d;
    const int D = dm.D, NDp = dm.NDp, ND = dm.ND;
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= Wc) return;
    const double* ld = sb.LOGDET + w * 2 * D * 3;
    // log-sum-exp over determinants (every lane redundantly; D is small)
    double mx = -INFINITY;
    for (int k = 0; k < D; ++k) {
        double l = ld[k * 3] + ld[(D + k) * 3];
        mx = fmax(mx, l);
    }
    cplx tot{0.0, 0.0};
    for (int k = 0; k < D; ++k) {
        double l = ld[k * 3] + ld[(D + k) * 3];
        cplx ph = cmul(cplx{ld[k * 3 + 1], ld[k * 3 + 2]}, cplx{ld[(D + k) * 3 + 1], ld[(D + k) * 3 + 2]});
        double m = exp(l - mx);
        tot.re += ph.re * m; tot.im += ph.im * m;
    }
    if (lane == 0) {
        if (log_abs) log_abs[w] = log(hypot(tot.re, tot.im)) + mx;
        if (phase) phase[w] = atan2(tot.im, tot.re);
    }
    if (!LAP) return;
    const cplx itot = cinv(tot);
    if (gx_abs != nullptr || gx_phase != nullptr) {
        // d log psi / d x_d = sum_k w_k sum_s tr(X dM_d): real part = grad log|psi|, imaginary part = grad of the phase
        for (int d = lane; d < ND; d += 32) {
            cplx gsum{0.0, 0.0};
            for (int k = 0; k < D; ++k) {
                double l = ld[k * 3] + ld[(D + k) * 3];
                cplx ph = cmul(cplx{ld[k * 3 + 1], ld[k * 3 + 2]}, cplx{ld[(D + k) * 3 + 1], ld[(D + k) * 3 + 2]});
                double m = exp(l - mx);
                cplx wk = cmul(cplx{ph.re * m, ph.im * m}, itot);
                const double* t0 = sb.TAU + ((w * 2 + 0) * D + k) * (long long)NDp * 2;
                const double* t1 = sb.TAU + ((w * 2 + 1) * D + k) * (long long)NDp * 2;
                cfma(gsum, wk, cplx{t0[2 * d] + t1[2 * d], t0[2 * d + 1] + t1[2 * d + 1]});
            }
            if (gx_abs) gx_abs[w * ND + d] = gsum.re;
            if (gx_phase) gx_phase[w * ND + d] = gsum.im;
        }
    }
    cplx ke{0.0, 0.0};
    for (int k = 0; k < D; ++k) {
        double l = ld[k * 3] + ld[(D + k) * 3];
        cplx ph = cmul(cplx{ld[k * 3 + 1], ld[k * 3 + 2]}, cplx{ld[(D + k) * 3 + 1], ld[(D + k) * 3 + 2]});
        double m = exp(l - mx);
        cplx wk = cmul(cplx{ph.re * m, ph.im * m}, itot);
        const double* t0 = sb.TAU + ((w * 2 + 0) * D + k) * (long long)NDp * 2;
        const double* t1 = sb.TAU + ((w * 2 + 1) * D + k) * (long long)NDp * 2;
        cplx sq{0.0, 0.0};
        for (int d = lane; d < ND; d += 32) {
            cplx t{t0[2 * d] + t1[2 * d], t0[2 * d + 1] + t1[2 * d + 1]};
            cfma(sq, t, t);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sq.re += __shfl_xor_sync(0xffffffffu, sq.re, off);
            sq.im += __shfl_xor_sync(0xffffffffu, sq.im, off);
        }
        const double* tl0 = sb.TRLAP + ((w * 2 + 0) * D + k) * 2;
        const double* tl1 = sb.TRLAP + ((w * 2 + 1) * D + k) * 2;
        const double* ts0 = sb.TRSQ + ((w * 2 + 0) * D + k) * 2;
        const double* ts1 = sb.TRSQ + ((w * 2 + 1) * D + k) * 2;
        cplx per{tl0[0] + tl1[0] - ts0[0] - ts1[0] + sq.re, tl0[1] + tl1[1] - ts0[1] - ts1[1] + sq.im};
        cfma(ke, wk, per);
    }
    if (lane == 0) {
        if (ke_re) ke_re[w] = -0.5 * ke.re;
        if (ke_im) ke_im[w] = -0.5 * ke.im;
    }
}

}  // namespace

static size_t g_det_smem[2] = {0, 0};     // largest dynamic shared memory configured for det_kernel<false/true>

int ds_launch_etab(const DsSys& sys, const SlaterBufs& sb, int Wc, int npar_max, bool jets, cudaStream_t stream) {
    dim3 grid((unsigned)((long long)Wc * sys.d.N));
    if (jets) etab_kernel<true><<<grid, 256, 0, stream>>>(sys, sb, npar_max);
    else etab_kernel<false><<<grid, 256, 0, stream>>>(sys, sb, npar_max);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_orb_assemble(const DsSys& sys, const SlaterBufs& sb, int Wc, int npar_max, bool jets,
                           cudaStream_t stream) {
    dim3 grid((unsigned)((long long)Wc * sys.d.N));
    if (jets) assemble_kernel<true><<<grid, 256, 0, stream>>>(sys, sb, npar_max);
    else assemble_kernel<false><<<grid, 256, 0, stream>>>(sys, sb, npar_max);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_orb_mean_addend(const DsSys& sys, const SlaterBufs& sb, int Wc, int npar_max, const double* GO0, const double* GO1,
                              int ldgo, bool jets, cudaStream_t stream) {
    dim3 grid((unsigned)((long long)Wc * sys.d.N));
    if (jets) orb_mean_addend_kernel<true><<<grid, 256, 0, stream>>>(sys, sb, npar_max, GO0, GO1, ldgo);
    else orb_mean_addend_kernel<false><<<grid, 256, 0, stream>>>(sys, sb, npar_max, GO0, GO1, ldgo);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_det(const DsSys& sys, const SlaterBufs& sb, int Wc, bool lap, cudaStream_t stream) {
    const int nmax = sys.d.full_det ? sys.d.N : (sys.d.n_up > sys.d.n_dn ? sys.d.n_up : sys.d.n_dn);
    DS_REQUIRE(nmax <= 128, "determinants larger than 128x128 are not supported (n_s=%d)", nmax);
    const int np = ((nmax + 2) / 3) * 3;
    const int nb = np / 3;
    dim3 grid((unsigned)((long long)Wc * ds_nblk(sys.d) * sys.d.D));
    static const bool use_v1 = getenv("DS_DET_V1") && atoi(getenv("DS_DET_V1")) != 0;
    static const bool use_v2 = getenv("DS_DET_V2") && atoi(getenv("DS_DET_V2")) != 0;
    if (lap && !use_v1 && !use_v2) {
        // fp64 tensor-path kernel: MT = m-tiles of the real embedding (2 n_s rows, padded to 8)
        const int mt = (2 * nmax + 7) / 8;
        bool done = false;
        int rc = 0;
        switch (mt) {
            case 1: rc = launch_det_dmma<1>(sys, sb, Wc, nmax, stream, &done); break;
            case 2: rc = launch_det_dmma<2>(sys, sb, Wc, nmax, stream, &done); break;
            case 3: rc = launch_det_dmma<3>(sys, sb, Wc, nmax, stream, &done); break;
            case 4: rc = launch_det_dmma<4>(sys, sb, Wc, nmax, stream, &done); break;
            case 5: case 6: rc = launch_det_dmma<6>(sys, sb, Wc, nmax, stream, &done); break;
            case 7: rc = launch_det_dmma<7>(sys, sb, Wc, nmax, stream, &done); break;
            case 8: rc = launch_det_dmma<8>(sys, sb, Wc, nmax, stream, &done); break;
            case 9: case 10: rc = launch_det_dmma<10>(sys, sb, Wc, nmax, stream, &done); break;
            case 11: case 12: rc = launch_det_dmma<12>(sys, sb, Wc, nmax, stream, &done); break;
            case 13: case 14: rc = launch_det_dmma<14>(sys, sb, Wc, nmax, stream, &done); break;
            case 15: case 16: rc = launch_det_dmma<16>(sys, sb, Wc, nmax, stream, &done); break;
            default: break;
        }
        if (rc) return rc;
        if (done) return 0;
    }
    if (lap && !use_v1) {
        // det_lap_kernel: G directions per group, two per thread
        const int nb2 = nb * nb;
        int G = 2 * (256 / nb2 > 1 ? 256 / nb2 : 1);
        auto smem_of = [&](int g) { return (size_t)(np * np + np + g * np * np + g * nb) * sizeof(cplx); };
        while (G > 2 && smem_of(G) > 100 * 1024) G -= 2;
        const int nd_even = (sys.d.ND + 1) & ~1;
        if (G > nd_even) G = nd_even;
        const int items = (G / 2) * nb2;
        int threads = ((items > G ? items : G) + 31) & ~31;
        const size_t smem = smem_of(G);
        if (threads <= 512 && smem <= 226 * 1024) {
            static size_t cfg = 0;
            if (smem > cfg) {
                DS_CUDA_CHECK(cudaFuncSetAttribute(det_lap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                cfg = smem;
            }
            det_lap_kernel<<<grid, threads, smem, stream>>>(sys, sb, G);
            DS_CUDA_CHECK(cudaGetLastError());
            return 0;
        }
    }
    static const bool cta_lu = getenv("DS_DET_CTA_LU") && atoi(getenv("DS_DET_CTA_LU")) != 0;
    if (!lap && nmax <= 32 && !cta_lu) {        // value-only: warp-per-matrix LU
        const long long n_mats = (long long)Wc * ds_nblk(sys.d) * sys.d.D;
        const unsigned blocks = (unsigned)((n_mats + 3) / 4);      // 4 warps per CTA: ~170 registers per thread
        if (nmax <= 8) det_warp_kernel<8><<<blocks, 128, 0, stream>>>(sys, sb, n_mats);
        else if (nmax <= 16) det_warp_kernel<16><<<blocks, 128, 0, stream>>>(sys, sb, n_mats);
        else if (nmax <= 24) det_warp_kernel<24><<<blocks, 128, 0, stream>>>(sys, sb, n_mats);
        else if (nmax <= 28) det_warp_kernel<28><<<blocks, 128, 0, stream>>>(sys, sb, n_mats);
        else det_warp_kernel<32><<<blocks, 128, 0, stream>>>(sys, sb, n_mats);
        DS_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
    int G = 1;
    if (lap) {
        G = (DET_THREADS + nb * nb - 1) / (nb * nb);
        if (G < 1) G = 1;
        // keep the CTA under ~100 KB so two fit per SM when matrices are small
        while (G > 1 && (size_t)(np * np + np + 2 * G * np * np) * sizeof(cplx) > 100 * 1024) --G;
    }
    size_t smem = (size_t)(np * np + np + (lap ? 2 * G * np * np : 0)) * sizeof(cplx);
    DS_REQUIRE(smem <= 226 * 1024, "determinant kernel needs %zu bytes of shared memory", smem);
    if (smem > g_det_smem[lap ? 1 : 0]) {
        if (lap) DS_CUDA_CHECK(cudaFuncSetAttribute(det_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else DS_CUDA_CHECK(cudaFuncSetAttribute(det_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g_det_smem[lap ? 1 : 0] = smem;
    }
    if (lap) det_kernel<true><<<grid, DET_THREADS, smem, stream>>>(sys, sb, G);
    else det_kernel<false><<<grid, DET_THREADS, smem, stream>>>(sys, sb, G);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_det_inverse(const DsSys& sys, const SlaterBufs& sb, int Wc, cudaStream_t stream) {
    const int nmax = sys.d.full_det ? sys.d.N : (sys.d.n_up > sys.d.n_dn ? sys.d.n_up : sys.d.n_dn);
    DS_REQUIRE(nmax <= 128, "determinants larger than 128x128 are not supported (n_s=%d)", nmax);
    DS_REQUIRE(sb.XINV[0] && (sys.d.full_det || sb.XINV[1]), "det_inverse: output buffers missing");
    const int np = ((nmax + 2) / 3) * 3;
    const size_t smem = (size_t)(np * np + np) * sizeof(cplx);
    DS_REQUIRE(smem <= 226 * 1024, "determinant kernel needs %zu bytes of shared memory", smem);
    if (smem > g_det_smem[1]) {
        DS_CUDA_CHECK(cudaFuncSetAttribute(det_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g_det_smem[1] = smem;
    }
    dim3 grid((unsigned)((long long)Wc * ds_nblk(sys.d) * sys.d.D));
    det_kernel<true><<<grid, DET_THREADS, smem, stream>>>(sys, sb, 0);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int ds_launch_combine(const DsSys& sys, const SlaterBufs& sb, int Wc, bool lap, double* log_abs, double* phase,
                      double* ke_re, double* ke_im, cudaStream_t stream, double* gx_abs, double* gx_phase) {
    dim3 grid((unsigned)((Wc + 3) / 4));
    if (lap) combine_kernel<true><<<grid, 128, 0, stream>>>(sys, sb, Wc, log_abs, phase, ke_re, ke_im, gx_abs, gx_phase);
    else combine_kernel<false><<<grid, 128, 0, stream>>>(sys, sb, Wc, log_abs, phase, ke_re, ke_im, nullptr, nullptr);
    DS_CUDA_CHECK(cudaGetLastError());
    return 0;
}

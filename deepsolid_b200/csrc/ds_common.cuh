// Common definitions for the deepsolid_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#define DS_PI 3.14159265358979323846
#define DS_MAX_ATOMS_PRIM 16
#define DS_RAE_STRIDE 20   // per (electron, atom): jets (v, g0, g1, g2, lap) of the distance and of the 3 relative-vector components

// ---------------------------------------------------------------------------
// Static description of the system + network, passed by value to kernels.
// ---------------------------------------------------------------------------
struct DsLattice {
    double lat[9];      // rows = lattice vectors
    double inv[9];      // inverse of lat (x . inv = fractional coordinates)
    double AV[9];       // lattice / 2pi        (supercell.py:131-139)
    double BV[9];       // reciprocal vectors incl. 2pi
    double an2[3];      // |AV_l|^2
    double metric[9];   // AV_l . AV_m
};

struct DsDims {
    int n_up, n_dn, N;      // electrons
    int A;                  // atoms in the primitive cell (network features)
    int H, P, D, L;         // stream widths, determinants, layers
    int ND, NDp, NDg;       // 3N, padded to a multiple of 8, NDp+8 (rows of the shared-mean matrix)
    int F;                  // features per electron-atom / electron-electron pair: 4 ('nu') or 7 ('tri')
    int dist_type;          // 0 = nu_distance (network.py:189-224), 1 = tri_distance (network.py:227-246)
    int env_type;           // 0 = isotropic, 1 = diagonal, 2 = full envelope (network.py:335-364)
    int full_det;           // 1: every spin channel produces N orbitals and ONE N x N determinant per k is taken
                            //    (network.py:552-559); 0: one n_s x n_s determinant per spin channel
    int C0, K0;             // F*A, F*A + 2F (layer-0 one-electron inputs, own + pair-mean)
    int K1;                 // H + 2P  (own + pair-mean columns of layers >= 1)
    int use_last;           // use_last_layer (network.py:129-134, 528-533): L pair layers instead of L-1, and the orbital
                            // projection takes the symmetric features of the last layer (own | spin means | pair means)
};
// number of two-electron feature levels (level 0 = input features; level l = output of pair layer l-1)
__host__ __device__ __forceinline__ int ds_pair_levels(const DsDims& d) { return d.L + (d.use_last ? 1 : 0); }

// orbitals per spin channel, and the Slater matrix an electron row goes to: block index, its size, the row index
__host__ __device__ __forceinline__ int ds_norb(const DsDims& d, int s) { return d.full_det ? d.N : (s ? d.n_dn : d.n_up); }
__host__ __device__ __forceinline__ int ds_nblk(const DsDims& d) { return d.full_det ? 1 : 2; }
__host__ __device__ __forceinline__ int ds_blk_n(const DsDims& d, int blk) { return d.full_det ? d.N : (blk ? d.n_dn : d.n_up); }

struct DsSys {
    DsDims d;
    DsLattice prim, sim;
    double atoms[DS_MAX_ATOMS_PRIM * 3];
};

// ---------------------------------------------------------------------------
// 3-variable Laplacian jets: value, gradient, Laplacian w.r.t. one 3-vector.
// ---------------------------------------------------------------------------
struct Jet {
    double v, g0, g1, g2, l;
};

__host__ __device__ __forceinline__ Jet jet_const(double c) { return Jet{c, 0.0, 0.0, 0.0, 0.0}; }
__host__ __device__ __forceinline__ Jet jet_add(const Jet& a, const Jet& b) {
    return Jet{a.v + b.v, a.g0 + b.g0, a.g1 + b.g1, a.g2 + b.g2, a.l + b.l};
}
__host__ __device__ __forceinline__ Jet jet_scale(const Jet& a, double c) {
    return Jet{a.v * c, a.g0 * c, a.g1 * c, a.g2 * c, a.l * c};
}
__host__ __device__ __forceinline__ void jet_axpy(Jet& acc, double c, const Jet& a) {
    acc.v = fma(c, a.v, acc.v); acc.g0 = fma(c, a.g0, acc.g0); acc.g1 = fma(c, a.g1, acc.g1);
    acc.g2 = fma(c, a.g2, acc.g2); acc.l = fma(c, a.l, acc.l);
}
__host__ __device__ __forceinline__ Jet jet_mul(const Jet& a, const Jet& b) {
    Jet r;
    r.v = a.v * b.v;
    r.g0 = a.v * b.g0 + b.v * a.g0;
    r.g1 = a.v * b.g1 + b.v * a.g1;
    r.g2 = a.v * b.g2 + b.v * a.g2;
    r.l = a.v * b.l + b.v * a.l + 2.0 * (a.g0 * b.g0 + a.g1 * b.g1 + a.g2 * b.g2);
    return r;
}
// phi(u) given phi, phi', phi'' evaluated at u.v
__host__ __device__ __forceinline__ Jet jet_chain(const Jet& u, double f, double df, double d2f) {
    Jet r;
    r.v = f;
    r.g0 = df * u.g0; r.g1 = df * u.g1; r.g2 = df * u.g2;
    r.l = df * u.l + d2f * (u.g0 * u.g0 + u.g1 * u.g1 + u.g2 * u.g2);
    return r;
}
__host__ __device__ __forceinline__ Jet jet_tanh(const Jet& z) {
    double t = tanh(z.v);
    double d1 = 1.0 - t * t;
    return jet_chain(z, t, d1, -2.0 * t * d1);
}

__host__ __device__ __forceinline__ double ds_sign(double w) { return (w > 0.0) - (w < 0.0); }

// Envelope value jet of one (electron, atom, parameter) term exp(-|S.rel|) for the anisotropic envelopes
// (network.py:340-364): y_m = sum_k S[k][m] rel_k, r = |y|; `rel` = jets of the 3 relative-vector components.
// diag: S[k][m] = delta_km sigma_m.  Also returns y and r for the parameter gradient.
__host__ __device__ __forceinline__ Jet ds_aniso_env(const Jet rel[3], const double S[9], bool diag, double y_out[3], double* r_out) {
    Jet q = jet_const(0.0);
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        Jet y = jet_const(0.0);
        if (diag) {
            y = jet_scale(rel[m], S[m * 3 + m]);
        } else {
#pragma unroll
            for (int k = 0; k < 3; ++k) jet_axpy(y, S[k * 3 + m], rel[k]);
        }
        y_out[m] = y.v;
        q = jet_add(q, jet_mul(y, y));
    }
    const double r = sqrt(q.v);
    *r_out = r;
    const Jet rj = jet_chain(q, r, 0.5 / r, -0.25 / (r * r * r));
    const double ex = exp(-r);
    return jet_chain(rj, ex, -ex, ex);
}

// network.enforce_pbc (network.py:42-57): wrap a position into the cell.
__host__ __device__ __forceinline__ void ds_wrap(const DsLattice& L, const double x[3], double out[3]) {
    double f[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double s = x[0] * L.inv[0 * 3 + k] + x[1] * L.inv[1 * 3 + k] + x[2] * L.inv[2 * 3 + k];
        f[k] = s - floor(s);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = f[0] * L.lat[0 * 3 + k] + f[1] * L.lat[1 * 3 + k] + f[2] * L.lat[2 * 3 + k];
}

// network.nu_distance (network.py:189-224) for one displacement d, with the analytic
// first/second derivatives w.r.t. d.  Derivative conventions follow JAX:
// d|w|/dw = sign(w) (0 at 0), floor has zero derivative.
// out[0] = sd, out[1..3] = rel.  When JETS is false only .v is meaningful.
template <bool JETS>
__host__ __device__ __forceinline__ void ds_nu_distance(const DsLattice& L, const double d[3], Jet out[4]) {
    Jet fj[3], gj[3];
    const double ipi = 1.0 / DS_PI;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        double w = d[0] * L.BV[l * 3 + 0] + d[1] * L.BV[l * 3 + 1] + d[2] * L.BV[l * 3 + 2];
        double m = floor((w + DS_PI) / (2.0 * DS_PI));
        w = w - m * 2.0 * DS_PI;
        double a = fabs(w);
        double ap = a * ipi;               // |w/pi|
        double f = a * (1.0 - ap * ap * ap / 4.0);
        double g = w * (1.0 - 1.5 * ap + 0.5 * ap * ap);
        if (JETS) {
            Jet wj{w, L.BV[l * 3 + 0], L.BV[l * 3 + 1], L.BV[l * 3 + 2], 0.0};
            double s = ds_sign(w);
            double df = s * (1.0 - ap * ap * ap);
            double d2f = -3.0 * ap * ap * ipi;
            double dg = 1.0 - 3.0 * ap + 1.5 * ap * ap;
            double d2g = -3.0 * s * ipi + 3.0 * w * ipi * ipi;
            fj[l] = jet_chain(wj, f, df, d2f);
            gj[l] = jet_chain(wj, g, dg, d2g);
        } else {
            fj[l] = jet_const(f);
            gj[l] = jet_const(g);
        }
    }
    Jet sd2 = jet_const(0.0);
    Jet rel[3] = {jet_const(0.0), jet_const(0.0), jet_const(0.0)};
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        if (JETS) {
            jet_axpy(sd2, L.an2[l], jet_mul(fj[l], fj[l]));
#pragma unroll
            for (int m = 0; m < 3; ++m)
                if (m != l) jet_axpy(sd2, L.metric[l * 3 + m], jet_mul(gj[l], gj[m]));
        } else {
            sd2.v += L.an2[l] * fj[l].v * fj[l].v;
#pragma unroll
            for (int m = 0; m < 3; ++m)
                if (m != l) sd2.v += L.metric[l * 3 + m] * gj[l].v * gj[m].v;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (JETS) jet_axpy(rel[j], L.AV[l * 3 + j], gj[l]);
            else rel[j].v += L.AV[l * 3 + j] * gj[l].v;
        }
    }
    double sd = sqrt(sd2.v);
    if (JETS) out[0] = jet_chain(sd2, sd, 0.5 / sd, -0.25 / (sd * sd * sd));
    else out[0] = jet_const(sd);
    out[1] = rel[0]; out[2] = rel[1]; out[3] = rel[2];
}

// network.tri_distance (network.py:227-246): w = d.BV^T, rel = [sum_l sin(w_l) AV_l, sum_l cos(w_l) AV_l] (6),
// sd^2 = sum_{lm} (AV_l.AV_m) [ (1-cos w_l)(1-cos w_m) + sin w_l sin w_m ].  out[0] = sd, out[1..6] = rel.
template <bool JETS>
__host__ __device__ __forceinline__ void ds_tri_distance(const DsLattice& L, const double d[3], Jet out[7]) {
    Jet sj[3], cj[3];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const double w = d[0] * L.BV[l * 3 + 0] + d[1] * L.BV[l * 3 + 1] + d[2] * L.BV[l * 3 + 2];
        double sn, cs;
        sincos(w, &sn, &cs);
        if (JETS) {
            Jet wj{w, L.BV[l * 3 + 0], L.BV[l * 3 + 1], L.BV[l * 3 + 2], 0.0};
            sj[l] = jet_chain(wj, sn, cs, -sn);
            cj[l] = jet_chain(wj, cs, -sn, -cs);
        } else {
            sj[l] = jet_const(sn);
            cj[l] = jet_const(cs);
        }
    }
    Jet sd2 = jet_const(0.0);
#pragma unroll
    for (int j = 0; j < 6; ++j) out[1 + j] = jet_const(0.0);
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        Jet oml = jet_const(1.0);
        jet_axpy(oml, -1.0, cj[l]);                                  // 1 - cos w_l
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            Jet omm = jet_const(1.0);
            jet_axpy(omm, -1.0, cj[m]);
            if (JETS) {
                jet_axpy(sd2, L.metric[l * 3 + m], jet_mul(oml, omm));
                jet_axpy(sd2, L.metric[l * 3 + m], jet_mul(sj[l], sj[m]));
            } else {
                sd2.v += L.metric[l * 3 + m] * (oml.v * omm.v + sj[l].v * sj[m].v);
            }
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (JETS) { jet_axpy(out[1 + j], L.AV[l * 3 + j], sj[l]); jet_axpy(out[4 + j], L.AV[l * 3 + j], cj[l]); }
            else { out[1 + j].v += L.AV[l * 3 + j] * sj[l].v; out[4 + j].v += L.AV[l * 3 + j] * cj[l].v; }
        }
    }
    const double sd = sqrt(sd2.v);
    if (JETS) out[0] = jet_chain(sd2, sd, 0.5 / sd, -0.25 / (sd * sd * sd));
    else out[0] = jet_const(sd);
}

// distance features of one displacement: returns nothing, fills out[0..F-1] (F = 4 or 7)
template <bool JETS>
__host__ __device__ __forceinline__ void ds_distance(int dist_type, const DsLattice& L, const double d[3], Jet out[7]) {
    if (dist_type == 1) ds_tri_distance<JETS>(L, d, out);
    else ds_nu_distance<JETS>(L, d, out);
}

// ---------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------
void ds_set_error(const char* fmt, ...);

#define DS_CUDA_CHECK(expr)                                                               \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            ds_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,    \
                         __LINE__, cudaGetErrorString(_e));                               \
            return -2;                                                                    \
        }                                                                                 \
    } while (0)

#define DS_REQUIRE(cond, ...)                                                             \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            ds_set_error(__VA_ARGS__);                                                    \
            return -1;                                                                    \
        }                                                                                 \
    } while (0)

struct __align__(16) cplx {
    double re, im;
};
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) {
    return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return cplx{a.re * s, a.im * s}; }
__host__ __device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
    acc.re = fma(a.re, b.re, acc.re); acc.re = fma(-a.im, b.im, acc.re);
    acc.im = fma(a.re, b.im, acc.im); acc.im = fma(a.im, b.re, acc.im);
}
__host__ __device__ __forceinline__ cplx cinv(cplx a) {
    // Smith's algorithm
    if (fabs(a.re) >= fabs(a.im)) {
        double r = a.im / a.re, den = a.re + a.im * r;
        return cplx{1.0 / den, -r / den};
    } else {
        double r = a.re / a.im, den = a.re * r + a.im;
        return cplx{r / den, -1.0 / den};
    }
}

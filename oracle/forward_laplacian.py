"""Second, independent CPU derivation of the kinetic energy: the forward-Laplacian
recursion the CUDA kernels implement, written in batched torch fp64.

TEST INFRASTRUCTURE (see oracle/deepsolid_oracle.py header).  It serves two
purposes: (1) it cross-validates the autodiff oracle (two derivations of the
same number agreeing to ~1e-11), (2) it exposes every intermediate the CUDA path
materialises (feature jets, per-layer value/Jacobian/Laplacian, orbital
derivatives, determinant traces) so a failing GPU parity test can be localised
stage by stage.

A "jet" of a quantity q(r), r in R^3, is (q, dq/dr_c [3], sum_c d2q/dr_c^2).
Default network options only: distance_type='nu', isotropic envelope,
full_det=False, use_last_layer=False, bias_orbitals=False.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

DT = torch.float64
PI = math.pi

#: optional hook ``f(A, W) -> A @ W`` applied to the two Jacobian-row contractions (one-electron
#: stream and orbital projection).  Used by scripts/ozaki_study.py and the tests of the int8
#: slice arithmetic to emulate the truncation of the tcgen05 path on the CPU.
JAC_MATMUL = None


def _jmm(A, W):
    return A @ W if JAC_MATMUL is None else JAC_MATMUL(A, W)


class Jet:
    """value v[...], gradient g[...,3], laplacian l[...] w.r.t. one 3-vector."""
    __slots__ = ("v", "g", "l")

    def __init__(self, v, g, l):
        self.v, self.g, self.l = v, g, l

    def __add__(self, o):
        if isinstance(o, Jet):
            return Jet(self.v + o.v, self.g + o.g, self.l + o.l)
        return Jet(self.v + o, self.g, self.l)

    def scale(self, c):
        return Jet(self.v * c, self.g * (c[..., None] if isinstance(c, torch.Tensor) and c.dim() else c), self.l * c)

    def __mul__(self, o):
        if isinstance(o, Jet):
            return Jet(self.v * o.v, self.v[..., None] * o.g + o.v[..., None] * self.g,
                       self.v * o.l + o.v * self.l + 2 * (self.g * o.g).sum(-1))
        return self.scale(o)

    def apply(self, f, df, d2f):
        """phi(u): value f, first derivative df, second derivative d2f (all evaluated at u)."""
        return Jet(f, df[..., None] * self.g, df * self.l + d2f * (self.g ** 2).sum(-1))


def _nu_distance_jet(d, AV, BV):
    """Jets of (sd, rel[3]) for displacement d[...,3]; network.py:189-224 differentiated by hand."""
    w = d @ BV.T                                    # (...,3)
    w = w - torch.floor((w + PI) / (2 * PI)) * 2 * PI
    shp = w.shape[:-1]
    zeros = torch.zeros(shp, dtype=DT)
    sd2 = Jet(zeros.clone(), torch.zeros(shp + (3,), dtype=DT), zeros.clone())
    rel = [Jet(zeros.clone(), torch.zeros(shp + (3,), dtype=DT), zeros.clone()) for _ in range(3)]
    an2 = (AV ** 2).sum(-1)
    metric = AV @ AV.T
    fj, gj = [], []
    for l in range(3):
        wl = Jet(w[..., l], BV[l].expand(shp + (3,)), zeros)
        a = wl.v.abs()
        s = torch.sign(wl.v)
        f = a * (1 - (a / PI) ** 3 / 4)
        df = s * (1 - a ** 3 / PI ** 3)
        d2f = -3 * a ** 2 / PI ** 3
        g = wl.v * (1 - 1.5 * a / PI + 0.5 * (a / PI) ** 2)
        dg = 1 - 3 * a / PI + 1.5 * a ** 2 / PI ** 2
        d2g = -3 * s / PI + 3 * wl.v / PI ** 2
        fj.append(wl.apply(f, df, d2f))
        gj.append(wl.apply(g, dg, d2g))
    for l in range(3):
        sd2 = sd2 + (fj[l] * fj[l]).scale(an2[l])
        for m in range(3):
            if m != l:
                sd2 = sd2 + (gj[l] * gj[m]).scale(metric[l, m])
        for j in range(3):
            rel[j] = rel[j] + gj[l].scale(AV[l, j])
    sd = torch.sqrt(sd2.v)
    sdj = sd2.apply(sd, 0.5 / sd, -0.25 / sd ** 3)
    return sdj, rel


def _wrap(x, lat):
    frac = x @ torch.linalg.inv(lat)
    return (frac - torch.floor(frac)) @ lat


def _stack_jets(js):
    return Jet(torch.stack([j.v for j in js], -1), torch.stack([j.g for j in js], -1),
               torch.stack([j.l for j in js], -1))          # g: (...,3,C)


def kinetic_forward_laplacian(params: Dict, X: torch.Tensor, sim_cell, klist, want=False):
    """Returns (log|psi| [B], phase angle [B], kinetic complex [B], intermediates dict)."""
    from .geometry import rederive
    sim_cell = rederive(sim_cell)
    prim = sim_cell.original_cell
    nu, nd = sim_cell.nelec
    N = nu + nd
    B = X.shape[0]
    A = prim.natm
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=DT)
    atoms = t(prim.atom_coords())
    x = X.reshape(B, N, 3)
    spin_of = torch.tensor([0] * nu + [1] * nd)
    nsp = torch.tensor([nu, nd], dtype=DT)
    inter = {}

    # ---- features ------------------------------------------------------
    px = _wrap(x, t(prim.a))
    sd_ae, rel_ae = _nu_distance_jet(px[:, :, None, :] - atoms, t(prim.AV), t(prim.BV))   # (B,N,A)
    # h0[i] = per atom [r, rel0, rel1, rel2]  -> (B,N,4A)
    feat = _stack_jets([sd_ae] + rel_ae)            # v (B,N,A,4) g (B,N,A,3,4)
    h_v = feat.v.reshape(B, N, 4 * A)
    h_g = feat.g.permute(0, 1, 3, 2, 4).reshape(B, N, 3, 4 * A)   # own-electron gradient only
    h_l = feat.l.reshape(B, N, 4 * A)
    sx = _wrap(x, t(sim_cell.a))
    dee = sx[:, :, None, :] - sx[:, None, :, :]
    eye = torch.eye(N, dtype=DT)
    sd_ee, rel_ee = _nu_distance_jet(dee + eye[None, :, :, None], t(sim_cell.AV), t(sim_cell.BV))
    p = _stack_jets([sd_ee] + rel_ee)               # pair jets w.r.t. r = x_i - x_j; v (B,N,N,4)
    off = (1 - eye)[None, :, :, None]
    p = Jet(p.v * off, p.g * off[..., None, :], p.l * off)
    r_ae = sd_ae                                    # for the envelope
    inter["ae_v"], inter["ee_v"] = h_v, p.v

    # ---- pair stream (no mixing: h2[i,j] is a function of x_i - x_j only) ----
    pair = [p]
    for li, layer in enumerate(params["double"]):
        W, b = layer["w"], layer["b"]
        cur = pair[-1]
        z = Jet(cur.v @ W + b, cur.g @ W, cur.l @ W)
        th = torch.tanh(z.v)
        tj = Jet(th, (1 - th ** 2)[..., None, :] * z.g,
                 (1 - th ** 2) * z.l - 2 * th * (1 - th ** 2) * (z.g ** 2).sum(-2))
        if cur.v.shape == tj.v.shape:
            s = 1 / math.sqrt(2.0)
            tj = Jet((cur.v + tj.v) * s, (cur.g + tj.g) * s, (cur.l + tj.l) * s)
        pair.append(tj)

    # ---- one-electron stream: value hv (B,N,C), Jacobian J (B,N,3N,C), Laplacian hl (B,N,C)
    C0 = 4 * A
    J = torch.zeros(B, N, N, 3, C0, dtype=DT)
    idx = torch.arange(N)
    J[:, idx, idx] = h_g                            # d h0_i / d x_i
    J = J.reshape(B, N, 3 * N, C0)
    hv, hl = h_v, h_l
    up = (spin_of == 0)
    masks = [up, ~up]

    def pair_means(pj):
        """m^s_i = mean_{j in s} h2[j,i]: value (B,N,2,P), Jacobian (B,N,3N,2,P), Laplacian (B,N,2,P)."""
        P = pj.v.shape[-1]
        mv = torch.stack([pj.v[:, m].sum(1) / n for m, n in zip(masks, nsp)], 2)        # (B,N,2,P)
        ml = torch.stack([2 * pj.l[:, m].sum(1) / n for m, n in zip(masks, nsp)], 2)
        mJ = torch.zeros(B, N, N, 3, 2, P, dtype=DT)        # [i, k, c, s]
        # k != i : d h2[k,i] / d x_k = +g(F[k,i]) ; only the spin of k
        g = pj.g                                              # (B, j, i, 3, P)
        for s, (m, n) in enumerate(zip(masks, nsp)):
            ks = torch.nonzero(m).flatten()
            mJs = mJ[:, :, :, :, s, :]                        # view (B,i,k,3,P)
            mJs[:, :, ks] = (g[:, ks] / n).permute(0, 2, 1, 3, 4)
            # k == i : - (1/n) sum_{j in s} g(F[j,i])
            mJs[:, idx, idx] = -(g[:, ks].sum(1) / n)
        return mv, mJ.reshape(B, N, 3 * N, 2, P), ml

    n_single = len(params["single"])
    for li in range(n_single):
        W, b = params["single"][li]["w"], params["single"][li]["b"]
        C = hv.shape[-1]
        pj = pair[li]
        P = pj.v.shape[-1]
        Wown, Wg, Wm = W[:C], W[C:3 * C].reshape(2, C, -1), W[3 * C:].reshape(2, P, -1)
        mv, mJ, ml = pair_means(pj)
        gv = torch.stack([hv[:, m].mean(1) for m in masks], 1)           # (B,2,C)
        gJ = torch.stack([J[:, m].mean(1) for m in masks], 2)            # (B,3N,2,C)
        gl = torch.stack([hl[:, m].mean(1) for m in masks], 1)
        zv = hv @ Wown + torch.einsum("bsc,sco->bo", gv, Wg)[:, None] + torch.einsum("bisp,spo->bio", mv, Wm) + b
        zJ = (_jmm(J, Wown) + torch.einsum("bdsc,sco->bdo", gJ, Wg)[:, None]
              + torch.einsum("bidsp,spo->bido", mJ, Wm))
        zl = hl @ Wown + torch.einsum("bsc,sco->bo", gl, Wg)[:, None] + torch.einsum("bisp,spo->bio", ml, Wm)
        th = torch.tanh(zv)
        d1 = 1 - th ** 2
        S = (zJ ** 2).sum(2)
        tv, tJ, tl = th, d1[:, :, None, :] * zJ, d1 * zl - 2 * th * d1 * S
        if tv.shape == hv.shape:
            s = 1 / math.sqrt(2.0)
            tv, tJ, tl = (hv + tv) * s, (J + tJ) * s, (hl + tl) * s
        hv, J, hl = tv, tJ, tl
        if want:
            inter[f"h{li + 1}_v"], inter[f"h{li + 1}_J"], inter[f"h{li + 1}_l"] = hv, J, hl

    # ---- orbitals, envelope, Bloch phase, determinants ------------------------
    D = params["orbital"][0]["w"].shape[1] // 2 // nu if nu > 0 else None
    lse_terms = []          # per spin: logdet (B,D) complex-phase & log-abs
    tr_lap = torch.zeros(B, 0, dtype=torch.complex128)
    taus, trsq, trlap, signs, logabs = [], [], [], [], []
    o = 0
    ch = 0
    for s, ns in enumerate((nu, nd)):
        if ns == 0:
            continue
        W = params["orbital"][ch]["w"]
        env = params["envelope"][ch]
        D = W.shape[1] // 2 // ns
        npar = ns * D
        sl = slice(o, o + ns)
        Yv = hv[:, sl] @ W                      # (B,ns,2npar)
        YJ = _jmm(J[:, sl], W)                       # (B,ns,3N,2npar)
        Yl = hl[:, sl] @ W
        cx = lambda y: torch.complex(y[..., :npar], y[..., npar:])
        Ov, OJ, Ol = cx(Yv), cx(YJ), cx(Yl)
        # envelope jets w.r.t. own electron
        r = Jet(r_ae.v[:, sl], r_ae.g[:, sl], r_ae.l[:, sl])            # (B,ns,A), g (B,ns,A,3)
        sig, pi_ = env["sigma"], env["pi"]                              # (A,npar)
        e = torch.exp(-torch.abs(r.v[..., None] * sig))                 # (B,ns,A,npar)
        asig = torch.abs(sig)
        env_v = (e * pi_).sum(2)
        env_g = ((-asig * e * pi_)[:, :, :, None, :] * r.g[..., None]).sum(2)      # (B,ns,3,npar)
        env_l = ((-asig * e * pi_) * r.l[..., None] + (sig ** 2 * e * pi_) * (r.g ** 2).sum(-1)[..., None]).sum(2)
        k = torch.as_tensor(np.asarray(klist[s]), dtype=DT)              # (ns orb,3)
        xs = x[:, sl]
        ph = torch.exp(1j * (xs @ k.T))                                  # (B,ns elec, ns orb)
        ph_g = 1j * k.T[None, None] * ph[:, :, None, :]                  # (B,ns,3,orb)
        ph_l = -(k ** 2).sum(-1)[None, None] * ph
        # broadcast phase over determinants: p = kdet*ns + orb
        tile = lambda a: a.unsqueeze(-2).expand(*a.shape[:-1], D, ns).reshape(*a.shape[:-1], npar)
        phv, phg, phl = tile(ph), tile(ph_g), tile(ph_l)
        Ev = env_v * phv
        Eg = env_g * phv[:, :, None, :] + env_v[:, :, None, :] * phg
        El = env_l * phv + env_v * phl + 2 * (env_g * phg).sum(2)
        Mv = Ov * Ev                                                    # (B,ns,npar)
        MJ = OJ * Ev[:, :, None, :]                                     # (B,ns,3N,npar)
        own = torch.arange(ns) + o
        MJv = MJ.reshape(B, ns, N, 3, npar)
        MJv[:, torch.arange(ns), own] += Ov[:, :, None, :] * Eg
        OJown = OJ.reshape(B, ns, N, 3, npar)[:, torch.arange(ns), own]          # (B,ns,3,npar)
        Ml = Ol * Ev + 2 * (OJown * Eg).sum(2) + Ov * El
        # to matrices (B,D,elec,orb)
        mat = lambda a: a.reshape(B, ns, D, ns).permute(0, 2, 1, 3)
        Amat = mat(Mv)
        dA = MJ.reshape(B, ns, 3 * N, D, ns).permute(0, 2, 3, 1, 4)      # (B,3N,D,elec,orb)
        lA = mat(Ml)
        Xinv = torch.linalg.inv(Amat)                                    # (B,D,orb,elec)
        sg, la = torch.linalg.slogdet(Amat)
        Yd = Xinv[:, None] @ dA                                          # (B,3N,D,orb,orb)
        tau = torch.diagonal(Yd, dim1=-2, dim2=-1).sum(-1)               # (B,3N,D)
        tsq = (Yd * Yd.transpose(-1, -2)).sum((-1, -2))                  # tr(Yd Yd)  (B,3N,D)
        tl_ = (Xinv * lA.transpose(-1, -2)).sum((-1, -2))                # tr(X lapA) (B,D)
        taus.append(tau); trsq.append(tsq.sum(1)); trlap.append(tl_); signs.append(sg); logabs.append(la)
        if want:
            inter[f"orb{s}"] = Amat
            inter[f"dorb{s}"] = dA
            inter[f"lorb{s}"] = lA
        o += ns
        ch += 1
    sign = signs[0]
    la = logabs[0]
    for sg2, la2 in zip(signs[1:], logabs[1:]):
        sign, la = sign * sg2, la + la2
    mx = la.max(dim=1, keepdim=True).values
    det = sign * torch.exp(la - mx)
    tot = det.sum(1)
    wk = det / tot[:, None]                                              # (B,D) complex weights
    logabs_psi = torch.log(tot.abs()) + mx[:, 0]
    angle = torch.angle(tot)
    tau_tot = sum(taus)                                                  # (B,3N,D)
    per_det = sum(trlap) - sum(trsq) + (tau_tot ** 2).sum(1)             # (B,D)
    kinetic = -0.5 * (wk * per_det).sum(1)
    if want:
        inter["weights"] = wk
        inter["grad_logpsi"] = (wk[:, None, :] * tau_tot).sum(-1)
    return logabs_psi, angle, kinetic, inter

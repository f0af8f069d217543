"""TEST INFRASTRUCTURE (never imported by the product).  A third, scalar restatement of the reference network in
40-digit mpmath arithmetic, with the Laplacian taken by high-order central differences instead of autodiff: it pins the
floating-point behaviour of oracle/deepsolid_oracle.py (torch fp64 + autograd) on a tiny system, as SURVEY.md §8(c)
asks ("high-precision (mpmath, 50 digits) evaluation of a tiny system").  Plain Python loops, one walker at a time.

Follows /root/reference/DeepSolid: network.py:42-57 (enforce_pbc), :189-224 (nu_distance), :249-302 (features),
:305-332 (symmetric features), :335-337 (isotropic envelope), :375-427 (slogdet, logdet_matmul), :449-458 (phase),
:461-560 (orbitals), hamiltonian.py:45-70 (kinetic energy from first and second derivatives of log psi).
"""
import mpmath as mp

mp.mp.dps = 40


def _m(a):
    return [[mp.mpf(float(v)) for v in row] for row in a]


def _inv3(a):
    return (mp.matrix(a) ** -1).tolist()


def enforce_pbc(latvec, x):                      # network.py:42-57
    inv = _inv3(latvec)
    frac = [sum(x[k] * inv[k][j] for k in range(3)) for j in range(3)]
    frac = [f - mp.floor(f) for f in frac]
    return [sum(frac[k] * latvec[k][j] for k in range(3)) for j in range(3)]


def scaled_f(w):                                 # network.py:189-195
    return abs(w) * (1 - abs(w / mp.pi) ** 3 / 4)


def scaled_g(w):                                 # network.py:198-204
    return w * (1 - mp.mpf(3) / 2 * abs(w / mp.pi) + mp.mpf(1) / 2 * abs(w / mp.pi) ** 2)


def nu_distance(d, a, b):                        # network.py:207-224
    w = [sum(d[k] * b[l][k] for k in range(3)) for l in range(3)]
    w = [wl - mp.floor((wl + mp.pi) / (2 * mp.pi)) * 2 * mp.pi for wl in w]
    norm = [mp.sqrt(sum(a[l][k] ** 2 for k in range(3))) for l in range(3)]
    r1 = sum((norm[l] * scaled_f(w[l])) ** 2 for l in range(3))
    sg = [scaled_g(wl) for wl in w]
    rel = [sum(sg[l] * a[l][j] for l in range(3)) for j in range(3)]
    r2 = sum(sum(a[l][k] * a[m][k] for k in range(3)) * sg[l] * sg[m] for l in range(3) for m in range(3) if l != m)
    return mp.sqrt(r1 + r2), rel


def tri_distance(d, a, b):                       # network.py:227-246
    w = [sum(d[k] * b[l][k] for k in range(3)) for l in range(3)]
    sg, cg = [mp.sin(wl) for wl in w], [mp.cos(wl) for wl in w]
    rel = [sum(sg[l] * a[l][j] for l in range(3)) for j in range(3)] + [sum(cg[l] * a[l][j] for l in range(3)) for j in range(3)]
    sd2 = sum(sum(a[l][k] * a[m][k] for k in range(3)) * ((1 - cg[l]) * (1 - cg[m]) + sg[l] * sg[m])
              for l in range(3) for m in range(3))
    return mp.sqrt(sd2), rel


def log_psi(params, x, cell, klist, spins, distance_type="nu", envelope_type="isotropic", full_det=False):
    """Complex log psi (method eval_logdet) of one walker x (flat list of 3N mpf)."""
    dist = nu_distance if distance_type == "nu" else tri_distance
    nf = 4 if distance_type == "nu" else 7
    from .geometry import rederive
    cell = rederive(cell)
    prim = cell.original_cell
    n_e = sum(spins)
    pa, pAV, pBV = _m(prim.a), _m(prim.AV), _m(prim.BV)
    sa, sAV, sBV = _m(cell.a), _m(cell.AV), _m(cell.BV)
    atoms = _m(prim.atom_coords())
    pos = [x[3 * i:3 * i + 3] for i in range(n_e)]
    ppos = [enforce_pbc(pa, p) for p in pos]
    spos = [enforce_pbc(sa, p) for p in pos]
    r_ae, ae_vec, h_one = [], [], []
    for i in range(n_e):                         # network.py:281-288, 485-487
        row, rr, vv = [], [], []
        for at in atoms:
            sd, rel = dist([ppos[i][k] - at[k] for k in range(3)], pAV, pBV)
            row += [sd] + rel
            rr.append(sd)
            vv.append(rel)
        h_one.append(row)
        r_ae.append(rr)
        ae_vec.append(vv)
    h_two = []
    for i in range(n_e):                         # network.py:290-300: h_two[i][j] from x_i - x_j, masked diagonal
        row = []
        for j in range(n_e):
            if i == j:
                row.append([mp.mpf(0)] * nf)
            else:
                sd, rel = dist([spos[i][k] - spos[j][k] for k in range(3)], sAV, sBV)
                row.append([sd] + rel)
        h_two.append(row)

    def sym(h1, h2):                             # network.py:305-332
        out = []
        n0 = spins[0]
        blocks = [range(0, n0), range(n0, n_e)]
        g1 = [[sum(h1[j][c] for j in blk) / len(blk) for c in range(len(h1[0]))] for blk in blocks if len(blk)]
        for i in range(n_e):
            g2 = [[sum(h2[j][i][c] for j in blk) / len(blk) for c in range(len(h2[0][0]))] for blk in blocks if len(blk)]
            out.append(h1[i] + sum(g1, []) + sum(g2, []))
        return out

    def dense(v, w, b):
        return [sum(v[k] * w[k][n] for k in range(len(v))) + (b[n] if b is not None else 0) for n in range(len(w[0]))]

    single = [(_m(p["w"]), [mp.mpf(float(v)) for v in p["b"]]) for p in params["single"]]
    double = [(_m(p["w"]), [mp.mpf(float(v)) for v in p["b"]]) for p in params["double"]]
    rs2 = mp.sqrt(2)
    for l in range(len(single)):                 # network.py:517-533
        inp = sym(h_one, h_two)
        nxt = [[mp.tanh(v) for v in dense(row, *single[l])] for row in inp]
        h_one = [[(a + b) / rs2 for a, b in zip(o, n)] if len(o) == len(n) else n for o, n in zip(h_one, nxt)]
        if l < len(double):
            nxt2 = [[[mp.tanh(v) for v in dense(h_two[i][j], *double[l])] for j in range(n_e)] for i in range(n_e)]
            h_two = [[[(a + b) / rs2 for a, b in zip(h_two[i][j], nxt2[i][j])] if len(h_two[i][j]) == len(nxt2[i][j])
                      else nxt2[i][j] for j in range(n_e)] for i in range(n_e)]
    # orbitals: network.py:536-557
    import numpy as _np
    off = 0
    logs = []
    n_orb_all = n_e if full_det else None
    kall = _m(_np.concatenate([_np.asarray(k).reshape(-1, 3) for k in klist], axis=0))
    rows_full = None
    for s, ns in enumerate(spins):
        if ns == 0:
            continue
        w = _m(params["orbital"][s]["w"])
        npar = len(w[0]) // 2
        pi = _m(params["envelope"][s]["pi"])
        sig_np = _np.asarray(params["envelope"][s]["sigma"], dtype=float)
        kl = kall if full_det else _m(klist[s])
        n_orb = n_orb_all if full_det else ns
        D = npar // n_orb
        mats = [[[None] * n_orb for _ in range(ns)] for _ in range(D)]
        for i in range(ns):
            e = off + i
            y = dense(h_one[e], w, None)
            for p in range(npar):
                env = mp.mpf(0)
                for a in range(len(atoms)):
                    if envelope_type == "isotropic":          # network.py:335-337
                        r = abs(r_ae[e][a] * mp.mpf(float(sig_np[a][p])))
                    elif envelope_type == "diagonal":         # network.py:340-343 (needs the 3-component relative vector)
                        r = mp.sqrt(sum((ae_vec[e][a][c] * mp.mpf(float(sig_np[a][c][p]))) ** 2 for c in range(3)))
                    else:                                     # network.py:349-364: r_m = sum_k ae_k sigma[k, m, a, p]
                        r = mp.sqrt(sum(sum(ae_vec[e][a][k] * mp.mpf(float(sig_np[k][m][a][p])) for k in range(3)) ** 2
                                        for m in range(3)))
                    env += mp.exp(-r) * pi[a][p]
                k, o = divmod(p, n_orb)
                phase = mp.expj(sum(kl[o][c] * pos[e][c] for c in range(3)))      # unwrapped x, network.py:449-458
                mats[k][i][o] = mp.mpc(y[p], y[npar + p]) * env * phase
        if full_det:                                          # network.py:552-559: rows of both spins, one N x N matrix
            rows_full = mats if rows_full is None else [a + b for a, b in zip(rows_full, mats)]
        else:
            logs.append([mp.det(mp.matrix(m)) for m in mats])
        off += ns
    if full_det:
        logs.append([mp.det(mp.matrix(m)) for m in rows_full])
    total = mp.mpc(0)
    for k in range(len(logs[0])):                # network.py:395-427 (uniform weights)
        prod = mp.mpc(1)
        for dl in logs:
            prod *= dl[k]
        total += prod
    return mp.log(total)


def kinetic(params, x, cell, klist, spins, h=mp.mpf("1e-8"), **net):
    """hamiltonian.py:45-70 with derivatives from 4th-order central differences in 40-digit arithmetic:
    KE = -1/2 sum_d [ d2 f + (d f)^2 ],  f = complex log psi."""
    x = [mp.mpf(float(v)) for v in x]
    f0 = log_psi(params, x, cell, klist, spins, **net)
    ke = mp.mpc(0)
    for d in range(len(x)):
        def at(t):
            y = list(x)
            y[d] += t
            return log_psi(params, y, cell, klist, spins, **net)
        fp1, fm1, fp2, fm2 = at(h), at(-h), at(2 * h), at(-2 * h)
        d1 = (-fp2 + 8 * fp1 - 8 * fm1 + fm2) / (12 * h)
        d2 = (-fp2 + 16 * fp1 - 30 * f0 + 16 * fm1 - fm2) / (12 * h * h)
        ke += d2 + d1 * d1
    return f0, -ke / 2

"""CPU oracle: torch-fp64 restatement of the reference's local-energy hot path.

THIS IS TEST INFRASTRUCTURE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product (``deepsolid_b200``) never does.

It stands alone: cell geometry (supercell atoms, AV / BV, k-points) comes from oracle/geometry.py, the Ewald
tables and the lattice classification from the classes below -- nothing is imported from the product, and a
cell object handed in by a test is re-derived from its primary inputs (geometry.rederive).

PINNING: the reference (bytedance/DeepSolid @ 812a2b8) is pure JAX + pyscf, neither is installable in this image,
and its own tests hold no golden numbers (test/test_network.py asserts three invariants only).  The oracle is a
line-by-line restatement pinned by (i) the reference's OWN SOURCE FILES executed on a torch stand-in for jax / pyscf
(tests/golden/torch_jax_shim.py; fixtures tests/golden/reference_shim_*.npz written by make_reference_golden.py
--backend shim; compared in tests/test_reference_golden.py: LiH test cell with every network option, and the
BASELINE configurations at full size), (ii) those three invariants, (iii) finite differences and a 40-digit mpmath
restatement of its own log psi, (iv) Madelung constants for the Ewald setup and (v) an independent forward-Laplacian
derivation (oracle/forward_laplacian.py).  NOT pinned: a run of the real JAX stack (XLA's evaluation order, JAX's
RNG stream); make_reference_golden.py --backend reference produces that file on a machine that has it.

Every function cites the reference lines (relative to /root/reference/DeepSolid/)
it follows.  The Laplacian is obtained with the *reference's algorithm*
(forward-over-reverse: ``jvp`` of ``grad``, hamiltonian.py:45-70 and :127-159),
via ``torch.func`` -- deliberately not the forward-Laplacian recursion the CUDA
kernels use.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np
import torch
from torch.func import grad, jvp, vmap

from .geometry import rederive

DT = torch.float64
CT = torch.complex128


def _t(x, dtype=DT):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


# ---------------------------------------------------------------------------
# network.py
# ---------------------------------------------------------------------------

def enforce_pbc(latvec: torch.Tensor, epos: torch.Tensor):
    """network.py:42-57."""
    recpvecs = torch.linalg.inv(latvec)
    frac = epos @ recpvecs
    wrap = torch.floor(frac)                # `// 1`; zero derivative like JAX
    return (frac - wrap) @ latvec, wrap


def scaled_f(w):
    """network.py:189-195."""
    return torch.abs(w) * (1 - torch.abs(w / math.pi) ** 3 / 4.0)


def scaled_g(w):
    """network.py:198-204."""
    return w * (1 - 3.0 / 2.0 * torch.abs(w / math.pi) + 1.0 / 2.0 * torch.abs(w / math.pi) ** 2)


def nu_distance(xea, a, b):
    """network.py:207-224."""
    w = torch.einsum("...ijk,lk->...ijl", xea, b)
    mod = torch.floor((w + math.pi) / (2 * math.pi))
    w = w - mod * 2 * math.pi
    r1 = (torch.linalg.norm(a, dim=-1) * scaled_f(w)) ** 2
    sg = scaled_g(w)
    rel = torch.einsum("...i,ij->...j", sg, a)
    r2 = torch.einsum("ij,kj->ik", a, a) * (sg[..., :, None] * sg[..., None, :])
    n = r2.shape[-1]
    off = torch.ones(n, n, dtype=r2.dtype) - torch.eye(n, dtype=r2.dtype)
    result = torch.sum(r1, dim=-1) + torch.sum(r2 * off, dim=(-1, -2))
    sd = result ** 0.5
    return sd, rel


def tri_distance(xea, a, b):
    """network.py:227-246."""
    w = torch.einsum("...ijk,lk->...ijl", xea, b)
    sg, cg = torch.sin(w), torch.cos(w)
    rel = torch.cat([torch.einsum("...i,ij->...j", sg, a),
                     torch.einsum("...i,ij->...j", cg, a)], dim=-1)
    metric = torch.einsum("ij,kj->ik", a, a)
    vector = (1 - cg[..., :, None]) * (1 - cg[..., None, :]) + sg[..., :, None] * sg[..., None, :]
    sd = torch.einsum("...ij,ij->...", vector, metric) ** 0.5
    return sd, rel


def construct_periodic_input_features(x, atoms, simulation_cell, distance_type="nu"):
    """network.py:249-302."""
    if distance_type == "nu":
        distance_func = nu_distance
    elif distance_type == "tri":
        distance_func = tri_distance
    else:
        raise ValueError("Unrecognized distance function.")
    prim = simulation_cell.original_cell
    x = x.reshape(-1, 3)
    n = x.shape[0]
    prim_x, _ = enforce_pbc(_t(prim.a), x)
    prim_xea = prim_x[..., None, :] - atoms
    sea, xea = distance_func(prim_xea, _t(prim.AV), _t(prim.BV))
    sea = sea[..., None]
    sim_x, _ = enforce_pbc(_t(simulation_cell.a), x)
    sim_xee = sim_x[:, None, :] - sim_x[None, :, :]
    eye = torch.eye(n, dtype=x.dtype)
    see, xee = distance_func(sim_xee + eye[..., None], _t(simulation_cell.AV), _t(simulation_cell.BV))
    see = (see * (1.0 - eye))[..., None]
    xee = xee * (1.0 - eye)[..., None]
    return xea, xee, sea, see


def construct_symmetric_features(h_one, h_two, spins):
    """network.py:305-332."""
    n0 = spins[0]
    h_ones = [h_one[:n0], h_one[n0:]]
    h_twos = [h_two[:n0], h_two[n0:]]
    g_one = [h.mean(dim=0, keepdim=True) for h in h_ones if h.numel() > 0]
    g_two = [h.mean(dim=0) for h in h_twos if h.numel() > 0]
    g_one = [g.expand(h_one.shape[0], -1) for g in g_one]
    return torch.cat([h_one] + g_one + g_two, dim=1)


def isotropic_envelope(ae, params):
    """network.py:335-337."""
    return torch.sum(torch.exp(-torch.abs(ae * params["sigma"])) * params["pi"], dim=1)


def diagonal_envelope(ae, params):
    """network.py:340-343."""
    r_ae = torch.linalg.norm(ae[..., None] * params["sigma"], dim=2)
    return torch.sum(torch.exp(-r_ae) * params["pi"], dim=1)


def full_envelope(ae, params):
    """network.py:349-364 (apply_covariance == einsum 'ijk,kmjn->ijmn')."""
    r_ae = torch.einsum("ijk,kmjn->ijmn", ae, params["sigma"])
    r_ae = torch.linalg.norm(r_ae, dim=2)
    return torch.sum(torch.exp(-r_ae) * params["pi"], dim=1)


def slogdet_op(x):
    """network.py:375-392."""
    if x.shape[-1] == 1:
        v = x[..., 0, 0]
        return torch.exp(1j * torch.angle(v)), torch.log(torch.abs(v))
    sign, logdet = torch.linalg.slogdet(x)
    return sign, logdet


def logdet_matmul(xs: Sequence[torch.Tensor]):
    """network.py:395-427 with w=None."""
    slogdets = [slogdet_op(x) for x in xs]
    sign_in, slogdet = slogdets[0]
    for s, l in slogdets[1:]:
        sign_in, slogdet = sign_in * s, slogdet + l
    # argmax gather: the max carries its own gradient, like slogdet[max_idx] in JAX
    slogdet_max = torch.gather(slogdet, 0, torch.argmax(slogdet.detach()).reshape(1))[0]     # (gather: vmap-able)
    det = sign_in * torch.exp(slogdet - slogdet_max)
    result = torch.sum(det)
    sign_out = torch.exp(1j * torch.angle(result))
    slog_out = torch.log(torch.abs(result)) + slogdet_max
    return sign_out, slog_out


_LAYER_TAP = None       # set by kfac_factors: callable (x, w, b, y) seeing every register_repeated_dense layer


def linear_layer(x, w, b=None):
    """network.py:430-443 (the KFAC tag is the identity on values)."""
    y = x @ w
    y = y + b if b is not None else y
    if _LAYER_TAP is not None:
        _LAYER_TAP(x, w, b, y)
    return y


def eval_phase(x, klist, spins, full_det=False):
    """network.py:449-458."""
    x = x.reshape(-1, 3)
    xs = [x[:spins[0]], x[spins[0]:]]
    if full_det:
        kall = torch.cat([_t(k) for k in klist], dim=0)
        kdot = [xx @ kall.T for xx, ne in zip(xs, spins) if ne > 0]
    else:
        kdot = [xx @ _t(k).T for xx, k, ne in zip(xs, klist, spins) if ne > 0]
    return [torch.exp(1j * kd) for kd in kdot]


def solid_fermi_net_orbitals(params, x, simulation_cell, klist, atoms, spins,
                             envelope_type="isotropic", full_det=False, distance_type="nu"):
    """network.py:461-560."""
    ae_, ee_, r_ae, r_ee = construct_periodic_input_features(
        x, atoms, simulation_cell, distance_type=distance_type)
    ae = torch.cat((r_ae, ae_), dim=2)
    ae = ae.reshape(ae.shape[0], -1)
    ee = torch.cat((r_ee, ee_), dim=2)
    to_env = r_ae if envelope_type == "isotropic" else ae_
    envelope = {"isotropic": isotropic_envelope, "diagonal": diagonal_envelope,
                "full": full_envelope}.get(envelope_type)
    h_one, h_two = ae, ee
    residual = lambda a, b: (a + b) / math.sqrt(2.0) if a.shape == b.shape else b
    nd, ns = len(params["double"]), len(params["single"])
    for i in range(nd):
        h_one_in = construct_symmetric_features(h_one, h_two, spins)
        h_one_next = torch.tanh(linear_layer(h_one_in, params["single"][i]["w"], params["single"][i]["b"]))
        h_two_next = torch.tanh(linear_layer(h_two, params["double"][i]["w"], params["double"][i]["b"]))
        h_one = residual(h_one, h_one_next)
        h_two = residual(h_two, h_two_next)
    if nd != ns:
        h_one_in = construct_symmetric_features(h_one, h_two, spins)
        h_one_next = torch.tanh(linear_layer(h_one_in, params["single"][-1]["w"], params["single"][-1]["b"]))
        h_one = residual(h_one, h_one_next)
        h_to_orbitals = h_one
    else:
        h_to_orbitals = construct_symmetric_features(h_one, h_two, spins)
    hs = [h_to_orbitals[:spins[0]], h_to_orbitals[spins[0]:]]
    active = [s for s in spins if s > 0]
    hs = [h for h, s in zip(hs, spins) if s > 0]
    orbitals = [linear_layer(h, p["w"], p.get("b")) for h, p in zip(hs, params["orbital"])]
    for i in range(len(active)):
        npar = params["orbital"][i]["w"].shape[-1] // 2
        orbitals[i] = orbitals[i][..., :npar] + 1j * orbitals[i][..., npar:]
    if envelope is not None:
        splits, o = [], 0
        for s in active:
            splits.append(to_env[o:o + s])
            o += s
        orbitals = [envelope(te, p) * orb for te, orb, p in zip(splits, orbitals, params["envelope"])]
    ntot = sum(spins)
    orbitals = [orb.reshape(s, -1, ntot if full_det else s).permute(1, 0, 2)
                for s, orb in zip(active, orbitals)]
    phases = eval_phase(x, klist, spins, full_det=full_det)
    orbitals = [orb * p[None, :, :] for orb, p in zip(orbitals, phases)]
    if full_det:
        orbitals = [torch.cat(orbitals, dim=1)]
    return orbitals, to_env


def eval_func(params, x, klist, simulation_cell, atoms, spins, envelope_type="isotropic",
              full_det=False, distance_type="nu", method_name="eval_slogdet"):
    """network.py:563-606."""
    orbitals, _ = solid_fermi_net_orbitals(params, x, simulation_cell, klist, atoms, spins,
                                           envelope_type, full_det, distance_type)
    if method_name == "eval_slogdet":
        return logdet_matmul(orbitals)[1]
    if method_name == "eval_logdet":
        sign, slog = logdet_matmul(orbitals)
        return torch.log(sign) + slog
    if method_name == "eval_phase_and_slogdet":
        return logdet_matmul(orbitals)
    if method_name == "eval_mats":
        return orbitals
    raise ValueError("Unrecognized method name")


def init_params(rng: np.random.Generator, natom: int, spins, envelope_type="isotropic",
                bias_orbitals=False, use_last_layer=False, full_det=False,
                hidden_dims=((256, 32),) * 3, determinants=8, distance_type="nu") -> Dict:
    """Shapes and distributions of network.py:60-186 (numpy RNG instead of jax.random)."""
    if distance_type == "nu":
        in_dims = (natom * 4, 4)
    elif distance_type == "tri":
        in_dims = (natom * 7, 7)
    else:
        raise ValueError("Unrecognized distance function.")
    active = [s for s in spins if s > 0]
    nch = len(active)
    dims_one_in = ([(nch + 1) * in_dims[0] + nch * in_dims[1]] +
                   [(nch + 1) * h[0] + nch * h[1] for h in hidden_dims])
    if not use_last_layer:
        dims_one_in[-1] = hidden_dims[-1][0]
    dims_one_out = [h[0] for h in hidden_dims]
    dims_two = [in_dims[1]] + [h[1] for h in hidden_dims]
    len_double = len(hidden_dims) if use_last_layer else len(hidden_dims) - 1
    P = {"single": [], "double": [], "orbital": [], "envelope": []}
    for s in active:
        npar = sum(spins) * determinants if full_det else s * determinants
        env = {"pi": np.ones((natom, npar))}
        if envelope_type == "isotropic":
            env["sigma"] = np.ones((natom, npar))
        elif envelope_type == "diagonal":
            env["sigma"] = np.ones((natom, 3, npar))
        elif envelope_type == "full":
            env["sigma"] = np.tile(np.eye(3)[..., None, None], [1, 1, natom, npar])
        P["envelope"].append(env)
    for i in range(len(hidden_dims)):
        P["single"].append({
            "w": rng.standard_normal((dims_one_in[i], dims_one_out[i])) / np.sqrt(float(dims_one_in[i])),
            "b": rng.standard_normal((dims_one_out[i],))})
        if i < len_double:
            P["double"].append({
                "w": rng.standard_normal((dims_two[i], dims_two[i + 1])) / np.sqrt(float(dims_two[i])),
                "b": rng.standard_normal((dims_two[i + 1],))})
    for s in active:
        npar = sum(spins) * determinants if full_det else s * determinants
        orb = {"w": rng.standard_normal((dims_one_in[-1], 2 * npar)) / np.sqrt(float(dims_one_in[-1]))}
        if bias_orbitals:
            orb["b"] = rng.standard_normal((2 * npar,))
        P["orbital"].append(orb)
    return P


def params_to_torch(params) -> Dict:
    def conv(v):
        if isinstance(v, dict):
            return {k: conv(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [conv(x) for x in v]
        return _t(v)
    return conv(params)


def make_solid_fermi_net(klist, simulation_cell, envelope_type="isotropic", bias_orbitals=False,
                         use_last_layer=False, full_det=False, hidden_dims=((256, 32),) * 3,
                         determinants=8, after_determinants=1, distance_type="nu",
                         method_name="eval_logdet") -> Callable:
    """network.py:609-667; returns apply(params, x) for ONE walker."""
    if method_name not in ["eval_slogdet", "eval_logdet", "eval_mats", "eval_phase_and_slogdet"]:
        raise ValueError("Method name is not in class dir.")
    simulation_cell = rederive(simulation_cell)      # the oracle's own supercell atoms, AV, BV (oracle/geometry.py)
    atoms = _t(simulation_cell.original_cell.atom_coords())
    spins = tuple(simulation_cell.nelec)

    def apply(params, x):
        return eval_func(params, x, klist=klist, simulation_cell=simulation_cell, atoms=atoms,
                         spins=spins, envelope_type=envelope_type, full_det=full_det,
                         distance_type=distance_type, method_name=method_name)
    return apply


# ---------------------------------------------------------------------------
# distance.py / ewaldsum.py
# ---------------------------------------------------------------------------

class MinimalImageDistance:
    """distance.py:32-141 (torch), including the classification of distance.py:41-59 with its `dot < tol` test
    that has no absolute value (an obtuse lattice is treated as orthogonal by the reference; kept)."""

    def __init__(self, latvec):
        lat = np.asarray(latvec, dtype=float)
        ortho_tol = 1e-10
        diagonal = bool(np.all(np.abs(lat - np.diag(np.diagonal(lat))) < ortho_tol))
        if diagonal:
            self.kind = 0                                            # diagonal_dist_i
        else:
            orthogonal = (np.dot(lat[0], lat[1]) < ortho_tol and np.dot(lat[1], lat[2]) < ortho_tol
                          and np.dot(lat[2], lat[0]) < ortho_tol)
            self.kind = 1 if orthogonal else 2                       # orthogonal_dist_i / general_dist_i
        self._latvec = _t(lat)
        self._invvec = torch.linalg.inv(self._latvec)
        # list of all 26 neighbouring cells, distance.py:64-66 (meshgrid with its default 'xy' indexing)
        mesh_grid = np.meshgrid(*[np.array([0, 1, 2]) for _ in range(3)])
        self.point_list = _t(np.stack([m.ravel() for m in mesh_grid], axis=0).T - 1)
        self.shifts = self.point_list @ self._latvec

    def dist_i(self, configs, vec):
        configs = configs.reshape(1, -1, 3)
        v = vec.reshape(-1, 1, 3)
        d1 = v - configs
        if self.kind == 0:      # distance.py:110-128
            diag = torch.diagonal(self._latvec)
            return torch.remainder(d1 + diag / 2, diag) - diag / 2
        if self.kind == 1:      # distance.py:91-108
            frac = d1 @ self._invvec
            return (torch.remainder(frac + 0.5, 1.0) - 0.5) @ self._latvec
        shifts = self.shifts.reshape(-1, 1, 1, 3)     # distance.py:70-89
        d1all = d1[None] + shifts
        dists = torch.linalg.norm(d1all, dim=-1)
        mininds = torch.argmin(dists, dim=0)
        return torch.gather(d1all, 0, mininds[None, ..., None].expand(1, *d1.shape))[0]

    def dist_matrix(self, configs):
        vs = self.dist_i(configs, configs)
        n = vs.shape[0]
        return vs * (1 - torch.eye(n, dtype=vs.dtype))[..., None]


_GPOINT_CACHE = {}


def _select_big(n0, n1, n2, cellvolume, recvec, alpha):
    """ewaldsum.py:194-200 on the integer mesh n0 x n1 x n2 ('ij' order)."""
    g = np.stack(np.meshgrid(n0, n1, n2, indexing="ij"), axis=0).reshape(3, -1).astype(float)
    gpoints = np.einsum("jn,jk->nk", g, recvec) * 2 * np.pi
    gsquared = np.einsum("nk,nk->n", gpoints, gpoints)
    gweight = 4 * np.pi * np.exp(-gsquared / (4 * alpha ** 2))
    gweight /= cellvolume * gsquared
    bigweight = gweight > 1e-12
    return gpoints[bigweight], gweight[bigweight]


def _reciprocal_points(latvec, ewald_gmax):
    """ewaldsum.py:58-89, the literal enumeration: every integer triple of the half space up to `ewald_gmax` is
    weighed (32 M points at 200; evaluated slab by slab along the first index, which keeps the reference's order)."""
    key = (latvec.tobytes(), int(ewald_gmax))
    if key not in _GPOINT_CACHE:
        cellvolume = np.linalg.det(latvec)
        recvec = np.linalg.inv(latvec).T
        smallestheight = np.amin(1 / np.linalg.norm(recvec, axis=1))
        alpha = 5.0 / smallestheight
        full = np.arange(-ewald_gmax, ewald_gmax + 1)
        pos = np.arange(1, ewald_gmax + 1)
        zero = np.array([0])
        parts = [_select_big(np.array([i]), full, full, cellvolume, recvec, alpha) for i in pos]     # gptsXpos
        parts.append(_select_big(zero, pos, full, cellvolume, recvec, alpha))                        # gptsX0Ypos
        parts.append(_select_big(zero, zero, pos, cellvolume, recvec, alpha))                        # gptsX0Y0Zpos
        _GPOINT_CACHE[key] = (float(alpha), float(cellvolume), np.concatenate([p[0] for p in parts], axis=0),
                              np.concatenate([p[1] for p in parts], axis=0))
    return _GPOINT_CACHE[key]


class EwaldSum:
    """ewaldsum.py:33-200 (torch / numpy), setup and per-walker halves, restated here on its own (nothing is shared
    with the product's host-side table builder, which `tests/test_geometry.py` checks against this class)."""

    def __init__(self, cell, ewald_gmax=200, nlatvec=1):
        cell = rederive(cell)
        self.nelec = tuple(cell.nelec)
        coords = np.asarray(cell.atom_coords(), dtype=float)
        charges = np.asarray(cell.atom_charges(), dtype=float)
        latvec = np.ascontiguousarray(np.asarray(cell.lattice_vectors(), dtype=float))
        self.atom_coords = _t(coords)
        self.atom_charges = _t(charges)
        self.latvec = _t(latvec)
        self.dist = MinimalImageDistance(latvec)
        # set_lattice_displacements, ewaldsum.py:48-56
        XYZ = np.meshgrid(*[np.arange(-nlatvec, nlatvec + 1)] * 3, indexing="ij")
        xyz = np.stack(XYZ, axis=-1).reshape((-1, 3))
        self.lattice_displacements = _t(np.dot(xyz, latvec))
        # set_up_reciprocal_ewald_sum, ewaldsum.py:58-90
        self.alpha, cellvolume, gpoints, gweight = _reciprocal_points(latvec, ewald_gmax)
        self.gpoints, self.gweight = _t(gpoints), _t(gweight)
        # set_ewald_constants, ewaldsum.py:92-101
        self.i_sum = float(np.sum(charges))
        ii_sum2 = float(np.sum(charges ** 2))
        ii_sum = (self.i_sum ** 2 - ii_sum2) / 2
        self.ijconst = -np.pi / (cellvolume * self.alpha ** 2)
        self.squareconst = -self.alpha / np.sqrt(np.pi) + self.ijconst / 2
        self.ii_const = ii_sum * self.ijconst + ii_sum2 * self.squareconst
        self.ion_ion = self.ewald_ion()

    def ee_const(self, ne):         # ewaldsum.py:109-110
        return ne * (ne - 1) / 2 * self.ijconst + ne * self.squareconst

    def ei_const(self, ne):         # ewaldsum.py:112-113
        return -ne * self.i_sum * self.ijconst

    def ewald_ion(self):            # ewaldsum.py:120-136
        if len(self.atom_charges) == 1:
            ion_ion_real = 0.0
        else:
            ion_distances = self.dist.dist_matrix(self.atom_coords.reshape(-1))
            rvec = ion_distances[None, :, :, :] + self.lattice_displacements[:, None, None, :]
            r = torch.linalg.norm(rvec, dim=-1)
            charge_ij = self.atom_charges[..., None] * self.atom_charges[None, ...]
            n = len(self.atom_charges)
            mask = torch.triu(torch.ones(n, n, dtype=torch.bool), diagonal=1)
            ion_ion_real = float(torch.sum(torch.where(mask[None], charge_ij * torch.erfc(self.alpha * r) / r,
                                                       torch.zeros((), dtype=DT))))
        GdotR = self.gpoints @ self.atom_coords.T
        self.ion_exp = torch.exp(1j * GdotR.to(CT)) @ self.atom_charges.to(CT)
        ion_ion_rec = float(torch.dot(self.gweight, self.ion_exp.abs() ** 2))
        return ion_ion_real + ion_ion_rec

    def _real_cij(self, dists):
        r = dists[:, :, None, :] + self.lattice_displacements
        r = torch.linalg.norm(r, dim=-1)
        return torch.sum(torch.erfc(self.alpha * r) / r, dim=-1)

    def ewald_electron(self, configs):
        nelec = sum(self.nelec)
        ei_d = self.dist.dist_i(self.atom_coords.reshape(-1), configs)
        ei_real = torch.sum(-self.atom_charges[None, :] * self._real_cij(ei_d))
        ee_real = torch.zeros((), dtype=DT)
        if nelec > 1:
            ee_d = self.dist.dist_matrix(configs)
            rvec = ee_d[None] + self.lattice_displacements[:, None, None, :]
            r = torch.linalg.norm(rvec, dim=-1)
            mask = torch.triu(torch.ones(nelec, nelec, dtype=torch.bool), diagonal=1)
            ee_real = torch.sum(torch.where(mask[None], torch.erfc(self.alpha * r) / r,
                                            torch.zeros((), dtype=DT)))
        ee_rec, ei_rec = self.reciprocal_space_electron(configs)
        return ee_real + ee_rec, ei_real + ei_rec

    def reciprocal_space_electron(self, configs):
        gr = configs.reshape(sum(self.nelec), -1) @ self.gpoints.T
        ssin, scos = torch.sin(gr).sum(dim=0), torch.cos(gr).sum(dim=0)
        ee = torch.dot(ssin ** 2 + scos ** 2, self.gweight)
        cs = -self.ion_exp.real * scos - self.ion_exp.imag * ssin
        return ee, 2 * torch.dot(cs, self.gweight)

    def energy(self, configs):      # ewaldsum.py:185-191
        ne = sum(self.nelec)
        ee, ei = self.ewald_electron(configs)
        return ee + self.ee_const(ne), ei + self.ei_const(ne), torch.tensor(self.ion_ion + self.ii_const, dtype=DT)


def enforce_pbc_batch(latvec, epos):
    """distance.py:144-163 (vmapped over the batch): divmod(frac, 1)."""
    lat = _t(latvec)
    B = epos.shape[0]
    frac = epos.reshape(B, -1, 3) @ torch.linalg.inv(lat)
    wrap = torch.floor(frac)
    return ((frac - wrap) @ lat).reshape(B, -1), wrap


# ---------------------------------------------------------------------------
# hamiltonian.py
# ---------------------------------------------------------------------------

def local_kinetic_energy_real_imag(f):
    """hamiltonian.py:45-70 ('for' mode): per direction jvp(grad(Re f)) and jvp(grad(Im f))."""
    def _lapl_over_f(params, x):
        ne = x.shape[-1]
        eye = torch.eye(ne, dtype=x.dtype)
        g_re = grad(lambda y: f(params, y).real)
        g_im = grad(lambda y: f(params, y).imag)
        re = torch.zeros((), dtype=DT)
        im = torch.zeros((), dtype=DT)
        for i in range(ne):
            p_re, t_re = jvp(g_re, (x,), (eye[i],))
            p_im, t_im = jvp(g_im, (x,), (eye[i],))
            re = re + t_re[i] + p_re[i] ** 2 - p_im[i] ** 2
            im = im + t_im[i] + 2 * p_re[i] * p_im[i]
        return [-0.5 * re, -0.5 * im * 1j]
    return _lapl_over_f


def local_kinetic_energy_partition(f, partition_number=3):
    """hamiltonian.py:127-159 ('partition' mode): vmapped jvp over chunks of eye(3N)."""
    def _lapl_over_f(params, x):
        n = x.shape[0]
        if n % partition_number:
            raise ValueError("partition_number must divide 3*N_elec")   # jnp.asarray(array_split) would fail
        eye = torch.eye(n, dtype=x.dtype)
        g_re = grad(lambda y: f(params, y).real)
        g_im = grad(lambda y: f(params, y).imag)
        vjvp = lambda g, e: vmap(lambda t: jvp(g, (x,), (t,)))(e)
        prs, pis, trs, tis = [], [], [], []
        for e in eye.reshape(partition_number, n // partition_number, n):
            pr, tr = vjvp(g_re, e)
            pi_, ti = vjvp(g_im, e)
            prs.append(pr); pis.append(pi_); trs.append(tr); tis.append(ti)
        primal = [torch.cat(prs), torch.cat(pis)]
        tangent = [torch.cat(trs), torch.cat(tis)]
        real = torch.trace(tangent[0]) + torch.trace(primal[0] ** 2) - torch.trace(primal[1] ** 2)
        imag = torch.trace(tangent[1]) + torch.trace(2 * primal[0] * primal[1])
        return [-0.5 * real, -0.5 * 1j * imag]
    return _lapl_over_f


def local_kinetic_energy_dim_batch(f):
    """hamiltonian.py:73-101."""
    return local_kinetic_energy_partition(f, partition_number=1)


def local_kinetic_energy_hessian(f):
    """hamiltonian.py:104-124."""
    from torch.func import hessian

    def _lapl_over_f(params, x):
        g_re = grad(lambda y: f(params, y).real)(x)
        g_im = grad(lambda y: f(params, y).imag)(x)
        h_re = hessian(lambda y: f(params, y).real)(x)
        h_im = hessian(lambda y: f(params, y).imag)(x)
        real = torch.trace(h_re) + torch.sum(g_re ** 2) - torch.sum(g_im ** 2)
        imag = torch.trace(h_im) + torch.sum(2 * g_re * g_im)
        return [-0.5 * real, -0.5 * 1j * imag]
    return _lapl_over_f


def local_ewald_energy(simulation_cell):
    """hamiltonian.py:163-179 (the pyscf energy_nuc assertion is replaced by the
    Madelung tests in tests/test_ewald_oracle.py)."""
    ewald = EwaldSum(simulation_cell)

    def _local_ewald_energy(x):
        return sum(ewald.energy(x))
    return _local_ewald_energy


def local_energy_seperate(f, simulation_cell, mode="for", partition_number=3):
    """hamiltonian.py:194-228."""
    if mode == "for":
        ke_ri = local_kinetic_energy_real_imag(f)
    elif mode == "hessian":
        ke_ri = local_kinetic_energy_hessian(f)
    elif mode == "dim_batch":
        ke_ri = local_kinetic_energy_dim_batch(f)
    elif mode == "partition":
        ke_ri = local_kinetic_energy_partition(f, partition_number=partition_number)
    else:
        raise ValueError("Unrecognized laplacian evaluation mode.")
    ew = local_ewald_energy(simulation_cell)

    def _local_energy(params, x):
        parts = ke_ri(params, x)
        return parts[0] + parts[1], ew(x)
    return _local_energy


# ---------------------------------------------------------------------------
# qmc.py / train.py
# ---------------------------------------------------------------------------

def mh_update(params, f_batch, x1, lp_1, num_accepts, latvec, stddev, xi, u):
    """qmc.py:153-224, symmetric branch, with the gaussian noise ``xi`` (B,3N) and the
    uniforms ``u`` (B,) supplied by the caller instead of jax.random."""
    x2 = x1 + stddev * xi
    x2, _ = enforce_pbc_batch(latvec, x2)
    lp_2 = 2.0 * f_batch(params, x2)
    ratio = lp_2 - lp_1
    rnd = torch.log(u)
    cond = ratio > rnd
    x_new = torch.where(cond[..., None], x2, x1)
    lp_new = torch.where(cond, lp_2, lp_1)
    return x_new, lp_new, num_accepts + cond.sum().to(DT), cond


def make_mcmc_step(batch_slog_network, batch_per_device, latvec, steps=10):
    """qmc.py:290-364 (Metropolis, all-electron moves).  ``noise`` = (xi[steps,B,3N], u[steps,B])
    replaces the PRNG key; returns (data, pmove, accept_masks[steps,B])."""
    def mcmc_step(params, data, noise, width):
        xi, u = noise
        logprob = 2.0 * batch_slog_network(params, data)
        n_acc = torch.zeros((), dtype=DT)
        masks = []
        for s in range(steps):
            data, logprob, n_acc, cond = mh_update(params, batch_slog_network, data, logprob, n_acc,
                                                   latvec, width, xi[s], u[s])
            masks.append(cond)
        pmove = n_acc / (steps * batch_per_device)
        return data, pmove, torch.stack(masks)
    return mcmc_step


def total_energy_stats(ke, ew):
    """train.py:74-89, single device: loss, imaginary, variance from per-walker (ke, ew)."""
    e_l = ke + ew
    mean = e_l.mean()
    variance = (e_l.abs() ** 2).mean() - mean.real.abs() ** 2
    return mean.real, mean.imag, variance


def batch_apply(apply, params, X):
    """process.py:116-118: the caller vmaps apply over the batch; a loop is the oracle."""
    return torch.stack([apply(params, x) for x in X])


# ---------------------------------------------------------------------------
# train.py:91-142 -- the energy-gradient estimator (custom JVP of total_energy), reverse mode via autograd
# ---------------------------------------------------------------------------
def _leaves(params):
    out = []
    for layer in params["single"]:
        out += [layer["w"], layer["b"]]
    for layer in params["double"]:
        out += [layer["w"], layer["b"]]
    for orb in params["orbital"]:
        out += [orb["w"]] + ([orb["b"]] if "b" in orb else [])
    for env in params["envelope"]:
        out += [env["pi"], env["sigma"]]
    return out


def _clone_params(params, requires_grad=True):
    def conv(v):
        if isinstance(v, dict):
            return {k: conv(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [conv(x) for x in v]
        return v.detach().clone().requires_grad_(requires_grad)
    return conv(params)


def logpsi_vjp(apply_phase_slog, params, X, cot_abs, cot_phase):
    """Pytree of d/dparams sum_b cot_abs[b] log|psi_b| + cot_phase[b] angle(psi_b); `apply_phase_slog` is the
    per-walker network with method_name='eval_phase_and_slogdet' (network.py:599-600)."""
    P = _clone_params(params)
    total = torch.zeros((), dtype=DT)
    for b, x in enumerate(X):
        sign, slog = apply_phase_slog(P, x)
        total = total + cot_abs[b] * slog + cot_phase[b] * torch.angle(sign)
    grads = torch.autograd.grad(total, _leaves(P), allow_unused=True)
    it = iter([g if g is not None else torch.zeros_like(l) for g, l in zip(grads, _leaves(P))])
    return {"single": [{"w": next(it), "b": next(it)} for _ in params["single"]],
            "double": [{"w": next(it), "b": next(it)} for _ in params["double"]],
            "orbital": [({"w": next(it), "b": next(it)} if "b" in o else {"w": next(it)}) for o in params["orbital"]],
            "envelope": [{"pi": next(it), "sigma": next(it)} for _ in params["envelope"]]}


def clip_difference(diff, clip_local_energy=5.0, clip_type="real"):
    """train.py:101-127 on one device."""
    if clip_local_energy <= 0.0:
        return diff
    if clip_type == "complex":
        radius, phase = diff.abs(), torch.angle(diff)
        radius_tv = radius.std(unbiased=False)
        radius_mean = torch.as_tensor(np.median(radius.numpy()))
        clip_radius = torch.clip(radius, radius_mean - radius_tv * clip_local_energy,
                                 radius_mean + radius_tv * clip_local_energy)
        return clip_radius * torch.exp(1j * phase)
    if clip_type == "real":
        tv_re = diff.real.abs().mean()
        tv_im = diff.imag.abs().mean()
        return torch.complex(torch.clip(diff.real, -clip_local_energy * tv_re, clip_local_energy * tv_re),
                             torch.clip(diff.imag, -clip_local_energy * tv_im, clip_local_energy * tv_im))
    raise ValueError("Unrecognized clip type.")


def total_energy_value_and_grad(apply_phase_slog, el_fun, params, X, clip_local_energy=5.0, clip_type="real"):
    """(loss, e_l, grads) of train.make_loss: loss = Re mean e_l, grads = the pullback of
    tangents_dot = mean(Re(clip_diff * conj(d log psi))) (train.py:129-137)."""
    out = [el_fun(params, x) for x in X]
    e_l = torch.stack([torch.as_tensor(complex(k) + float(e), dtype=torch.complex128) for k, e in out])
    loss = e_l.mean().real
    clip_diff = clip_difference(e_l - loss, clip_local_energy, clip_type)
    n = len(X)
    grads = logpsi_vjp(apply_phase_slog, params, X, clip_diff.real / n, clip_diff.imag / n)
    return loss, e_l, grads


# ---------------------------------------------------------------------------
# qmc.py:63-150, 227-287 -- one-electron moves and importance sampling (caller-supplied noise)
# ---------------------------------------------------------------------------
def limdrift(g, cutoff=1.0):
    """qmc.py:63-82."""
    shape = g.shape
    g3 = g.reshape(-1, 3)
    tot = torch.linalg.norm(g3, dim=-1)
    normalize = torch.clip(tot, min=cutoff, max=float(tot.max()) if tot.numel() else cutoff)
    return (cutoff * g3 / normalize[:, None]).reshape(shape)


def mh_one_electron_update(params, f_batch, x1, lp_1, num_accepts, latvec, stddev, xi3, u, i):
    """qmc.py:227-287: move electron i % N of every walker by stddev * xi3 (B,3), re-wrap the configuration."""
    n = x1.shape[0]
    x = x1.reshape(n, -1, 3)
    ii = i % x.shape[1]
    x2 = x.clone()
    x2[:, ii] = x2[:, ii] + stddev * xi3
    x2, _ = enforce_pbc_batch(latvec, x2.reshape(n, -1))
    lp_2 = 2.0 * f_batch(params, x2)
    cond = (lp_2 - lp_1) > torch.log(u)
    return torch.where(cond[..., None], x2, x1), torch.where(cond, lp_2, lp_1), num_accepts + cond.sum().to(DT), cond


def make_mcmc_step_one_electron(batch_slog_network, batch_per_device, latvec, steps=10):
    """qmc.py:355-358 with one_electron_moves=True; noise = (xi[steps*N,B,3], u[steps*N,B])."""
    def mcmc_step(params, data, noise, width):
        xi, u = noise
        nelec = data.shape[-1] // 3
        nsteps = nelec * steps
        logprob = 2.0 * batch_slog_network(params, data)
        n_acc = torch.zeros((), dtype=DT)
        masks = []
        for s in range(nsteps):
            data, logprob, n_acc, cond = mh_one_electron_update(params, batch_slog_network, data, logprob, n_acc,
                                                                latvec, width, xi[s], u[s], s)
            masks.append(cond)
        return data, n_acc / (nsteps * batch_per_device), torch.stack(masks)
    return mcmc_step


def value_and_grad_x(slog_apply, params, X):
    """jax.vmap(jax.value_and_grad(slog_network, argnums=1)) (qmc.py:325)."""
    vals, grads = [], []
    for x in X:
        xr = x.detach().clone().requires_grad_(True)
        v = slog_apply(params, xr)
        g, = torch.autograd.grad(v, xr)
        vals.append(v.detach()); grads.append(g)
    return torch.stack(vals), torch.stack(grads)


def make_mcmc_step_importance(slog_apply, batch_per_device, latvec, steps=10):
    """importance_update (qmc.py:83-150, atoms=None) driven by make_mcmc_step; noise = (xi[steps,B,3N], u[steps,B]).
    The reference recomputes grad(x1) every step; it equals the stored gradient of the accepted configuration."""
    def mcmc_step(params, data, noise, width):
        xi, u = noise
        x1 = data
        lpsi, _ = value_and_grad_x(slog_apply, params, x1)
        lp_1 = 2.0 * lpsi
        n_acc = torch.zeros((), dtype=DT)
        masks = []
        for s in range(steps):
            _, grad = value_and_grad_x(slog_apply, params, x1)
            grad = limdrift(grad)
            gauss = width * xi[s]
            x2 = x1 + gauss + width ** 2 * grad
            x2, _ = enforce_pbc_batch(latvec, x2)
            lpsi_2, new_grad = value_and_grad_x(slog_apply, params, x2)
            new_grad = limdrift(new_grad)
            forward = (gauss ** 2).sum(-1)
            backward = ((gauss + width ** 2 * (grad + new_grad)) ** 2).sum(-1)
            lp_2 = 2.0 * lpsi_2 + (forward - backward) / (2.0 * width ** 2)
            cond = (lp_2 - lp_1) > torch.log(u[s])
            x1 = torch.where(cond[..., None], x2, x1)
            lp_1 = torch.where(cond, lp_2, lp_1)
            n_acc = n_acc + cond.sum()
            masks.append(cond)
        return x1, n_acc / (steps * batch_per_device), torch.stack(masks)
    return mcmc_step


# ---------------------------------------------------------------------------
# KFAC curvature statistics of the tagged layers, as the reference's estimator forms them for one batch:
# train.py:128-133 (loss tag on conj(log psi), variance 0.5), utils/kfac_ferminet_alpha/estimator.py:284-320
# (fisher_exact, one index), loss_functions.py:529-537 (tangent = 1/sqrt(variance)), tracer.py:196-332 + vjp_rc.py
# (complex cotangent dy = sqrt2 (d Re F/dy + i d Im F/dy), F = conj(log psi)), curvature_blocks.py:262-281 and
# curvature_tags_and_blocks.py:142-156 (RepeatedDenseBlock: rows = every leading index), curvature_blocks.py:111-133
# (NaiveDiagonal for the untagged envelope leaves).
# ---------------------------------------------------------------------------
def kfac_factors(apply_phase_slog, params, X):
    """-> dict(single=[...], double=[...], orbital=[...], envelope=[...]): per tagged layer
    {inputs_factor, outputs_factor, extra_scale}; per envelope spin {pi, sigma} complex diagonal factors."""
    global _LAYER_TAP
    P = _clone_params(params)
    ids = {}
    for kind in ("single", "double", "orbital"):
        for i, layer in enumerate(P[kind]):
            ids[id(layer["w"])] = (kind, i)
    acc = {}
    env_leaves = [t for env in P["envelope"] for t in (env["pi"], env["sigma"])]
    dw = [torch.zeros_like(t, dtype=torch.complex128) for t in env_leaves]
    sqrt2 = math.sqrt(2.0)                                   # 1 / sqrt(variance = 0.5)
    for x in X:
        taps = []
        _LAYER_TAP = lambda xx, w, b, y: taps.append((ids[id(w)], xx, b is not None, y))
        try:
            sign, slog = apply_phase_slog(P, x)
        finally:
            _LAYER_TAP = None
        re_f, im_f = slog, -torch.angle(sign)               # F = conj(log psi)
        ys = [t[3] for t in taps]
        g_re = torch.autograd.grad(re_f, ys + env_leaves, retain_graph=True, allow_unused=True)
        g_im = torch.autograd.grad(im_f, ys + env_leaves, retain_graph=False, allow_unused=True)
        z = lambda g, ref: torch.zeros_like(ref) if g is None else g
        for k, (key, xx, has_b, y) in enumerate(taps):
            dy = sqrt2 * (z(g_re[k], y) + 1j * z(g_im[k], y))
            x2 = xx.reshape(-1, xx.shape[-1])
            if has_b:
                x2 = torch.cat([x2, torch.ones_like(x2[:, :1])], dim=1)
            d2 = dy.reshape(-1, dy.shape[-1])
            a = acc.setdefault(key, {"xx": 0.0, "dd": 0.0, "xdy": 0.0, "rows": 0})
            a["xx"] = a["xx"] + x2.T @ x2
            a["dd"] = a["dd"] + (d2.conj().T @ d2).real
            a["xdy"] = a["xdy"] + x2.to(d2.dtype).T @ d2
            a["rows"] += x2.shape[0]
        for k, leaf in enumerate(env_leaves):
            dw[k] = dw[k] + sqrt2 * (z(g_re[len(ys) + k], leaf) + 1j * z(g_im[len(ys) + k], leaf))
    B = len(X)
    out = {"single": [], "double": [], "orbital": [], "envelope": []}
    for kind in ("single", "double", "orbital"):
        for i in range(len(P[kind])):
            a = acc[(kind, i)]
            out[kind].append({"inputs_factor": (a["xx"] / a["rows"]).detach(),
                              "outputs_factor": (a["dd"] / a["rows"]).detach(),
                              "extra_scale": a["rows"] // B,
                              # sum_rows (x,1)^T dy = sqrt2 (d sum Re F + i d sum Im F) / d(w; b): a check of the taps
                              "xdy": a["xdy"].detach()})
    for s in range(len(P["envelope"])):
        out["envelope"].append({"pi": (dw[2 * s] * dw[2 * s] / B).detach(),
                                "sigma": (dw[2 * s + 1] * dw[2 * s + 1] / B).detach()})
    return out


# ---------------------------------------------------------------------------
# estimator.py:15-85 -- observables (single process: pmean is the identity)
# ---------------------------------------------------------------------------
def make_complex_polarization(simulation_cell, direction=0, ndim=3):
    """estimator.py:15-40."""
    simulation_cell = rederive(simulation_cell)
    rec_vec = _t(np.asarray(simulation_cell.reciprocal_vectors())[direction])

    def complex_polarization(data):
        data = _t(data)
        data = data.reshape(list(data.shape[:-1]) + [-1, ndim])
        dots = torch.einsum("i,...i->...", rec_vec, data).sum(dim=-1)
        return torch.exp(1j * dots).mean(dim=-1)

    return complex_polarization


def make_structure_factor(simulation_cell, nq=4, ndim=3):
    """estimator.py:42-85."""
    simulation_cell = rederive(simulation_cell)
    mesh = np.meshgrid(*[np.arange(nq) for _ in range(3)])
    point_list = np.stack([m.ravel() for m in mesh], axis=0).T
    qvecs = _t(point_list @ np.asarray(simulation_cell.reciprocal_vectors()))
    nelec = simulation_cell.nelectron

    def structure_factor(data):
        data = _t(data)
        data = data.reshape(list(data.shape[:-1]) + [-1, ndim])
        dots = torch.einsum("kj,...j->...k", qvecs, data)          # batch, ne, npoint
        rho_k = torch.exp(1j * dots).sum(dim=1)
        rho_k_one = rho_k.mean(dim=0)
        rho_k_two = (rho_k.abs() ** 2).mean(dim=0)
        return (rho_k_two - rho_k_one.abs() ** 2) / nelec

    return structure_factor


# ---------------------------------------------------------------------------
# pretrain.py:70-89 -- the Hartree-Fock pretraining loss and its parameter gradient (autograd through eval_mats)
# ---------------------------------------------------------------------------
def pretrain_loss_and_grad(apply_mats, params, X, target, full_det=False):
    """target: list of (B, n_s, n_s) complex tensors.  -> (loss, grads pytree)."""
    P = _clone_params(params)
    predict = None
    for x in X:
        mats = apply_mats(P, x)                              # list of (D, n, n)
        predict = [[m] for m in mats] if predict is None else [p + [m] for p, m in zip(predict, mats)]
    predict = [torch.stack(p) for p in predict]              # (B, D, n, n)
    if full_det:
        B, na, nb = target[0].shape[0], target[0].shape[1], target[1].shape[1]
        t = torch.zeros(B, na + nb, na + nb, dtype=torch.complex128)
        t[:, :na, :na] = target[0]
        t[:, na:, na:] = target[1]
        target = [t]
    loss = torch.stack([(torch.abs(tar[:, None, ...] - pre) ** 2).mean() for tar, pre in zip(target, predict)]).mean()
    grads = torch.autograd.grad(loss, _leaves(P), allow_unused=True)
    it = iter([g if g is not None else torch.zeros_like(l) for g, l in zip(grads, _leaves(P))])
    return loss.detach(), {"single": [{"w": next(it), "b": next(it)} for _ in params["single"]],
                           "double": [{"w": next(it), "b": next(it)} for _ in params["double"]],
                           "orbital": [({"w": next(it), "b": next(it)} if "b" in o else {"w": next(it)}) for o in params["orbital"]],
                           "envelope": [{"pi": next(it), "sigma": next(it)} for _ in params["envelope"]]}

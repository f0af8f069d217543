"""Host-side (once per system) Ewald tables for the periodic Coulomb term.

Restates the *setup* half of the reference ``EwaldSum`` (ewaldsum.py:33-136,194-200)
in numpy; the per-walker half (ewaldsum.py:138-191) is the CUDA kernel
``ewald_kernel`` in csrc/ewald.cu.  Also classifies the lattice the way
``distance.MinimalImageDistance.__init__`` does (distance.py:35-68), including its
quirk of testing ``dot < tol`` without an absolute value.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.special import erfc

DIST_DIAGONAL, DIST_ORTHOGONAL, DIST_GENERAL = 0, 1, 2


def classify_lattice(latvec: np.ndarray) -> int:
    """distance.py:41-59."""
    latvec = np.asarray(latvec, dtype=float)
    tol = 1e-10
    if np.all(np.abs(latvec - np.diag(np.diagonal(latvec))) < tol):
        return DIST_DIAGONAL
    if (np.dot(latvec[0], latvec[1]) < tol and np.dot(latvec[1], latvec[2]) < tol
            and np.dot(latvec[2], latvec[0]) < tol):
        return DIST_ORTHOGONAL
    return DIST_GENERAL


def min_image_point_list() -> np.ndarray:
    """27 neighbour offsets in the reference's order (distance.py:64-66).

    The reference uses ``meshgrid`` with its default 'xy' indexing; the order only
    matters for argmin tie-breaking, which we keep identical.
    """
    mesh = np.meshgrid(*[np.array([0, 1, 2]) for _ in range(3)])
    return np.stack([m.ravel() for m in mesh], axis=0).T - 1


def _select_big(n_ranges, recvec, cellvolume, alpha):
    """ewaldsum.py:194-200 on an integer box (C-order enumeration like meshgrid 'ij')."""
    mesh = np.meshgrid(*n_ranges, indexing="ij")
    ints = np.stack([m.ravel() for m in mesh], axis=-1).astype(float)
    gpoints = ints @ recvec * 2 * np.pi
    gsq = np.einsum("ik,ik->i", gpoints, gpoints)
    with np.errstate(divide="ignore", invalid="ignore"):
        gweight = 4 * np.pi * np.exp(-gsq / (4 * alpha ** 2)) / (cellvolume * gsq)
    big = gweight > 1e-12
    return gpoints[big], gweight[big]


@dataclass
class EwaldTables:
    latvec: np.ndarray              # (3,3)
    invvec: np.ndarray              # (3,3) inverse of latvec
    atom_coords: np.ndarray         # (A_sc,3)
    atom_charges: np.ndarray        # (A_sc,)
    nelec: tuple
    dist_kind: int
    mi_shifts: np.ndarray           # (27,3) minimal-image candidate shifts (general lattices)
    lattice_displacements: np.ndarray  # (27,3) real-space sum images
    alpha: float
    gpoints: np.ndarray             # (nG,3)
    gweight: np.ndarray             # (nG,)
    ion_exp: np.ndarray             # (nG,) complex
    ijconst: float
    squareconst: float
    ii_const: float
    ion_ion: float
    i_sum: float

    def ee_const(self, ne):         # ewaldsum.py:109-110
        return ne * (ne - 1) / 2 * self.ijconst + ne * self.squareconst

    def ei_const(self, ne):         # ewaldsum.py:112-113
        return -ne * self.i_sum * self.ijconst

    @property
    def ii_total(self):             # ewaldsum.py:190
        return self.ion_ion + self.ii_const


def _min_image(kind, latvec, invvec, shifts, d):
    """distance.py:70-128 for displacement array d (...,3); numpy, used only for the ion-ion term."""
    if kind == DIST_DIAGONAL:
        diag = np.diagonal(latvec)
        return (d + diag / 2) % diag - diag / 2
    if kind == DIST_ORTHOGONAL:
        frac = d @ invvec
        return ((frac + 0.5) % 1 - 0.5) @ latvec
    dall = d[None] + shifts.reshape((-1,) + (1,) * (d.ndim - 1) + (3,))
    idx = np.argmin(np.linalg.norm(dall, axis=-1), axis=0)
    return np.take_along_axis(dall, idx[None, ..., None], axis=0)[0]


def build_ewald_tables(cell, ewald_gmax: int = 200, nlatvec: int = 1) -> EwaldTables:
    latvec = np.asarray(cell.lattice_vectors(), dtype=float)
    coords = np.asarray(cell.atom_coords(), dtype=float)
    charges = np.asarray(cell.atom_charges(), dtype=float)
    invvec = np.linalg.inv(latvec)
    kind = classify_lattice(latvec)
    plist = min_image_point_list()
    mi_shifts = plist @ latvec

    # ewaldsum.py:48-56
    rng = np.arange(-nlatvec, nlatvec + 1)
    xyz = np.stack(np.meshgrid(rng, rng, rng, indexing="ij"), axis=-1).reshape(-1, 3)
    disp = xyz @ latvec

    # ewaldsum.py:58-90
    cellvolume = np.linalg.det(latvec)
    recvec = invvec.T
    smallestheight = np.amin(1 / np.linalg.norm(recvec, axis=1))
    alpha = 5.0 / smallestheight
    # The reference enumerates |n| <= 200 and keeps weight > 1e-12.  The same set is
    # obtained from a box that just covers the sphere |G| < Gmax where the weight
    # drops below threshold (weight is monotone in |G|); the C-order of survivors is
    # unchanged because the small box is a sub-box of the big one.
    g = np.linspace(1e-3, 200.0, 400001)
    w = 4 * np.pi * np.exp(-g ** 2 / (4 * alpha ** 2)) / (abs(cellvolume) * g ** 2)
    ok = np.nonzero(w > 1e-12)[0]
    gmax_len = g[ok[-1]] * 1.001 + 1e-6 if len(ok) else 0.0
    nmax = np.minimum(np.ceil(gmax_len * np.linalg.norm(latvec, axis=1) / (2 * np.pi)).astype(int) + 1,
                      ewald_gmax)
    nx, ny, nz = [int(v) for v in nmax]
    zero = np.array([0])
    groups = [
        (np.arange(1, nx + 1), np.arange(-ny, ny + 1), np.arange(-nz, nz + 1)),
        (zero, np.arange(1, ny + 1), np.arange(-nz, nz + 1)),
        (zero, zero, np.arange(1, nz + 1)),
    ]
    sel = [_select_big(r, recvec, cellvolume, alpha) for r in groups]
    gpoints = np.concatenate([s[0] for s in sel], axis=0)
    gweight = np.concatenate([s[1] for s in sel], axis=0)

    # ewaldsum.py:92-101
    i_sum = float(np.sum(charges))
    ii_sum2 = float(np.sum(charges ** 2))
    ii_sum = (i_sum ** 2 - ii_sum2) / 2
    ijconst = -np.pi / (cellvolume * alpha ** 2)
    squareconst = -alpha / np.sqrt(np.pi) + ijconst / 2
    ii_const = ii_sum * ijconst + ii_sum2 * squareconst

    # ewaldsum.py:120-136
    if len(charges) == 1:
        ion_real = 0.0
    else:
        d = coords[:, None, :] - coords[None, :, :]
        d = _min_image(kind, latvec, invvec, mi_shifts, d)
        d = d * (1 - np.eye(len(charges)))[..., None]
        rvec = d[None] + disp[:, None, None, :]
        r = np.linalg.norm(rvec, axis=-1)
        cij = charges[:, None] * charges[None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            e = cij * erfc(alpha * r) / r
        ion_real = float(np.sum(np.triu(e, k=1)))
    gdotr = gpoints @ coords.T
    ion_exp = np.exp(1j * gdotr) @ charges
    ion_rec = float(gweight @ (np.abs(ion_exp) ** 2))

    return EwaldTables(latvec=latvec, invvec=invvec, atom_coords=coords, atom_charges=charges,
                       nelec=tuple(cell.nelec), dist_kind=kind, mi_shifts=mi_shifts,
                       lattice_displacements=disp, alpha=float(alpha), gpoints=gpoints,
                       gweight=gweight, ion_exp=ion_exp, ijconst=float(ijconst),
                       squareconst=float(squareconst), ii_const=float(ii_const),
                       ion_ion=ion_real + ion_rec, i_sum=i_sum)

"""Host-side cell geometry for the local-energy hot path (numpy only, no pyscf).

The reference hands a ``pyscf.pbc.gto.Cell`` decorated by ``supercell.get_supercell``
to every hot-path function.  The hot path only ever reads a handful of attributes
from it; this module provides a plain record with exactly those attributes and
the numpy arithmetic that produces them:

* ``lattice_vectors() / a``, ``atom_coords()``, ``atom_charges()``, ``nelec``,
  ``reciprocal_vectors()``           (pyscf Cell API used at ewaldsum.py:40-43,
                                      network.py:278-296,643, supercell.py:106-107)
* ``original_cell, S, scale``        (supercell.py:90-92)
* ``AV, BV`` on both cells           (supercell.py:98-140, ``sym_type='minimal'``
                                      plus the fcc/bcc/hexagonal row sets)
* supercell k-points                 (supercell.py:32-48)

All lengths are Bohr.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np

ANGSTROM_BOHR = 0.52917721067  # reference: utils/units.py:25


def angstrom2bohr(x):
    return x / ANGSTROM_BOHR


@dataclass
class Cell:
    """The subset of ``pyscf.pbc.gto.Cell`` the hot path reads."""

    a: np.ndarray                       # (3,3) rows = lattice vectors, Bohr
    coords: np.ndarray                  # (A,3) Bohr
    charges: np.ndarray                 # (A,)  (effective) nuclear charges
    nelec: Tuple[int, int]              # (n_up, n_dn)
    symbols: Sequence[str] = ()
    # attributes attached by get_supercell / set_symmetry_lat in the reference
    original_cell: Optional["Cell"] = None
    S: Optional[np.ndarray] = None
    scale: int = 1
    AV: Optional[np.ndarray] = None
    BV: Optional[np.ndarray] = None
    name: str = ""
    extra: dict = field(default_factory=dict)

    def __post_init__(self):
        self.a = np.asarray(self.a, dtype=np.float64).reshape(3, 3)
        self.coords = np.asarray(self.coords, dtype=np.float64).reshape(-1, 3)
        self.charges = np.asarray(self.charges, dtype=np.float64).reshape(-1)
        if self.coords.shape[0] != self.charges.shape[0]:
            raise ValueError("coords and charges disagree on the number of atoms")
        self.nelec = (int(self.nelec[0]), int(self.nelec[1]))

    # --- pyscf-compatible accessors -------------------------------------
    def lattice_vectors(self) -> np.ndarray:
        return self.a

    def atom_coords(self) -> np.ndarray:
        return self.coords

    def atom_charges(self) -> np.ndarray:
        return self.charges

    def reciprocal_vectors(self) -> np.ndarray:
        # pyscf: b = 2 pi inv(a).T  (a_i . b_j = 2 pi delta_ij)
        return 2.0 * np.pi * np.linalg.inv(self.a).T

    @property
    def nelectron(self) -> int:
        return self.nelec[0] + self.nelec[1]

    @property
    def natm(self) -> int:
        return self.coords.shape[0]

    @property
    def vol(self) -> float:
        return float(abs(np.linalg.det(self.a)))


_SYM_MATS = {
    # supercell.py:109-129
    "minimal": np.eye(3),
    "fcc": np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]], dtype=float),
    "bcc": np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, -1, 0], [1, 0, -1], [0, 1, -1]], dtype=float),
    "hexagonal": np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, -1, 0]], dtype=float),
}


def set_symmetry_lat(supercell: Cell, sym_type: str = "minimal") -> Cell:
    """AV/BV for both cells (reference supercell.py:98-140)."""
    mat = _SYM_MATS.get(sym_type, np.eye(3))
    prim_bv = mat @ supercell.original_cell.reciprocal_vectors()
    sim_bv = mat @ supercell.reciprocal_vectors()
    supercell.BV = sim_bv
    supercell.AV = np.linalg.pinv(sim_bv).T
    supercell.original_cell.BV = prim_bv
    supercell.original_cell.AV = np.linalg.pinv(prim_bv).T
    return supercell


def _integer_matrix(S) -> np.ndarray:
    Si = np.rint(np.asarray(S, dtype=float)).astype(np.int64).reshape(3, 3)
    if not np.allclose(Si, np.asarray(S, dtype=float), atol=1e-9) or round(float(np.linalg.det(Si))) == 0:
        raise ValueError("the supercell matrix S must be a non-singular integer matrix")
    return Si


def _adjugate(Si: np.ndarray) -> np.ndarray:
    """Integer adjugate: Si @ adj == det(Si) * I, exactly."""
    adj = np.empty((3, 3), dtype=np.int64)
    for r in range(3):
        for c in range(3):
            m = np.delete(np.delete(Si, c, axis=0), r, axis=1)
            adj[r, c] = (-1) ** (r + c) * (m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0])
    return adj


def _coset_representatives(M: np.ndarray) -> np.ndarray:
    """Integer row vectors n with n . inv(M) in [0, 1)^3, i.e. one representative of every coset of the lattice
    spanned by the rows of M inside Z^3, in lexicographic order of n.

    The reference finds the same points (in the same order) by scanning the bounding box of the M-image of the
    unit cube and testing the floating-point fractional coordinates against [0, 1 - 1e-12)
    (supercell.py:36-45 with M = S^T, :52-60 with M = S); here membership is decided in exact integer
    arithmetic: n . adj(M) is det(M) times the fractional coordinate.
    """
    det = int(round(float(np.linalg.det(M))))
    adj = _adjugate(M) * (1 if det > 0 else -1)
    corners = np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)], dtype=np.int64) @ M
    lo, hi = corners.min(axis=0), corners.max(axis=0)
    reps = []
    for n0 in range(lo[0], hi[0]):
        for n1 in range(lo[1], hi[1]):
            for n2 in range(lo[2], hi[2]):
                f = np.array([n0, n1, n2], dtype=np.int64) @ adj
                if np.all(f >= 0) and np.all(f < abs(det)):
                    reps.append((n0, n1, n2))
    if len(reps) != abs(det):
        raise ValueError(f"found {len(reps)} cosets for |det S| = {abs(det)}")
    return np.array(reps, dtype=np.int64).reshape(-1, 3)


def get_supercell_copies(latvec: np.ndarray, S: np.ndarray) -> np.ndarray:
    """Translations R = n . latvec of the primitive cell that tile the supercell S . latvec (what
    supercell.py:51-61 returns; |det S| of them)."""
    return _coset_representatives(_integer_matrix(S)).astype(float) @ np.asarray(latvec, dtype=float)


def get_supercell_kpts(supercell: Cell) -> np.ndarray:
    """Supercell reciprocal-lattice points inside the primitive reciprocal unit cell (what supercell.py:32-48
    returns): k = (m . inv(S)^T) . B_prim for the coset representatives m of S^T."""
    Si = _integer_matrix(supercell.S)
    m = _coset_representatives(Si.T).astype(float)
    frac = m @ np.linalg.inv(Si.astype(float)).T
    return frac @ supercell.original_cell.reciprocal_vectors()


def get_supercell(cell: Cell, S, sym_type: str = "minimal", spin: int = 0) -> Cell:
    """Simulation cell = S x primitive cell (reference supercell.py:64-95).

    ``spin`` is the primitive cell's n_up - n_dn; the supercell carries spin*scale.
    """
    S = np.asarray(S, dtype=float).reshape(3, 3)
    scale = int(abs(int(np.round(np.linalg.det(S)))))
    superlattice = np.dot(S, cell.lattice_vectors())
    rpts = get_supercell_copies(cell.lattice_vectors(), S)
    coords, charges, symbols = [], [], []
    syms = list(cell.symbols) if len(cell.symbols) else ["X"] * cell.natm
    for sym, xyz, z in zip(syms, cell.coords, cell.charges):
        for R in rpts:
            coords.append(xyz + R)
            charges.append(z)
            symbols.append(sym)
    nel = int(round(float(np.sum(cell.charges)))) * scale
    sp = spin * scale
    if (nel + sp) % 2:
        raise ValueError("electron count and spin are inconsistent")
    sc = Cell(a=superlattice, coords=np.array(coords), charges=np.array(charges),
              nelec=((nel + sp) // 2, (nel - sp) // 2), symbols=symbols,
              name=cell.name)
    sc.original_cell = cell
    sc.S = S
    sc.scale = scale
    sc.extra["sym_type"] = sym_type
    return set_symmetry_lat(sc, sym_type)


def make_klist(simulation_cell: Cell, twist=(0.0, 0.0, 0.0)):
    """Synthetic stand-in for ``hf.SCF.klist`` (hf.py:84-104).

    The reference occupies Hartree-Fock orbitals per supercell k-point and repeats
    each k-point once per occupied orbital.  Without an SCF we occupy ``n_s // nk``
    orbitals on every k-point and give the remainder to the lowest-index k-points
    (SURVEY section 8d).  ``twist`` is in fractions of the supercell reciprocal
    vectors (base_config.py:140, process.py).
    """
    kpts = get_supercell_kpts(simulation_cell)
    tw = np.asarray(twist, dtype=float) @ simulation_cell.reciprocal_vectors()
    kpts = kpts + tw[None, :]
    nk = kpts.shape[0]
    out = []
    for ns in simulation_cell.nelec:
        base, rem = divmod(ns, nk)
        occ = [base + (1 if k < rem else 0) for k in range(nk)]
        rows = [np.tile(kpts[k][None, :], (occ[k], 1)) for k in range(nk) if occ[k] > 0]
        out.append(np.concatenate(rows, axis=0) if rows else np.zeros((0, 3)))
    return out


# ---------------------------------------------------------------------------
# The five BASELINE.json systems (SURVEY section 8d defines the synthetic contract)
# ---------------------------------------------------------------------------

def hydrogen_chain(n_cells: int = 5, L: float = 2.0) -> Cell:
    """config/two_hydrogen_cell.py:27-43 with ``H,n,1,1,L,0,ccpvdz``."""
    prim = Cell(a=np.diag([2 * L, 100.0, 100.0]),
                coords=[[L, 0, 0], [0, 0, 0]], charges=[1.0, 1.0], nelec=(1, 1),
                symbols=["H", "H"], name="H-chain")
    return get_supercell(prim, np.diag([n_cells, 1, 1]))


def bcc_lithium(S=(2, 2, 1)) -> Cell:
    """config/poscar/bcc_li.vasp (cubic a = 3.4268178940 A, 2 Li) tiled by diag(S)."""
    a = angstrom2bohr(3.4268178940)
    h = angstrom2bohr(1.713408947)
    prim = Cell(a=np.eye(3) * a, coords=[[0, 0, 0], [h, h, h]], charges=[3.0, 3.0],
                nelec=(3, 3), symbols=["Li", "Li"], name="bcc-Li")
    return get_supercell(prim, np.diag(S))


def graphene(S=(3, 3, 1), L_ang: float = 2.46, z_bohr: float = 20.0, Z: float = 3.0) -> Cell:
    """config/graphene.py:25-37 lattice; effective charge Z per atom."""
    L = angstrom2bohr(L_ang)
    prim = Cell(a=[[L * np.cos(np.pi / 6), -L * 0.5, 0],
                   [L * np.cos(np.pi / 6), L * 0.5, 0],
                   [0, 0, z_bohr]],
                coords=[[3 ** (-0.5) * L, 0, 0], [2 * 3 ** (-0.5) * L, 0, 0]],
                charges=[Z, Z], nelec=(int(Z), int(Z)), symbols=["C", "C"], name="graphite")
    return get_supercell(prim, np.diag(S))


def diamond(S: int = 2, L_ang: float = 3.567, Z: float = 4.0) -> Cell:
    """config/diamond.py:24-33 fcc primitive cell, ECP-like charge Z."""
    L = angstrom2bohr(L_ang)
    prim = Cell(a=(np.ones((3, 3)) - np.eye(3)) * L / 2,
                coords=[[0, 0, 0], [0.25 * L] * 3], charges=[Z, Z],
                nelec=(int(Z), int(Z)), symbols=["C", "C"], name="diamond")
    return get_supercell(prim, np.eye(3) * S)


def rock_salt(S: int = 3, L_ang: float = 4.0, Zx: float = 3.0, Zy: float = 1.0) -> Cell:
    """config/rock_salt.py:24-35 (LiH)."""
    L = angstrom2bohr(L_ang)
    ne = int(Zx + Zy)
    prim = Cell(a=(np.ones((3, 3)) - np.eye(3)) * L / 2,
                coords=[[0, 0, 0], [0.5 * L] * 3], charges=[Zx, Zy],
                nelec=(ne // 2, ne - ne // 2), symbols=["Li", "H"], name="LiH")
    return get_supercell(prim, np.eye(3) * S)


#: name -> (builder, batch, laplacian mode, partition_number); BASELINE.json ``configs``
SYSTEMS = {
    "h10": (lambda: hydrogen_chain(5, 2.0), 256, "for", 3),
    "li24": (lambda: bcc_lithium((2, 2, 1)), 4096, "for", 3),
    # the literal read_poscar.py tiling S = 2I of config/poscar/bcc_li.vasp (SURVEY section 8d, config 2 note)
    "li48": (lambda: bcc_lithium((2, 2, 2)), 4096, "for", 3),
    "graphite54": (lambda: graphene((3, 3, 1)), 4096, "for", 3),
    "diamond64": (lambda: diamond(2), 4096, "partition", 3),
    "lih108": (lambda: rock_salt(3), 2048, "for", 3),
    # small extra systems for tests
    "h4": (lambda: hydrogen_chain(2, 2.0), 64, "for", 3),
    "lih_prim": (lambda: rock_salt(1), 64, "for", 3),
    "graphene8": (lambda: graphene((2, 1, 1), Z=2.0), 64, "for", 3),
}


def build_system(name: str) -> Cell:
    if name not in SYSTEMS:
        raise ValueError(f"unknown system {name!r}; choose from {sorted(SYSTEMS)}")
    cell = SYSTEMS[name][0]()
    cell.name = name
    return cell


def init_walkers(cell: Cell, batch: int, seed: int = 666, init_width: float = 0.8) -> np.ndarray:
    """Gaussian blobs around atoms, wrapped into the cell (init_guess.py:69-80).

    Electrons are assigned round-robin over the simulation-cell atoms, spin-up block
    first.  Returns (batch, 3N) float64.
    """
    rng = np.random.default_rng(seed)
    pos = []
    for ns in cell.nelec:
        for e in range(ns):
            pos.append(cell.coords[e % cell.natm])
    pos = np.concatenate(pos)
    guess = pos[None, :] + init_width * rng.standard_normal((batch, pos.size))
    frac = guess.reshape(batch, -1, 3) @ np.linalg.inv(cell.a)
    frac = frac - np.floor(frac)
    return (frac @ cell.a).reshape(batch, -1)

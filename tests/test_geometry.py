"""CPU: the product's host-side geometry and Ewald tables (deepsolid_b200.cell / ewald_tables, written in their own
form) against the oracle's literal restatement of supercell.py / ewaldsum.py / distance.py (oracle/geometry.py,
oracle.deepsolid_oracle.EwaldSum), which shares no code with the product."""
import numpy as np
import pytest
import torch

from deepsolid_b200 import cell as C
from deepsolid_b200.ewald_tables import build_ewald_tables
from oracle import deepsolid_oracle as O
from oracle import geometry as G

NAMES = ["h10", "h4", "li24", "li48", "graphite54", "graphene8", "diamond64", "lih108", "lih_prim"]


def _same_cell(a, b, tol=1e-13):
    assert np.abs(a.lattice_vectors() - b.lattice_vectors()).max() < tol
    assert a.atom_coords().shape == b.atom_coords().shape
    assert np.abs(a.atom_coords() - b.atom_coords()).max() < tol            # same atoms in the same order
    assert np.array_equal(a.atom_charges(), b.atom_charges()) and tuple(a.nelec) == tuple(b.nelec)
    for x, y in [(a, b), (a.original_cell, b.original_cell)]:
        assert np.abs(x.AV - y.AV).max() < tol and np.abs(x.BV - y.BV).max() < tol
    assert a.scale == b.scale and np.array_equal(np.asarray(a.S), np.asarray(b.S))


@pytest.mark.parametrize("name", NAMES)
def test_named_systems_match_the_reference_construction(name):
    a, b = C.build_system(name), G.build_system(name)
    _same_cell(a, b)
    assert np.abs(C.get_supercell_kpts(a) - G.get_supercell_kpts(b)).max() < 1e-14
    for tw in [(0.0, 0.0, 0.0), (0.25, 0.0, 0.5)]:
        for x, y in zip(C.make_klist(a, tw), G.make_klist(b, tw)):
            assert x.shape == y.shape and np.abs(x - y).max() < 1e-14
    assert np.abs(C.init_walkers(a, 5, seed=9) - G.init_walkers(b, 5, seed=9)).max() < 1e-12


@pytest.mark.parametrize("S", [np.diag([2, 3, 1]), [[1, 1, 0], [-1, 1, 0], [0, 0, 2]], [[2, -1, 0], [1, 1, 1], [0, 1, -2]],
                               [[0, 1, 1], [1, 0, 1], [1, 1, 0]], [[1, 0, 0], [2, 3, 0], [-1, 1, 2]]])
@pytest.mark.parametrize("sym_type", ["minimal", "fcc", "bcc", "hexagonal"])
def test_general_supercell_matrices(S, sym_type):
    lat = np.array([[3.1, 0.2, 0.0], [0.4, 2.9, 0.1], [-0.3, 0.5, 3.4]])
    atoms = [[0.0, 0.0, 0.0], [1.0, 1.2, 0.7]]
    a = C.get_supercell(C.Cell(a=lat, coords=atoms, charges=[2.0, 2.0], nelec=(2, 2), symbols=["X", "Y"]), S, sym_type)
    b = G.get_supercell(G.RefCell(lat, [("X", atoms[0]), ("Y", atoms[1])], {"X": 2.0, "Y": 2.0}), np.asarray(S), sym_type)
    _same_cell(a, b, tol=1e-12)
    assert np.abs(C.get_supercell_kpts(a) - G.get_supercell_kpts(b)).max() < 1e-13


def test_non_integer_supercell_matrix_raises():
    prim = C.Cell(a=np.eye(3), coords=[[0, 0, 0]], charges=[2.0], nelec=(1, 1))
    with pytest.raises(ValueError):
        C.get_supercell(prim, np.diag([1.5, 1, 1]))
    with pytest.raises(ValueError):
        C.get_supercell(prim, np.diag([1, 0, 1]))


LATTICES = {
    "orthogonal": (np.array([[3.0, 3.0, 0.0], [-2.0, 2.0, 0.0], [0.0, 0.0, 5.0]]), 1),       # orthogonal, not diagonal
    "obtuse": (np.array([[4.0, 0.0, 0.0], [-1.0, 4.2, 0.0], [-0.5, -0.8, 5.0]]), 1),          # the reference calls it orthogonal
    "general": (np.array([[4.0, 0.3, 0.1], [1.2, 4.5, 0.2], [0.5, 0.7, 5.0]]), 2),
    "diagonal": (np.diag([4.0, 5.0, 6.0]), 0),
}


@pytest.mark.parametrize("name", NAMES + list(LATTICES))
def test_ewald_tables_match_the_literal_enumeration(name):
    if name in LATTICES:
        lat, kind = LATTICES[name]
        atoms = np.array([[0.1, 0.2, 0.3], [2.0, 2.5, 2.2]])
        a = C.Cell(a=lat, coords=atoms, charges=[2.0, 1.0], nelec=(2, 1))
        b = G.RefCell(lat, [("X", atoms[0]), ("Y", atoms[1])], {"X": 2.0, "Y": 1.0}, spin=1)
    else:
        a, b, kind = C.build_system(name), G.build_system(name), None
    tb = build_ewald_tables(a)                    # product: box just covering the weight > 1e-12 sphere
    ew = O.EwaldSum(b)                            # oracle: every integer triple up to ewald_gmax = 200
    if kind is not None:
        assert tb.dist_kind == ew.dist.kind == kind
    assert tb.dist_kind == ew.dist.kind
    assert abs(tb.alpha - ew.alpha) < 1e-14 * ew.alpha
    assert tb.gpoints.shape == tuple(ew.gpoints.shape)
    assert np.abs(tb.gpoints - ew.gpoints.numpy()).max() < 1e-12            # same points in the same order
    assert np.abs(tb.gweight - ew.gweight.numpy()).max() < 1e-15
    assert np.abs(tb.ion_exp - ew.ion_exp.numpy()).max() < 1e-10
    assert np.abs(tb.lattice_displacements - ew.lattice_displacements.numpy()).max() < 1e-13
    assert np.abs(tb.mi_shifts - ew.dist.shifts.numpy()).max() < 1e-13
    ne = sum(a.nelec)
    assert abs(tb.ee_const(ne) - ew.ee_const(ne)) < 1e-12 and abs(tb.ei_const(ne) - ew.ei_const(ne)) < 1e-12
    assert abs(tb.ii_total - (ew.ion_ion + ew.ii_const)) < 1e-10

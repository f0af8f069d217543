"""Known answers for the Ewald setup/energy (replaces the pyscf energy_nuc assertion of
hamiltonian.py:170-172): Madelung energies of point-charge lattices in a neutralising
background, the NaCl Madelung constant, alpha-independence against a naive Ewald sum."""
import math
import numpy as np
import pytest
import torch
from scipy.special import erfc

from deepsolid_b200.cell import Cell
from deepsolid_b200.ewald_tables import build_ewald_tables, classify_lattice
from oracle import deepsolid_oracle as O

# Wigner-crystal Madelung energies, Rydberg * r_s  (Fuchs 1935; Ceperley & Alder tables)
WIGNER = {"sc": -1.760118890, "bcc": -1.791858511, "fcc": -1.791747230}


def _lattice(kind, a=3.0):
    if kind == "sc":
        return np.eye(3) * a
    if kind == "bcc":
        return np.array([[-1, 1, 1], [1, -1, 1], [1, 1, -1]]) * a / 2
    return (np.ones((3, 3)) - np.eye(3)) * a / 2


@pytest.mark.parametrize("kind", ["sc", "bcc", "fcc"])
def test_wigner_madelung(kind):
    lat = _lattice(kind)
    c = Cell(a=lat, coords=[[0, 0, 0]], charges=[1.0], nelec=(1, 0))
    tb = build_ewald_tables(c)
    vol = abs(np.linalg.det(lat))
    rs = (3 * vol / (4 * math.pi)) ** (1 / 3)
    assert abs(tb.ii_total - WIGNER[kind] / 2 / rs) < 2e-7


def test_nacl_madelung():
    a = 5.0
    c = Cell(a=(np.ones((3, 3)) - np.eye(3)) * a / 2, coords=[[0, 0, 0], [a / 2] * 3], charges=[1.0, -1.0], nelec=(1, 1))
    tb = build_ewald_tables(c)
    assert abs(tb.ii_total - (-1.7475645946331822 / (a / 2))) < 1e-8


def test_lattice_classification_quirk():
    """distance.py:49-53 tests dot < tol WITHOUT abs: obtuse lattices count as 'orthogonal'."""
    assert classify_lattice(np.diag([1.0, 2, 3])) == 0
    assert classify_lattice(np.array([[1.0, 1, 0], [-1, 1, 0], [0, 0, 1]])) == 1
    assert classify_lattice(np.array([[1.0, 0.2, 0], [0.2, 1, 0], [0, 0, 1]])) == 2
    assert classify_lattice(np.array([[1.0, 0, 0], [-0.5, 1, 0], [0, 0, 1]])) == 1      # obtuse: misclassified like the reference


def _naive_ewald_total(lat, pos, q, alpha, nreal=4, nrec=9):
    """Plain Ewald sum of a neutral set of point charges (independent of the restated tables)."""
    vol = abs(np.linalg.det(lat))
    rec = 2 * np.pi * np.linalg.inv(lat).T
    n = len(q)
    rng = np.arange(-nreal, nreal + 1)
    R = np.stack(np.meshgrid(rng, rng, rng, indexing="ij"), -1).reshape(-1, 3) @ lat
    e = 0.0
    for i in range(n):
        for j in range(n):
            d = pos[i] - pos[j] + R
            r = np.linalg.norm(d, axis=1)
            if i == j:
                r = r[r > 1e-12]
            e += 0.5 * q[i] * q[j] * np.sum(erfc(alpha * r) / r)
    rng = np.arange(-nrec, nrec + 1)
    G = np.stack(np.meshgrid(rng, rng, rng, indexing="ij"), -1).reshape(-1, 3) @ rec
    g2 = (G ** 2).sum(1)
    G, g2 = G[g2 > 1e-12], g2[g2 > 1e-12]
    S = (q[None, :] * np.exp(1j * G @ pos.T)).sum(1)
    e += 2 * np.pi / vol * np.sum(np.exp(-g2 / (4 * alpha ** 2)) / g2 * np.abs(S) ** 2)
    e -= alpha / math.sqrt(math.pi) * np.sum(q ** 2)
    e -= np.pi / (2 * vol * alpha ** 2) * np.sum(q) ** 2
    return e


@pytest.mark.parametrize("kind", ["diag", "general"])
def test_total_energy_matches_naive_ewald(kind):
    if kind == "diag":
        lat = np.diag([4.0, 5.0, 6.0])
    else:
        lat = np.array([[4.0, 0.3, 0.1], [1.2, 4.5, 0.2], [0.5, 0.7, 5.0]])
    atoms = np.array([[0.1, 0.2, 0.3], [2.0, 2.5, 2.2]])
    c = Cell(a=lat, coords=atoms, charges=[2.0, 1.0], nelec=(2, 1))
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 1, (3, 3)) @ lat
    ew = O.EwaldSum(c)
    ee, ei, ii = ew.energy(torch.as_tensor(x.reshape(-1)))
    pos = np.concatenate([x, atoms])
    q = np.array([-1.0, -1.0, -1.0, 2.0, 1.0])
    for alpha in (0.9, 1.3):
        ref = _naive_ewald_total(lat, pos, q, alpha)
        assert abs(float(ee + ei + ii) - ref) < 1e-8

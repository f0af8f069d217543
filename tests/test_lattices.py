"""Minimal-image branches of distance.MinimalImageDistance (distance.py:41-59, 91-108) that none of the BASELINE cells
reaches: an orthogonal-but-not-diagonal lattice and an OBTUSE lattice, which the reference classifies as 'orthogonal'
because it tests `dot < tol` without an absolute value (and then applies the fractional-coordinate wrap to a lattice
where that is not the minimal image -- reproduced, not corrected).  GPU: Ewald terms, log psi and the local energy
against the oracle; CPU: the oracle really takes the branch and differs from the general search on the obtuse cell."""
import numpy as np
import pytest
import torch

from deepsolid_b200 import cell as C
from deepsolid_b200 import network, hamiltonian
from oracle import deepsolid_oracle as O

LATTICES = {
    "orthogonal": np.array([[3.0, 3.0, 0.0], [-2.0, 2.0, 0.0], [0.0, 0.0, 5.0]]),
    "obtuse": np.array([[4.0, 0.0, 0.0], [-1.9, 4.2, 0.0], [-1.5, -1.8, 5.0]]),
}


def build(kind, S=(2, 1, 1)):
    lat = LATTICES[kind]
    frac = np.array([[0.1, 0.15, 0.2], [0.6, 0.55, 0.7]])
    prim = C.Cell(a=lat, coords=frac @ lat, charges=[2.0, 2.0], nelec=(2, 2), symbols=["He", "He"], name=kind)
    sc = C.get_supercell(prim, np.diag(S))
    sc.name = kind
    kl = C.make_klist(sc)
    pn = O.init_params(np.random.default_rng(31), prim.natm, sc.nelec)
    return sc, kl, O.params_to_torch(pn)


@pytest.mark.parametrize("kind", ["orthogonal", "obtuse"])
def test_oracle_takes_the_orthogonal_branch(kind):
    sc, kl, P = build(kind)
    ew = O.EwaldSum(sc)
    assert ew.dist.kind == 1
    X = torch.as_tensor(C.init_walkers(sc, 6, seed=3, init_width=1.5))
    gen = O.MinimalImageDistance(sc.lattice_vectors())
    gen.kind = 2                                   # what a corrected classification would use
    differs = 0
    for b in range(6):
        a = ew.dist.dist_i(ew.atom_coords.reshape(-1), X[b])
        g = gen.dist_i(ew.atom_coords.reshape(-1), X[b])
        # both are lattice images of the same displacement ...
        fr = (a - g) @ torch.linalg.inv(torch.as_tensor(sc.lattice_vectors()))
        assert float((fr - fr.round()).abs().max()) < 1e-10
        differs += int((torch.linalg.norm(a, dim=-1) > torch.linalg.norm(g, dim=-1) + 1e-9).sum())
    # ... equal on the truly orthogonal cell, NOT always minimal on the obtuse one (the quirk is observable)
    assert (differs == 0) if kind == "orthogonal" else (differs > 0)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["orthogonal", "obtuse"])
def test_gpu_orthogonal_branch_matches_oracle(kind):
    sc, kl, P = build(kind)
    dev = torch.device("cuda", 0)
    kw = dict(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8)
    ld = network.make_solid_fermi_net(method_name="eval_logdet", **kw)
    hp = ld.apply.hotpath()
    assert hp.tables.dist_kind == 1
    X = torch.as_tensor(C.init_walkers(sc, 6, seed=3, init_width=1.5))
    ke, ew = hamiltonian.local_energy_seperate(ld.apply, sc, mode="for")(P, X.to(dev))
    ee, ei, ii = hp.ewald(X.to(dev))
    v = ld.apply(P, X.to(dev)).cpu()
    f = O.make_solid_fermi_net(kl, sc, determinants=8, method_name="eval_logdet")
    elo = O.local_energy_seperate(f, sc, mode="dim_batch")
    oew = O.EwaldSum(sc)
    for b in range(6):
        ko, eo = elo(P, X[b])
        e0, e1, e2 = oew.energy(X[b])
        assert abs(float(e0) - float(ee[b])) < 1e-10 and abs(float(e1) - float(ei[b])) < 1e-10
        assert abs(float(e2) - float(torch.as_tensor(ii).reshape(-1)[0])) < 1e-10
        assert abs(float(eo) - float(ew[b])) < 1e-10
        assert abs(complex(ko) - complex(ke[b].cpu())) < 1e-8
        assert abs(float(f(P, X[b]).real) - float(v[b].real)) < 1e-10

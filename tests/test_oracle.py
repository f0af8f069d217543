"""CPU tests of the oracle: the reference's own invariants (test/test_network.py:65-122),
equivalence of the Laplacian modes (hamiltonian.py:45-159), finite differences, the
independent forward-Laplacian derivation, and the committed golden vectors."""
import math
import numpy as np
import pytest
import torch

from conftest import system, golden, angle_diff
from deepsolid_b200 import cell as C
from oracle import deepsolid_oracle as O
from oracle import forward_laplacian as FL


@pytest.mark.parametrize("name", ["lih_prim", "graphene8"])
def test_periodic_bc(name):
    """test_network.py:65-83: translating ALL electrons by a primitive lattice vector keeps
    log|psi| and multiplies the phase by exp(i sum_k k.t)."""
    sc, kl, _, P = system(name)
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet")
    x = torch.as_tensor(C.init_walkers(sc, 1, seed=3)[0])
    t = torch.as_tensor(sc.original_cell.lattice_vectors()[2])
    p1, s1 = f(P, x)
    p2, s2 = f(P, x + t.repeat(sc.nelectron))
    kp = sum(torch.as_tensor(k).sum(0) for k in kl)
    assert abs(float(s1 - s2)) < 1e-10
    assert abs(complex(p1 * torch.exp(1j * torch.dot(kp, t)) - p2)) < 1e-10


@pytest.mark.parametrize("name", ["lih_prim", "graphene8"])
def test_twisted_bc(name):
    """test_network.py:86-106 with twist 0: moving ONE electron by a supercell vector is a symmetry."""
    sc, kl, _, P = system(name)
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet")
    x = torch.as_tensor(C.init_walkers(sc, 1, seed=4)[0])
    x2 = x.clone()
    x2[:3] += torch.as_tensor(sc.lattice_vectors()[1])
    p1, s1 = f(P, x)
    p2, s2 = f(P, x2)
    assert abs(float(s1 - s2)) < 1e-10
    assert abs(complex(p2 / p1) - 1.0) < 1e-9


def test_antisymmetry():
    """test_network.py:109-122."""
    sc, kl, _, P = system("graphene8")
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet")
    x = torch.as_tensor(C.init_walkers(sc, 1, seed=5)[0])
    x2 = torch.cat([x[3:6], x[:3], x[6:]])
    p1, s1 = f(P, x)
    p2, s2 = f(P, x2)
    assert abs(complex(p1 + p2)) < 1e-10 and abs(float(s1 - s2)) < 1e-10


def test_laplacian_modes_agree_and_match_finite_differences():
    sc, kl, _, P = system("lih_prim")
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_logdet")
    x = torch.as_tensor(C.init_walkers(sc, 1, seed=6)[0])
    vals = [sum(O.local_kinetic_energy_real_imag(f)(P, x)),
            sum(O.local_kinetic_energy_partition(f, 3)(P, x)),
            sum(O.local_kinetic_energy_dim_batch(f)(P, x)),
            sum(O.local_kinetic_energy_hessian(f)(P, x))]
    for v in vals[1:]:
        assert abs(complex(v - vals[0])) < 1e-11
    h, lap, g2, f0 = 1e-4, 0, 0, f(P, x)
    for d in range(x.numel()):
        e = torch.zeros_like(x); e[d] = h
        fp, fm = f(P, x + e), f(P, x - e)
        lap += (fp - 2 * f0 + fm) / h ** 2
        g2 += ((fp - fm) / (2 * h)) ** 2
    assert abs(complex(-0.5 * (lap + g2) - vals[0])) < 2e-5      # O(h^2) + roundoff/h^2 of central differences
    with pytest.raises(ValueError):
        O.local_energy_seperate(f, sc, mode="nope")
    with pytest.raises(ValueError):
        O.local_kinetic_energy_partition(f, 5)(P, x)       # 5 does not divide 3N = 12


@pytest.mark.parametrize("name", ["h4", "graphene8"])
def test_forward_laplacian_matches_autodiff(name):
    """Two independent derivations of the kinetic energy (autodiff jvp-of-grad vs the analytic
    forward-Laplacian recursion the CUDA kernels implement)."""
    sc, kl, _, P = system(name)
    X = torch.as_tensor(C.init_walkers(sc, 2, seed=7))
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_logdet")
    la, ang, ke, _ = FL.kinetic_forward_laplacian(P, X, sc, kl)
    for b in range(2):
        v = f(P, X[b])
        k = sum(O.local_kinetic_energy_dim_batch(f)(P, X[b]))
        assert abs(float(v.real - la[b])) < 1e-11
        assert float(angle_diff(v.imag, ang[b])) < 1e-11
        assert abs(complex(k - ke[b])) < 1e-9


@pytest.mark.parametrize("name", ["h4", "lih_prim", "graphene8", "h10"])
def test_oracle_reproduces_golden(name):
    sc, kl, _, P = system(name)
    g = golden(name)
    X = torch.as_tensor(g["x"])
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet")
    for b in range(min(2, X.shape[0])):
        s, l = f(P, X[b])
        assert abs(float(l) - g["logabs"][b]) < 1e-11
        assert float(angle_diff(torch.angle(s), g["phase"][b])) < 1e-11
    ew = O.EwaldSum(sc)
    ee, ei, ii = ew.energy(X[0])
    assert abs(float(ee) - g["ee"][0]) < 1e-10 and abs(float(ei) - g["ei"][0]) < 1e-10
    assert abs(float(ii) - float(g["ii"])) < 1e-10
    fl_la, fl_ang, fl_ke, _ = FL.kinetic_forward_laplacian(P, X, sc, kl)
    assert np.abs(fl_ke.numpy() - g["ke"]).max() < 1e-9       # golden ke came from autodiff


def test_mcmc_oracle_reproduces_golden():
    sc, kl, _, P = system("h4")
    g = golden("h4")
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_slogdet")
    steps, B = g["xi"].shape[0], g["x"].shape[0]
    mc = O.make_mcmc_step(lambda p, xx: O.batch_apply(f, p, xx), B, sc.lattice_vectors(), steps=steps)
    xn, pmove, masks = mc(P, torch.as_tensor(g["x"]), (torch.as_tensor(g["xi"]), torch.as_tensor(g["u"])), float(g["width"]))
    assert (masks.numpy() == g["masks"]).all()
    assert np.abs(xn.numpy() - g["x_new"]).max() < 1e-12
    assert abs(float(pmove) - float(g["pmove"])) < 1e-15


def test_total_energy_stats_quirk():
    """train.py:76-80: variance subtracts |Re mean|^2 only."""
    ke = torch.tensor([1 + 2j, 3 - 1j], dtype=torch.complex128)
    ew = torch.tensor([-1.0, 0.5], dtype=torch.float64)
    loss, im, var = O.total_energy_stats(ke, ew)
    e = ke + ew
    assert abs(float(loss) - float(e.real.mean())) < 1e-15
    assert abs(float(var) - float((e.abs() ** 2).mean() - e.real.mean() ** 2)) < 1e-15

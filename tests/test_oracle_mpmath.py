"""Pins the floating-point behaviour of the torch oracle: a scalar 40-digit mpmath restatement of the network
(oracle/mp_reference.py; Laplacian by high-order central differences, no autodiff) against oracle/deepsolid_oracle.py
on a tiny system (H4 chain: 2+2 electrons, two 16/8-wide layers, 2 determinants).  SURVEY.md §8(c)."""
import numpy as np
import torch

from deepsolid_b200 import cell as C
from oracle import deepsolid_oracle as O
from oracle import mp_reference as MP


def test_torch_oracle_matches_40_digit_restatement():
    sc = C.build_system("h4")
    kl = C.make_klist(sc)
    hidden = ((16, 8), (16, 8))
    pn = O.init_params(np.random.default_rng(5), sc.original_cell.natm, sc.nelec, hidden_dims=hidden, determinants=2)
    P = O.params_to_torch(pn)
    X = C.init_walkers(sc, 2, seed=9)
    f = O.make_solid_fermi_net(kl, sc, hidden_dims=hidden, determinants=2, method_name="eval_logdet")
    ke_fun = O.local_kinetic_energy_real_imag(f)
    for x in X:
        want_f, want_ke = MP.kinetic(pn, list(x), sc, kl, sc.nelec)
        got_f = f(P, torch.as_tensor(x))
        got_ke = ke_fun(P, torch.as_tensor(x))
        got_ke = complex(got_ke[0]) + complex(got_ke[1]) if isinstance(got_ke, (tuple, list)) else complex(got_ke)
        assert abs(float(got_f.real) - float(want_f.real)) < 1e-12
        dphi = float(got_f.imag) - float(want_f.imag)
        assert abs((dphi + np.pi) % (2 * np.pi) - np.pi) < 1e-12
        assert abs(got_ke - complex(want_ke)) < 1e-12, (got_ke, complex(want_ke))     # measured 3e-15 .. 6e-15


import pytest


@pytest.mark.parametrize("distance_type,envelope_type,full_det", [
    ("tri", "isotropic", False), ("nu", "diagonal", False), ("nu", "full", False), ("nu", "isotropic", True)])
def test_torch_oracle_variants_match_40_digit_restatement(distance_type, envelope_type, full_det):
    """The secondary variants (tri features, anisotropic envelopes, full determinants) of the torch oracle against the
    same scalar mpmath restatement; envelope parameters randomised so that sigma / pi matter."""
    sc = C.build_system("h4")
    kl = C.make_klist(sc)
    hidden = ((12, 6), (12, 6))
    rng = np.random.default_rng(17)
    pn = O.init_params(rng, sc.original_cell.natm, sc.nelec, hidden_dims=hidden, determinants=2,
                       distance_type=distance_type, envelope_type=envelope_type, full_det=full_det)
    for env in pn["envelope"]:
        env["pi"] = env["pi"] * (1.0 + 0.3 * rng.standard_normal(env["pi"].shape))
        env["sigma"] = env["sigma"] + 0.2 * rng.standard_normal(env["sigma"].shape)
    P = O.params_to_torch(pn)
    x = C.init_walkers(sc, 1, seed=5)[0]
    kw = dict(hidden_dims=hidden, determinants=2, distance_type=distance_type, envelope_type=envelope_type, full_det=full_det)
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_logdet", **kw)
    want_f, want_ke = MP.kinetic(pn, list(x), sc, kl, sc.nelec, distance_type=distance_type, envelope_type=envelope_type,
                                 full_det=full_det)
    got_f = f(P, torch.as_tensor(x))
    got_ke = O.local_kinetic_energy_real_imag(f)(P, torch.as_tensor(x))
    got_ke = complex(got_ke[0]) + complex(got_ke[1]) if isinstance(got_ke, (tuple, list)) else complex(got_ke)
    assert abs(float(got_f.real) - float(want_f.real)) < 1e-12
    dphi = float(got_f.imag) - float(want_f.imag)
    assert abs((dphi + np.pi) % (2 * np.pi) - np.pi) < 1e-12
    assert abs(got_ke - complex(want_ke)) < 1e-11, (got_ke, complex(want_ke))


def test_oracle_parameter_gradient_matches_40_digit_finite_differences():
    """d log|psi| / d theta and d angle(psi) / d theta of the torch oracle (autograd, the pullback the energy-gradient
    estimator and the KFAC statistics are built on) against central differences of the mpmath restatement."""
    import copy
    import mpmath as mp
    sc = C.build_system("h4")
    kl = C.make_klist(sc)
    hidden = ((12, 6), (12, 6))
    rng = np.random.default_rng(3)
    pn = O.init_params(rng, sc.original_cell.natm, sc.nelec, hidden_dims=hidden, determinants=2)
    P = O.params_to_torch(pn)
    x = C.init_walkers(sc, 1, seed=8)
    f_ps = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet", hidden_dims=hidden, determinants=2)
    one, zero = torch.ones(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64)
    g_abs = O.logpsi_vjp(f_ps, P, torch.as_tensor(x), one, zero)
    g_ph = O.logpsi_vjp(f_ps, P, torch.as_tensor(x), zero, one)
    picks = [("single", 0, "w"), ("single", 1, "b"), ("double", 0, "w"), ("orbital", 1, "w"), ("envelope", 0, "sigma"),
             ("envelope", 1, "pi")]
    h = 1e-7                                              # parameters are fp64 inputs: the step must be representable
    for kind, i, leaf in picks:
        arr = pn[kind][i][leaf]
        idx = tuple(int(rng.integers(0, s)) for s in arr.shape)
        vals = []
        for sgn in (+1, -1, +2, -2):
            q = copy.deepcopy(pn)
            q[kind][i][leaf][idx] = arr[idx] + sgn * h
            vals.append(MP.log_psi(q, [mp.mpf(float(v)) for v in x[0]], sc, kl, sc.nelec))
        d = (-vals[2] + 8 * vals[0] - 8 * vals[1] + vals[3]) / (12 * mp.mpf(h))
        assert abs(float(d.real) - float(g_abs[kind][i][leaf][idx])) < 1e-9 * max(1.0, abs(float(d.real))), (kind, i, leaf)
        assert abs(float(d.imag) - float(g_ph[kind][i][leaf][idx])) < 1e-9 * max(1.0, abs(float(d.imag))), (kind, i, leaf)

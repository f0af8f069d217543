"""Secondary Metropolis variants of qmc.py: one-electron moves (qmc.py:227-287) and drift-diffusion importance
sampling (qmc.py:63-150), plus the position gradient they need.  Accept masks must be bit-identical to the
oracle's for identical (x1, xi, u); gradients within 1e-9."""
import numpy as np
import pytest
import torch

from conftest import system
from deepsolid_b200 import cell as C
from oracle import deepsolid_oracle as O


def test_limdrift_and_pbc_mirror_the_oracle():
    from deepsolid_b200 import qmc
    torch.manual_seed(1)
    g = 3.0 * torch.randn(5, 12, dtype=torch.float64)
    assert torch.allclose(qmc.limdrift(g), O.limdrift(g))
    n = torch.linalg.norm(qmc.limdrift(g).reshape(-1, 3), dim=-1)
    assert float(n.max()) <= 1.0 + 1e-12
    sc, *_ = system("graphene8")
    lat = torch.as_tensor(sc.lattice_vectors())
    x = 20.0 * torch.randn(4, 3 * sum(sc.nelec), dtype=torch.float64)
    assert torch.allclose(qmc.enforce_pbc(lat, x), O.enforce_pbc_batch(lat, x)[0], atol=1e-12)


def _nets(name):
    from deepsolid_b200 import network
    sc, kl, pn, P = system(name)
    kw = dict(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8)
    sl = network.make_solid_fermi_net(method_name="eval_slogdet", **kw)
    return sc, kl, P, sl


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["h4", "graphene8", "lih_prim"])
def test_gpu_position_gradient_matches_autograd(name):
    sc, kl, P, sl = _nets(name)
    hp = sl.apply.hotpath()
    hp.set_params(P)
    X = torch.as_tensor(C.init_walkers(sc, 4, seed=4))
    la, ph, ga, gp = hp.logpsi_grad_x(X.cuda(), want_phase_grad=True)
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_slogdet")
    vo, go = O.value_and_grad_x(f, P, X)
    assert float((la.cpu() - vo).abs().max()) < 1e-10
    assert float((ga.cpu() - go).abs().max()) < 1e-9 * max(1.0, float(go.abs().max()))
    fp = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet")
    vp, gpo = O.value_and_grad_x(lambda p, x: torch.angle(fp(p, x)[0]), P, X)
    assert float((gp.cpu() - gpo).abs().max()) < 1e-9 * max(1.0, float(gpo.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["h4", "graphene8"])
def test_gpu_one_electron_moves_bit_exact_masks(name):
    from deepsolid_b200 import qmc
    sc, kl, P, sl = _nets(name)
    B, steps = 6, 2
    N = sum(sc.nelec)
    X = torch.as_tensor(C.init_walkers(sc, B, seed=9))
    g = torch.Generator().manual_seed(5)
    xi = torch.randn(steps * N, B, 3, dtype=torch.float64, generator=g)
    u = torch.rand(steps * N, B, dtype=torch.float64, generator=g)
    lat = torch.as_tensor(sc.lattice_vectors())
    step = qmc.make_mcmc_step(sl.apply, B, lat, steps=steps, one_electron_moves=True)
    xn, pmove, masks = step(P, X.cuda(), (xi, u), 0.3, return_masks=True)
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_slogdet")
    ostep = O.make_mcmc_step_one_electron(lambda p, x: O.batch_apply(f, p, x), B, lat, steps=steps)
    xo, po, mo = ostep(P, X, (xi, u), 0.3)
    assert masks.shape == (steps * N, B)
    assert torch.equal(masks.cpu().bool(), mo)
    assert float((xn.cpu() - xo).abs().max()) < 1e-12
    assert abs(float(pmove) - float(po)) < 1e-15
    assert 0 < int(mo.sum()) < mo.numel()
    # device RNG path: reproducible, moves something
    a1, p1 = step(P, X.cuda(), 7, 0.3)
    a2, p2 = step(P, X.cuda(), 7, 0.3)
    assert torch.equal(a1, a2) and 0.0 < float(p1) <= 1.0


@pytest.mark.gpu
def test_gpu_importance_sampling_matches_oracle():
    from deepsolid_b200 import qmc
    sc, kl, P, sl = _nets("h4")
    B, steps = 5, 3
    X = torch.as_tensor(C.init_walkers(sc, B, seed=3))
    g = torch.Generator().manual_seed(11)
    xi = torch.randn(steps, B, X.shape[1], dtype=torch.float64, generator=g)
    u = torch.rand(steps, B, dtype=torch.float64, generator=g)
    lat = torch.as_tensor(sc.lattice_vectors())
    step = qmc.make_mcmc_step(sl.apply, B, lat, steps=steps, importance_sampling=sl.apply)
    xn, pmove, masks = step(P, X.cuda(), (xi, u), 0.2, return_masks=True)
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_slogdet")
    ostep = O.make_mcmc_step_importance(f, B, lat, steps=steps)
    xo, po, mo = ostep(P, X, (xi, u), 0.2)
    assert torch.equal(masks.cpu().bool(), mo)
    assert float((xn.cpu() - xo).abs().max()) < 1e-9
    assert abs(float(pmove) - float(po)) < 1e-15
    with pytest.raises(ValueError):
        qmc.make_mcmc_step(sl.apply, B, lat, importance_sampling=sl.apply, one_electron_moves=True)

"""Hartree-Fock pretraining (pretrain.py:43-168).  CPU: the Adam mirror against torch.optim.Adam and the
loss / cotangent algebra against autograd.  GPU: loss and parameter gradient through ds_orbitals_vjp against the
oracle's autograd (1e-9 relative), and a short pretraining run towards the free-electron determinant."""
import numpy as np
import pytest
import torch

from conftest import system
from deepsolid_b200 import cell as C
from deepsolid_b200 import pretrain
from oracle import deepsolid_oracle as O


def test_adam_matches_torch_and_loss_cotangent_is_the_gradient():
    torch.manual_seed(0)
    p = {"single": [{"w": torch.randn(3, 2, dtype=torch.float64), "b": torch.randn(2, dtype=torch.float64)}],
         "double": [], "orbital": [{"w": torch.randn(2, 4, dtype=torch.float64)}, {"w": torch.randn(2, 4, dtype=torch.float64)}],
         "envelope": [{"pi": torch.ones(1, 2, dtype=torch.float64), "sigma": torch.ones(1, 2, dtype=torch.float64)}] * 2}
    from deepsolid_b200.hotpath import flatten_params
    ref = [t.clone().requires_grad_(True) for t in flatten_params(p)]
    topt = torch.optim.Adam(ref, lr=5e-3)
    opt = pretrain.Adam(5e-3)
    state = opt.init(p)
    cur = [t.clone() for t in flatten_params(p)]
    for it in range(3):
        g = [torch.randn_like(t) for t in cur]
        for r, gg in zip(ref, g):
            r.grad = gg.clone()
        topt.step()
        from deepsolid_b200.hotpath import unflatten_params
        upd, state = opt.update(unflatten_params(g, 1, False), state)
        cur = [c + u for c, u in zip(cur, upd)]
        for a, b in zip(cur, ref):
            assert torch.allclose(a, b.detach(), rtol=1e-12, atol=1e-14)
    # cotangent: d loss = sum Re(conj(cot) d predict)
    pre = [torch.randn(4, 3, 2, 2, dtype=torch.complex128, requires_grad=True) for _ in range(2)]
    tar = [torch.randn(4, 2, 2, dtype=torch.complex128) for _ in range(2)]
    loss, cots = pretrain.pretrain_loss_cotangent(pre, tar)
    loss.backward()
    for q, c in zip(pre, cots):
        assert torch.allclose(q.grad, c.detach())          # torch's complex gradient convention is d loss / d conj(z)


def test_plane_wave_scf_shapes():
    sc, kl, pn, P = system("h4")
    scf = pretrain.PlaneWaveSCF(sc, kl)
    X = C.init_walkers(sc, 5, seed=1).reshape(5, -1, 3)
    up, dn = scf.eval_orb_mat(X)
    assert up.shape == (5, sc.nelec[0], sc.nelec[0]) and dn.shape == (5, sc.nelec[1], sc.nelec[1])
    ph, sl = scf.eval_slogdet(X)
    assert np.allclose(np.abs(ph), 1.0) and sl.shape == (5,)


@pytest.mark.gpu
@pytest.mark.parametrize("name,full_det", [("h4", False), ("graphene8", False), ("h4", True)])
def test_gpu_pretrain_loss_and_gradient_match_oracle(name, full_det):
    from deepsolid_b200 import network
    sc, kl, pn, P = system(name)
    if full_det:
        P = O.params_to_torch(O.init_params(np.random.default_rng(888), sc.original_cell.natm, sc.nelec, full_det=True))
    mats = network.make_solid_fermi_net(envelope_type="isotropic", full_det=full_det, klist=kl, simulation_cell=sc,
                                        determinants=8, method_name="eval_mats")
    hp = mats.apply.hotpath()
    X = torch.as_tensor(C.init_walkers(sc, 4, seed=13))
    scf = pretrain.PlaneWaveSCF(sc, kl)
    target = [torch.as_tensor(m) for m in scf.eval_orb_mat(X.numpy().reshape(4, -1, 3))]
    hp.set_params(P)
    loss, cots = pretrain.pretrain_loss_cotangent(hp.orbitals(X.cuda()), [t.cuda() for t in target], full_det)
    got = hp.orbitals_vjp(X.cuda(), cots)
    f = O.make_solid_fermi_net(kl, sc, full_det=full_det, method_name="eval_mats")
    want_loss, want = O.pretrain_loss_and_grad(f, P, X, target, full_det)
    assert abs(float(loss) - float(want_loss)) < 1e-12 * max(1.0, abs(float(want_loss)))
    for a, b in zip(O._leaves(got), O._leaves(want)):
        assert float((a.cpu() - b).abs().max()) <= 1e-9 * max(float(b.abs().max()), 1e-30)


@pytest.mark.gpu
def test_gpu_pretraining_lowers_the_loss():
    from deepsolid_b200 import network
    sc, kl, pn, P = system("h4")
    kw = dict(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8)
    mats = network.make_solid_fermi_net(method_name="eval_mats", **kw)
    hp = mats.apply.hotpath()
    slog = network.make_solid_fermi_net(method_name="eval_slogdet", hotpath=hp, **kw)
    data = torch.as_tensor(C.init_walkers(sc, 64, seed=3)).cuda()
    lines = []
    params, data = pretrain.pretrain_hartree_fock(P, data, slog.apply, mats.apply, 7, sc, pretrain.PlaneWaveSCF(sc, kl),
                                                  iterations=40, learning_rate=5e-3, log=lines.append)
    losses = [float(l.split("Loss=")[1].split(",")[0]) for l in lines]
    assert len(losses) == 40 and all(np.isfinite(losses))
    assert np.mean(losses[-5:]) < 0.7 * np.mean(losses[:5]), (losses[:5], losses[-5:])
    assert torch.isfinite(data).all() and data.shape == (64, 3 * sum(sc.nelec))

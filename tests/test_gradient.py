"""Energy-gradient estimator (train.py:91-142).  CPU: the oracle's reverse-mode pullback against central finite
differences of its own log psi, and the clipping rules.  GPU: ds_logpsi_vjp (through the C ABI) against the oracle
on the same walkers, cotangents and parameters; tolerance 1e-9 relative to the largest entry of each leaf."""
import numpy as np
import pytest
import torch

from conftest import system
from deepsolid_b200 import cell as C
from oracle import deepsolid_oracle as O


def _flat(tree):
    return O._leaves(tree)


def test_oracle_vjp_matches_finite_differences():
    sc, kl, pn, P = system("h4")
    X = torch.as_tensor(C.init_walkers(sc, 2, seed=11))
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet")
    ca = torch.tensor([0.7, -0.4], dtype=torch.float64)
    cp = torch.tensor([-0.2, 0.9], dtype=torch.float64)
    g = O.logpsi_vjp(f, P, X, ca, cp)

    def total(params):
        t = 0.0
        for b, x in enumerate(X):
            sign, slog = f(params, x)
            t += float(ca[b] * slog) + float(cp[b] * torch.angle(sign))
        return t

    rng = np.random.default_rng(0)
    P2 = O._clone_params(P, requires_grad=False)
    for leaf, gleaf in zip(_flat(P2), _flat(g)):
        idx = tuple(int(rng.integers(0, s)) for s in leaf.shape)
        h = 1e-6
        old = float(leaf[idx])
        leaf[idx] = old + h
        up = total(P2)
        leaf[idx] = old - h
        dn = total(P2)
        leaf[idx] = old
        fd = (up - dn) / (2 * h)
        assert abs(fd - float(gleaf[idx])) < 1e-6 * max(1.0, abs(fd)), (idx, fd, float(gleaf[idx]))


def test_clip_rules():
    torch.manual_seed(0)
    d = torch.complex(torch.randn(64, dtype=torch.float64), 0.1 * torch.randn(64, dtype=torch.float64))
    d[3] = 50.0 + 7.0j
    r = O.clip_difference(d, 5.0, "real")
    assert float(r.real.abs().max()) <= 5.0 * float(d.real.abs().mean()) + 1e-12
    assert float(r.imag.abs().max()) <= 5.0 * float(d.imag.abs().mean()) + 1e-12
    c = O.clip_difference(d, 5.0, "complex")
    assert torch.allclose(torch.angle(c), torch.angle(d))
    assert float(c.abs().max()) < float(d.abs().max())
    assert torch.equal(O.clip_difference(d, 0.0, "real"), d)
    with pytest.raises(ValueError):
        O.clip_difference(d, 5.0, "polar")
    # the product mirror applies the same rules (host tensors, single process)
    from deepsolid_b200 import train
    assert torch.allclose(train.clip_difference(d, 5.0, "real"), r)
    assert torch.allclose(train.clip_difference(d, 5.0, "complex"), c)
    with pytest.raises(ValueError):
        train.clip_difference(d, 5.0, "polar")


def _rel_err(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-300)


@pytest.mark.gpu
@pytest.mark.parametrize("name,nw", [("h4", 5), ("lih_prim", 4), ("graphene8", 3), ("h10", 3), ("graphite54", 2)])
def test_gpu_logpsi_vjp_matches_oracle(name, nw):
    from deepsolid_b200 import network
    sc, kl, pn, P = system(name)
    dev = torch.device("cuda", 0)
    net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                       determinants=8, method_name="eval_logdet")
    hp = net.apply.hotpath()
    X = torch.as_tensor(C.init_walkers(sc, nw, seed=21))
    rng = np.random.default_rng(3)
    ca = torch.as_tensor(rng.standard_normal(nw))
    cp = torch.as_tensor(rng.standard_normal(nw))
    hp.set_params(P)
    n0 = hp.launch_count()
    g = hp.logpsi_vjp(X.to(dev), ca, cp)
    assert hp.launch_count() > n0
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet")
    go = O.logpsi_vjp(f, P, X, ca, cp)
    names = [f"single{i}.{k}" for i in range(3) for k in "wb"] + [f"double{i}.{k}" for i in range(2) for k in "wb"] + \
            ["orbital0.w", "orbital1.w", "env0.pi", "env0.sigma", "env1.pi", "env1.sigma"]
    for nm, a, b in zip(names, _flat(g), _flat(go)):
        assert tuple(a.shape) == tuple(b.shape), nm
        assert _rel_err(a.cpu(), b) < 1e-9, (nm, _rel_err(a.cpu(), b))


@pytest.mark.gpu
def test_gpu_vjp_is_chunk_invariant_and_linear():
    """Workspace chunking must not change the batch sum; the pullback is linear in the cotangents."""
    from deepsolid_b200 import network
    sc, kl, pn, P = system("h4")
    dev = torch.device("cuda", 0)
    net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                       determinants=8, method_name="eval_logdet")
    hp = net.apply.hotpath()
    hp.set_params(P)
    X = torch.as_tensor(C.init_walkers(sc, 37, seed=2)).to(dev)
    ca = torch.linspace(-1, 1, 37, dtype=torch.float64)
    cp = torch.linspace(0.5, -0.5, 37, dtype=torch.float64)
    g1 = _flat(hp.logpsi_vjp(X, ca, cp))
    hp.set_workspace_limit(8 << 20)            # forces several chunks
    g2 = _flat(hp.logpsi_vjp(X, ca, cp))
    hp.set_workspace_limit(8 << 30)
    for a, b in zip(g1, g2):
        assert _rel_err(a, b) < 1e-11
    ga = _flat(hp.logpsi_vjp(X, ca, torch.zeros_like(cp)))
    gp = _flat(hp.logpsi_vjp(X, torch.zeros_like(ca), cp))
    for a, b, c in zip(g1, ga, gp):
        assert _rel_err(b + c, a) < 1e-11
    empty = _flat(hp.logpsi_vjp(X[:0], ca[:0], cp[:0]))
    assert all(float(t.abs().max()) == 0.0 for t in empty)


@pytest.mark.gpu
def test_gpu_value_and_grad_matches_oracle():
    from deepsolid_b200 import network, train, hamiltonian
    sc, kl, pn, P = system("h4")
    dev = torch.device("cuda", 0)
    net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                       determinants=8, method_name="eval_logdet")
    X = torch.as_tensor(C.init_walkers(sc, 6, seed=8))
    f_ps = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet")
    f_ld = O.make_solid_fermi_net(kl, sc, method_name="eval_logdet")
    el = O.local_energy_seperate(f_ld, sc, mode="dim_batch")
    for clip_type in ("real", "complex"):
        te = train.make_loss(net.apply, None, sc, clip_local_energy=1.0, clip_type=clip_type)
        (loss, aux), g = te.value_and_grad(P, X.to(dev))
        lo, e_l, go = O.total_energy_value_and_grad(f_ps, el, P, X, clip_local_energy=1.0, clip_type=clip_type)
        assert abs(float(loss) - float(lo)) < 1e-8
        assert float((aux.local_energy.cpu() - e_l).abs().max()) < 1e-8
        for a, b in zip(_flat(g), _flat(go)):
            assert _rel_err(a.cpu(), b) < 1e-7
    with pytest.raises(ValueError):
        train.make_loss(net.apply, None, sc, clip_type="polar")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["h4", "graphene8"])
def test_gpu_orbitals_vjp_matches_oracle(name):
    """eval_mats pullback (network.py:601-602; pretrain.py:70-89): gradient of sum Re(conj(cot) M)."""
    from deepsolid_b200 import network
    sc, kl, pn, P = system(name)
    dev = torch.device("cuda", 0)
    net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                       determinants=8, method_name="eval_mats")
    hp = net.apply.hotpath()
    hp.set_params(P)
    nw = 3
    X = torch.as_tensor(C.init_walkers(sc, nw, seed=31))
    mats = net.apply(P, X.to(dev))
    gen = torch.Generator().manual_seed(3)
    cots = [torch.complex(torch.randn(m.shape, dtype=torch.float64, generator=gen),
                          torch.randn(m.shape, dtype=torch.float64, generator=gen)) for m in mats]
    g = hp.orbitals_vjp(X.to(dev), cots)
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_mats")
    Pc = O._clone_params(P)
    total = torch.zeros((), dtype=torch.float64)
    for b, x in enumerate(X):
        ms = f(Pc, x)
        for m, c in zip(ms, cots):
            assert float((m.detach() - mats[[id(q) for q in cots].index(id(c))][b].cpu()).abs().max()) < 1e-10
            total = total + (c[b].real * m.real + c[b].imag * m.imag).sum()
    go = torch.autograd.grad(total, O._leaves(Pc), allow_unused=True)
    for a, b in zip(_flat(g), go):
        b = torch.zeros_like(a.cpu()) if b is None else b
        assert _rel_err(a.cpu(), b) < 1e-9 or float(b.abs().max()) == 0.0 and float(a.abs().max()) == 0.0

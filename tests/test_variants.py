"""Secondary structural variants on the same kernels: the `tri` distance features
(network.py:227-246; 7 features per electron-atom / electron-electron pair).  Same tolerances as the default path."""
import numpy as np
import pytest
import torch

from conftest import angle_diff, system
from deepsolid_b200 import cell as C
from oracle import deepsolid_oracle as O


def _setup(name):
    sc = C.build_system(name)
    kl = C.make_klist(sc)
    pn = O.init_params(np.random.default_rng(888), sc.original_cell.natm, sc.nelec, distance_type="tri")
    return sc, kl, O.params_to_torch(pn)


def test_oracle_tri_features_are_periodic_and_shaped():
    sc, kl, P = _setup("graphene8")
    assert P["single"][0]["w"].shape[0] == 3 * 7 * sc.original_cell.natm + 2 * 7
    assert P["double"][0]["w"].shape[0] == 7
    f = O.make_solid_fermi_net(kl, sc, distance_type="tri", method_name="eval_slogdet")
    x = torch.as_tensor(C.init_walkers(sc, 1, seed=3))[0]
    shift = torch.as_tensor(sc.lattice_vectors())[0].repeat(sum(sc.nelec))
    assert abs(float(f(P, x)) - float(f(P, x + shift))) < 1e-10          # periodic under a supercell lattice vector


def test_unknown_distance_raises():
    from deepsolid_b200 import network
    sc, kl, P = _setup("h4")
    with pytest.raises(ValueError):
        network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                     determinants=8, distance_type="cubic")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["h4", "graphene8", "lih_prim"])
def test_gpu_tri_distance_matches_oracle(name):
    from deepsolid_b200 import network, hamiltonian, qmc
    sc, kl, P = _setup(name)
    dev = torch.device("cuda", 0)
    kw = dict(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8, distance_type="tri")
    ld = network.make_solid_fermi_net(method_name="eval_logdet", **kw)
    hp = ld.apply.hotpath()
    sl = network.make_solid_fermi_net(method_name="eval_slogdet", hotpath=hp, **kw)
    nw = 3
    X = torch.as_tensor(C.init_walkers(sc, nw, seed=17))
    f_ld = O.make_solid_fermi_net(kl, sc, distance_type="tri", method_name="eval_logdet")
    f_ps = O.make_solid_fermi_net(kl, sc, distance_type="tri", method_name="eval_phase_and_slogdet")
    f_sl = O.make_solid_fermi_net(kl, sc, distance_type="tri", method_name="eval_slogdet")
    v = ld.apply(P, X.to(dev)).cpu()
    vo = torch.stack([f_ld(P, x) for x in X])
    assert float((v.real - vo.real).abs().max()) < 1e-10
    assert float(angle_diff(v.imag, vo.imag).max()) < 1e-10
    el = hamiltonian.local_energy_seperate(ld.apply, sc, mode="for")
    ke, ew = el(P, X.to(dev))
    elo = O.local_energy_seperate(f_ld, sc, mode="dim_batch")
    for b in range(nw):
        ko, eo = elo(P, X[b])
        assert abs(complex(ko) - complex(ke[b].cpu())) < 1e-8
        assert abs(float(eo) - float(ew[b])) < 1e-9
    # parameter gradient
    rng = np.random.default_rng(1)
    ca, cp = torch.as_tensor(rng.standard_normal(nw)), torch.as_tensor(rng.standard_normal(nw))
    g = hp.logpsi_vjp(X.to(dev), ca, cp)
    go = O.logpsi_vjp(f_ps, P, X, ca, cp)
    for a, b in zip(O._leaves(g), O._leaves(go)):
        assert tuple(a.shape) == tuple(b.shape)
        assert float((a.cpu() - b).abs().max()) < 1e-9 * max(1.0, float(b.abs().max()))
    # Metropolis accept masks
    B, steps = 4, 3
    Xm = torch.as_tensor(C.init_walkers(sc, B, seed=5))
    gen = torch.Generator().manual_seed(2)
    xi = torch.randn(steps, B, Xm.shape[1], dtype=torch.float64, generator=gen)
    u = torch.rand(steps, B, dtype=torch.float64, generator=gen)
    lat = torch.as_tensor(sc.lattice_vectors())
    xn, pm, masks = qmc.make_mcmc_step(sl.apply, B, lat, steps=steps)(P, Xm.to(dev), (xi, u), 0.3, return_masks=True)
    xo, po, mo = O.make_mcmc_step(lambda p, x: O.batch_apply(f_sl, p, x), B, lat, steps=steps)(P, Xm, (xi, u), 0.3)
    assert torch.equal(masks.cpu().bool(), mo)


# ---------------------------------------------------------------------------
# diagonal / full envelopes (network.py:340-364)
# ---------------------------------------------------------------------------
def _env_setup(name, envelope_type):
    sc = C.build_system(name)
    kl = C.make_klist(sc)
    rng = np.random.default_rng(888)
    pn = O.init_params(rng, sc.original_cell.natm, sc.nelec, envelope_type=envelope_type)
    for env in pn["envelope"]:          # the initial values (ones / identity) would hide index mistakes
        env["sigma"] = env["sigma"] * (0.6 + 0.8 * rng.random(env["sigma"].shape)) + 0.15 * rng.standard_normal(env["sigma"].shape)
        env["pi"] = env["pi"] * (0.5 + rng.random(env["pi"].shape))
    return sc, kl, O.params_to_torch(pn)


def test_envelope_needs_nu_distance():
    from deepsolid_b200 import network
    sc, kl, P = _env_setup("h4", "diagonal")
    with pytest.raises(ValueError):
        network.make_solid_fermi_net(envelope_type="cubic", full_det=False, klist=kl, simulation_cell=sc, determinants=8)


@pytest.mark.gpu
@pytest.mark.parametrize("envelope_type", ["diagonal", "full"])
@pytest.mark.parametrize("name", ["h4", "graphene8"])
def test_gpu_anisotropic_envelopes_match_oracle(name, envelope_type):
    from deepsolid_b200 import network, hamiltonian
    sc, kl, P = _env_setup(name, envelope_type)
    dev = torch.device("cuda", 0)
    kw = dict(envelope_type=envelope_type, full_det=False, klist=kl, simulation_cell=sc, determinants=8)
    ld = network.make_solid_fermi_net(method_name="eval_logdet", **kw)
    hp = ld.apply.hotpath()
    nw = 3
    X = torch.as_tensor(C.init_walkers(sc, nw, seed=23))
    f_ld = O.make_solid_fermi_net(kl, sc, envelope_type=envelope_type, method_name="eval_logdet")
    f_ps = O.make_solid_fermi_net(kl, sc, envelope_type=envelope_type, method_name="eval_phase_and_slogdet")
    v = ld.apply(P, X.to(dev)).cpu()
    vo = torch.stack([f_ld(P, x) for x in X])
    assert float((v.real - vo.real).abs().max()) < 1e-10
    assert float(angle_diff(v.imag, vo.imag).max()) < 1e-10
    ke, ew = hamiltonian.local_energy_seperate(ld.apply, sc, mode="for")(P, X.to(dev))
    elo = O.local_energy_seperate(f_ld, sc, mode="dim_batch")
    for b in range(nw):
        ko, eo = elo(P, X[b])
        assert abs(complex(ko) - complex(ke[b].cpu())) < 1e-8
    rng = np.random.default_rng(4)
    ca, cp = torch.as_tensor(rng.standard_normal(nw)), torch.as_tensor(rng.standard_normal(nw))
    g = hp.logpsi_vjp(X.to(dev), ca, cp)
    go = O.logpsi_vjp(f_ps, P, X, ca, cp)
    for a, b in zip(O._leaves(g), O._leaves(go)):
        assert tuple(a.shape) == tuple(b.shape)
        assert float((a.cpu() - b).abs().max()) < 1e-9 * max(1.0, float(b.abs().max()))


# ---------------------------------------------------------------------------
# bias_orbitals=True (network.py:177-179, 538-541)
# ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["h4", "graphene8"])
def test_gpu_orbital_bias_matches_oracle(name):
    from deepsolid_b200 import network, hamiltonian
    sc = C.build_system(name)
    kl = C.make_klist(sc)
    pn = O.init_params(np.random.default_rng(888), sc.original_cell.natm, sc.nelec, bias_orbitals=True)
    P = O.params_to_torch(pn)
    assert "b" in P["orbital"][0]
    dev = torch.device("cuda", 0)
    kw = dict(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8, bias_orbitals=True)
    ld = network.make_solid_fermi_net(method_name="eval_logdet", **kw)
    hp = ld.apply.hotpath()
    nw = 3
    X = torch.as_tensor(C.init_walkers(sc, nw, seed=29))
    f_ld = O.make_solid_fermi_net(kl, sc, bias_orbitals=True, method_name="eval_logdet")
    f_ps = O.make_solid_fermi_net(kl, sc, bias_orbitals=True, method_name="eval_phase_and_slogdet")
    v = ld.apply(P, X.to(dev)).cpu()
    vo = torch.stack([f_ld(P, x) for x in X])
    assert float((v.real - vo.real).abs().max()) < 1e-10
    assert float(angle_diff(v.imag, vo.imag).max()) < 1e-10
    ke, ew = hamiltonian.local_energy_seperate(ld.apply, sc, mode="for")(P, X.to(dev))
    elo = O.local_energy_seperate(f_ld, sc, mode="dim_batch")
    for b in range(nw):
        ko, eo = elo(P, X[b])
        assert abs(complex(ko) - complex(ke[b].cpu())) < 1e-8
    rng = np.random.default_rng(6)
    ca, cp = torch.as_tensor(rng.standard_normal(nw)), torch.as_tensor(rng.standard_normal(nw))
    g = hp.logpsi_vjp(X.to(dev), ca, cp)
    Pc = O._clone_params(P)
    total = torch.zeros((), dtype=torch.float64)
    for b, x in enumerate(X):
        sign, slog = f_ps(Pc, x)
        total = total + ca[b] * slog + cp[b] * torch.angle(sign)
    from deepsolid_b200.hotpath import flatten_params
    go = torch.autograd.grad(total, flatten_params(Pc))
    for a, b in zip(flatten_params(g), go):
        assert tuple(a.shape) == tuple(b.shape)
        assert float((a.cpu() - b).abs().max()) < 1e-9 * max(1.0, float(b.abs().max()))


# ---------------------------------------------------------------------------
# full_det=True (network.py:552-559): one N x N determinant per k instead of one per spin channel
# ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["h4", "graphene8", "lih_prim"])
def test_gpu_full_det_matches_oracle(name):
    from deepsolid_b200 import network, hamiltonian
    from deepsolid_b200.hotpath import flatten_params
    sc = C.build_system(name)
    kl = C.make_klist(sc)
    N = sum(sc.nelec)
    pn = O.init_params(np.random.default_rng(888), sc.original_cell.natm, sc.nelec, full_det=True)
    P = O.params_to_torch(pn)
    assert P["orbital"][0]["w"].shape[1] == 2 * N * 8
    dev = torch.device("cuda", 0)
    kw = dict(envelope_type="isotropic", full_det=True, klist=kl, simulation_cell=sc, determinants=8)
    ld = network.make_solid_fermi_net(method_name="eval_logdet", **kw)
    hp = ld.apply.hotpath()
    mt = network.make_solid_fermi_net(method_name="eval_mats", hotpath=hp, **kw)
    nw = 3
    X = torch.as_tensor(C.init_walkers(sc, nw, seed=37))
    f_ld = O.make_solid_fermi_net(kl, sc, full_det=True, method_name="eval_logdet")
    f_ps = O.make_solid_fermi_net(kl, sc, full_det=True, method_name="eval_phase_and_slogdet")
    f_mt = O.make_solid_fermi_net(kl, sc, full_det=True, method_name="eval_mats")
    mats = mt.apply(P, X.to(dev))
    assert len(mats) == 1 and tuple(mats[0].shape) == (nw, 8, N, N)
    for b in range(nw):
        assert float((mats[0][b].cpu() - f_mt(P, X[b])[0]).abs().max()) < 1e-10
    v = ld.apply(P, X.to(dev)).cpu()
    vo = torch.stack([f_ld(P, x) for x in X])
    assert float((v.real - vo.real).abs().max()) < 1e-10
    assert float(angle_diff(v.imag, vo.imag).max()) < 1e-10
    ke, ew = hamiltonian.local_energy_seperate(ld.apply, sc, mode="for")(P, X.to(dev))
    elo = O.local_energy_seperate(f_ld, sc, mode="dim_batch")
    for b in range(nw):
        ko, eo = elo(P, X[b])
        assert abs(complex(ko) - complex(ke[b].cpu())) < 1e-8
    rng = np.random.default_rng(8)
    ca, cp = torch.as_tensor(rng.standard_normal(nw)), torch.as_tensor(rng.standard_normal(nw))
    g = hp.logpsi_vjp(X.to(dev), ca, cp)
    go = O.logpsi_vjp(f_ps, P, X, ca, cp)
    for a, b in zip(flatten_params(g), flatten_params(go)):
        assert tuple(a.shape) == tuple(b.shape)
        assert float((a.cpu() - b).abs().max()) < 1e-9 * max(1.0, float(b.abs().max()))


# ---------------------------------------------------------------------------
# spin-polarised occupations (n_up != n_dn): different matrix sizes per spin channel everywhere
# ---------------------------------------------------------------------------
def _polarised(name, nelec):
    sc = C.build_system(name)
    assert sum(nelec) == sum(sc.nelec)
    sc.nelec = tuple(nelec)
    kl = C.make_klist(sc)
    pn = O.init_params(np.random.default_rng(888), sc.original_cell.natm, sc.nelec)
    return sc, kl, O.params_to_torch(pn)


@pytest.mark.gpu
@pytest.mark.parametrize("name,nelec,full_det", [("h4", (3, 1), False), ("graphene8", (5, 3), False), ("graphene8", (3, 5), False),
                                                 ("h4", (1, 3), True)])
def test_gpu_spin_polarised_matches_oracle(name, nelec, full_det):
    from deepsolid_b200 import network, hamiltonian, qmc
    from deepsolid_b200.hotpath import flatten_params
    sc, kl, P = _polarised(name, nelec)
    if full_det:
        P = O.params_to_torch(O.init_params(np.random.default_rng(888), sc.original_cell.natm, sc.nelec, full_det=True))
    dev = torch.device("cuda", 0)
    kw = dict(envelope_type="isotropic", full_det=full_det, klist=kl, simulation_cell=sc, determinants=8)
    ld = network.make_solid_fermi_net(method_name="eval_logdet", **kw)
    hp = ld.apply.hotpath()
    sl = network.make_solid_fermi_net(method_name="eval_slogdet", hotpath=hp, **kw)
    nw = 5
    X = torch.as_tensor(C.init_walkers(sc, nw, seed=41))
    f_ld = O.make_solid_fermi_net(kl, sc, full_det=full_det, method_name="eval_logdet")
    f_ps = O.make_solid_fermi_net(kl, sc, full_det=full_det, method_name="eval_phase_and_slogdet")
    f_sl = O.make_solid_fermi_net(kl, sc, full_det=full_det, method_name="eval_slogdet")
    v = ld.apply(P, X.to(dev)).cpu()
    vo = torch.stack([f_ld(P, x) for x in X])
    assert float((v.real - vo.real).abs().max()) < 1e-10
    assert float(angle_diff(v.imag, vo.imag).max()) < 1e-10
    ke, ew = hamiltonian.local_energy_seperate(ld.apply, sc, mode="for")(P, X.to(dev))
    elo = O.local_energy_seperate(f_ld, sc, mode="dim_batch")
    for b in range(nw):
        ko, eo = elo(P, X[b])
        assert abs(complex(ko) - complex(ke[b].cpu())) < 1e-8
        assert abs(float(eo) - float(ew[b])) < 1e-9
    rng = np.random.default_rng(12)
    ca, cp = torch.as_tensor(rng.standard_normal(nw)), torch.as_tensor(rng.standard_normal(nw))
    g = hp.logpsi_vjp(X.to(dev), ca, cp)
    go = O.logpsi_vjp(f_ps, P, X, ca, cp)
    for a, b in zip(flatten_params(g), flatten_params(go)):
        assert float((a.cpu() - b).abs().max()) < 1e-9 * max(1.0, float(b.abs().max()))
    if not full_det:        # KFAC statistics with different row counts per spin channel
        from deepsolid_b200 import kfac
        gf, wf = kfac.curvature_estimate(hp, P, X.to(dev), sync=False), O.kfac_factors(f_ps, P, X)
        for kind in ("single", "double", "orbital"):
            for gb, wb in zip(gf[kind], wf[kind]):
                assert gb["extra_scale"] == wb["extra_scale"]
                for key in ("inputs_factor", "outputs_factor"):
                    assert float((gb[key].cpu() - wb[key]).abs().max()) <= 1e-9 * max(float(wb[key].abs().max()), 1e-30)
    B, steps = 4, 2
    gen = torch.Generator().manual_seed(6)
    xi = torch.randn(steps, B, X.shape[1], dtype=torch.float64, generator=gen)
    u = torch.rand(steps, B, dtype=torch.float64, generator=gen)
    lat = torch.as_tensor(sc.lattice_vectors())
    xn, pm, masks = qmc.make_mcmc_step(sl.apply, B, lat, steps=steps)(P, X[:B].to(dev), (xi, u), 0.3, return_masks=True)
    xo, po, mo = O.make_mcmc_step(lambda p, x: O.batch_apply(f_sl, p, x), B, lat, steps=steps)(P, X[:B], (xi, u), 0.3)
    assert torch.equal(masks.cpu().bool(), mo)


@pytest.mark.gpu
@pytest.mark.parametrize("name,opts", [("h4", {}), ("graphene8", {}), ("lih_prim", {"bias_orbitals": True}),
                                       ("h4", {"full_det": True}), ("graphene8", {"i8": False}),
                                       ("li24", {"hidden_dims": ((256, 32), (256, 32))})])
def test_gpu_use_last_layer_matches_oracle(name, opts):
    """use_last_layer=True (network.py:129-134, 528-533): one more pair layer, and the orbital projection takes the
    832-wide symmetric features of the last layer (own | spin means | pair means).  Against the oracle: log psi, phase,
    orbital matrices, kinetic energy, accept masks, parameter gradient (the Kronecker-factor statistics: test_kfac.py)."""
    from deepsolid_b200 import network, hamiltonian, qmc
    opts = dict(opts)
    i8 = opts.pop("i8", True)
    hd = opts.pop("hidden_dims", ((256, 32),) * 3)
    sc, kl, _, _ = system(name)
    P = O.params_to_torch(O.init_params(np.random.default_rng(77), sc.original_cell.natm, sc.nelec, use_last_layer=True,
                                        hidden_dims=hd, **opts))
    assert len(P["double"]) == len(P["single"]) and P["orbital"][0]["w"].shape[0] == 3 * hd[0][0] + 2 * hd[0][1]
    dev = torch.device("cuda", 0)
    kw = dict(envelope_type="isotropic", klist=kl, simulation_cell=sc, determinants=8, use_last_layer=True, hidden_dims=hd,
              full_det=opts.get("full_det", False), bias_orbitals=opts.get("bias_orbitals", False))
    ld = network.make_solid_fermi_net(method_name="eval_logdet", **kw)
    hp = ld.apply.hotpath()
    sl = network.make_solid_fermi_net(method_name="eval_slogdet", hotpath=hp, **kw)
    mt = network.make_solid_fermi_net(method_name="eval_mats", hotpath=hp, **kw)
    okw = dict(full_det=opts.get("full_det", False), bias_orbitals=opts.get("bias_orbitals", False), hidden_dims=hd)
    f_ld = O.make_solid_fermi_net(kl, sc, method_name="eval_logdet", **okw)
    f_mt = O.make_solid_fermi_net(kl, sc, method_name="eval_mats", **okw)
    f_sl = O.make_solid_fermi_net(kl, sc, method_name="eval_slogdet", **okw)
    nw = 4
    X = torch.as_tensor(C.init_walkers(sc, nw, seed=19))
    if not i8:
        hp.debug_set("i8", 0)
    v = ld.apply(P, X.to(dev)).cpu()
    mats = mt.apply(P, X.to(dev))
    ke, ew = hamiltonian.local_energy_seperate(ld.apply, sc, mode="for")(P, X.to(dev))
    elo = O.local_energy_seperate(f_ld, sc, mode="dim_batch")
    for b in range(nw):
        vo = f_ld(P, X[b])
        assert abs(float(vo.real) - float(v[b].real)) < 1e-10
        assert float(angle_diff(v[b].imag, vo.imag)) < 1e-10
        mo = f_mt(P, X[b])
        for s in range(len(mo)):
            assert float((mats[s][b].cpu() - mo[s]).abs().max()) < 1e-11
        ko, eo = elo(P, X[b])
        assert abs(complex(ko) - complex(ke[b].cpu())) < 1e-8
    steps = 3
    rng = np.random.default_rng(5)
    xi = torch.as_tensor(rng.standard_normal((steps, nw, X.shape[1])))
    u = torch.as_tensor(rng.uniform(size=(steps, nw)))
    xn, pm, masks = qmc.make_mcmc_step(sl.apply, nw, sc.lattice_vectors(), steps=steps)(P, X.to(dev), (xi, u), 0.3, return_masks=True)
    xo, po, mo = O.make_mcmc_step(lambda p, x: O.batch_apply(f_sl, p, x), nw, sc.lattice_vectors(), steps=steps)(P, X, (xi, u), 0.3)
    assert (masks.cpu().numpy().astype(bool) == mo.numpy().astype(bool)).all()
    assert float((xn.cpu() - xo).abs().max()) < 1e-12
    # parameter gradient (reverse sweep through the extra pair layer and the 832-row orbital projection) vs oracle autograd
    from deepsolid_b200.hotpath import flatten_params
    f_ps = O.make_solid_fermi_net(kl, sc, method_name="eval_phase_and_slogdet", **okw)
    ca, cp = torch.as_tensor(rng.standard_normal(nw)), torch.as_tensor(rng.standard_normal(nw))
    g = hp.logpsi_vjp(X.to(dev), ca, cp)
    go = O.logpsi_vjp(f_ps, P, X, ca, cp)
    for a, b in zip(flatten_params(g), flatten_params(go)):
        scale = max(1.0, float(b.abs().max()))
        assert tuple(a.shape) == tuple(b.shape) and float((a.cpu() - b).abs().max()) < 1e-9 * scale

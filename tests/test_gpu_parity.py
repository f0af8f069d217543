"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs, against the committed golden fixtures, and -- at BASELINE sizes -- through
size-independent invariants.  Tolerances (north_star):
  log|psi|, phase: 1e-10 abs;  kinetic / local energy: 1e-8 Ha per walker;
  Metropolis accept masks: identical given identical (x1, xi, u)."""
import numpy as np
import pytest
import torch

from conftest import system, golden, angle_diff
from deepsolid_b200 import cell as C
from deepsolid_b200 import network, hamiltonian, qmc, train
from oracle import deepsolid_oracle as O

pytestmark = pytest.mark.gpu

TOL_LOG, TOL_E = 1e-10, 1e-8
_nets = {}


def nets(name):
    if name not in _nets:
        sc, kl, _, P = system(name)
        kw = dict(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8)
        ld = network.make_solid_fermi_net(method_name="eval_logdet", **kw)
        hp = ld.apply.hotpath()
        sl = network.make_solid_fermi_net(method_name="eval_slogdet", hotpath=hp, **kw)
        ps = network.make_solid_fermi_net(method_name="eval_phase_and_slogdet", hotpath=hp, **kw)
        mt = network.make_solid_fermi_net(method_name="eval_mats", hotpath=hp, **kw)
        _nets[name] = (ld, sl, ps, mt, hp)
    return _nets[name]


def dev():
    return torch.device("cuda", 0)


def test_extension_is_loaded_and_launches_kernels():
    ld, _, _, _, hp = nets("h4")
    sc, kl, _, P = system("h4")
    n0 = hp.launch_count()
    ld.apply(P, torch.as_tensor(C.init_walkers(sc, 2)).to(dev()))
    assert hp.launch_count() > n0
    assert "libdeepsolid_b200.so" in open("/proc/self/maps").read()


@pytest.mark.parametrize("name", ["h4", "lih_prim", "graphene8", "h10", "li24"])
def test_golden_logpsi_kinetic_ewald(name):
    ld, sl, ps, mt, hp = nets(name)
    sc, kl, _, P = system(name)
    g = golden(name)
    X = torch.as_tensor(g["x"]).to(dev())
    v = ld.apply(P, X).cpu()
    assert np.abs(v.real.numpy() - g["logabs"]).max() < TOL_LOG
    assert float(angle_diff(v.imag, g["phase"]).max()) < TOL_LOG
    assert np.abs(sl.apply(P, X).cpu().numpy() - g["logabs"]).max() < TOL_LOG
    sign, slog = ps.apply(P, X)
    assert float((sign.cpu() - torch.exp(1j * torch.as_tensor(g["phase"]))).abs().max()) < TOL_LOG
    el = hamiltonian.local_energy_seperate(ld.apply, sc, mode="for")
    ke, ew = el(P, X)
    assert np.abs(ke.cpu().numpy() - g["ke"]).max() < TOL_E
    assert np.abs(ew.cpu().numpy() - (g["ee"] + g["ei"] + float(g["ii"]))).max() < 1e-10
    ee, ei, ii = hp.ewald(X)
    assert np.abs(ee.cpu().numpy() - g["ee"]).max() < 1e-10 and np.abs(ei.cpu().numpy() - g["ei"]).max() < 1e-10
    # single-walker (unbatched) call, the reference's per-walker signature
    k1, e1 = el(P, X[0])
    assert abs(complex(k1.cpu()) - g["ke"][0]) < TOL_E and k1.dim() == 0


@pytest.mark.parametrize("name", ["h4", "graphene8"])
def test_live_oracle_parity(name):
    """Fresh seeded walkers (not the fixture): CUDA vs the oracle evaluated now."""
    ld, sl, ps, mt, hp = nets(name)
    sc, kl, _, P = system(name)
    X = torch.as_tensor(C.init_walkers(sc, 3, seed=2024))
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_logdet")
    fm = O.make_solid_fermi_net(kl, sc, method_name="eval_mats")
    elo = O.local_energy_seperate(f, sc, mode="partition", partition_number=3)
    ke, ew = hamiltonian.local_energy_seperate(ld.apply, sc, mode="partition", partition_number=3)(P, X.to(dev()))
    mats = mt.apply(P, X.to(dev()))
    v = ld.apply(P, X.to(dev())).cpu()
    for b in range(3):
        ko, eo = elo(P, X[b])
        vo = f(P, X[b])
        assert abs(complex(ko) - complex(ke[b].cpu())) < TOL_E
        assert abs(float(eo) - float(ew[b])) < 1e-10
        assert abs(float(vo.real) - float(v[b].real)) < TOL_LOG
        mo = fm(P, X[b])
        for s in range(2):
            assert float((mats[s][b].cpu() - mo[s]).abs().max()) < 1e-11


def test_modes_are_the_same_quantity_and_bad_arguments_raise():
    ld, _, _, _, hp = nets("lih_prim")
    sc, kl, _, P = system("lih_prim")
    X = torch.as_tensor(C.init_walkers(sc, 4, seed=11)).to(dev())
    ref, _ = hamiltonian.local_energy_seperate(ld.apply, sc, mode="for")(P, X)
    for mode, pn in [("hessian", 3), ("dim_batch", 3), ("partition", 3), ("partition", 12)]:
        ke, _ = hamiltonian.local_energy_seperate(ld.apply, sc, mode=mode, partition_number=pn)(P, X)
        assert float((ke - ref).abs().max()) < 1e-12
    with pytest.raises(ValueError):       # 5 does not divide 3N = 12 (hamiltonian.py:131-133,145)
        hamiltonian.local_energy_seperate(ld.apply, sc, mode="partition", partition_number=5)(P, X)
    with pytest.raises(ValueError):
        ld.apply(P, torch.zeros(2, 7, dtype=torch.float64, device=dev()))
    bad = {k: v for k, v in P.items()}
    bad["single"] = P["single"][:2]
    with pytest.raises(ValueError):
        ld.apply(bad, X)
    ld.apply(P, X)                                                           # context still usable


@pytest.mark.parametrize("name", ["h4", "lih_prim", "graphene8", "h10", "li24"])
def test_mcmc_accept_masks_bit_exact(name):
    _, sl, _, _, hp = nets(name)
    sc, kl, _, P = system(name)
    g = golden(name)
    steps, B = g["xi"].shape[0], g["x"].shape[0]
    step = qmc.make_mcmc_step(sl.apply, B, sc.lattice_vectors(), steps=steps)
    xn, pmove, masks = step(P, torch.as_tensor(g["x"]).to(dev()), (torch.as_tensor(g["xi"]), torch.as_tensor(g["u"])),
                            float(g["width"]), return_masks=True)
    assert (masks.cpu().numpy().astype(bool) == g["masks"]).all()
    assert np.abs(xn.cpu().numpy() - g["x_new"]).max() < 1e-12
    assert abs(float(pmove) - float(g["pmove"])) < 1e-15
    # host-buffer entry point gives the same thing
    xh, pmh, mh = step(P, torch.as_tensor(g["x"]), (torch.as_tensor(g["xi"]), torch.as_tensor(g["u"])),
                       float(g["width"]), return_masks=True)
    assert (mh.numpy().astype(bool) == g["masks"]).all() and np.abs(xh.numpy() - g["x_new"]).max() < 1e-12


def test_mcmc_device_rng_is_reproducible_and_samples():
    _, sl, _, _, hp = nets("h10")
    sc, kl, _, P = system("h10")
    X = torch.as_tensor(C.init_walkers(sc, 64, seed=1)).to(dev())
    step = qmc.make_mcmc_step(sl.apply, 64, sc.lattice_vectors(), steps=10)
    a, pa = step(P, X, 123, 0.1)
    b, pb = step(P, X, 123, 0.1)
    c, pc = step(P, X, 124, 0.1)
    assert torch.equal(a, b) and float(pa) == float(pb) and not torch.equal(a, c)
    assert 0.05 < float(pa) < 0.999
    frac = a.reshape(64, -1, 3) @ torch.linalg.inv(torch.as_tensor(sc.a)).to(dev())
    assert float(frac.min()) >= 0.0 and float(frac.max()) < 1.0 + 1e-12       # wrapped into the cell


def test_stats_and_total_energy():
    ld, _, _, _, hp = nets("h10")
    sc, kl, _, P = system("h10")
    X = torch.as_tensor(C.init_walkers(sc, 37, seed=3)).to(dev())
    loss, aux = train.make_loss(ld.apply, ld.apply, sc, mode="for")(P, X)
    e = aux.local_energy.cpu()
    lo, imo, varo = O.total_energy_stats(aux.kinetic.cpu(), aux.ewald.cpu())
    assert abs(float(loss) - float(lo)) < 1e-11 and abs(float(aux.imaginary) - float(imo)) < 1e-11
    assert abs(float(aux.variance) - float(varo)) < 1e-9
    # host walkers -> host entry point -> same numbers
    loss_h, aux_h = train.make_loss(ld.apply, ld.apply, sc, mode="for")(P, X.cpu())
    assert abs(float(loss_h) - float(loss)) < 1e-11 and not aux_h.local_energy.is_cuda


def test_empty_and_ragged_batches():
    ld, sl, _, _, hp = nets("h10")
    sc, kl, _, P = system("h10")
    el = hamiltonian.local_energy_seperate(ld.apply, sc)
    ke, ew = el(P, torch.zeros(0, 30, dtype=torch.float64, device=dev()))
    assert ke.shape == (0,) and ew.shape == (0,)
    X = torch.as_tensor(C.init_walkers(sc, 23, seed=8)).to(dev())
    ref_ke, ref_ew = el(P, X)
    ref_lp = sl.apply(P, X)
    hp.set_workspace_limit(6 << 20)        # forces several chunks, the last one ragged
    try:
        ke, ew = el(P, X)
        lp = sl.apply(P, X)
    finally:
        hp.set_workspace_limit(8 << 30)
    assert float((ke - ref_ke).abs().max()) < 1e-12 and float((ew - ref_ew).abs().max()) < 1e-12
    assert float((lp - ref_lp).abs().max()) < 1e-12


# ---- BASELINE-size property tests (no oracle at these sizes: it would take minutes per walker)

@pytest.mark.parametrize("name,batch", [("graphite54", 48), ("diamond64", 24), ("lih108", 8)])
def test_full_size_invariants(name, batch):
    ld, sl, ps, _, hp = nets(name)
    sc, kl, _, P = system(name)
    mode, pn = C.SYSTEMS[name][2], C.SYSTEMS[name][3]
    X = torch.as_tensor(C.init_walkers(sc, batch, seed=5)).to(dev())
    el = hamiltonian.local_energy_seperate(ld.apply, sc, mode=mode, partition_number=pn)
    ke, ew = el(P, X)
    assert torch.isfinite(ke.real).all() and torch.isfinite(ew).all()
    N = sc.nelectron
    # (1) periodic BC (test_network.py:65-83): all electrons + primitive lattice vector
    t = torch.as_tensor(sc.original_cell.lattice_vectors()[1]).to(dev())
    X2 = X + t.repeat(N)
    v1, v2 = ld.apply(P, X), ld.apply(P, X2)
    kp = sum(torch.as_tensor(k).sum(0) for k in kl).to(dev())
    assert float((v1.real - v2.real).abs().max()) < 1e-9
    assert float(angle_diff((v2.imag - v1.imag).cpu(), float(torch.dot(kp, t))).max()) < 1e-9
    ke2, ew2 = el(P, X2)
    assert float((ke2 - ke).abs().max()) < 1e-7 * max(1.0, float(ke.abs().max()))
    assert float((ew2 - ew).abs().max()) < 1e-9
    # (2) twisted BC, twist 0 (test_network.py:86-106): one electron + supercell vector
    X3 = X.clone(); X3[:, 3:6] += torch.as_tensor(sc.lattice_vectors()[0]).to(dev())
    v3 = ld.apply(P, X3)
    assert float((v3.real - v1.real).abs().max()) < 1e-9
    assert float(angle_diff(v3.imag.cpu(), v1.imag.cpu()).max()) < 1e-8
    # (3) antisymmetry (test_network.py:109-122) and permutation invariance of E_L
    X4 = torch.cat([X[:, 3:6], X[:, :3], X[:, 6:]], dim=1)
    v4 = ld.apply(P, X4)
    assert float((v4.real - v1.real).abs().max()) < 1e-9
    assert float(angle_diff(v4.imag.cpu(), (v1.imag + np.pi).cpu()).max()) < 1e-8
    ke4, ew4 = el(P, X4)
    assert float((ke4 - ke).abs().max()) < 1e-7 * max(1.0, float(ke.abs().max()))
    # (4) the kinetic energy agrees with central finite differences of the CUDA log psi (one walker)
    if name == "graphite54":
        x = X[:1]
        h = 2e-4
        eye = torch.eye(3 * N, dtype=torch.float64, device=dev()) * h
        fp = ld.apply(P, x + eye); fm = ld.apply(P, x - eye); f0 = ld.apply(P, x)[0]
        dphi = lambda a: torch.complex(a.real - f0.real, torch.angle(torch.exp(1j * (a.imag - f0.imag))))
        dp, dm = dphi(fp), dphi(fm)
        lap = ((dp + dm) / h ** 2).sum()
        g2 = (((dp - dm) / (2 * h)) ** 2).sum()
        assert abs(complex((-0.5 * (lap + g2)).cpu()) - complex(ke[0].cpu())) < 5e-4 * max(1.0, abs(complex(ke[0].cpu())))


def test_non_finite_walkers_propagate_and_stay_local():
    """The reference never raises on numerics (NaN/Inf propagate, process.py:307-318 optionally catches them): a walker
    with a NaN / Inf coordinate gives NaN outputs for THAT walker only, and a Metropolis proposal landing on it is
    rejected (qmc.py:217-221: `nan > log u` is False)."""
    ld, sl, ps, _, hp = nets("graphene8")
    sc, kl, _, P = system("graphene8")
    X = torch.as_tensor(C.init_walkers(sc, 9, seed=31)).to(dev())
    el = hamiltonian.local_energy_seperate(ld.apply, sc, mode="for")
    ke0, ew0 = el(P, X)
    la0 = sl.apply(P, X)
    Xb = X.clone()
    Xb[3, 4] = float("nan")
    Xb[6, 0] = float("inf")
    ke, ew = el(P, Xb)
    la = sl.apply(P, Xb)
    good = [i for i in range(9) if i not in (3, 6)]
    assert torch.isnan(la[3]) and not torch.isfinite(la[6])
    assert not torch.isfinite(ke[3].real) and not torch.isfinite(ke[6].real)
    assert torch.equal(la[good], la0[good])                      # bitwise: no cross-walker contamination
    assert torch.equal(ke[good], ke0[good])                      # the sweep has a fixed summation order: bitwise too
    assert torch.equal(ew[good], ew0[good])
    # Metropolis: a NaN proposal is rejected, the walker keeps its position
    steps, B, n3 = 2, 9, X.shape[1]
    xi = torch.zeros(steps, B, n3, dtype=torch.float64)
    xi[0, 2, 5] = float("nan")
    u = torch.full((steps, B), 0.5, dtype=torch.float64)
    step = qmc.make_mcmc_step(sl.apply, B, sc.lattice_vectors(), steps=steps)
    xn, pmove, masks = step(P, X, (xi, u), 0.02, return_masks=True)
    assert not bool(masks[0, 2]) and torch.isfinite(xn).all()


@pytest.mark.parametrize("name,batch", [("graphene8", 33), ("li24", 21), ("h10", 40)])
def test_sweep_is_bit_reproducible_and_fused_digits_match_the_unfused_path(name, batch):
    """The reference is deterministic on fixed inputs; so is the sweep (per-8-row partial sums of zJ^2 reduced in a fixed
    order, ordered cross-warp sums in the feature kernel -- no floating-point atomics on the default path).  The
    optional fused-digit GEMM epilogue (OZ_JACD: digits of the next operand formed in the epilogue of a cluster pair,
    residual rows read back from the input digits, no fp64 Jacobian in HBM) agrees with the default path that
    materialises fp64 Jacobian rows and slices them in a separate pass."""
    ld, sl, _, _, hp = nets(name)
    sc, kl, _, P = system(name)
    X = torch.as_tensor(C.init_walkers(sc, batch, seed=123)).to(dev())
    el = hamiltonian.local_energy_seperate(ld.apply, sc, mode="for")
    ke1, ew1 = el(P, X)
    ke2, ew2 = el(P, X)
    assert torch.equal(ke1, ke2) and torch.equal(ew1, ew2)
    hp.set_workspace_limit(64 << 20)           # other chunking, same bits: walkers never interact
    try:
        ke3, _ = el(P, X)
    finally:
        hp.set_workspace_limit(24 << 30)
    assert torch.equal(ke1, ke3)
    hp.debug_set("fused_digits", 1)
    try:
        ke4, _ = el(P, X)
        ke5, _ = el(P, X)
    finally:
        hp.debug_set("fused_digits", 0)
    assert torch.equal(ke4, ke5)
    assert float((ke4 - ke1).abs().max()) < 1e-9

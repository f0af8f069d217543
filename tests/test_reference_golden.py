"""Consumes `tests/golden/reference_*.npz` written by `tests/golden/make_reference_golden.py`:
  * `reference_shim_*.npz` (committed): outputs of the reference's OWN SOURCE FILES -- network.py, hamiltonian.py,
    ewaldsum.py, distance.py, supercell.py, qmc.mh_update, imported unmodified from /root/reference in the build
    container -- executed on a torch stand-in for jax / pyscf (`tests/golden/torch_jax_shim.py`; neither can be
    installed in this image).  Cells: the LiH cell of the reference's test/test_cell.py (S = I, diag(2,1,1), every
    structural option of make_solid_fermi_net) and the BASELINE configurations at full size;
  * `reference_*.npz` with source="reference": the same quantities from the real DeepSolid under JAX, for a maintainer
    whose machine has it (`--backend reference`).
The plumbing test keeps writer and reader in step by round-tripping a file written from the oracle (never accepted
as a pin: it is stamped source="oracle-selftest").

Checked per file: oracle (CPU) and CUDA path (GPU) against log|psi|, phase, kinetic energy in the three Laplacian
modes, ee / ei / ii Ewald terms, orbital matrices, accept masks and final walkers of the recorded Metropolis moves;
and the repo's geometry (AV, BV, supercell atoms) against the pyscf-built cell stored in the file."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "reference_*.npz")))

from deepsolid_b200 import cell as C                      # noqa: E402
from oracle import deepsolid_oracle as O                  # noqa: E402


def case_opts(g):
    """Network options of the fixture (make_solid_fermi_net keywords; {} = base_config defaults)."""
    import json
    return json.loads(str(g["opts"])) if "opts" in g.files else {}


def tangent_of(g, sc):
    """The fixed parameter tangent of the gradient pin: a parameter-shaped numpy draw from the stored seed."""
    o = case_opts(g)
    ikw = {k: o[k] for k in ("envelope_type", "bias_orbitals", "use_last_layer", "full_det", "distance_type") if k in o}
    return O.params_to_torch(O.init_params(np.random.default_rng(int(g["tangent_seed"])), sc.original_cell.natm, sc.nelec, **ikw))


def case_modes(g):
    return tuple(str(g["modes"]).split(",")) if "modes" in g.files else ("for", "partition", "dim_batch")


def load_case(path):
    g = np.load(path, allow_pickle=False)
    prim = C.Cell(a=g["prim_a"], coords=g["prim_atoms"], charges=g["prim_charges"],
                  nelec=tuple(int(v) // int(round(abs(np.linalg.det(g["S"])))) for v in g["nelec"]))
    scale = int(round(abs(np.linalg.det(g["S"]))))
    sc = C.get_supercell(prim, g["S"], spin=(int(g["nelec"][0]) - int(g["nelec"][1])) // scale)
    params = {"single": [], "double": [], "orbital": [], "envelope": []}
    for key in g.files:
        if key.startswith("param/"):
            _, group, idx, leaf = key.split("/")
            while len(params[group]) <= int(idx):
                params[group].append({})
            params[group][int(idx)][leaf] = torch.as_tensor(g[key])
    if "param_seed" in g.files:          # shim fixtures store the seed of the numpy parameter draw instead of 0.5 M doubles
        o = case_opts(g)
        ikw = {k: o[k] for k in ("envelope_type", "bias_orbitals", "use_last_layer", "full_det", "distance_type") if k in o}
        pn = O.init_params(np.random.default_rng(int(g["param_seed"])), prim.natm, sc.nelec, **ikw)
        flat = [v for grp in ("single", "double", "orbital", "envelope") for d in pn[grp] for v in d.values()]
        assert abs(sum(float(np.abs(v).sum()) for v in flat) - float(g["param_checksum"])) < 1e-9 * float(g["param_checksum"])
        params = O.params_to_torch(pn)
    klist = [g["klist0"], g["klist1"]]
    return g, sc, klist, params


# what counts as a pin: outputs of the real DeepSolid under JAX ("reference"), or of the reference's own source files
# executed on the torch stand-in for jax / pyscf (tests/golden/torch_jax_shim.py; JAX cannot be installed in this image)
PIN_SOURCES = ("reference", "reference-source/torch-shim")


def check_geometry(g, sc):
    assert np.abs(sc.lattice_vectors() - g["sim_a"]).max() < 1e-12
    assert np.abs(sc.atom_coords() - g["sim_atoms"]).max() < 1e-10 and np.array_equal(sc.atom_charges(), g["sim_charges"])
    assert np.abs(sc.AV - g["sim_AV"]).max() < 1e-12 and np.abs(sc.BV - g["sim_BV"]).max() < 1e-12
    assert np.abs(sc.original_cell.AV - g["prim_AV"]).max() < 1e-12
    assert tuple(sc.nelec) == tuple(int(v) for v in g["nelec"])


def check_oracle(g, sc, klist, P, tol_e=1e-9, max_walkers=None, kinetic=True):
    nets = {m: O.make_solid_fermi_net(klist, sc, method_name=m, **case_opts(g))
            for m in ("eval_logdet", "eval_slogdet", "eval_phase_and_slogdet", "eval_mats")}
    pnum = int(g["partition_number"]) if "partition_number" in g.files else 3
    X = torch.as_tensor(g["x"])
    ew = O.EwaldSum(sc)
    assert abs(ew.alpha - float(g["ewald_alpha"])) < 1e-12 * ew.alpha and ew.gweight.shape[0] == int(g["ewald_ng"])
    assert abs((ew.ion_ion + ew.ii_const) - float(g["energy_nuc"])) < 1e-5          # hamiltonian.py:170-172
    for b in range(X.shape[0] if max_walkers is None else min(X.shape[0], max_walkers)):
        sign, slog = nets["eval_phase_and_slogdet"](P, X[b])
        assert abs(float(slog) - g["logabs"][b]) < 1e-10
        assert abs(np.angle(np.exp(1j * (float(torch.angle(sign)) - g["phase"][b])))) < 1e-10
        mats = nets["eval_mats"](P, X[b])
        for s in range(len(mats)):
            assert np.abs(mats[s].numpy() - g[f"mats{s}"][b]).max() < 1e-11
        for mode in (case_modes(g) if kinetic else ()):
            ke, e = O.local_energy_seperate(nets["eval_logdet"], sc, mode=mode, partition_number=pnum)(P, X[b])
            assert abs(complex(ke) - g[f"ke_{mode}"][b]) < tol_e and abs(float(e) - g[f"ewald_{mode}"][b]) < 1e-10
        ee, ei, ii = ew.energy(X[b])
        assert abs(float(ee) - g["ee"][b]) < 1e-10 and abs(float(ei) - g["ei"][b]) < 1e-10 and abs(float(ii) - g["ii"][b]) < 1e-10
    if "te_loss" in g.files and kinetic:      # train.make_loss(...).total_energy forward (train.py:66-89), whole batch
        el = O.local_energy_seperate(nets["eval_logdet"], sc, mode=case_modes(g)[0], partition_number=pnum)
        kes, ews = zip(*[el(P, X[b]) for b in range(X.shape[0])])
        loss, imag, var = O.total_energy_stats(torch.stack([torch.as_tensor(k) for k in kes]), torch.stack([torch.as_tensor(e) for e in ews]))
        assert abs(float(loss) - float(g["te_loss"])) < 1e-9 and abs(float(imag) - float(g["te_imaginary"])) < 1e-9
        assert abs(float(var) - float(g["te_variance"])) < 1e-8 * max(1.0, abs(float(g["te_variance"])))
    if "te_jvp_real" in g.files and kinetic:      # train.py:90-142: <gradient estimator, fixed tangent>
        T = tangent_of(g, sc)
        f = O.make_solid_fermi_net(klist, sc, method_name="eval_phase_and_slogdet", **case_opts(g))
        el = O.local_energy_seperate(nets["eval_logdet"], sc, mode=case_modes(g)[0], partition_number=pnum)
        for clip_type in ("real", "complex"):
            _, _, grads = O.total_energy_value_and_grad(f, el, P, X, clip_local_energy=5.0, clip_type=clip_type)
            dot = sum(float((a * b).sum()) for a, b in zip(O._leaves(grads), O._leaves(T)))
            want = float(g[f"te_jvp_{clip_type}"])
            assert abs(dot - want) < 1e-8 * max(1.0, abs(want)), (clip_type, dot, want)
    if "pt_loss" in g.files:             # pretrain.py:43-107: loss and parameter gradient of the orbital matching
        full_det = bool(case_opts(g).get("full_det", False))
        target = [torch.as_tensor(g["pt_target0"]), torch.as_tensor(g["pt_target1"])]
        loss, grads = O.pretrain_loss_and_grad(nets["eval_mats"], P, X, target, full_det)
        T = tangent_of(g, sc)
        dot = sum(float((a * b).sum()) for a, b in zip(O._leaves(grads), O._leaves(T)))
        assert abs(float(loss) - float(g["pt_loss"])) < 1e-11 * max(1.0, abs(float(g["pt_loss"])))
        assert abs(dot - float(g["pt_dot"])) < 1e-9 * max(1.0, abs(float(g["pt_dot"])))
        norms = np.asarray([float(a.norm()) for a in O._leaves(grads)])
        assert np.abs(norms - g["pt_norms"]).max() < 1e-9 * max(1.0, float(g["pt_norms"].max()))
    if "obs_sk" in g.files:              # estimator.py:15-85
        assert np.abs(O.make_structure_factor(sc, nq=3)(X).numpy() - g["obs_sk"]).max() < 1e-12
        for d in range(3):
            assert abs(complex(O.make_complex_polarization(sc, direction=d)(X)) - g["obs_pol"][d]) < 1e-12
    if "oe_x_new" in g.files:            # one-electron moves and importance sampling (qmc.py:63-150, 227-287)
        lat = torch.as_tensor(sc.lattice_vectors())
        B0 = X.shape[0]
        N = sum(sc.nelec)
        bs = lambda p, x: O.batch_apply(nets["eval_slogdet"], p, x)
        xo, po, _ = O.make_mcmc_step_one_electron(bs, B0, lat, steps=g["oe_u"].shape[0] // N)(
            P, X, (torch.as_tensor(g["oe_xi"]), torch.as_tensor(g["oe_u"])), float(g["oe_width"]))
        assert np.abs(xo.numpy() - g["oe_x_new"]).max() < 1e-12 and abs(float(po) - float(g["oe_pmove"])) < 1e-15
        xo, po, _ = O.make_mcmc_step_importance(nets["eval_slogdet"], B0, lat, steps=g["imp_u"].shape[0])(
            P, X, (torch.as_tensor(g["imp_xi"]), torch.as_tensor(g["imp_u"])), float(g["imp_width"]))
        assert np.abs(xo.numpy() - g["imp_x_new"]).max() < 1e-9 and abs(float(po) - float(g["imp_pmove"])) < 1e-15
    steps, B = g["u"].shape
    mc = O.make_mcmc_step(lambda p, x: O.batch_apply(nets["eval_slogdet"], p, x), B, sc.lattice_vectors(), steps=steps)
    xn, pmove, masks = mc(P, X, (torch.as_tensor(g["xi"]), torch.as_tensor(g["u"])), float(g["width"]))
    assert (masks.numpy().astype(bool) == g["masks"].astype(bool)).all()
    assert np.abs(xn.numpy() - g["x_new"]).max() < 1e-12 and abs(float(pmove) - float(g["pmove"])) < 1e-15


def check_gpu(g, sc, klist, P):
    from deepsolid_b200 import network, hamiltonian, qmc
    dev = torch.device("cuda", 0)
    kw = dict(envelope_type="isotropic", full_det=False, klist=klist, simulation_cell=sc, determinants=8)
    kw.update(case_opts(g))
    pnum = int(g["partition_number"]) if "partition_number" in g.files else 3
    ld = network.make_solid_fermi_net(method_name="eval_logdet", **kw)
    hp = ld.apply.hotpath()
    sl = network.make_solid_fermi_net(method_name="eval_slogdet", hotpath=hp, **kw)
    mt = network.make_solid_fermi_net(method_name="eval_mats", hotpath=hp, **kw)
    X = torch.as_tensor(g["x"]).to(dev)
    v = ld.apply(P, X).cpu()
    assert np.abs(v.real.numpy() - g["logabs"]).max() < 1e-10
    assert np.abs(np.angle(np.exp(1j * (v.imag.numpy() - g["phase"])))).max() < 1e-10
    mats = mt.apply(P, X)
    for s in range(len(mats)):
        assert np.abs(mats[s].cpu().numpy() - g[f"mats{s}"]).max() < 1e-11
    for mode in case_modes(g):
        ke, ew = hamiltonian.local_energy_seperate(ld.apply, sc, mode=mode, partition_number=pnum)(P, X)
        assert np.abs(ke.cpu().numpy() - g[f"ke_{mode}"]).max() < 1e-8
        assert np.abs(ew.cpu().numpy() - g[f"ewald_{mode}"]).max() < 1e-10
    if "te_loss" in g.files:
        from deepsolid_b200 import train
        loss, aux = train.make_loss(ld.apply, ld.apply, sc, mode=case_modes(g)[0], partition_number=pnum)(P, X)
        assert abs(float(loss) - float(g["te_loss"])) < 1e-8 and abs(float(aux.imaginary) - float(g["te_imaginary"])) < 1e-8
        assert abs(float(aux.variance) - float(g["te_variance"])) < 1e-7 * max(1.0, abs(float(g["te_variance"])))
        assert np.abs(aux.local_energy.cpu().numpy() - g["te_local_energy"]).max() < 1e-8
    if "te_jvp_real" in g.files:
        from deepsolid_b200 import train
        T = tangent_of(g, sc)
        for clip_type in ("real", "complex"):
            lossf = train.make_loss(ld.apply, ld.apply, sc, clip_local_energy=5.0, clip_type=clip_type,
                                    mode=case_modes(g)[0], partition_number=pnum)
            _, grads = lossf.value_and_grad(P, X)
            dot = sum(float((torch.as_tensor(a).cpu() * b).sum()) for a, b in zip(O._leaves(grads), O._leaves(T)))
            want = float(g[f"te_jvp_{clip_type}"])
            assert abs(dot - want) < 1e-7 * max(1.0, abs(want)), (clip_type, dot, want)
    ee, ei, ii = hp.ewald(X)
    assert np.abs(ee.cpu().numpy() - g["ee"]).max() < 1e-10 and np.abs(ei.cpu().numpy() - g["ei"]).max() < 1e-10
    if "pt_loss" in g.files:
        from deepsolid_b200 import pretrain
        full_det = bool(case_opts(g).get("full_det", False))
        target = [torch.as_tensor(g["pt_target0"]).to(dev), torch.as_tensor(g["pt_target1"]).to(dev)]
        hp.set_params(P)
        loss, cots = pretrain.pretrain_loss_cotangent(hp.orbitals(X), target, full_det)
        grads = hp.orbitals_vjp(X, cots)
        T = tangent_of(g, sc)
        dot = sum(float((torch.as_tensor(a).cpu() * b).sum()) for a, b in zip(O._leaves(grads), O._leaves(T)))
        assert abs(float(loss) - float(g["pt_loss"])) < 1e-10 * max(1.0, abs(float(g["pt_loss"])))
        assert abs(dot - float(g["pt_dot"])) < 1e-8 * max(1.0, abs(float(g["pt_dot"])))
        norms = np.asarray([float(torch.as_tensor(a).norm()) for a in O._leaves(grads)])
        assert np.abs(norms - g["pt_norms"]).max() < 1e-8 * max(1.0, float(g["pt_norms"].max()))
    if "obs_sk" in g.files:
        from deepsolid_b200 import estimator
        assert np.abs(estimator.make_structure_factor(sc, nq=3, hotpath=hp)(X).cpu().numpy() - g["obs_sk"]).max() < 1e-12
        for d in range(3):
            pol = estimator.make_complex_polarization(sc, direction=d, hotpath=hp)(X).cpu()
            assert abs(complex(pol) - g["obs_pol"][d]) < 1e-12
    if "oe_x_new" in g.files:
        lat = torch.as_tensor(sc.lattice_vectors())
        B0, N = X.shape[0], sum(sc.nelec)
        st = qmc.make_mcmc_step(sl.apply, B0, lat, steps=g["oe_u"].shape[0] // N, one_electron_moves=True)
        xn, pm = st(P, X, (torch.as_tensor(g["oe_xi"]), torch.as_tensor(g["oe_u"])), float(g["oe_width"]))
        assert np.abs(xn.cpu().numpy() - g["oe_x_new"]).max() < 1e-12 and abs(float(pm) - float(g["oe_pmove"])) < 1e-15
        st = qmc.make_mcmc_step(sl.apply, B0, lat, steps=g["imp_u"].shape[0], importance_sampling=sl.apply)
        xn, pm = st(P, X, (torch.as_tensor(g["imp_xi"]), torch.as_tensor(g["imp_u"])), float(g["imp_width"]))
        assert np.abs(xn.cpu().numpy() - g["imp_x_new"]).max() < 1e-9 and abs(float(pm) - float(g["imp_pmove"])) < 1e-15
    steps, B = g["u"].shape
    step = qmc.make_mcmc_step(sl.apply, B, sc.lattice_vectors(), steps=steps)
    xn, pmove, masks = step(P, X, (torch.as_tensor(g["xi"]), torch.as_tensor(g["u"])), float(g["width"]), return_masks=True)
    assert (masks.cpu().numpy().astype(bool) == g["masks"].astype(bool)).all()
    assert np.abs(xn.cpu().numpy() - g["x_new"]).max() < 1e-12


@pytest.mark.parametrize("path", FILES or [None])
def test_oracle_matches_reference_outputs(path):
    if path is None:
        pytest.skip("no tests/golden/reference_*.npz: the reference (jax + pyscf) cannot run in this image; "
                    "generate it with tests/golden/make_reference_golden.py -- parity stays unpinned until then")
    g, sc, klist, P = load_case(path)
    assert str(g["source"]) in PIN_SOURCES, "only files written by the reference's own code pin parity"
    check_geometry(g, sc)
    # (CPU budget: the forward-over-reverse Laplacian of the oracle costs seconds per walker beyond ~20 electrons; the
    #  GPU test below compares every walker of every file)
    n = sum(sc.nelec)
    modes = case_modes(g)
    # the 2 x 3N forward-over-reverse sweeps of mode 'for' cost the oracle ~50 s per walker at 54 electrons: the CPU suite
    # checks the kinetic energy up to 48 electrons ('for') / 64 ('partition'); beyond that log|psi|, orbital matrices,
    # Ewald terms, masks and geometry here, and the kinetic energy in the GPU test (oracle == fixture at those sizes was
    # verified when the fixtures were written: graphite-54 49 s, LiH-108 27 s)
    kinetic = n <= 48 or (modes == ("partition",) and n <= 64)
    check_oracle(g, sc, klist, P, max_walkers=1 if (n >= 20 or case_opts(g)) else 2, kinetic=kinetic)


@pytest.mark.gpu
@pytest.mark.parametrize("path", FILES or [None])
def test_gpu_matches_reference_outputs(path):
    if path is None:
        pytest.skip("no tests/golden/reference_*.npz (see tests/golden/make_reference_golden.py)")
    g, sc, klist, P = load_case(path)
    assert str(g["source"]) in PIN_SOURCES
    check_gpu(g, sc, klist, P)


def test_writer_and_reader_plumbing_roundtrip(tmp_path):
    """The generator's file layout is what this reader expects (oracle backend; NOT a parity pin)."""
    r = subprocess.run([sys.executable, os.path.join(GOLD, "make_reference_golden.py"), "--backend", "oracle",
                        "--out", str(tmp_path), "--batch", "2", "--steps", "2", "--burn", "2", "--cases", "reference_lih_s211"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    g, sc, klist, P = load_case(str(tmp_path / "reference_lih_s211.npz"))
    assert str(g["source"]) == "oracle-selftest"
    check_geometry(g, sc)
    # (CPU budget: the forward-over-reverse Laplacian of the oracle costs seconds per walker beyond ~20 electrons; the
    #  GPU test below compares every walker of every file)
    n = sum(sc.nelec)
    modes = case_modes(g)
    # the 2 x 3N forward-over-reverse sweeps of mode 'for' cost the oracle ~50 s per walker at 54 electrons: the CPU suite
    # checks the kinetic energy up to 48 electrons ('for') / 64 ('partition'); beyond that log|psi|, orbital matrices,
    # Ewald terms, masks and geometry here, and the kinetic energy in the GPU test (oracle == fixture at those sizes was
    # verified when the fixtures were written: graphite-54 49 s, LiH-108 27 s)
    kinetic = n <= 48 or (modes == ("partition",) and n <= 64)
    check_oracle(g, sc, klist, P, max_walkers=1 if (n >= 20 or case_opts(g)) else 2, kinetic=kinetic)


@pytest.mark.gpu
def test_gpu_reader_plumbing_roundtrip(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(GOLD, "make_reference_golden.py"), "--backend", "oracle",
                        "--out", str(tmp_path), "--batch", "3", "--steps", "3", "--burn", "5", "--cases", "reference_lih_s211"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    g, sc, klist, P = load_case(str(tmp_path / "reference_lih_s211.npz"))
    check_gpu(g, sc, klist, P)

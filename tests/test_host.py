"""CPU tests of the host logic: geometry, Ewald tables, error behaviour of the closures,
the C-ABI surface (symbols, no-GPU failure), and the 2-rank gloo path of the statistics."""
import ctypes
import os
import re
import subprocess
import sys
import numpy as np
import pytest
import torch

from conftest import ROOT, system
from deepsolid_b200 import cell as C
from deepsolid_b200 import _lib, network, hamiltonian, qmc, train
from deepsolid_b200.ewald_tables import build_ewald_tables


def test_cells_match_baseline_configs():
    want = {"h10": ((5, 5), 10, 2), "li24": ((12, 12), 8, 2), "graphite54": ((27, 27), 18, 2),
            "diamond64": ((32, 32), 16, 2), "lih108": ((54, 54), 54, 2)}
    for name, (nelec, natm, nprim) in want.items():
        sc = C.build_system(name)
        assert sc.nelec == nelec and sc.natm == natm and sc.original_cell.natm == nprim
        for cc in (sc, sc.original_cell):       # a_i . b_j = 2 pi delta_ij ; AV = a / 2 pi
            assert np.allclose(cc.a @ cc.BV.T, 2 * np.pi * np.eye(3), atol=1e-12)
            assert np.allclose(cc.AV, cc.a / (2 * np.pi), atol=1e-12)
        kl = C.make_klist(sc)
        assert [k.shape for k in kl] == [(nelec[0], 3), (nelec[1], 3)]
        assert abs(sc.vol - sc.scale * sc.original_cell.vol) < 1e-9
        assert abs(sum(sc.charges) - sum(nelec)) < 1e-12        # neutral


def test_supercell_kpts_are_supercell_reciprocal_vectors():
    sc = C.build_system("graphite54")
    k = C.get_supercell_kpts(sc)
    assert k.shape == (9, 3)
    frac = k @ sc.a.T / (2 * np.pi)       # integer combinations of the supercell reciprocal lattice
    assert np.allclose(frac, np.round(frac), atol=1e-10)


def test_ewald_table_sizes():
    tb = build_ewald_tables(C.build_system("diamond64"))
    assert tb.dist_kind == 2 and tb.lattice_displacements.shape == (27, 3)
    assert len(tb.gweight) == len(tb.gpoints) == len(tb.ion_exp) and (tb.gweight > 1e-12).all()
    # half space: no G and -G together
    s = {tuple(np.round(g, 8)) for g in tb.gpoints}
    assert not any(tuple(np.round(-np.array(g), 8)) in s for g in list(s)[:200])
    assert build_ewald_tables(C.build_system("h10")).dist_kind == 0


def test_walker_init_inside_cell():
    sc = C.build_system("graphite54")
    X = C.init_walkers(sc, 5)
    frac = X.reshape(5, -1, 3) @ np.linalg.inv(sc.a)
    assert X.shape == (5, 162) and (frac >= 0).all() and (frac < 1).all()


def test_constructor_errors_match_reference():
    sc, kl, _, _ = system("h4")
    with pytest.raises(ValueError, match="Method name"):
        network.make_solid_fermi_net(klist=kl, simulation_cell=sc, method_name="nope")
    with pytest.raises(ValueError, match="distance"):
        network.make_solid_fermi_net(klist=kl, simulation_cell=sc, distance_type="l2", envelope_type="isotropic", full_det=False)
    with pytest.raises(ValueError, match="at most 3 layers"):
        network.make_solid_fermi_net(klist=kl, simulation_cell=sc, use_last_layer=True, hidden_dims=((256, 32),) * 4)
    network.make_solid_fermi_net(klist=kl, simulation_cell=sc, use_last_layer=True)    # constructs (forward paths)
    with pytest.raises(ValueError, match="not implemented"):
        network.make_solid_fermi_net(klist=kl, simulation_cell=sc, envelope_type="output")
    network.make_solid_fermi_net(klist=kl, simulation_cell=sc)                  # reference defaults (full / full_det) construct
    net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8)
    with pytest.raises(ValueError, match="laplacian"):
        hamiltonian.local_energy_seperate(net.apply, sc, mode="nope")
    with pytest.raises(ValueError, match="one elec"):
        qmc.make_mcmc_step(net.apply, 4, sc.a, importance_sampling=lambda *a: 0, one_electron_moves=True)
    with pytest.raises(TypeError):
        hamiltonian.local_energy_seperate(lambda p, x: 0, sc)(None, None)


def test_init_params_shapes():
    sc, kl, _, _ = system("graphite54")
    net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8)
    p = net.init(0)
    assert [tuple(l["w"].shape) for l in p["single"]] == [(32, 256), (832, 256), (832, 256)]
    assert [tuple(l["w"].shape) for l in p["double"]] == [(4, 32), (32, 32)]
    assert [tuple(o["w"].shape) for o in p["orbital"]] == [(256, 432), (256, 432)]
    assert tuple(p["envelope"][0]["sigma"].shape) == (2, 216)
    from deepsolid_b200.hotpath import flatten_params
    assert len(flatten_params(p)) == 16


def test_c_abi_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "deepsolid_b200.h")).read()
    declared = set(re.findall(r"DS_API[^;(]*?\b(ds_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert declared <= exported
    assert not {e for e in exported if not e.startswith("ds_") and not e.startswith("_")}   # only the C ABI


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    sc, kl, _, P = system("h4")
    net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8)
    with pytest.raises(RuntimeError, match="CUDA"):
        net.apply(P, torch.zeros(12, dtype=torch.float64))
    lib = _lib.load()
    h = ctypes.c_void_p()
    sd, nd = _lib.SystemDesc(), _lib.NetDesc()
    rc = lib.ds_ctx_create(ctypes.byref(sd), ctypes.byref(nd), 0, ctypes.byref(h))
    assert rc != 0 and b"no CPU fallback" in lib.ds_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "deepsolid_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("oracle/", ""), fn


_WORKER = r"""
import os, sys, torch, torch.distributed as td
sys.path.insert(0, sys.argv[1])
from deepsolid_b200 import train, dist
td.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2)
r = td.get_rank()
e = torch.tensor([[1 + 2j, 3 - 1j, 0.5 + 0j], [2 - 1j, -1 + 0.5j, 4 + 4j]], dtype=torch.complex128)[r]
stats = torch.stack([e.real.sum(), e.imag.sum(), (e.abs() ** 2).sum(), e.real.sum(), torch.zeros(()).double(), torch.tensor(3.0).double()])
loss, im, var = train.reduce_energy_stats(stats)
pm = dist.pmean(torch.tensor([0.25 + 0.5 * r], dtype=torch.float64))
if r == 0:
    print("RESULT", float(loss), float(im), float(var), float(pm))
td.destroy_process_group()
"""


def test_two_rank_gloo_statistics(tmp_path):
    """world_size-2 path of train.py:76-80 / qmc.py:360-361 on CPU (gloo)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29577", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    vals = [float(v) for v in [l for l in outs[0][0].splitlines() if l.startswith("RESULT")][0].split()[1:]]
    e = torch.tensor([[1 + 2j, 3 - 1j, 0.5 + 0j], [2 - 1j, -1 + 0.5j, 4 + 4j]], dtype=torch.complex128)
    means = e.mean(1)
    loss = float(means.real.mean()); im = float(means.imag.mean())
    var = float(((e.abs() ** 2).mean(1) - means.real.abs() ** 2).mean())      # mean of per-device variances
    assert abs(vals[0] - loss) < 1e-14 and abs(vals[1] - im) < 1e-14 and abs(vals[2] - var) < 1e-14
    assert abs(vals[3] - 0.5) < 1e-15


def test_training_step_composition_and_adam_update():
    """train.make_training_step (train.py:147-185) with stub closures: order of the calls and what is returned."""
    import torch
    from deepsolid_b200 import train
    calls = []
    P = {"single": [{"w": torch.ones(2, 2, dtype=torch.float64), "b": torch.zeros(2, dtype=torch.float64)}] * 2,
         "double": [{"w": torch.ones(1, 1, dtype=torch.float64), "b": torch.zeros(1, dtype=torch.float64)}],
         "orbital": [{"w": torch.ones(2, 2, dtype=torch.float64)}] * 2,
         "envelope": [{"pi": torch.ones(1, 1, dtype=torch.float64), "sigma": torch.ones(1, 1, dtype=torch.float64)}] * 2}

    def mcmc_step(params, data, key, width):
        calls.append("mcmc")
        return data + width, torch.tensor(0.5)

    def val_and_grad(params, data):
        calls.append("grad")
        g = {k: [{kk: torch.full_like(v, 2.0) for kk, v in d.items()} for d in params[k]] for k in params}
        return (torch.tensor(-1.0), "aux"), g

    init, opt_update = train.make_adam_update(train.learning_rate_schedule(rate=0.1, decay=1.0, delay=10.0))
    step = train.make_training_step(mcmc_step, val_and_grad, opt_update)
    state = init(P)
    data = torch.zeros(3, 6, dtype=torch.float64)
    data, P1, state, loss, aux, pmove, sd = step(0, data, P, state, 7, 0.02)
    assert calls == ["mcmc", "grad"] and float(loss) == -1.0 and aux == "aux" and float(pmove) == 0.5
    assert torch.allclose(data, torch.full((3, 6), 0.02, dtype=torch.float64))
    # first Adam step moves every parameter by -lr * sign(g) (bias-corrected m / sqrt(v) = 1)
    assert torch.allclose(P1["single"][0]["w"], P["single"][0]["w"] - 0.1, atol=1e-7)
    _, P2, state, *_ = step(1, data, P1, state, 8, 0.02)
    assert torch.allclose(P2["single"][0]["w"], P1["single"][0]["w"] - 0.1 / (1 + 1 / 10.0), atol=1e-7)
    assert abs(train.learning_rate_schedule()(10000.0) - 0.025) < 1e-15


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU restatement timed on the host cores) runs without a GPU and prints ONE
    JSON line with the contract's keys."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--system", "h4",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "local_energies_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]


_WORKER_KFAC = r"""
import os, sys, torch, torch.distributed as td
sys.path.insert(0, sys.argv[1])
from deepsolid_b200 import kfac, estimator
td.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2)
r = td.get_rank()

class StubHotPath:                      # raw sums of ds_kfac_factors for a 1-layer toy, different on each rank
    def set_params(self, p): pass
    def kfac_factors(self, x):
        s = float(r + 1)
        blk = lambda n_in, n_out, rows: {"a": s * torch.ones(n_in + 1, n_in + 1, dtype=torch.float64),
                                         "g": s * torch.eye(n_out, dtype=torch.float64), "rows": rows}
        env = lambda v: [{"pi": torch.full((1, 2), v, dtype=torch.float64), "sigma": torch.full((1, 2), 2 * v, dtype=torch.float64)}] * 2
        return {"single": [blk(3, 2, 8)], "double": [], "orbital": [blk(2, 4, 4), blk(2, 4, 4)],
                "envelope_abs": env(s), "envelope_phase": env(0.5 * s), "batch": 4}

P = {"single": [{"w": torch.zeros(3, 2), "b": torch.zeros(2)}], "double": [],
     "orbital": [{"w": torch.zeros(2, 4)}, {"w": torch.zeros(2, 4)}], "envelope": []}
est = kfac.curvature_estimate(StubHotPath(), P, None, sync=True)
z = estimator._pmean_c(torch.tensor([1.0 + 2.0j, -1.0j], dtype=torch.complex128) * (r + 1))
if r == 0:
    print("RESULT", float(est["single"][0]["inputs_factor"][0, 0]), tuple(est["single"][0]["inputs_factor"].shape),
          float(est["single"][0]["outputs_factor"][1, 1]), est["single"][0]["extra_scale"],
          tuple(est["orbital"][0]["inputs_factor"].shape), float(est["orbital"][1]["outputs_factor"][0, 0]),
          complex(est["envelope"][0]["pi"][0, 0]), complex(z[0]), complex(z[1]))
td.destroy_process_group()
"""


def test_two_rank_gloo_kfac_and_observable_means(tmp_path):
    """Cross-rank mean of the per-rank curvature factors (utils.py:293-294) and of a complex observable, on CPU."""
    script = tmp_path / "worker_kfac.py"
    script.write_text(_WORKER_KFAC)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29578", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT")][0]
    got = eval("(" + line[len("RESULT "):].replace(") (", "), (").replace(" ", ", ").replace(",,", ",") + ")")
    # rank r holds s = r + 1: inputs factor s / rows, outputs factor 2 s / rows; mean over ranks = 1.5 x the s = 1 value
    assert abs(got[0] - 1.5 / 8) < 1e-15 and got[1] == (4, 4)          # bias layer keeps the homogeneous coordinate
    assert abs(got[2] - 2 * 1.5 / 8) < 1e-15 and got[3] == 2           # extra_scale = rows / batch
    assert got[4] == (2, 2)                                            # orbital layer without bias: last row / column dropped
    assert abs(got[5] - 2 * 1.5 / 4) < 1e-15
    # NaiveDiagonal statistic dw dw / B with dw = sqrt2 (g_abs - i g_phase): mean over ranks of (s^2) = 2.5
    want_pi = 2.0 * complex(1.0, -0.5) ** 2 / 4 * 2.5
    assert abs(got[6] - want_pi) < 1e-14
    assert abs(got[7] - 1.5 * (1 + 2j)) < 1e-15 and abs(got[8] - 1.5 * (-1j)) < 1e-15


_WORKER_KFAC_STEP = r"""
import os, sys, torch, torch.distributed as td
sys.path.insert(0, sys.argv[1])
from deepsolid_b200 import kfac
from deepsolid_b200.hotpath import flatten_params
td.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2)
r = td.get_rank()

class StubHotPath:
    tdev = torch.device("cpu")
    def set_params(self, p): pass
    def kfac_factors(self, x):
        s = float(r + 1)
        blk = lambda n_in, n_out, rows: {"a": s * (torch.ones(n_in + 1, n_in + 1, dtype=torch.float64) + torch.eye(n_in + 1, dtype=torch.float64)),
                                         "g": s * torch.eye(n_out, dtype=torch.float64), "rows": rows}
        env = lambda v: [{"pi": torch.full((1, 2), v, dtype=torch.float64), "sigma": torch.full((1, 2), 2 * v, dtype=torch.float64)}] * 2
        return {"single": [blk(3, 2, 8)], "double": [], "orbital": [blk(2, 4, 4), blk(2, 4, 4)],
                "envelope_abs": env(s), "envelope_phase": env(0.5 * s), "batch": 4}

def params(v):
    f = lambda *shape: torch.full(shape, v, dtype=torch.float64)
    return {"single": [{"w": f(3, 2), "b": f(2)}], "double": [], "orbital": [{"w": f(2, 4)}, {"w": f(2, 4)}],
            "envelope": [{"pi": f(1, 2), "sigma": f(1, 2)}, {"pi": f(1, 2), "sigma": f(1, 2)}]}

def vag(scale):
    return lambda p, d: ((torch.tensor(0.0), None), params(scale))

opt = kfac.Optimizer(vag(float(r + 1)), StubHotPath(), norm_constraint=1e-3)      # per-rank gradients 1 and 2
new, stats = opt.step(params(0.5), None, learning_rate=0.1, damping=1e-3)
mine = torch.cat([t.reshape(-1) for t in flatten_params(new)])
both = [torch.zeros_like(mine) for _ in range(2)]
td.all_gather(both, mine)
if r == 0:
    print("RESULT", bool(torch.equal(both[0], both[1])), float((mine - 0.5).abs().max()), stats["coefficient"])
td.destroy_process_group()
"""


def test_two_rank_gloo_kfac_step_keeps_replicas_identical(tmp_path):
    """Optimizer.step averages the per-rank energy gradient and the norm-constraint scalar over ranks
    (kfac_ferminet_alpha optimizer.py:423, :593): ranks that saw different walkers end with the SAME parameters."""
    script = tmp_path / "worker_kfac_step.py"
    script.write_text(_WORKER_KFAC_STEP)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29581", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT")][0].split()
    assert line[1] == "True", "parameter replicas diverged"
    assert float(line[2]) > 0.0 and 0.0 < float(line[3]) <= 1.0

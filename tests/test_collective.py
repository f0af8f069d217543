"""The hot path's collective behind the C ABI (ds_stats_allreduce; replaces pmean_if_pmap in train.py:78-80 and
qmc.py:360-361).  Single rank on any GPU box; the NCCL reduction itself when >= 2 GPUs are visible: two processes,
each with half of the walkers, against the single-process computation on the concatenated set -- including the
reference's mean-of-local-variances quirk (train.py:76-80) and the torch.distributed path of dist.py."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import system
from deepsolid_b200 import cell as C

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _expected(e_parts, pmoves):
    """train.py:74-80 / qmc.py:360-361 with pmean over the parts."""
    loss = np.mean([np.mean(e.real) for e in e_parts])
    imag = np.mean([np.mean(e.imag) for e in e_parts])
    var_quirk = np.mean([np.mean(np.abs(e) ** 2) - abs(np.mean(e.real)) ** 2 for e in e_parts])
    var_glob = np.mean([np.mean(np.abs(e) ** 2) for e in e_parts]) - loss ** 2
    return loss, imag, var_quirk, var_glob, float(np.mean(pmoves))


def test_single_rank_statistics_and_determinism():
    from deepsolid_b200 import network, hamiltonian
    sc, kl, _, P = system("h10")
    dev = torch.device("cuda", 0)
    ld = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc,
                                      determinants=8, method_name="eval_logdet")
    hp = ld.apply.hotpath()
    X = torch.as_tensor(C.init_walkers(sc, 301, seed=12)).to(dev)
    ke, ew = hamiltonian.local_energy_seperate(ld.apply, sc)(P, X)
    s6 = hp.energy_stats(ke, ew)
    assert torch.equal(s6, hp.energy_stats(ke, ew))                # one block, fixed order: bit-reproducible
    nacc = torch.tensor([1234.0], dtype=torch.float64, device=dev)
    e = (ke + ew).cpu().numpy()
    for gv in (False, True):
        out = hp.stats_allreduce(s6, None, nacc, moves_per_rank=20 * 301, global_variance=gv).cpu().numpy()
        loss, imag, vq, vg, _ = _expected([e], [0.0])
        assert abs(out[0] - loss) < 1e-11 and abs(out[1] - imag) < 1e-11
        assert abs(out[2] - (vg if gv else vq)) < 1e-9               # one rank: both definitions coincide
        assert out[5] == 1.0 and out[6] == 301.0 and abs(out[7] - 1234.0 / (20 * 301)) < 1e-15


_WORKER = r"""
import os, sys, json
import numpy as np, torch, torch.distributed as td
sys.path.insert(0, sys.argv[1])
from deepsolid_b200 import cell as C, network, hamiltonian, train, qmc, dist
from oracle import deepsolid_oracle as O
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
td.init_process_group("nccl", device_id=dev)
sc = C.build_system("h10"); kl = C.make_klist(sc)
P = O.params_to_torch(O.init_params(np.random.default_rng(888), sc.original_cell.natm, sc.nelec))
kw = dict(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8, device=local)
ld = network.make_solid_fermi_net(method_name="eval_logdet", **kw)
hp = ld.apply.hotpath()
sl = network.make_solid_fermi_net(method_name="eval_slogdet", hotpath=hp, **{k: v for k, v in kw.items() if k != "device"})
Xall = torch.as_tensor(C.init_walkers(sc, 2 * 96, seed=77))
X = Xall[rank * 96:(rank + 1) * 96].to(dev)
ke, ew = hamiltonian.local_energy_seperate(ld.apply, sc)(P, X)
s6 = hp.energy_stats(ke, ew)
comm = dist.native_comm(local)
assert comm is not None
xn, nacc, _ = hp.mcmc(X, 5, 0.1, seed=100 + rank)
out_q = hp.stats_allreduce(s6, comm, nacc, moves_per_rank=5 * 96, global_variance=False).cpu().numpy()
out_g = hp.stats_allreduce(s6, comm, nacc, moves_per_rank=5 * 96, global_variance=True).cpu().numpy()
# the Python closures (train.make_loss, qmc.make_mcmc_step) route through the same entry point / torch.distributed
loss, aux = train.make_loss(ld.apply, ld.apply, sc)(P, X)
py = train.reduce_energy_stats(s6)                                    # torch.distributed all-reduce of dist.py
_, pmove = qmc.make_mcmc_step(sl.apply, 96, sc.lattice_vectors(), steps=5)(P, X, 100 + rank, 0.1)
res = {"rank": rank, "e_re": (ke + ew).real.cpu().tolist(), "e_im": (ke + ew).imag.cpu().tolist(),
       "pm_local": float(nacc) / (5 * 96), "out_q": out_q.tolist(), "out_g": out_g.tolist(),
       "loss": float(loss), "var": float(aux.variance), "imag": float(aux.imaginary),
       "py": [float(v) for v in py], "pmove": float(pmove)}
open(os.path.join(sys.argv[2], f"rank{rank}.json"), "w").write(json.dumps(res))
dist.destroy_native_comm()
td.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs >= 2 visible GPUs")
def test_two_gpu_nccl_reduction_matches_single_process(tmp_path):
    import json
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29591", str(script), ROOT, str(tmp_path)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    res = [json.loads((tmp_path / f"rank{k}.json").read_text()) for k in range(2)]
    parts = [np.array(q["e_re"]) + 1j * np.array(q["e_im"]) for q in res]
    loss, imag, vq, vg, pm = _expected(parts, [q["pm_local"] for q in res])
    assert abs(vq - vg) > 1e-12                                     # the quirk is observable on this split
    for q in res:                                                   # every rank holds the reduced values
        assert abs(q["out_q"][0] - loss) < 1e-11 and abs(q["out_q"][1] - imag) < 1e-11
        assert abs(q["out_q"][2] - vq) < 1e-9 and abs(q["out_g"][2] - vg) < 1e-9
        assert q["out_q"][5] == 2.0 and q["out_q"][6] == 192.0 and abs(q["out_q"][7] - pm) < 1e-15
        assert abs(q["loss"] - loss) < 1e-11 and abs(q["var"] - vq) < 1e-9 and abs(q["imag"] - imag) < 1e-11
        assert abs(q["py"][0] - loss) < 1e-11 and abs(q["py"][2] - vq) < 1e-9
        assert abs(q["pmove"] - pm) < 1e-15
    assert res[0]["out_q"] == res[1]["out_q"]

#!/usr/bin/env python
"""bench.py -- local-energies/sec of the DeepSolid local-energy hot path on B200.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` (torchrun for N>1)
prints ONE JSON line on rank 0.  A "step" is one pass of the hot path over one batch of
synthetic walkers: ``total_energy(params, data)`` = per-walker kinetic energy
(forward-Laplacian sweep) + Ewald energy + packed statistics + one all-reduce
(train.py:66-89 forward).  ``value`` times it with walkers resident in HBM; ``e2e``
times the same call with pinned HOST walkers through the C-ABI host entry point
(H2D of walkers and D2H of the per-walker energies inside the timed region).

``--impl reference`` times the CPU oracle (torch-fp64 restatement of the reference
algorithm, hamiltonian.py:127-159 'partition'/dim_batch mode = jvp-of-grad over all 3N
directions) on all host threads; the reference itself needs JAX + pyscf which do not
exist in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from deepsolid_b200 import cell as C

METRIC = "local_energies_per_sec"
UNIT = "local-energies/s"
DEFAULT_SYSTEM = "graphite54"       # BASELINE.json configs[2]: the config `metric` is quoted on


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--system", default=DEFAULT_SYSTEM)
    ap.add_argument("--batch", type=int, default=0, help="walkers per GPU (default: BASELINE batch of the system)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="walkers per CPU-baseline sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mcmc", action="store_true", help="also time the Metropolis step (extras)")
    ap.add_argument("--grad", action="store_true", help="also time energy + parameter gradient (value_and_grad; extras)")
    ap.add_argument("--kfac", action="store_true", help="also time the KFAC curvature statistics and the forward (extras)")
    ap.add_argument("--equil", type=int, default=2, help="Metropolis calls (20 moves each) used to equilibrate walkers")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: BASELINE batch per GPU (headline); strong: the BASELINE batch is global and split over "
                         "the GPUs (process.py:96,134).  The weak line also carries the strong measurement as `strong`.")
    ap.add_argument("--no-probes", action="store_true", help="skip the cuBLAS DGEMM / int8 MMA peak probes (ncu launch lists)")
    ap.add_argument("--cpu-worker", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--threads", type=int, default=1, help=argparse.SUPPRESS)
    return ap.parse_args()


def work_model(cell, D=8, H=256, P=32):
    """SURVEY section 8d / BASELINE.md section 4 flop model per walker."""
    nu, nd = cell.nelec
    N = nu + nd
    A = cell.original_cell.natm
    f_one = 2 * N * H * ((12 * A + 8) + (3 * H + 2 * P) * 2)
    f_two = 2 * N * N * (4 * P + P * P)
    f_orb = sum(2 * ns * H * (2 * ns * D) for ns in (nu, nd))
    f_det = sum(D * (8 / 3) * ns ** 3 for ns in (nu, nd))
    f_fwd = f_one + f_two + f_orb + f_det
    f_lap = (3 * N + 2) * (f_one + f_orb) + 8 * f_two + sum(D * 3 * N * 8 * ns ** 3 for ns in (nu, nd))
    return {"F_fwd": f_fwd, "F_lap": f_lap, "F_EL": f_fwd + f_lap}


def make_inputs(system, batch, seed=666):
    cell = C.build_system(system)
    klist = C.make_klist(cell)
    X = C.init_walkers(cell, batch, seed=seed)
    return cell, klist, X


def make_params(cell):
    """Random-init parameters of the named architecture (product-side init, numpy Generator seed 888; the draw
    order is the reference's, network.py:135-184)."""
    from deepsolid_b200 import network
    return network.init_solid_fermi_net_params(888, atoms=cell.original_cell.atom_coords(), spins=cell.nelec,
                                               envelope_type="isotropic", full_det=False, determinants=8)


# ---------------------------------------------------------------------------
ORACLE_MODE = ("partition", 3)      # the reference's own faster Laplacian mode on CPU (hamiltonian.py:127-159)


def cpu_worker(args):
    """One single-purpose process of the CPU arm: builds the oracle (reference algorithm, torch fp64) for `--system`,
    then evaluates one walker per request line {"x": [...]} -> {"ke_re", "ke_im", "ew", "dt"}.  The oracle is imported
    here only: it is the CPU baseline / checker, never part of the measured GPU path."""
    torch.set_num_threads(max(1, args.threads))
    from oracle import deepsolid_oracle as O
    cell = C.build_system(args.system)
    klist = C.make_klist(cell)
    P = make_params(cell)
    f = O.make_solid_fermi_net(klist, cell, method_name="eval_logdet")
    el = O.local_energy_seperate(f, cell, mode=ORACLE_MODE[0], partition_number=ORACLE_MODE[1])
    sys.stdout.write("ready\n"); sys.stdout.flush()
    for line in sys.stdin:
        req = json.loads(line)
        if "threads" in req:
            torch.set_num_threads(int(req["threads"]))
            continue
        if "quit" in req:
            break
        x = torch.as_tensor(np.asarray(req["x"], dtype=np.float64))
        t0 = time.perf_counter()
        ke, ew = el(P, x)
        dt = time.perf_counter() - t0
        sys.stdout.write(json.dumps({"ke_re": float(ke.real), "ke_im": float(ke.imag), "ew": float(ew), "dt": dt}) + "\n")
        sys.stdout.flush()


class OraclePool:
    """The CPU arm: one oracle process per host core (single-threaded BLAS each -- measured 1.9x the throughput of
    one process with all threads on an 8-core host, because the per-direction GEMMs are small), each evaluating one
    walker per step.  The number of processes is capped by free memory (~0.3 + 3.4e-4 N^2 GB each)."""

    def __init__(self, system, n_elec, cores=None, max_procs=None):
        cores = cores or (os.cpu_count() or 1)
        self.cores = cores
        est_gb = 0.3 + 3.4e-4 * n_elec ** 2
        try:
            import psutil
            by_mem = max(1, int(0.5 * psutil.virtual_memory().available / 2 ** 30 / est_gb))
        except Exception:
            by_mem = cores
        # large systems: fewer processes with more threads each keep one step within tens of seconds
        t_single = 14.0 * (n_elec / 54.0) ** 2.75
        threads = 1
        while t_single / threads ** 0.8 > 30.0 and threads * 2 <= cores:
            threads *= 2
        n = max(1, min(cores // threads, by_mem, max_procs or cores))
        self.threads = threads
        env = dict(os.environ, OMP_NUM_THREADS=str(threads), MKL_NUM_THREADS=str(threads), CUDA_VISIBLE_DEVICES="")
        self.procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--cpu-worker", "--system", system,
                                        "--threads", str(threads)], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                       stderr=subprocess.DEVNULL, text=True, bufsize=1, env=env) for _ in range(n)]
        for p in self.procs:
            line = p.stdout.readline()
            if line.strip() != "ready":
                self.close()
                raise RuntimeError("a CPU oracle worker failed to start")

    def __len__(self):
        return len(self.procs)

    def shrink(self, n):
        """Keep n processes; the freed cores go to the survivors as BLAS threads."""
        n = max(1, min(n, len(self.procs)))
        for p in self.procs[n:]:
            self._stop(p)
        self.procs = self.procs[:n]
        self.threads = max(1, self.cores // n)
        for p in self.procs:
            p.stdin.write(json.dumps({"threads": self.threads}) + "\n"); p.stdin.flush()

    def step(self, X):
        """One walker per process, all concurrently -> (wall seconds, [(ke, ew)] in walker order)."""
        X = np.asarray(X, dtype=np.float64)
        assert X.shape[0] == len(self.procs)
        t0 = time.perf_counter()
        for p, x in zip(self.procs, X):
            p.stdin.write(json.dumps({"x": x.tolist()}) + "\n"); p.stdin.flush()
        out = []
        for p in self.procs:
            r = json.loads(p.stdout.readline())
            out.append((complex(r["ke_re"], r["ke_im"]), r["ew"]))
        return time.perf_counter() - t0, out

    @staticmethod
    def _stop(p):
        try:
            p.stdin.write(json.dumps({"quit": 1}) + "\n"); p.stdin.flush()
            p.wait(timeout=10)
        except Exception:
            p.kill()

    def close(self):
        for p in self.procs:
            self._stop(p)
        self.procs = []


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/ds_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)),
                       reasons=sorted(reasons), samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def measure_fp64_peak(dev):
    """cuBLAS DGEMM 8192^3 (MEASURED_PEAKS.json carries no fp64 figure)."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize(dev)
    best = 1e30
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12      # TFLOP/s


def measure_int8_mma_rate(hp, dev):
    """tcgen05 kind::i8 issue rate of THIS GPU: the int8-slice GEMM kernel of the sweep at the benchmark's layer shape
    with TMA and the epilogue switched off (DS_OZ_DBG=5: operands stay in shared memory, accumulators are never
    read), i.e. the tensor pipe fed from shared memory at the kernel's own instruction mix.  Tera-int8-ops/s."""
    import ctypes as Ct
    m, n, k = 148 * 64 * 80, 256, 320
    a = torch.randn(m, k, dtype=torch.float64, device=dev)
    b = torch.randn(k, n, dtype=torch.float64, device=dev)
    c = torch.empty(m, n, dtype=torch.float64, device=dev)
    gm, sm = Ct.c_double(), Ct.c_double()
    old = os.environ.get("DS_OZ_DBG")
    os.environ["DS_OZ_DBG"] = "5"
    try:
        rc = hp.lib.ds_ozaki_dgemm_probe(dev.index or 0, a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, 5,
                                         Ct.byref(gm), Ct.byref(sm), None)
    finally:
        if old is None:
            os.environ.pop("DS_OZ_DBG", None)
        else:
            os.environ["DS_OZ_DBG"] = old
    torch.cuda.synchronize(dev)
    del a, b, c
    if rc or gm.value <= 0:
        return None
    return 21 * 2.0 * m * n * k / (gm.value * 1e-3) / 1e12


def git_sha():
    try:
        sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True,
                             timeout=10).stdout.strip()
        if sha:
            return sha
    except Exception:
        pass
    try:        # snapshot on a GPU box (no .git): the revision the build recorded
        return open(os.path.join(ROOT, "deepsolid_b200", "_build", "git_sha.txt")).read().strip() or None
    except Exception:
        return None


def committed_profile(kind):
    """Newest profiles/r*_{kind}.json with its provenance (git sha, chunk size) -- ncu numbers cannot be taken in a
    timed run, so they come from the capture committed with the code and say which code that was."""
    import glob
    import re
    files = sorted(f for f in glob.glob(os.path.join(ROOT, "profiles", f"r*_{kind}.json"))
                   if re.fullmatch(rf"r\d+_{kind}\.json", os.path.basename(f)))       # default-path captures only
    for f in reversed(files):
        try:
            d = json.load(open(f))
            d["file"] = os.path.relpath(f, ROOT)
            return d
        except Exception:
            continue
    return None


def run_ours(args):
    import torch.distributed as td
    from deepsolid_b200 import network, train, qmc
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    system = args.system
    global_batch = args.batch or C.SYSTEMS[system][1]
    # weak: every GPU holds the BASELINE batch; strong: the BASELINE batch is global, split as (world, batch/world, 3N)
    batch_weak = global_batch
    batch_strong = max(1, global_batch // world)
    batch = batch_strong if args.scaling == "strong" else batch_weak
    mode, pn = C.SYSTEMS[system][2], C.SYSTEMS[system][3]
    cell, klist, X = make_inputs(system, batch, seed=666 + rank)        # every rank its own walkers
    P = make_params(cell)
    net = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=klist,
                                       simulation_cell=cell, determinants=8, method_name="eval_logdet", device=local)
    slog = network.make_solid_fermi_net(envelope_type="isotropic", full_det=False, klist=klist,
                                        simulation_cell=cell, determinants=8, method_name="eval_slogdet",
                                        hotpath=net.apply.hotpath())
    hp = net.apply.hotpath()
    total_energy = train.make_loss(net.apply, net.apply, cell, mode=mode, partition_number=pn)
    mcmc_step = qmc.make_mcmc_step(slog.apply, batch, cell.lattice_vectors(), steps=20)

    Xd = torch.as_tensor(X).to(dev)
    # short equilibration with the (parity-tested) GPU Metropolis kernel; mirrors the burn-in of process.py:256-260
    for it in range(args.equil):
        Xd, pm = mcmc_step(P, Xd, 1000 + it + 17 * rank, 0.05)
    Xh = Xd.cpu().pin_memory()
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)     # 256 MB > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            td.barrier()
            torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize(dev)
            tot += e0.elapsed_time(e1)
        barrier()
        t = torch.tensor([tot], dtype=torch.float64, device=dev)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    fp64_peak = measure_fp64_peak(dev) if (rank == 0 and not args.no_probes) else None
    i8_measured = measure_int8_mma_rate(hp, dev) if (rank == 0 and not args.no_probes) else None
    keep = {}

    def step_dev():
        loss, aux = total_energy(P, Xd)
        keep["loss"] = loss

    def step_host():
        loss, aux = total_energy(P, Xh)
        keep["loss_h"] = float(loss)

    sampler = ClockSampler(local)
    for _ in range(max(args.warmup, 3)):
        step_dev()
    torch.cuda.synchronize(dev)
    hp.profile(True)
    l0 = hp.launch_count()
    if rank == 0:
        sampler.start()
    ms = timed(step_dev, args.steps, 0)
    clocks = sampler.stop() if rank == 0 else {}
    launches = hp.launch_count() - l0
    prof = hp.profile_get()
    hp.profile(False)
    value = world * batch * args.steps / (ms / 1e3)

    strong = None
    if args.scaling == "weak":
        # the same step on the strong-scaling share of the named global batch (at N = 1 it is the same measurement)
        if world == 1:
            strong = {"value": value, "ms_per_step": ms / args.steps, "global_batch": global_batch, "batch_per_gpu": batch}
        else:
            Xs = Xd[:batch_strong].contiguous()

            def step_strong():
                loss_s, _ = total_energy(P, Xs)
                keep["loss_s"] = loss_s
            ms_s = timed(step_strong, args.steps, 2)
            strong = {"value": world * batch_strong * args.steps / (ms_s / 1e3), "ms_per_step": ms_s / args.steps,
                      "global_batch": batch_strong * world, "batch_per_gpu": batch_strong}

    e2e = None
    if not args.no_e2e:
        ms_h = timed(step_host, args.steps, 1)
        e2e = {"value": world * batch * args.steps / (ms_h / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(batch * X.shape[1] * 8), "d2h_bytes_per_step": int(3 * batch * 8),
               "ms_per_step": ms_h / args.steps}
    extras = {}
    if args.mcmc:
        keepx = {"x": Xd}

        def step_mcmc():
            keepx["x"], _ = mcmc_step(P, keepx["x"], 7, 0.02)
        ms_m = timed(step_mcmc, max(1, args.steps // 2), 1)
        extras["mcmc_moves_per_sec"] = world * batch * 20 * max(1, args.steps // 2) / (ms_m / 1e3)

    if args.grad:
        def step_grad():
            (loss_g, aux_g), grads = total_energy.value_and_grad(P, Xd)
            keep["gnorm"] = grads["single"][1]["w"]
        ms_g = timed(step_grad, max(1, args.steps // 2), 1)
        extras["value_and_grad_walkers_per_sec"] = world * batch * max(1, args.steps // 2) / (ms_g / 1e3)
        extras["grad_single1_w_norm"] = float(keep["gnorm"].norm())
    if args.kfac:
        from deepsolid_b200 import kfac as kfac_mod

        def step_kfac():
            keep["kf"] = kfac_mod.curvature_estimate(hp, P, Xd, sync=world > 1)
        ms_k = timed(step_kfac, max(1, args.steps // 2), 1)
        extras["kfac_factor_walkers_per_sec"] = world * batch * max(1, args.steps // 2) / (ms_k / 1e3)

        def step_fwd():
            keep["lp"] = hp.logpsi(Xd)
        ms_f = timed(step_fwd, max(1, args.steps // 2) * 4, 1)
        extras["forwards_per_sec"] = world * batch * max(1, args.steps // 2) * 4 / (ms_f / 1e3)
    if rank != 0:
        if world > 1:
            td.destroy_process_group()
        return
    wm = work_model(cell)
    jac_tf = prof["jac_flops"] / (prof["jac_ms"] / 1e3) / 1e12 if prof["jac_ms"] > 0 else None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    i8 = os.environ.get("DS_NO_I8", "0") in ("", "0")
    # ncu-measured DRAM bytes (per launch of the dominant kernel; per walker over the whole pass) from the capture
    # committed with this code: each file records the git sha and chunk size it was taken at
    traffic = hbm = None
    if i8 and system == DEFAULT_SYSTEM:
        tr = committed_profile("traffic")
        if tr and "kernels" in tr:
            traffic = {"per_launch_dram_bytes": {k: v.get("dram_bytes_per_launch") for k, v in tr["kernels"].items()},
                       "chunk_walkers": tr.get("chunk_walkers"), "git_sha": tr.get("git_sha"), "source": tr["file"]}
        hb = committed_profile("hbm")
        if hb and "dram_bytes_per_walker" in hb:
            gbs = hb["dram_bytes_per_walker"] * batch * args.steps / (ms / 1e3) / 1e9
            hbm = {"achieved_gbs": gbs, "frac": gbs / peaks["hbm_gbs"] if peaks.get("hbm_gbs") else None,
                   "dram_bytes_per_walker": hb["dram_bytes_per_walker"], "git_sha": hb.get("git_sha"),
                   "chunk_walkers": hb.get("chunk_walkers"),
                   "source": hb["file"] + " (ncu dram__bytes_read+write summed over every kernel of one pass) / step time"}
    bf16_peak = peaks.get("bf16_tflops_sustained") or peaks.get("bf16_tflops") or 1590.0
    peak_note = ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks.get("bf16_tflops_sustained")
                 else ("MEASURED_PEAKS.json bf16_tflops" if peaks.get("bf16_tflops") else "fallback 1.59 PFLOP/s (MEASURED_PEAKS.json absent)"))
    if i8:
        # tcgen05 kind::i8 runs at twice the dense bf16 rate; one fp64-accurate product costs 21 int8 slice products
        oz_products = 21
        i8_peak = 2.0 * bf16_peak
        eq_peak = i8_peak / oz_products
        roofline = {
            "bound": "tensor",
            "kernel": "oz_gemm_kernel<JAC|ORBJ> (tcgen05 kind::i8 sliced-integer fp64 GEMM of the Jacobian sweep)",
            "achieved": jac_tf, "peak": eq_peak, "unit": "TFLOP/s",
            "frac": (jac_tf / eq_peak) if jac_tf else None,
            "achieved_def": "algorithmic fp64 flops 2*M*N*K of the Jacobian-sweep GEMM launches / their CUDA-event time",
            "peak_source": f"fp64-equivalent ceiling of the int8 tensor pipe = 2 x {bf16_peak:.1f} TF/s ({peak_note}; kind::i8 = 2x bf16 rate) "
                           f"/ {oz_products} slice products (6x6 digits, diagonals s+t<6)",
            "int8_tops_executed": (jac_tf * oz_products) if jac_tf else None, "int8_tops_peak": i8_peak,
            "int8_tops_measured": i8_measured,
            "int8_tops_measured_def": "this kernel's MMA issue loop alone at 757760x256x320 (no TMA, no epilogue), same run",
            "frac_of_measured_int8": (jac_tf * oz_products / i8_measured) if (jac_tf and i8_measured) else None,
            "fp64_dmma_peak_tflops": fp64_peak,
            "frac_of_fp64_dmma_peak": (jac_tf / fp64_peak) if (jac_tf and fp64_peak) else None,
            "hbm_gbs_peak": peaks.get("hbm_gbs"), "hbm": hbm,
            "launches": prof["jac_launches"], "kernel_ms_per_step": prof["jac_ms"] / args.steps,
            "kernel_share_of_step": prof["jac_ms"] / ms,
            "traffic": traffic,
            "survey_model_tflops": wm["F_EL"] * batch * args.steps / (ms / 1e3) / 1e12,
        }
    else:
        roofline = {
            "bound": "tensor", "kernel": "gemm_f64_kernel<JAC|ORBJ> (fp64 DMMA Jacobian sweep)",
            "achieved": jac_tf, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": (jac_tf / fp64_peak) if jac_tf else None,
            "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json has no fp64 entry; "
                           f"its bf16 burst figure is {peaks.get('bf16_tflops')} TF/s, hbm {peaks.get('hbm_gbs')} GB/s)",
            "launches": prof["jac_launches"], "kernel_ms_per_step": prof["jac_ms"] / args.steps,
            "kernel_share_of_step": prof["jac_ms"] / ms,
            "traffic": None,
            "survey_model_tflops": wm["F_EL"] * batch * args.steps / (ms / 1e3) / 1e12,
        }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{system}: {cell.nelectron} e-, {cell.natm} atoms, batch {batch}/GPU, "
                               f"laplacian mode {mode}, 8 dets, hidden ((256,32),)*3",
                   "system": system, "batch_per_gpu": batch, "global_batch": batch * world,
                   "walkers": f"gaussian init (seed 666+rank) + {20 * args.equil} GPU Metropolis moves; params N(0,1)/sqrt(fan_in) seed 888",
                   "l2": "256 MB flush between timed steps; per-step workspace >> L2",
                   "chunk_walkers": hp.workspace_info()["chunk_walkers"], "git_sha": git_sha(),
                   "gemm_arith": ("fp64 results; Jacobian-sweep GEMMs as error-free int8 digit slices on tcgen05 (int32 accumulation, "
                                  "one fp64 rounding per output)") if i8 else "fp64 DMMA",
                   "parallelism": f"walker-sharded dp{world}, one all-reduce of 4 doubles per step"},
        "clocks": clocks, "gpu_launches": int(launches),
        "roofline": roofline, "loss": float(keep["loss"]),
    }
    if e2e:
        line["e2e"] = e2e
    if strong:
        line["strong"] = strong
    if extras:
        line["extras"] = extras
    if not args.no_cpu_baseline:
        pool = OraclePool(system, cell.nelectron)
        try:
            S = len(pool)
            Xc = Xh.numpy()
            pool.step(Xc[S:2 * S] if Xc.shape[0] >= 2 * S else Xc[:S])       # untimed: first-call costs of every process
            dt, out = pool.step(Xc[:S])
        finally:
            pool.close()
        # the same walkers through the GPU: report the agreement next to the rate
        ke_g, ew_g = net.apply.hotpath().local_energy(Xd[:S])
        d = max(abs(k + e - complex(kg) - float(eg)) for (k, e), kg, eg in zip(out, ke_g.cpu(), ew_g.cpu()))
        line["cpu_baseline"] = {"value": S / dt, "unit": UNIT, "cores": S * pool.threads, "kind": "port",
                                "sample": f"{S} walkers of the same batch evaluated concurrently by {S} oracle processes x "
                                          f"{pool.threads} thread(s) (reference algorithm, Laplacian mode "
                                          f"'{ORACLE_MODE[0]}' P={ORACLE_MODE[1]}: vmapped jvp-of-grad), {dt:.1f} s; "
                                          f"host has {os.cpu_count()} cores",
                                "max_abs_diff_vs_gpu_Ha": d}
    print(json.dumps(line), flush=True)
    if world > 1:
        td.destroy_process_group()


def run_reference(args):
    """The reference's algorithm on the host cores (CPU oracle; JAX + pyscf are absent from the image): rank 0 only.
    A step = one walker on each of S single-thread oracle processes, all concurrent; S starts at the core count and is
    lowered (the survivors get the freed cores as BLAS threads) only if W + K steps would not finish in ~2.5 minutes."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    system = args.system
    batch = args.batch or C.SYSTEMS[system][1]
    cell, klist, X = make_inputs(system, batch)
    pool = OraclePool(system, cell.nelectron)
    try:
        S = len(pool)
        t_first, _ = pool.step(X[:S])                    # untimed: first-call costs of every process; calibrates the step time
        budget = 200.0
        n_steps = args.warmup + args.steps
        if t_first * n_steps > budget and S > 1:
            pool.shrink(max(1, int(S * budget / (t_first * n_steps))))
            S = len(pool)
        off, times = S, []
        for it in range(n_steps):
            if off + S > X.shape[0]:
                off = 0
            dt, _ = pool.step(X[off:off + S])
            off += S
            if it >= args.warmup:
                times.append(dt)
        threads = pool.threads
    finally:
        pool.close()
    t = sum(times)
    value = S * len(times) / t
    mode = C.SYSTEMS[system][2]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / len(times), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{system}: {cell.nelectron} e-, {cell.natm} atoms, batch {batch}/GPU, "
                               f"laplacian mode {mode}, 8 dets, hidden ((256,32),)*3",
                   "system": system, "batch_per_gpu": batch, "global_batch": batch * args.gpus,
                   "note": "CPU torch-fp64 restatement of the reference algorithm (JAX/pyscf absent from the image); "
                           "each step = a bounded sample of the batch, rate is per walker"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": S * threads, "kind": "port",
                         "sample": f"{S} walkers per step, one per oracle process ({S} processes x {threads} thread(s), "
                                   f"Laplacian mode '{ORACLE_MODE[0]}' P={ORACLE_MODE[1]}); host has {os.cpu_count()} cores"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.cpu_worker:
        cpu_worker(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

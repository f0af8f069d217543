"""GPU parity at the sizes the benchmark is quoted on (BASELINE.json configs 3, 4, 5 and 2).

(a) (kinetic, ewald) of equilibrated walkers against the CPU oracle (reference algorithm: forward-over-reverse
    sweeps, hamiltonian.py:73-101 `dim_batch` / :127-159 `partition`), tolerance 1e-8 Ha per walker;
(b) the whole int8-slice path (tcgen05 kind::i8 Jacobian sweep) against the repo's fp64 DMMA path on >= 1024
    equilibrated walkers per configuration: max |dE_L| <= 1e-8 Ha, the 99.9-percentile is printed (SURVEY 8c:
    walkers near a node amplify error, so the tail of the distribution is what matters).
Walkers are equilibrated with the (mask-bit-exact) GPU Metropolis kernel from the benchmark's seeded blobs
(SURVEY 8d), `burn` moves of width 0.1; the oracle gets the same float64 positions."""
import numpy as np
import pytest
import torch

from conftest import system
from deepsolid_b200 import cell as C
from deepsolid_b200 import network, hamiltonian, qmc
from oracle import deepsolid_oracle as O

pytestmark = pytest.mark.gpu
TOL_E = 1e-8
_state = {}


def dev():
    return torch.device("cuda", 0)


def setup(name, n_walkers, burn=120):
    """(logdet net, hotpath, params, equilibrated walkers on the device), cached per system."""
    if name not in _state:
        sc, kl, _, P = system(name)
        kw = dict(envelope_type="isotropic", full_det=False, klist=kl, simulation_cell=sc, determinants=8)
        ld = network.make_solid_fermi_net(method_name="eval_logdet", **kw)
        hp = ld.apply.hotpath()
        sl = network.make_solid_fermi_net(method_name="eval_slogdet", hotpath=hp, **kw)
        _state[name] = [ld, sl, hp, None]
    ld, sl, hp, X = _state[name]
    sc, kl, _, P = system(name)
    if X is None or X.shape[0] < n_walkers:
        X = torch.as_tensor(C.init_walkers(sc, n_walkers, seed=777)).to(dev())
        step = qmc.make_mcmc_step(sl.apply, n_walkers, sc.lattice_vectors(), steps=20)
        pm = 0.0
        for it in range(burn // 20):
            X, pm = step(P, X, 1000 + it, 0.1)
        assert 0.02 < float(pm) < 0.98, f"equilibration does not move / never rejects: pmove = {float(pm)}"
        _state[name][3] = X
    return ld, hp, P, _state[name][3][:n_walkers]


@pytest.mark.parametrize("name,n,mode,pn,omode", [
    ("graphite54", 8, "for", 3, "dim_batch"),
    ("diamond64", 4, "partition", 3, "partition"),      # BASELINE config 4: the partition_number = 3 path on both sides
    ("lih108", 2, "for", 3, "dim_batch"),
    ("li24", 8, "for", 3, "dim_batch"),
])
def test_local_energy_matches_oracle_at_baseline_size(name, n, mode, pn, omode):
    ld, hp, P, X = setup(name, max(n, 64))
    sc, kl, _, _ = system(name)
    X = X[:n]
    ke, ew = hamiltonian.local_energy_seperate(ld.apply, sc, mode=mode, partition_number=pn)(P, X)
    f = O.make_solid_fermi_net(kl, sc, method_name="eval_logdet")
    elo = O.local_energy_seperate(f, sc, mode=omode, partition_number=pn)
    v = ld.apply(P, X).cpu()
    worst = 0.0
    for b in range(n):
        xo = X[b].cpu()
        ko, eo = elo(P, xo)
        vo = f(P, xo)
        d_ke = abs(complex(ko) - complex(ke[b].cpu()))
        d_ew = abs(float(eo) - float(ew[b]))
        worst = max(worst, d_ke + d_ew)
        assert d_ke < TOL_E, f"{name} walker {b}: |d kinetic| = {d_ke:.3e} Ha"
        assert d_ew < 1e-9, f"{name} walker {b}: |d ewald| = {d_ew:.3e} Ha"
        assert abs(float(vo.real) - float(v[b].real)) < 1e-9
    print(f"\n[{name}] {n} equilibrated walkers vs oracle ({omode}): max |dE_L| = {worst:.3e} Ha")


@pytest.mark.parametrize("name,n", [("graphite54", 1024), ("diamond64", 1024), ("lih108", 1024), ("li24", 1024)])
def test_int8_slice_path_matches_fp64_dmma_path_full_batch(name, n):
    ld, hp, P, X = setup(name, n)
    sc, kl, _, _ = system(name)
    mode, pn = C.SYSTEMS[name][2], C.SYSTEMS[name][3]
    el = hamiltonian.local_energy_seperate(ld.apply, sc, mode=mode, partition_number=pn)
    ke8, ew8 = el(P, X)
    hp.debug_set("i8", 0)                       # every GEMM of the sweep on the fp64 DMMA kernels (same as DS_NO_I8=1)
    try:
        ke64, ew64 = el(P, X)
    finally:
        hp.debug_set("i8", 1)
    assert torch.isfinite(ke64.real).all() and torch.isfinite(ke8.real).all()
    d = (ke8 - ke64).abs().double().cpu().numpy()
    q999 = float(np.quantile(d, 0.999))
    print(f"\n[{name}] int8-slice vs fp64 DMMA on {n} equilibrated walkers: max |dE_L| = {d.max():.3e} Ha, "
          f"99.9 % = {q999:.3e}, median = {np.median(d):.3e}; max |E_kin| = {float(ke64.abs().max()):.3e}")
    assert d.max() < TOL_E
    assert torch.equal(ew8, ew64)


@pytest.mark.parametrize("name,n", [("graphite54", 512), ("lih108", 128)])
def test_fused_digit_path_matches_default_at_baseline_size(name, n):
    """The fused-digit sweep (cluster pairs, DSMEM row-max exchange, MN-major operands) on equilibrated walkers of the
    benchmark systems against the default path: same tolerance as the int8-vs-fp64 comparison."""
    ld, hp, P, X = setup(name, n)
    sc, kl, _, _ = system(name)
    el = hamiltonian.local_energy_seperate(ld.apply, sc, mode=C.SYSTEMS[name][2], partition_number=C.SYSTEMS[name][3])
    ke0, _ = el(P, X)
    hp.debug_set("fused_digits", 1)
    try:
        ke1, _ = el(P, X)
        ke2, _ = el(P, X)
    finally:
        hp.debug_set("fused_digits", 0)
    d = (ke1 - ke0).abs().double().cpu().numpy()
    print(f"\n[{name}] fused-digit vs default sweep on {n} walkers: max |dE_L| = {d.max():.3e} Ha, median = {np.median(d):.3e}")
    assert torch.equal(ke1, ke2) and d.max() < TOL_E

"""Checkpoint interop (checkpoint.py:92-165): round trip, the reference's shape checks, and reading a file whose
parameter leaves were pickled as jax DeviceArrays -- without jax installed."""
import pickle
import sys
import types
import zipfile

import numpy as np
import pytest
import torch

from conftest import system
from deepsolid_b200 import cell as C, checkpoint


def _tree_equal(a, b):
    if isinstance(a, dict):
        return all(_tree_equal(a[k], b[k]) for k in a)
    if isinstance(a, list):
        return all(_tree_equal(x, y) for x, y in zip(a, b))
    return torch.equal(torch.as_tensor(a), torch.as_tensor(b))


def test_round_trip_and_checks(tmp_path):
    sc, kl, pn, P = system("h4")
    data = torch.as_tensor(C.init_walkers(sc, 8, seed=1)).reshape(2, 4, -1)
    f1 = checkpoint.save(str(tmp_path), 7, data, P, opt_state=None, mcmc_width=0.03)
    checkpoint.save(str(tmp_path), 3, data, P)
    assert checkpoint.find_last_checkpoint(str(tmp_path)) == f1
    assert checkpoint.find_last_checkpoint(str(tmp_path / "missing")) is None
    t, d, params, opt_state, width = checkpoint.restore(f1, batch_size=8, n_devices=2)
    assert t == 8 and width == pytest.approx(0.03) and opt_state is None
    assert torch.equal(d, data) and _tree_equal(params, P)
    with pytest.raises(ValueError):
        checkpoint.restore(f1, batch_size=16)
    with pytest.raises(ValueError):
        checkpoint.restore(f1, n_devices=1)
    # the file is a plain npz with the reference's keys and a device axis on every parameter leaf
    z = np.load(f1, allow_pickle=True)
    assert sorted(z.files) == ["data", "mcmc_width", "opt_state", "params", "t"]
    assert z["params"].tolist()["single"][1]["w"].shape == (2, 832, 256)


class _FakeDeviceArray:
    """Pickles like jax 0.2.x DeviceArray: (reconstruct_device_array, (fun, args, arr_state, aval_state))."""
    def __init__(self, value):
        self.value = np.asarray(value)

    def __reduce__(self):
        fun, args, state = self.value.__reduce__()
        return (sys.modules["jax._src.device_array"].reconstruct_device_array, (fun, args, state, {"weak_type": False}))


def test_reads_jax_pickled_leaves_without_jax(tmp_path):
    assert "jax" not in sys.modules or isinstance(sys.modules["jax"], types.ModuleType)
    sc, kl, pn, P = system("h4")
    fake = types.ModuleType("jax._src.device_array")

    def reconstruct_device_array(*a):      # what jax would call; must never run here
        raise AssertionError("jax reconstructor must not be executed")
    reconstruct_device_array.__module__ = "jax._src.device_array"
    reconstruct_device_array.__qualname__ = "reconstruct_device_array"
    fake.reconstruct_device_array = reconstruct_device_array
    mods = {"jax": types.ModuleType("jax"), "jax._src": types.ModuleType("jax._src"), "jax._src.device_array": fake}
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        def wrap(v):
            if isinstance(v, dict):
                return {k: wrap(x) for k, x in v.items()}
            if isinstance(v, list):
                return [wrap(x) for x in v]
            return _FakeDeviceArray(np.broadcast_to(v.numpy(), (1,) + tuple(v.shape)).copy())
        data = C.init_walkers(sc, 4, seed=2).reshape(1, 4, -1)
        fname = str(tmp_path / "qmcjax_ckpt_000011.npz")
        with open(fname, "wb") as f:
            np.savez(f, t=11, data=data, params=np.asarray(wrap(P), dtype=object), opt_state=np.asarray(None, dtype=object),
                     mcmc_width=_FakeDeviceArray(np.asarray([0.02])))
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    with zipfile.ZipFile(fname) as z:
        assert b"jax._src.device_array" in z.read("params.npy")
    t, d, params, _, width = checkpoint.restore(fname, batch_size=4, n_devices=1)
    assert t == 12 and width == pytest.approx(0.02)
    assert _tree_equal(params, P) and np.array_equal(d.numpy(), data)


@pytest.mark.parametrize("envelope_type", ["diagonal", "full"])
@pytest.mark.parametrize("ndev", [1, 2])
def test_anisotropic_envelope_leaves_round_trip(tmp_path, envelope_type, ndev):
    """sigma is (A, 3, q) / (3, 3, A, q) for the diagonal / full envelopes (network.py:146-152): the device axis is
    dropped whatever the leaf's own rank is."""
    from deepsolid_b200 import network
    sc, kl, pn, _ = system("h4")
    P = network.init_solid_fermi_net_params(3, atoms=sc.original_cell.atom_coords(), spins=sc.nelec,
                                            envelope_type=envelope_type)
    data = torch.as_tensor(C.init_walkers(sc, 4, seed=1)).reshape(ndev, 4 // ndev, -1)
    f = checkpoint.save(str(tmp_path), 1, data, P)
    _, d, params, _, _ = checkpoint.restore(f, batch_size=4, n_devices=ndev)
    assert _tree_equal(params, P)
    want = (2, 3, 16) if envelope_type == "diagonal" else (3, 3, 2, 16)      # A = 2, q = n_s * D = 2 * 8
    assert tuple(params["envelope"][0]["sigma"].shape) == want


def test_diverged_replicas_and_foreign_globals_are_refused(tmp_path):
    sc, kl, pn, P = system("h4")
    data = C.init_walkers(sc, 4, seed=2).reshape(2, 2, -1)

    def rep(v, bump=0.0):
        if isinstance(v, dict):
            return {k: rep(x, bump) for k, x in v.items()}
        if isinstance(v, list):
            return [rep(x, bump) for x in v]
        a = np.broadcast_to(v.numpy(), (2,) + tuple(v.shape)).copy()
        a[1] += bump
        return a
    fname = str(tmp_path / "qmcjax_ckpt_000001.npz")
    np.savez(open(fname, "wb"), t=1, data=data, params=np.asarray(rep(P, 1e-3), dtype=object),
             opt_state=np.asarray(None, dtype=object), mcmc_width=np.asarray(0.02))
    with pytest.raises(ValueError, match="diverged"):
        checkpoint.restore(fname)

    class Evil:
        def __reduce__(self):
            import os
            return (os.system, ("true",))
    # a foreign global in the parameters is refused outright ...
    np.savez(open(fname, "wb"), t=1, data=data, params=np.asarray({"single": Evil()}, dtype=object),
             opt_state=np.asarray(None, dtype=object), mcmc_width=np.asarray(0.02))
    with pytest.raises(pickle.UnpicklingError, match="allow-list"):
        checkpoint.restore(fname)
    # ... and one in the optimiser state only costs the optimiser state
    np.savez(open(fname, "wb"), t=1, data=data, params=np.asarray(rep(P), dtype=object),
             opt_state=np.asarray({"s": Evil()}, dtype=object), mcmc_width=np.asarray(0.02))
    t, d, params, opt_state, _ = checkpoint.restore(fname)
    assert opt_state is None and _tree_equal(params, P)

"""Checkpoint interop with DeepSolid (checkpoint.py:92-165): ``qmcjax_ckpt_%06d.npz`` files holding
``t, data, params, opt_state, mcmc_width``.

* ``data``: walkers ``(n_devices, batch/n_devices, 3N)`` (process.py:96,134);
* ``params``: the network pytree, every leaf replicated over a leading device axis
  (constants.replicate_all_local_devices, process.py:137), pickled inside the npz;
* files written by the reference pickle ``jax`` DeviceArrays; they are read here WITHOUT jax: the
  unpickler maps jax's array reconstructors onto plain numpy arrays.

``restore`` returns torch tensors laid out for this package (parameters without the device axis,
walkers with it); ``save`` writes files the reference's own ``checkpoint.restore`` can read."""
from __future__ import annotations

import io
import os
import pickle
import zipfile
from typing import Optional

import numpy as np
import torch

PREFIX = "qmcjax_ckpt_"


def find_last_checkpoint(ckpt_path: Optional[str] = None) -> Optional[str]:
    """checkpoint.py:42-68: newest readable checkpoint of a directory, or None."""
    if ckpt_path and os.path.exists(ckpt_path):
        files = [f for f in os.listdir(ckpt_path) if PREFIX in f]
        for file in sorted(files, reverse=True):
            fname = os.path.join(ckpt_path, file)
            try:
                with zipfile.ZipFile(fname) as z:
                    if z.testzip() is None:
                        return fname
            except (OSError, EOFError, zipfile.BadZipFile):
                continue
    return None


def _np_reconstruct(fun, args, arr_state, *unused):
    """Stand-in for jax's (_)reconstruct_device_array: rebuild the numpy value, do not device_put it."""
    value = fun(*args)
    value.__setstate__(arr_state)
    return value


_SAFE_GLOBALS = {
    ("builtins", n) for n in ("dict", "list", "tuple", "set", "frozenset", "int", "float", "complex", "bool", "str",
                              "bytes", "bytearray", "slice", "range")
} | {("collections", "OrderedDict"), ("collections", "defaultdict"), ("collections", "namedtuple")}
_SAFE_MODULE_ROOTS = ("numpy",)


class _NoJaxUnpickler(pickle.Unpickler):
    """Unpickler for the object members of a checkpoint: numpy reconstructors, plain containers and jax's
    device-array reconstructor (mapped to numpy) only.  Anything else raises -- a checkpoint is data, and
    resolving arbitrary globals would execute code chosen by whoever wrote the file."""

    def find_class(self, module, name):
        root = module.split(".")[0]
        if root in ("jax", "jaxlib"):
            if "reconstruct" in name:
                return _np_reconstruct
            raise pickle.UnpicklingError(f"checkpoint references {module}.{name}, which cannot be read without jax")
        if root in _SAFE_MODULE_ROOTS or (module, name) in _SAFE_GLOBALS:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"checkpoint references {module}.{name}: not on the allow-list of this reader")


def _load_member(z: zipfile.ZipFile, key: str):
    with z.open(key + ".npy") as f:
        raw = io.BytesIO(f.read())
    version = np.lib.format.read_magic(raw)
    if version == (1, 0):
        shape, fortran, dtype = np.lib.format.read_array_header_1_0(raw)
    else:
        shape, fortran, dtype = np.lib.format.read_array_header_2_0(raw)
    if dtype.hasobject:
        obj = _NoJaxUnpickler(raw).load()
        return obj if isinstance(obj, np.ndarray) else np.asarray(obj, dtype=object)
    raw.seek(0)
    return np.load(raw, allow_pickle=False)


def _to_native(a):
    return a.tolist() if isinstance(a, np.ndarray) else a


def _strip_device_axis(tree, n_devices: int):
    """Every parameter leaf of a reference checkpoint is replicated over a leading device axis
    (process.py:136 `replicate_all_local_devices`; checkpoint.py:92-122 saves it as is): check and drop it, whatever
    the leaf's own rank is (envelope sigma is (A, q), (A, 3, q) or (3, 3, A, q) depending on envelope_type)."""
    def conv(v, path):
        if isinstance(v, dict):
            return {k: conv(x, f"{path}/{k}") for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [conv(x, f"{path}/{i}") for i, x in enumerate(v)]
        a = np.asarray(v, dtype=np.float64)
        if a.ndim < 1 or a.shape[0] != n_devices:
            raise ValueError(f"parameter leaf {path} has shape {a.shape}: no leading axis over the {n_devices} device(s) "
                             "the walkers were saved from")
        if n_devices > 1 and not np.array_equal(a[0], a[-1]):
            raise ValueError(f"parameter leaf {path} differs between devices: replicas have diverged")
        return torch.as_tensor(np.ascontiguousarray(a[0]))
    return {k: conv(v, k) for k, v in tree.items()}


def restore(restore_filename: str, batch_size: Optional[int] = None, n_devices: Optional[int] = None,
            shape_check: bool = True):
    """checkpoint.py:125-165.  Returns ``(t, data, params, opt_state, mcmc_width)``: ``t`` = iterations
    completed, ``data`` float64 tensor ``(n_devices, batch/n_devices, 3N)``, ``params`` the torch pytree
    (device axis removed).  ``n_devices`` (default: the file's) and ``batch_size`` are checked like the
    reference does (ValueError on mismatch)."""
    with zipfile.ZipFile(restore_filename) as z:
        t = int(_to_native(_load_member(z, "t"))) + 1
        data = np.asarray(_load_member(z, "data"), dtype=np.float64)
        params = _to_native(_load_member(z, "params"))
        try:        # optimiser state may pickle classes this reader refuses or cannot import: params / walkers do not need it
            opt_state = _to_native(_load_member(z, "opt_state"))
        except (pickle.UnpicklingError, ImportError, AttributeError, KeyError):
            opt_state = None
        mcmc_width = _to_native(_load_member(z, "mcmc_width"))
    if data.ndim != 3:
        raise ValueError(f"walkers in a checkpoint must have shape (devices, batch/devices, 3N); found {data.shape}")
    if shape_check:
        if n_devices is not None and data.shape[0] != n_devices:
            raise ValueError("Incorrect number of devices found. Expected {}, found {}.".format(data.shape[0], n_devices))
        if batch_size and data.shape[0] * data.shape[1] != batch_size:
            raise ValueError("Wrong batch size in loaded data. Expected {}, found {}.".format(
                batch_size, data.shape[0] * data.shape[1]))
    if not isinstance(params, dict) or "single" not in params:
        raise ValueError("checkpoint does not hold a solid-FermiNet parameter pytree")
    params = _strip_device_axis(params, data.shape[0])
    if isinstance(mcmc_width, (list, np.ndarray)):
        mcmc_width = float(np.asarray(mcmc_width).reshape(-1)[0])
    return t, torch.as_tensor(data), params, opt_state, mcmc_width


def save(save_path: str, t: int, data, params, opt_state=None, mcmc_width=None) -> str:
    """checkpoint.py:92-122: ``save_path/qmcjax_ckpt_%06d.npz``.  Parameters are written as numpy arrays
    replicated over the leading device axis of ``data`` so that the reference can restore the file."""
    os.makedirs(save_path, exist_ok=True)
    data = np.asarray(torch.as_tensor(data).detach().cpu(), dtype=np.float64)
    if data.ndim == 2:
        data = data[None]
    ndev = data.shape[0]

    def conv(v):
        if isinstance(v, dict):
            return {k: conv(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)):
            return [conv(x) for x in v]
        a = np.asarray(torch.as_tensor(v).detach().cpu(), dtype=np.float64)
        return np.broadcast_to(a, (ndev,) + a.shape).copy()

    fname = os.path.join(save_path, f"{PREFIX}{t:06d}.npz")
    with open(fname, "wb") as f:
        np.savez(f, t=t, data=data, params=np.asarray(conv(params), dtype=object), opt_state=np.asarray(opt_state, dtype=object),
                 mcmc_width=np.asarray(mcmc_width if mcmc_width is not None else np.nan))
    return fname

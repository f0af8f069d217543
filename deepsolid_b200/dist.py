"""torch.distributed plumbing that replaces constants.pmean_if_pmap (constants.py:30-45):
one process per GPU, NCCL (or gloo on CPU) all-reduce of small statistics vectors."""
from __future__ import annotations

import torch
import torch.distributed as td


def world_size() -> int:
    return td.get_world_size() if td.is_available() and td.is_initialized() else 1


def psum(t: torch.Tensor) -> torch.Tensor:
    """Sum over ranks.  Host tensors (the host-buffer entry points return host statistics) are staged
    through the current CUDA device when the process group is NCCL, which only reduces device memory."""
    if world_size() > 1:
        if not t.is_cuda and td.get_backend() == "nccl":
            d = t.to(torch.device("cuda", torch.cuda.current_device()))
            td.all_reduce(d, op=td.ReduceOp.SUM)
            return d.to(t.device)
        t = t.clone()
        td.all_reduce(t, op=td.ReduceOp.SUM)
    return t


def pmean(t: torch.Tensor) -> torch.Tensor:
    n = world_size()
    return psum(t) / n if n > 1 else t


# ---------------------------------------------------------------------------
# the C-ABI collective (ds_stats_allreduce): one NCCL communicator per process, created from an id that rank 0
# obtains through the library and ships to the other ranks over the torch.distributed control plane
# ---------------------------------------------------------------------------
_native = {"comm": None, "tried": False}


def native_comm(device_index: int):
    """ncclComm_t (as a ctypes void pointer) for ds_stats_allreduce, or None when the process group is not NCCL
    (single process, or the gloo groups of the CPU tests)."""
    import ctypes as C

    from . import _lib
    if world_size() == 1 or td.get_backend() != "nccl":
        return None
    if not _native["tried"]:
        _native["tried"] = True
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        if td.get_rank() == 0:
            _lib.check(lib.ds_nccl_unique_id(buf))
        ids = [buf.raw]
        td.broadcast_object_list(ids, src=0)
        comm = C.c_void_p()
        _lib.check(lib.ds_nccl_comm_init(C.byref(comm), world_size(), ids[0], td.get_rank(), int(device_index)))
        _native["comm"] = comm
    return _native["comm"]


def destroy_native_comm():
    from . import _lib
    if _native["comm"] is not None:
        _lib.load().ds_nccl_comm_destroy(_native["comm"])
    _native["comm"], _native["tried"] = None, False

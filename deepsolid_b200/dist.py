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

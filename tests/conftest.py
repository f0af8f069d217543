import os, sys
import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


from deepsolid_b200 import cell as C            # noqa: E402
from oracle import deepsolid_oracle as O        # noqa: E402

_cache = {}


def system(name):
    """(cell, klist, numpy params, torch params) with the benchmark's seeds."""
    if name not in _cache:
        sc = C.build_system(name)
        kl = C.make_klist(sc)
        pn = O.init_params(np.random.default_rng(888), sc.original_cell.natm, sc.nelec)
        _cache[name] = (sc, kl, pn, O.params_to_torch(pn))
    return _cache[name]


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz"))


def angle_diff(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return torch.angle(torch.exp(1j * (a - b))).abs()

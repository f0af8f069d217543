"""GPU check of the tcgen05 int8-slice GEMM probe against torch fp64 matmul (debug tool)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from deepsolid_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda", 0)
cases = [(64, 128, 64), (64, 128, 320), (128, 256, 320), (100, 200, 256), (1000, 256, 320), (4536, 432, 256),
         (148 * 64 * 8, 256, 320)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
torch.manual_seed(0)
for (m, n, k) in cases:
    a = torch.randn(m, k, dtype=torch.float64, device=dev) * torch.exp(2 * torch.randn(m, 1, dtype=torch.float64, device=dev))
    b = torch.randn(k, n, dtype=torch.float64, device=dev) / k ** 0.5
    c = torch.full((m, n), float("nan"), dtype=torch.float64, device=dev)
    gm, sm = C.c_double(), C.c_double()
    reps = 5 if m * n * k > 1e9 else 1
    rc = lib.ds_ozaki_dgemm_probe(0, a.data_ptr(), b.data_ptr(), c.data_ptr(), m, n, k, reps, C.byref(gm), C.byref(sm), None)
    if rc:
        print(f"{m}x{n}x{k}: rc={rc} {lib.ds_last_error().decode()}", flush=True)
        break
    torch.cuda.synchronize()
    ref = a @ b
    scale = (a.abs().amax(1, keepdim=True) * b.abs().amax(0, keepdim=True)) * k
    err = ((c - ref).abs() / scale).max().item()
    nan = int(torch.isnan(c).sum())
    tf = 2.0 * m * n * k / (gm.value * 1e-3) / 1e12
    print(f"{m}x{n}x{k}: max |dC|/(rowmax*colmax*K) = {err:.3e}  nan={nan}  gemm {gm.value:.3f} ms ({tf:.1f} TF-equivalent)  "
          f"slice {sm.value:.3f} ms", flush=True)
    if nan or not err < 1e-9:
        bad = ((c - ref).abs() / scale)
        idx = torch.nonzero(~(bad < 1e-9))[:8]
        print("  first bad entries (row, col):", idx.tolist(), flush=True)
        print("  c  :", c[:2, :6].tolist(), flush=True)
        print("  ref:", ref[:2, :6].tolist(), flush=True)

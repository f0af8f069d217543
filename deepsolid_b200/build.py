"""In-tree build of libdeepsolid_b200.so (hand-written CUDA for sm_100a, plain C ABI).

``python -m deepsolid_b200.build`` or ``build()``: every csrc/*.cu is compiled with
``nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo`` (cross-compiles without a
GPU) and linked into ``deepsolid_b200/libdeepsolid_b200.so``.  Objects are cached by
mtime under ``deepsolid_b200/_build``.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OBJ = HERE / "_build"
LIB = HERE / "libdeepsolid_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]
FLAGS += os.environ.get("DS_EXTRA_NVCC_FLAGS", "").split()      # e.g. -DDS_OZ_PROF for the role clocks of oz_gemm_kernel


def _deps():
    return [p.stat().st_mtime for p in list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "deepsolid_b200.h"]]


def _compile(src: Path, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    newest = max([src.stat().st_mtime] + _deps())
    if obj.exists() and obj.stat().st_mtime >= newest:
        return obj
    cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        sys.stderr.write(r.stderr)
    return obj


def build(verbose: bool = False, force: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    if force:
        for o in OBJ.glob("*.o"):
            o.unlink()
    try:        # the GPU box gets a snapshot without .git: leave the revision where bench.py can read it
        sha = subprocess.run(["git", "-C", str(HERE.parent), "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True).stdout.strip()
        dirty = subprocess.run(["git", "-C", str(HERE.parent), "status", "--porcelain", "--untracked-files=no"], capture_output=True, text=True).stdout.strip()
        if sha:
            (OBJ / "git_sha.txt").write_text(sha + ("+dirty" if dirty else "") + "\n")
    except Exception:
        pass
    srcs = sorted(CSRC.glob("*.cu"))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    if (not LIB.exists()) or LIB.stat().st_mtime < max(o.stat().st_mtime for o in objs):
        cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))

#!/bin/bash
# double-buffered TMEM oz kernel (32-row tiles): probe correctness + timing vs the 64-row kernel, then suite + bench
set -x
mkdir -p gpurun_out
timeout 120 python scripts/oz_check.py 2>&1 | tail -12
DS_OZ_TN=64 timeout 120 python scripts/oz_check.py 771120x256x320 75776x256x320 2>&1 | tail -3
timeout 120 python scripts/oz_check.py 771120x256x320 75776x256x320 2>&1 | tail -3
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-400
DS_OZ_TN=64 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-200

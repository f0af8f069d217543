#!/bin/bash
# progressive diagonal completion in the oz kernel: probe (correctness + time, A/B), suite, bench A/B
set -x
mkdir -p gpurun_out
timeout 120 python scripts/oz_check.py 2>&1 | tail -8
timeout 120 python scripts/oz_check.py 771120x256x320 385560x432x256 2>&1 | tail -2
DS_OZ_PROG=0 timeout 120 python scripts/oz_check.py 771120x256x320 385560x432x256 2>&1 | tail -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-170
DS_OZ_PROG=0 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-170

#!/bin/bash
# value-only features kernel: parity, forward profile, mcmc bench
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/fwd_profile.py 4096 2>&1 | tail -25
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --mcmc 2>&1 | tail -3

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size or golden" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/lih108_launches.csv \
  python bench.py --system lih108 --batch 128 --steps 1 --warmup 1 --equil 0 --no-cpu-baseline --no-e2e > gpurun_out/lih108_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/lih108_launches.csv 2>/dev/null | head -6
timeout 900 python bench.py --system lih108 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-200

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or live or full_size" 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"l0_jac2|slice_means|features_pair" -c 60 --csv --log-file gpurun_out/l0_launches.csv \
  python bench.py --batch 1024 --steps 1 --warmup 1 --equil 0 --no-cpu-baseline --no-e2e > gpurun_out/l0_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/l0_launches.csv 2>/dev/null | head -5
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-160

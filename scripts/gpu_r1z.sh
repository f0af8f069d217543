#!/bin/bash
# where does LiH-108 (config 5) spend its step?
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/lih108_launches.csv \
  python bench.py --system lih108 --batch 128 --steps 1 --warmup 1 --equil 0 --no-cpu-baseline --no-e2e > gpurun_out/lih108_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/lih108_launches.csv 2>/dev/null | head -16
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/diamond64_launches.csv \
  python bench.py --system diamond64 --batch 256 --steps 1 --warmup 1 --equil 0 --no-cpu-baseline --no-e2e > gpurun_out/diamond64_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/diamond64_launches.csv 2>/dev/null | head -12

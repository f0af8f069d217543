#!/bin/bash
mkdir -p gpurun_out
T=r1i
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
DS_L0_ONE=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_ab.json 2>> gpurun_out/${T}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --batch 1024 --steps 1 --warmup 3 --equil 0 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_launch.log 2>&1
python scripts/launch_summary.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launch_summary.txt; head -22 gpurun_out/${T}_launch_summary.txt

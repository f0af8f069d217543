#!/bin/bash
set -x
mkdir -p gpurun_out
T=r1f
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
tail -12 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
cat gpurun_out/${T}_bench.json
DS_DET_V2=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${T}_bench_detv2.json 2>> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_bench_detv2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --batch 1024 --steps 1 --warmup 3 --equil 0 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_launch.log 2>&1
python scripts/launch_summary.py gpurun_out/${T}_launches.csv > gpurun_out/${T}_launch_summary.txt; head -18 gpurun_out/${T}_launch_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'det_dmma_kernel' -s 4 -c 2 \
  -o gpurun_out/${T}_prof python bench.py --batch 256 --steps 1 --warmup 3 --equil 0 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_full.log 2>&1

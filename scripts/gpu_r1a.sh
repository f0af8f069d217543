#!/bin/bash
# round-1 GPU pass: parity tests, bench line, launch list, ncu --set full of the hot kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1a_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1a_tests.log
tail -3 gpurun_out/r1a_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r1a_bench.json 2> gpurun_out/r1a_bench.err; echo "bench rc=$?"
cat gpurun_out/r1a_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r1a_launches.csv \
  python bench.py --batch 1024 --steps 1 --warmup 3 --equil 0 --no-cpu-baseline --no-e2e > gpurun_out/r1a_ncu_launch.log 2>&1
python scripts/launch_summary.py gpurun_out/r1a_launches.csv > gpurun_out/r1a_launch_summary.txt; cat gpurun_out/r1a_launch_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'oz_gemm_kernel|det_kernel|slice_rows_kernel|features_pair_kernel' -s 12 -c 10 \
  -o gpurun_out/r1a_prof python bench.py --batch 256 --steps 1 --warmup 3 --equil 0 --no-cpu-baseline --no-e2e > gpurun_out/r1a_ncu_full.log 2>&1
ls -la gpurun_out

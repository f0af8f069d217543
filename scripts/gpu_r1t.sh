#!/bin/bash
# final round-1 measurement pass: tests, bench (+cpu baseline), reference arm, launch list with DRAM bytes, ncu --set full of the hot kernels
mkdir -p gpurun_out
T=r1t
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --batch 1024 --steps 1 --warmup 3 --equil 0 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'oz_gemm_kernel|det_dmma_kernel|slice_means_kernel|slice_rows_kernel|features_pair_kernel|l0_jac2_kernel' -s 12 -c 14 \
  -o gpurun_out/${T}_prof python bench.py --batch 256 --steps 1 --warmup 3 --equil 0 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
ls -la gpurun_out | tail -6
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2>> gpurun_out/${T}_bench.err; tail -c 600 gpurun_out/${T}_bench_reference.json
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --mcmc --grad --kfac 2>/dev/null | tail -1 > gpurun_out/${T}_bench_extras.json

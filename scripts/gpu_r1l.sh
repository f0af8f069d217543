#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1l_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r1l_tests.log
tail -6 gpurun_out/r1l_tests.log
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --mcmc --grad > gpurun_out/r1l_bench.json 2> gpurun_out/r1l_bench.err; echo "bench rc=$?"
tail -c 700 gpurun_out/r1l_bench.json; tail -3 gpurun_out/r1l_bench.err

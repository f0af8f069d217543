#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python scripts/fwd_profile.py 4096 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fwd_launches.csv python scripts/fwd_profile.py 4096 > gpurun_out/fwd_l.log 2>&1
python scripts/launch_summary.py gpurun_out/fwd_launches.csv 2>&1 | head -14
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --mcmc --grad --kfac 2>&1 | tail -1 > gpurun_out/r1r_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r1r_bench.json')); print(d['value'], d['extras'])"

#!/bin/bash
# forward (value-only) launch list + full ncu capture of the lane-per-pair features kernel
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fwd_launches.csv python scripts/fwd_profile.py 4096 > gpurun_out/fwd_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:features_value -s 2 -c 1 -o gpurun_out/fv python scripts/fwd_profile.py 4096 > gpurun_out/fv.log 2>&1
python scripts/launch_summary.py gpurun_out/fwd_launches.csv 2>&1 | head -30

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:det_dmma -s 1 -c 1 -o gpurun_out/det14 \
  python bench.py --system lih108 --batch 64 --steps 1 --warmup 1 --equil 0 --no-cpu-baseline --no-e2e > gpurun_out/det14.log 2>&1
ls -la gpurun_out/det14.ncu-rep

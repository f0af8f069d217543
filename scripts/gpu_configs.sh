#!/bin/bash
# the other BASELINE.json configurations (parity-tested elsewhere; one throughput line each for the record)
mkdir -p gpurun_out
for s in li24 diamond64 lih108 h10; do
  timeout 900 python bench.py --system $s --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1_cfg_$s.json 2> gpurun_out/r1_cfg_$s.err
  tail -c 600 gpurun_out/r1_cfg_$s.json | head -c 300; echo
done

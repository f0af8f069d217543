#!/bin/bash
# component limits of the 32-row double-buffered kernel (same flags as gpu_probe.sh)
mkdir -p gpurun_out
for dbg in 0 1 2 3 4 5 6; do
  echo "== DS_OZ_DBG=$dbg"
  DS_OZ_DBG=$dbg timeout 120 python scripts/oz_check.py 771120x256x320 2>&1 | grep -v "first bad\|  c  :\|  ref:" | cut -c1-160
done > gpurun_out/r1_probe_tn32.log 2>&1
cat gpurun_out/r1_probe_tn32.log

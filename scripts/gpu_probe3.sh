#!/bin/bash
mkdir -p gpurun_out
for dbg in 0 1 4 5; do
  echo "== PROG=1 DS_OZ_DBG=$dbg"
  DS_OZ_DBG=$dbg timeout 120 python scripts/oz_check.py 771120x256x320 2>&1 | grep -v "first bad\|  c  :\|  ref:" | cut -c1-160
done > gpurun_out/r1_probe_prog.log 2>&1
cat gpurun_out/r1_probe_prog.log

#!/bin/bash
# component limits of the tcgen05 int8-slice GEMM: DS_OZ_DBG bit 1 = no epilogue TMEM reads/stores, 2 = no MMA issue, 4 = no TMA
mkdir -p gpurun_out
for dbg in 0 1 2 3 4 5; do
  echo "== DS_OZ_DBG=$dbg"
  DS_OZ_DBG=$dbg timeout 120 python scripts/oz_check.py 771120x256x320 385560x432x256 2>&1 | grep -v "first bad\|  c  :\|  ref:" | cut -c1-200
done > gpurun_out/r1_probe.log 2>&1
cat gpurun_out/r1_probe.log

#!/bin/bash
# ncu --set full of the two value-only kernels of the Metropolis forward (final code)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'features_value_kernel|det_warp_kernel' -s 4 -c 2 -o gpurun_out/fwd_final \
  python scripts/fwd_profile.py 4096 > gpurun_out/fwd_final.log 2>&1
ls -la gpurun_out/fwd_final.ncu-rep

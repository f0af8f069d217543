#!/bin/bash
# a longer VMC optimisation on the H10 chain (config 1): 300 KFAC iterations, 1024 walkers
mkdir -p gpurun_out
timeout 900 python examples/train_vmc.py --system h10 --batch 1024 --iterations 300 --burn-in 50 > gpurun_out/r1_train_h10.log 2>&1
grep -E "^Step 000(00|01|50)|^Step 00(1|2)[05]0|^Step 0029|^energy" gpurun_out/r1_train_h10.log | cut -c1-150

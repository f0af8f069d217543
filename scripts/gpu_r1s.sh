#!/bin/bash
# 2-GPU checks: walker-sharded bench with extras, and the example VMC loop with the KFAC statistics averaged over ranks
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --mcmc --grad --kfac 2>&1 | tail -1 > gpurun_out/r1s_bench_2gpu.json
python -c "
import json; d=json.load(open('gpurun_out/r1s_bench_2gpu.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['extras'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 examples/train_vmc.py --system h4 --batch 256 --iterations 30 --burn-in 10 2>&1 | tail -6
DS_WS_GIB=48 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-160

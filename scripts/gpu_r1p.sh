#!/bin/bash
set -x
mkdir -p gpurun_out
python examples/train_vmc.py --system h4 --batch 256 --iterations 40 --burn-in 10 2>&1 | tail -45
python -m pytest tests/test_train_loop.py tests/test_observables.py -m gpu -x -q 2>&1 | tail -8

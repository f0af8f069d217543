#!/bin/bash
set -x
mkdir -p gpurun_out
T=r1e
timeout 900 python -m pytest tests/test_gradient.py -m gpu -x -q > gpurun_out/${T}_grad_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_grad_tests.log
tail -30 gpurun_out/${T}_grad_tests.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
tail -5 gpurun_out/${T}_tests.log

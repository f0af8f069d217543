#!/bin/bash
# KFAC factor statistics: parity tests, then the whole GPU suite
set -x
mkdir -p gpurun_out
python -m pytest tests/test_kfac.py -m gpu -x -q 2>&1 | tail -15
python -m pytest tests -m gpu -x -q 2>&1 | tail -4

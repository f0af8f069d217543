#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_variants.py -m gpu -q > gpurun_out/r1k_new.log 2>&1; echo "rc=$?" >> gpurun_out/r1k_new.log
tail -40 gpurun_out/r1k_new.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r1k_tests.log
tail -4 gpurun_out/r1k_tests.log

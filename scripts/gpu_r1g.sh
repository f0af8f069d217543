#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_variants.py -m gpu -x -q > gpurun_out/r1g_variants.log 2>&1; echo "rc=$?" >> gpurun_out/r1g_variants.log
tail -30 gpurun_out/r1g_variants.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r1g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r1g_tests.log
tail -4 gpurun_out/r1g_tests.log

#!/bin/bash
# memcheck of the hot path on small systems (every kernel family: sweep, gradient, moves, variants)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py tests/test_gradient.py tests/test_moves.py tests/test_variants.py -m gpu -q -x \
  -k "h4 or lih_prim or chunk or value_and_grad or importance" > gpurun_out/r1_memcheck.log 2>&1
echo "rc=$?" >> gpurun_out/r1_memcheck.log
grep -E "ERROR SUMMARY|Invalid|passed|failed|rc=" gpurun_out/r1_memcheck.log | head -20

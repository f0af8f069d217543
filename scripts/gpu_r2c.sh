#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_variants.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --system lih108 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-160
timeout 900 python bench.py --system diamond64 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-160
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-160

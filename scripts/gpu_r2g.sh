#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or live or full_size" 2>&1 | tail -2
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-160
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-160

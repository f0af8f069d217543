#!/bin/bash
# in-call A/B on LiH-108 (54 x 54 determinants): committed HEAD (ab_old/) against the working tree
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "full_size" 2>&1 | tail -1
for i in 1 2; do
  (cd ab_old && timeout 600 python bench.py --system lih108 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-120 | sed 's/^/OLD /')
  timeout 600 python bench.py --system lih108 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-120 | sed 's/^/NEW /'
done

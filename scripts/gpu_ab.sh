#!/bin/bash
# in-call A/B: committed HEAD (ab_old/) against the working tree, alternating runs on the same box
for i in 1 2; do
  (cd ab_old && timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-120 | sed 's/^/OLD /')
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-120 | sed 's/^/NEW /'
done

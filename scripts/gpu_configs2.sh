#!/bin/bash
# the other BASELINE.json configurations with the final round-1 code (one throughput line each for the record)
mkdir -p gpurun_out
rm -f gpurun_out/r1_final2_other_configs.jsonl
for s in h10 li24 li48 diamond64 lih108; do
  timeout 900 python bench.py --system $s --steps 2 --warmup 3 --no-cpu-baseline 2> gpurun_out/r1_cfg_$s.err | tail -1 >> gpurun_out/r1_final2_other_configs.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/r1_final2_other_configs.jsonl'):
    d=json.loads(l); print(d['config']['system'], d['config']['batch_per_gpu'], round(d['value'],1), round(d['ms_per_step'],1))
PY

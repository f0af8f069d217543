#!/bin/bash
# final verification of the round: full GPU suite, smoke, default bench (+ cpu baseline, e2e), extras
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --mcmc --grad --kfac 2>/dev/null | tail -1 > gpurun_out/final_bench_extras.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline'].get('max_abs_diff_vs_gpu_Ha'))
e=json.loads(open('gpurun_out/final_bench_extras.json').read().strip().splitlines()[-1]); print(e['value'], e['extras'])
PY

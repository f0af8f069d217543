#!/bin/bash
# One runner for every GPU session:  gpurun --timeout N -- 'bash scripts/gpu.sh <task> [<task> ...]'
# Each task writes its artefacts under gpurun_out/ with the tag $TAG (default r2).  Tasks:
#   tests        full `pytest -m gpu` suite                      newtests   the BASELINE-size parity files only (-s: prints the dE_L distributions)
#   smoke        __graft_entry__.smoke()                         bench      default bench line (+ cpu baseline, e2e)
#   quick        short bench without cpu baseline / e2e          extras     bench with --mcmc --grad --kfac
#   reference    bench.py --impl reference                       configs    one bench line per other BASELINE configuration
#   launches     ncu launch list (time + DRAM bytes) of one pass profile    ncu --set full of the hot kernels
#   probe        component limits of oz_gemm_kernel (DS_OZ_DBG)  ab         committed HEAD (ab_old/, see make_ab_old.sh) vs working tree
#   diag5        accuracy + speed of the 5-diagonal experiment   sanitize   compute-sanitizer memcheck of small systems
#   strong       global batch 4096 split over the visible GPUs (run under gpurun --gpus N)
#   adopt        copy this call's traffic / HBM JSONs into profiles/ before `bench` (run after launches + profile)
#   opts         in-call A/B of environment knobs ($OPTS, ';'-separated)     partests   test_gpu_parity + test_variants only
TAG=${TAG:-r2}
O=gpurun_out
mkdir -p $O
B="python bench.py"
for task in "$@"; do
  echo "=== task $task"
  case $task in
    tests)     timeout 1700 python -m pytest tests -m gpu -q -x > $O/${TAG}_tests.log 2>&1; echo "rc=$?" >> $O/${TAG}_tests.log; tail -4 $O/${TAG}_tests.log ;;
    newtests)  timeout 1500 python -m pytest tests/test_baseline_parity.py tests/test_lattices.py tests/test_reference_golden.py -m gpu -q -s > $O/${TAG}_newtests.log 2>&1
               echo "rc=$?" >> $O/${TAG}_newtests.log; grep -E "^\[|passed|failed|rc=|Error|assert" $O/${TAG}_newtests.log | head -40 ;;
    smoke)     python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
    bench)     timeout 900 $B > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; tail -c 1500 $O/${TAG}_bench.json ;;
    quick)     timeout 300 $B --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | tee $O/${TAG}_quick.json | cut -c1-200 ;;
    extras)    timeout 600 $B --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --mcmc --grad --kfac 2>/dev/null | tail -1 > $O/${TAG}_bench_extras.json; tail -c 700 $O/${TAG}_bench_extras.json ;;
    reference) timeout 600 $B --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err; tail -c 600 $O/${TAG}_bench_reference.json ;;
    configs)   rm -f $O/${TAG}_other_configs.jsonl
               for s in h10 li24 li48 diamond64 lih108; do
                 timeout 900 $B --system $s --steps 2 --warmup 3 2> $O/${TAG}_cfg_$s.err | tail -1 >> $O/${TAG}_other_configs.jsonl
               done
               python -c "
import json
for l in open('$O/${TAG}_other_configs.jsonl'):
    d=json.loads(l); print(d['config']['system'], d['config']['batch_per_gpu'], round(d['value'],1), round(d['ms_per_step'],1), (d.get('cpu_baseline') or {}).get('max_abs_diff_vs_gpu_Ha'))" ;;
    launches)  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $O/${TAG}_launches.csv \
                 $B --batch 1024 --steps 1 --warmup 3 --equil 0 --no-cpu-baseline --no-e2e --no-probes > $O/${TAG}_ncu_launch.log 2>&1
               python scripts/launch_summary.py $O/${TAG}_launches.csv > $O/${TAG}_launch_summary.txt; head -14 $O/${TAG}_launch_summary.txt
               read CH SHA <<< $(python -c "
import json
d=[json.loads(l) for l in open('$O/${TAG}_ncu_launch.log') if l.startswith('{')][-1]['config']; print(d['chunk_walkers'], d['git_sha'])")
               python scripts/hbm_from_launches.py $O/${TAG}_launches.csv 4096 $O/${TAG}_hbm.json $CH $SHA ;;
    profile)   timeout 900 ncu --set full --clock-control none --import-source on \
                 -k regex:'oz_gemm_kernel|det_dmma_kernel|slice_means_kernel|slice_rows_kernel|features_pair_kernel|l0_jac2_kernel|means_digits_kernel' -s 14 -c 16 \
                 -o $O/${TAG}_prof $B --batch 256 --steps 1 --warmup 3 --equil 0 --no-cpu-baseline --no-e2e --no-probes > $O/${TAG}_ncu_full.log 2>&1; ls -la $O/${TAG}_prof*
               read CH SHA <<< $(python -c "
import json
d=[json.loads(l) for l in open('$O/${TAG}_ncu_full.log') if l.startswith('{')][-1]['config']; print(d['chunk_walkers'], d['git_sha'])")
               python scripts/traffic_from_ncu.py $O/${TAG}_prof.ncu-rep $O/${TAG}_traffic.json $CH $SHA
               python scripts/ncu_summary.py $O/${TAG}_prof.ncu-rep > $O/${TAG}_ncu_full_summary.txt ;;
    probe)     for dbg in 0 1 2 3 4 5; do
                 echo "== DS_OZ_DBG=$dbg"
                 DS_OZ_DBG=$dbg timeout 120 python scripts/oz_check.py 771120x256x320 385560x432x256 2>&1 | grep -v "first bad\|  c  :\|  ref:" | cut -c1-200
               done > $O/${TAG}_probe.log 2>&1; cat $O/${TAG}_probe.log ;;
    ab)        for i in 1 2; do
                 (cd ab_old && timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-120 | sed 's/^/OLD /')
                 timeout 300 $B --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-120 | sed 's/^/NEW /'
               done ;;
    diag5)     DS_OZ_DIAGS=5 timeout 900 python -m pytest tests/test_baseline_parity.py -m gpu -q -s -k "int8" 2>&1 | grep -E "^\[|passed|failed" | sed 's/^/DIAG5 /'
               DS_OZ_DIAGS=5 timeout 300 $B --steps 4 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | cut -c1-160 | sed 's/^/DIAG5 /' ;;
    sanitize)  timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
                 python -m pytest tests/test_gpu_parity.py tests/test_gradient.py tests/test_moves.py tests/test_variants.py -m gpu -q -x \
                 -k "h4 or lih_prim or chunk or value_and_grad or importance" > $O/${TAG}_memcheck.log 2>&1
               echo "rc=$?" >> $O/${TAG}_memcheck.log; grep -E "ERROR SUMMARY|Invalid|passed|failed|rc=" $O/${TAG}_memcheck.log | head -20 ;;
    strong)    N=$(nvidia-smi -L | wc -l)
               timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
                 bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --no-cpu-baseline 2> $O/${TAG}_strong_$N.err | tail -1 | tee $O/${TAG}_strong_$N.json | cut -c1-300 ;;
    opts)      # in-call A/B of environment knobs: OPTS="DS_OZ_OPT=0;DS_OZ_OPT=8;..." (each entry may hold several VAR=VALUE words)
               IFS=';' read -ra LIST <<< "$OPTS"
               for rep in 1 2; do for o in "${LIST[@]}"; do
                 echo -n "[$o] "; env $o timeout 300 $B --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --no-probes ${BARGS:-} 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'le/s', round(d['ms_per_step'],1), 'ms  gemm', round(d['roofline']['kernel_ms_per_step'],1), 'ms frac', round(d['roofline']['frac'],3), 'sm', d['clocks']['sm_mhz'], 'MHz', d['clocks'].get('power_w_max'), 'W chunk', d['config'].get('chunk_walkers'))"
               done; done | tee -a $O/${TAG}_opts.log ;;
    adopt)     # make the bench line of THIS call quote the ncu numbers of THIS call: copy the fresh traffic / HBM JSONs
               # over the committed ones (the same files are committed from gpurun_out/ afterwards)
               for f in hbm traffic; do [ -f $O/${TAG}_$f.json ] && cp $O/${TAG}_$f.json profiles/${TAG}_$f.json; done; ls -la profiles/${TAG}_hbm.json profiles/${TAG}_traffic.json ;;
    partests)  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_variants.py -m gpu -q -x 2>&1 | tail -3 ;;
    *)         echo "unknown task $task" ;;
  esac
done

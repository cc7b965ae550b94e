#!/bin/bash
# A/B of environment switches on the bench frame: gpu_env_ab.sh <tag> "" "VAR=1" "VAR2=1 VAR3=2" ...   (each argument = one env set)
TAG=$1; shift
mkdir -p gpurun_out
LOG=gpurun_out/r2_envab_$TAG.log
: > $LOG
for rep in 1 2; do
for E in "$@"; do
  echo "== [$E] (pass $rep)" >> $LOG
  env $E timeout 600 python bench.py --steps 60 --warmup 20 --no-cpu-baseline --no-4k 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); c = d['config']
        print('fps %.2f ms %.3f e2e %.2f (blocking %.2f)  Mrays/s %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('blocking_readback_value', 0), c['mrays_per_s_per_gpu']))
        print('  kern', {k: round(v['ms_per_frame'], 3) for k, v in c['kernels'].items()})
        print('  pass', {k: round(v, 3) for k, v in c['pass_ms'].items()})
" >> $LOG
done
done
cat $LOG

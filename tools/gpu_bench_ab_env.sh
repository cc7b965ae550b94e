#!/bin/bash
# gpu_bench_ab.sh with an environment set applied to every build: gpu_bench_ab_env.sh <tag> "VAR=1 VAR2=2" default lib_x ...
TAG=$1; shift; ENVSET=$1; shift
mkdir -p gpurun_out
LOG=gpurun_out/r2_benchab_$TAG.log
: > $LOG
for rep in 1 2; do
for L in "$@"; do
  D=""; [ "$L" != "default" ] && D="$PWD/$L"
  echo "== $L [$ENVSET] (pass $rep)" >> $LOG
  env $ENVSET RPT_LIB_DIR=$D timeout 600 python bench.py --steps 60 --warmup 20 --no-cpu-baseline --no-4k 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); c = d['config']
        print('fps %.2f ms %.3f e2e %.2f  Mrays/s %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], c['mrays_per_s_per_gpu']))
        print('  kern', {k: round(v['ms_per_frame'], 3) for k, v in c['kernels'].items()})
        print('  pass', {k: round(v, 3) for k, v in c['pass_ms'].items()})
" >> $LOG
done
done
cat $LOG

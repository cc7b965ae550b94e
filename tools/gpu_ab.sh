#!/bin/bash
# Round-2 A/B of experiment builds (development aid).  For each LIBDIR ("default" = the product build): traversal / GRIS
# parity subset, the traversal microbenchmark, a short bench.py.  Usage: gpu_ab.sh <tag> default lib_x ...
TAG=$1; shift
mkdir -p gpurun_out
LOG=gpurun_out/r2_ab_$TAG.log
: > $LOG
for L in "$@"; do
  D=""; [ "$L" != "default" ] && D="$PWD/$L"
  echo "== $L" >> $LOG
  env RPT_LIB_DIR=$D timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -x -q -k "closest or queue or shadow or degenerate or gris or golden or traversal" 2>&1 | tail -3 >> $LOG
  env RPT_LIB_DIR=$D timeout 300 python tools/gpu_tracebench.py 2>&1 | tail -4 >> $LOG
  env RPT_LIB_DIR=$D timeout 600 python bench.py --steps 60 --warmup 20 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); c = d['config']
        print('fps %.2f ms %.3f e2e %.2f  Mrays/s %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], c['mrays_per_s_per_gpu']))
        print('  kern', {k: round(v['ms_per_frame'], 3) for k, v in c['kernels'].items()})
    else:
        print(ln, end='')
" >> $LOG
done
cat $LOG

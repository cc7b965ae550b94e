#!/bin/bash
# A/B of experiment builds: for each LIBDIR ("default" = the product build) the traversal / queue parity tests, the traversal
# microbenchmark and a short bench.py (development aid): gpu_ab6.sh default lib_coop ...
mkdir -p gpurun_out
: > gpurun_out/ab6.log
for L in "$@"; do
  D=""; [ "$L" != "default" ] && D="$PWD/$L"
  echo "== $L" >> gpurun_out/ab6.log
  env RPT_LIB_DIR=$D timeout 600 python -m pytest tests -m gpu -x -q -k "closest or queue or shadow or degenerate or gris or fullsize or golden" 2>&1 | tail -3 >> gpurun_out/ab6.log
  [ -n "$TRACEBENCH" ] && env RPT_LIB_DIR=$D timeout 300 python tools/gpu_tracebench.py 2>&1 | tail -4 >> gpurun_out/ab6.log
  env RPT_LIB_DIR=$D timeout 600 python bench.py --steps 60 --warmup 20 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); c = d['config']
        print('fps %.2f ms %.3f e2e %.2f  Mrays/s %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], c['mrays_per_s_per_gpu']))
        print('  kern', {k: round(v['ms_per_frame'], 3) for k, v in c['kernels'].items()})
    else:
        print(ln, end='')
" >> gpurun_out/ab6.log
done
cat gpurun_out/ab6.log

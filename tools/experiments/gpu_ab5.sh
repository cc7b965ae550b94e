#!/bin/bash
# parity tests, then bench.py under "LIBDIR|VAR=value" settings (development aid): gpu_ab5.sh "default|X=0" "lib_foo|RPT_Y=1" ...
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
: > gpurun_out/ab5.log
for spec in "$@"; do
  L=${spec%%|*}; V=${spec#*|}
  echo "== $L $V" >> gpurun_out/ab5.log
  D=""; [ "$L" != "default" ] && D="$PWD/$L"
  env RPT_LIB_DIR=$D $V timeout 600 python bench.py --steps 60 --warmup 20 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); c = d['config']
        print('fps %.2f ms %.3f e2e %.2f  Mrays/s %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value'], c['mrays_per_s_per_gpu']))
        print('  kern', {k: round(v['ms_per_frame'], 3) for k, v in c['kernels'].items()})
    else:
        print(ln, end='')
" >> gpurun_out/ab5.log
done
cat gpurun_out/ab5.log

#!/bin/bash
# A/B of library builds: bench.py frame (Ajar 1080p) + the 51 M-triangle field (development aid)
mkdir -p gpurun_out
: > gpurun_out/ab4.log
for L in "$@"; do
  echo "== $L" >> gpurun_out/ab4.log
  D=""; [ "$L" != "default" ] && D="$PWD/$L"
  RPT_LIB_DIR=$D timeout 600 python bench.py --steps 60 --warmup 20 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); c = d['config']
        print('  ajar fps %.2f ms %.3f' % (d['value'], d['ms_per_step']), {k: round(v['ms_per_frame'], 3) for k, v in c['kernels'].items() if k.startswith('trace')})
" >> gpurun_out/ab4.log
  RPT_LIB_DIR=$D timeout 600 python tools/gpu_configs.py field 5 28 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('  field %-5s %.3f ms/frame  %.0f Mrays/s' % (d['method'], d['ms_per_frame'], d['mrays_per_s']))
" >> gpurun_out/ab4.log
done
cat gpurun_out/ab4.log

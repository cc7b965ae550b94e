#!/bin/bash
# A/B of library builds on the traversal microbenchmark (Ajar), the Ajar frame and the 51 M-triangle field (development aid)
mkdir -p gpurun_out
: > gpurun_out/ab3.log
for L in "$@"; do
  echo "== $L" >> gpurun_out/ab3.log
  D=""; [ "$L" != "default" ] && D="$PWD/$L"
  RPT_LIB_DIR=$D timeout 300 python tools/gpu_tracebench.py >> gpurun_out/ab3.log 2>&1
  RPT_LIB_DIR=$D timeout 300 python tools/gpu_quick.py 1920 1080 ajar 30 2>&1 | grep -E "GRIS:" >> gpurun_out/ab3.log
  RPT_LIB_DIR=$D timeout 600 python tools/gpu_configs.py field 5 28 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('  field %-5s %.3f ms/frame  %.0f Mrays/s  build %.0f ms' % (d['method'], d['ms_per_frame'], d['mrays_per_s'], d['bvh_build_ms']))
" >> gpurun_out/ab3.log
done
cat gpurun_out/ab3.log

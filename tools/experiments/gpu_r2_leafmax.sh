#!/bin/bash
# leaf size of the wide BVH (build-time switch RPT_LEAF_MAX, default 3) on the traversal microbenchmark and the frame
mkdir -p gpurun_out
: > gpurun_out/r2_leafmax.log
for L in 1 2 3; do
  echo "== RPT_LEAF_MAX=$L" >> gpurun_out/r2_leafmax.log
  env RPT_LEAF_MAX=$L timeout 300 python tools/gpu_tracebench.py 2>&1 | tail -4 >> gpurun_out/r2_leafmax.log
  env RPT_LEAF_MAX=$L timeout 300 python bench.py --steps 40 --warmup 10 --no-cpu-baseline --no-4k 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); c = d['config']
        print('fps %.2f ms %.3f' % (d['value'], d['ms_per_step']), {k: round(v['ms_per_frame'], 3) for k, v in c['kernels'].items()})
" >> gpurun_out/r2_leafmax.log
done
cat gpurun_out/r2_leafmax.log

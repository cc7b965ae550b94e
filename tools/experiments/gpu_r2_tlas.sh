#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --durations=6 2>&1 | tail -25 > gpurun_out/r2_tlas_tests.log
cat gpurun_out/r2_tlas_tests.log
timeout 600 python bench.py --steps 60 --warmup 20 --no-cpu-baseline --no-4k 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); c = d['config']
        print('fps %.2f ms %.3f e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k: round(v['ms_per_frame'], 3) for k, v in c['kernels'].items()})
    else:
        print(ln, end='')
"

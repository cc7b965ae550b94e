#!/bin/bash
# bench N=1 under a list of VAR=value settings (development aid)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
: > gpurun_out/ab2.log
for V in "$@"; do
  echo "== $V" >> gpurun_out/ab2.log
  env $V timeout 600 python bench.py --steps 60 --warmup 20 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); c = d['config']
        print('fps %.2f ms %.3f e2e %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value']))
        print('  pass', {k: round(v, 3) for k, v in c['pass_ms'].items()})
        print('  kern', {k: round(v['ms_per_frame'], 3) for k, v in c['kernels'].items()}, 'wait', round(c['tail_wait_ms_per_frame'], 3))
    else:
        print(ln, end='')
" >> gpurun_out/ab2.log
done
cat gpurun_out/ab2.log

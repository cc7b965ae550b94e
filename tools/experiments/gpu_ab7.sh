#!/bin/bash
# short bench under a list of VAR=value settings, traversal figures only (development aid): gpu_ab7.sh X=0 RPT_FOO=1 ...
for v in "$@"; do
  echo "== $v"
  env $v timeout 120 python bench.py --steps 60 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); c = d['config']
        print('fps %.2f ms %.3f' % (d['value'], d['ms_per_step']), {k: round(v['ms_per_frame'], 3) for k, v in c['kernels'].items() if k.startswith('trace')}, 'gris_pathtrace %.3f' % c['pass_ms']['gris_pathtrace'])
"
done

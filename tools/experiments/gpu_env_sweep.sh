#!/bin/bash
# traversal microbenchmark + frame timing under a list of VAR=value settings (development aid)
mkdir -p gpurun_out
: > gpurun_out/env_sweep.log
for V in "$@"; do
  echo "== $V" >> gpurun_out/env_sweep.log
  env $V timeout 300 python tools/gpu_tracebench.py >> gpurun_out/env_sweep.log 2>&1
  env $V timeout 300 python tools/gpu_quick.py 1920 1080 ajar 30 2>&1 | grep -E "bvh build|GRIS:|gris_|gbuffer|rays/frame" >> gpurun_out/env_sweep.log
done
cat gpurun_out/env_sweep.log

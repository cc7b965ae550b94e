#!/bin/bash
# quick GPU iteration: parity tests + per-pass timing (development aid); extra args: VAR=value settings to A/B
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_quick.py 1920 1080 ajar 30 > gpurun_out/quick_ajar.log 2>&1
cat gpurun_out/quick_ajar.log
for V in "$@"; do
  echo "== $V"
  env $V timeout 300 python tools/gpu_quick.py 1920 1080 ajar 30 2>&1 | tee gpurun_out/quick_ajar_ab.log | grep -E "GRIS:|gris_|gbuffer"
done

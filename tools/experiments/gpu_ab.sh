#!/bin/bash
# parity tests + A/B of library builds (RPT_LIB_DIR) on the traversal microbenchmark and the frame (development aid)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
: > gpurun_out/ab.log
for L in "$@"; do
  echo "== $L" >> gpurun_out/ab.log
  D=""; [ "$L" != "default" ] && D="$PWD/$L"
  RPT_LIB_DIR=$D timeout 300 python tools/gpu_tracebench.py >> gpurun_out/ab.log 2>&1
  RPT_LIB_DIR=$D timeout 300 python tools/gpu_quick.py 1920 1080 ajar 30 2>&1 | grep -E "GRIS:|gris_|gbuffer" >> gpurun_out/ab.log
done
cat gpurun_out/ab.log
timeout 900 python tools/gpu_configs.py field 5 28 > gpurun_out/config_field.log 2>&1
cat gpurun_out/config_field.log | cut -c1-900

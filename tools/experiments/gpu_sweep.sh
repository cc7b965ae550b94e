#!/bin/bash
# parity tests, then a sweep of the traversal kernel's tunables on the microbenchmark and the frame (development aid)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
: > gpurun_out/sweep.log
for T in 0 4 8 12 16 20 24; do
  echo "== RPT_TRI_THRESHOLD=$T" >> gpurun_out/sweep.log
  RPT_TRI_THRESHOLD=$T timeout 300 python tools/gpu_tracebench.py >> gpurun_out/sweep.log 2>&1
  RPT_TRI_THRESHOLD=$T timeout 300 python tools/gpu_quick.py 1920 1080 ajar 30 2>&1 | grep -E "GRIS:|gris_|gbuffer" >> gpurun_out/sweep.log
done
for F in 2 4 12 16; do
  echo "== RPT_FETCH_THRESHOLD=$F (tri 12)" >> gpurun_out/sweep.log
  RPT_FETCH_THRESHOLD=$F timeout 300 python tools/gpu_tracebench.py >> gpurun_out/sweep.log 2>&1
  RPT_FETCH_THRESHOLD=$F timeout 300 python tools/gpu_quick.py 1920 1080 ajar 30 2>&1 | grep -E "GRIS:|gris_|gbuffer" >> gpurun_out/sweep.log
done
cat gpurun_out/sweep.log

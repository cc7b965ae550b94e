#!/bin/bash
# GPU tests + bench (both stream modes) + config 5 field in both stream modes (development aid)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 60 --warmup 20 --no-cpu-baseline > gpurun_out/bench_r3.json 2> gpurun_out/bench_r3.err
cut -c1-2500 gpurun_out/bench_r3.json; tail -3 gpurun_out/bench_r3.err
RPT_TRACE_ONE_STREAM=1 timeout 900 python tools/gpu_configs.py field 5 28 > gpurun_out/config_field_onestream.log 2>&1
timeout 900 python tools/gpu_configs.py field 5 28 > gpurun_out/config_field.log 2>&1
echo "== one stream"; cut -c1-700 gpurun_out/config_field_onestream.log; echo "== overlap"; cut -c1-700 gpurun_out/config_field.log

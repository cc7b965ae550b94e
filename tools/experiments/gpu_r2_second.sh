#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -q --durations=10 2>&1 | tail -25 > gpurun_out/r2_second.log
cat gpurun_out/r2_second.log
tools/gpu_r2_prof_trav.sh lib_base lib_compact

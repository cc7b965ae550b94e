#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_first.log; nproc >> gpurun_out/r2_first.log
timeout 900 python -m pytest tests/test_gpu_baseline_configs.py -m gpu -x -q --durations=10 2>&1 | tail -25 >> gpurun_out/r2_first.log
cat gpurun_out/r2_first.log
tools/gpu_r2_ab.sh trav1 lib_base lib_compact lib_compact8 lib_compactstk lib_compactrcp

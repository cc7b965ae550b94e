#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -22 > gpurun_out/r2_third_tests.log
cat gpurun_out/r2_third_tests.log
timeout 900 python bench.py --steps 60 --warmup 20 > gpurun_out/r2_third_bench.json 2> gpurun_out/r2_third_bench.err
tail -c 6000 gpurun_out/r2_third_bench.json; tail -5 gpurun_out/r2_third_bench.err

#!/bin/bash
# ncu launch list of a few steady-state frames (development aid): gpu_launches.sh [skip] [count]
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gris|gbuffer|postProcess|traceQueue" -s ${1:-600} -c ${2:-130} --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 6 --warmup 12 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/launches.csv | cut -c1-200

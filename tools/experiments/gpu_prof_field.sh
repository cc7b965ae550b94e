#!/bin/bash
# ncu capture of the traversal kernels on the 51 M-triangle field (config 5, HBM-resident BVH) (development aid)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:traceQueueKernel" -s 40 -c 2 -f -o gpurun_out/prof_field_tq \
   python tools/gpu_configs.py field 5 28 > gpurun_out/prof_field_tq.log 2>&1
tail -2 gpurun_out/prof_field_tq.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gi|traceQueue|gbuffer" -s 200 -c 60 --csv --log-file gpurun_out/field_launches.csv \
   python tools/gpu_configs.py field 5 28 > gpurun_out/field_launches.log 2>&1
tail -1 gpurun_out/field_launches.log | cut -c1-200

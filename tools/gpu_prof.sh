#!/bin/bash
# ncu evidence for the frame kernels (development aid): launch list + one full capture of the top kernel
mkdir -p gpurun_out
KRE='regex:gris|gBuffer|postProcess|gbuffer'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -s 40 -c 40 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 12 --warmup 10 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:grisPathTrace -s 12 -c 1 -f -o gpurun_out/prof_pathtrace \
   python bench.py --steps 4 --warmup 10 --no-cpu-baseline > gpurun_out/prof_pathtrace.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:grisSpatial -s 12 -c 1 -f -o gpurun_out/prof_spatial \
   python bench.py --steps 4 --warmup 10 --no-cpu-baseline > gpurun_out/prof_spatial.log 2>&1
ls -la gpurun_out

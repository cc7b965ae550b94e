#!/bin/bash
# ncu --set full of one bench frame's traversal launches (see tools/ncu_traffic.py) + the launch list of 4 frames
mkdir -p gpurun_out
timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k "regex:traceQueue" -s 39 -c 13 -f \
   -o gpurun_out/r2_frame_trace python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-4k > gpurun_out/r2_frame_trace.log 2>&1
tail -1 gpurun_out/r2_frame_trace.log | cut -c1-150
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(?!.*(ploc|Ploc|cub|flatten|morton|initLeaves|collapse))" -c 400 --csv --log-file gpurun_out/r2_launches.csv \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-4k > gpurun_out/r2_launches.log 2>&1
tail -3 gpurun_out/r2_launches.csv | cut -c1-200

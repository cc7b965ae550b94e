#!/bin/bash
# Round-end ncu captures of the bench frame's dominant kernels (one frame's launches each), see tools/ncu_traffic.py
mkdir -p gpurun_out
L=${1:-default}; D=""; [ "$L" != "default" ] && D="$PWD/$L"
# (RPT_INLINE_TAIL=1: the tail as one kernel, so that a frame has exactly 13 traversal and 6 bounce launches for the -s / -c windows;
#  the launches captured — bounces 1..6 and the reuse passes' visibility rays — are the same kernels on the same rays either way)
export RPT_INLINE_TAIL=1
env RPT_LIB_DIR=$D timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k "regex:traceQueue" -s 39 -c 13 -f \
   -o gpurun_out/r2_frame_trace python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-4k > gpurun_out/r2_frame_trace.log 2>&1
tail -1 gpurun_out/r2_frame_trace.log | cut -c1-150
env RPT_LIB_DIR=$D timeout 900 ncu --set full --cache-control none --clock-control none --import-source on -k "regex:grisBounceKernel" -s 18 -c 6 -f \
   -o gpurun_out/r2_frame_bounce python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-4k > gpurun_out/r2_frame_bounce.log 2>&1
tail -1 gpurun_out/r2_frame_bounce.log | cut -c1-150
env RPT_LIB_DIR=$D timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:^(?!.*(ploc|Ploc|flatten|morton|initLeaves|collapse|DeviceRadixSort|DeviceScan))" -c 400 --csv --log-file gpurun_out/r2_launches.csv \
   python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-4k > gpurun_out/r2_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches.csv

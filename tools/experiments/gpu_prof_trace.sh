#!/bin/bash
# full ncu capture of the queue traversal kernels inside the frame (development aid)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:traceQueueKernel" -s 338 -c 2 -f -o gpurun_out/prof_tq \
   python bench.py --steps 4 --warmup 10 --no-cpu-baseline > gpurun_out/prof_tq.log 2>&1
tail -2 gpurun_out/prof_tq.log | cut -c1-200

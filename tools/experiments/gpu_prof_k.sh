#!/bin/bash
# full ncu capture of one kernel by regex + launch skip (development aid): gpu_prof_k.sh <regex> <skip> <out> [count]
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$1" -s $2 -c ${4:-1} -f -o gpurun_out/$3 \
   python bench.py --steps 4 --warmup 10 --no-cpu-baseline > gpurun_out/$3.log 2>&1
tail -1 gpurun_out/$3.log | cut -c1-200

#!/bin/bash
# one full ncu capture of a named kernel (development aid): gpu_prof1.sh <kernel regex> <out name>
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$1" -s 12 -c 1 -f -o gpurun_out/$2 \
   python bench.py --steps 4 --warmup 10 --no-cpu-baseline > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log | cut -c1-300

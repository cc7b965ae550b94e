#!/bin/bash
# full ncu captures of several kernels of the bench frame, one launch each (development aid): gpu_prof_multi.sh name:regex:skip ...
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r name regex skip <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c 1 -f -o gpurun_out/prof_$name \
     python bench.py --steps 3 --warmup 8 --no-cpu-baseline > gpurun_out/prof_$name.log 2>&1
  echo "$name: $(tail -1 gpurun_out/prof_$name.log | cut -c1-120)"
done
ls -la gpurun_out/*.ncu-rep

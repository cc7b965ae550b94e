#!/bin/bash
# ncu --set full of the queue traversal kernel on the microbenchmark's shuffled bounce rays (launch 15) and shuffled shadow rays
# (launch 41), for each experiment build given: gpu_r2_prof_trav.sh lib_base lib_compact ...
mkdir -p gpurun_out
for L in "$@"; do
  D=""; [ "$L" != "default" ] && D="$PWD/$L"
  for S in 14 40; do
    env RPT_LIB_DIR=$D timeout 600 ncu --set full --clock-control none --import-source on -k "regex:traceQueue" -s $S -c 1 -f \
       -o gpurun_out/r2_trav_${L}_$S python tools/gpu_tracebench.py > gpurun_out/r2_trav_${L}_$S.log 2>&1
    tail -2 gpurun_out/r2_trav_${L}_$S.log | cut -c1-200
  done
done
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# compute-sanitizer over the GPU tests (development aid): memcheck on everything, initcheck on the parity subset
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
   python -m pytest tests -x -q -m gpu > gpurun_out/memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/memcheck.log
grep -E "ERROR SUMMARY|Invalid|passed|failed|memcheck exit|out of bounds|misaligned" gpurun_out/memcheck.log | head -12
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 7 --print-limit 12 \
   python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cornell and (gi_bit or gris_bit or naive or di_bit)" > gpurun_out/initcheck.log 2>&1
echo "initcheck exit $?" >> gpurun_out/initcheck.log
grep -E "ERROR SUMMARY|Uninitialized|passed|failed|initcheck exit" gpurun_out/initcheck.log | head -12
grep -A12 "Uninitialized" gpurun_out/initcheck.log | head -60

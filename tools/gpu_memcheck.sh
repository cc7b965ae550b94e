#!/bin/bash
# compute-sanitizer over the GPU tests (development aid): memcheck on everything but the full-size BASELINE configurations,
# initcheck on a parity subset that includes the two-level scenes
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
   python -m pytest tests -x -q -m gpu -k "not baseline_configs and not fullsize and not cli" > gpurun_out/memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/memcheck.log
grep -E "ERROR SUMMARY|Invalid|passed|failed|memcheck exit|out of bounds|misaligned" gpurun_out/memcheck.log | head -12
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 7 --print-limit 12 \
   python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "(cornell or field_tlas) and (gi_bit or gris_bit or naive or di_bit or closest or queue)" > gpurun_out/initcheck.log 2>&1
echo "initcheck exit $?" >> gpurun_out/initcheck.log
grep -E "ERROR SUMMARY|Uninitialized|passed|failed|initcheck exit" gpurun_out/initcheck.log | head -12
grep -A12 "Uninitialized" gpurun_out/initcheck.log | head -60
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 12 \
   python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "field_tlas and (gris_bit or closest or queue)" > gpurun_out/racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/racecheck.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|passed|failed|racecheck exit" gpurun_out/racecheck.log | head -12

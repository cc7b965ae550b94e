#!/bin/bash
# round-end evidence: GPU tests, smoke, bench (both arms), launch list of the bench frame, full capture of the spatial shift kernels
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:gris|gbuffer|gBuffer|postProcess|traceQueue" -s 300 -c 128 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 6 --warmup 8 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:grisSpatialShift" -s 20 -c 2 -f -o gpurun_out/prof_shift \
   python bench.py --steps 4 --warmup 10 --no-cpu-baseline > gpurun_out/prof_shift.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_n1.json; cut -c1-300 gpurun_out/bench_ref.json; ls -la gpurun_out/prof_shift.ncu-rep

#!/bin/bash
# full GPU test suite + the driver's two bench arms at N=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16 > gpurun_out/r2_final_tests.log
cat gpurun_out/r2_final_tests.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err
tail -c 600 gpurun_out/r2_final_reference.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
tail -c 1500 gpurun_out/r2_final_bench_n1.json; tail -3 gpurun_out/r2_final_bench_n1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

#!/bin/bash
# N-GPU bench as the driver launches it (torchrun, one rank per GPU): strips image check + fixed-4K line
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 40 --warmup 10 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo rc=$?
tail -c 5000 gpurun_out/r2_bench_n$N.json; tail -20 gpurun_out/r2_bench_n$N.err

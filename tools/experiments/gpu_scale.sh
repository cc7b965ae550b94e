#!/bin/bash
# bench.py at N GPUs of one box (development aid): gpu_scale.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 60 --warmup 20 \
   > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?"
cat gpurun_out/bench_n$N.json | cut -c1-2500; tail -5 gpurun_out/bench_n$N.err

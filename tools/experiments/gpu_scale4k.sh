#!/bin/bash
# config 4: the 3840x2160 film on N GPUs (strong scaling; development aid): gpu_scale4k.sh N
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --film 3840x2160 --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/bench4k_n1.json 2> gpurun_out/bench4k_n1.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --film 3840x2160 --steps 40 --warmup 10 \
     > gpurun_out/bench4k_n$N.json 2> gpurun_out/bench4k_n$N.err
fi
echo "exit $?"; grep '^{' gpurun_out/bench4k_n$N.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); c = d['config']
print('4K on', d['n_gpus'], 'GPU(s): film frames/s %.2f  ms %.3f  value %.1f  e2e %.1f' % (c['film_frames_per_s'], d['ms_per_step'], d['value'], d['e2e']['value']))
print(c['workload'][:200])"; tail -3 gpurun_out/bench4k_n$N.err

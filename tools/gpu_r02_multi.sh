#!/bin/bash
# N-GPU pass: torchrun bench with the device-side peer hop vs host-issued NCCL send/recv.  gpurun --gpus N --timeout 900 -- 'bash tools/gpu_r02_multi.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
for hop in peer nccl; do
  echo "== N=$N hop=$hop"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 64 --warmup 8 --hop $hop \
      > gpurun_out/bench_r02_gpus${N}_${hop}.json 2> gpurun_out/bench_multi_${hop}.err
  tail -c 1200 gpurun_out/bench_r02_gpus${N}_${hop}.json; tail -4 gpurun_out/bench_multi_${hop}.err | cut -c1-300
done

#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 64 --warmup 8 \
    > gpurun_out/bench_r02_gpus${N}_final.json 2> gpurun_out/bench_multi_final.err
tail -c 1500 gpurun_out/bench_r02_gpus${N}_final.json; tail -3 gpurun_out/bench_multi_final.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>/dev/null | tail -c 600

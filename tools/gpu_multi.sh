#!/bin/bash
# N-GPU pass (gpurun --gpus N): layer-split pipeline through bench.py (one process per GPU, NCCL hop) and through the reference's own scheduler
# (one process, -sm layer, peer copies), plus parity of the 2-stage split against the CPU.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 64 --warmup 8 \
    > gpurun_out/bench_gpus$N.json 2> gpurun_out/bench_gpus$N.err
tail -c 1800 gpurun_out/bench_gpus$N.json; tail -3 gpurun_out/bench_gpus$N.err
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
python tools/make_gguf.py /tmp/f32.gguf --layers 4 --vocab 8192 --ftype f32 2>&1 | tail -1
GGML_BACKEND_PATH= timeout 600 oracle/_ref/bin/llama-quantize /tmp/f32.gguf /tmp/q4l.gguf q4_k_m $(nproc) > gpurun_out/quantize.log 2>&1; rm -f /tmp/f32.gguf
echo "== $N-GPU layer split vs CPU (llama_parity)"; timeout 300 oracle/_ref/bin/llama_parity /tmp/q4l.gguf 48 48 $(nproc) 1 2>&1 | tail -2 | tee gpurun_out/llama_parity_gpus$N.json
python tools/make_gguf.py /tmp/q8b.gguf 2>&1 | tail -1
timeout 900 oracle/_ref/bin/llama-bench -m /tmp/q8b.gguf -p 2048 -n 64 -d 0 -fa 1 -ngl 99 -sm layer -r 2 -o md 2> gpurun_out/llama_bench_multi.err | tee gpurun_out/llama_bench_gpus$N.md
tail -3 gpurun_out/llama_bench_multi.err

#!/bin/bash
set -u
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
echo "== 2 GPUs, -sm layer, host timing"
GGML_B200_HOST_TIMING=1 timeout 300 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 0 -fa 1 -ngl 99 -sm layer -r 3 -o md 2> gpurun_out/sm2.err | grep "pp"; grep "host time" gpurun_out/sm2.err
echo "== 2 GPUs, pp512 (one ubatch) and pp4096 (8 ubatches)"
timeout 300 oracle/_ref/bin/llama-bench -m $M -p 512,4096 -n 0 -fa 1 -ngl 99 -sm layer -r 3 -o md 2>/dev/null | grep "pp"
echo "== 2 GPUs, fusions off"
GGML_B200_NO_TILE_FUSION=1 timeout 300 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 0 -fa 1 -ngl 99 -sm layer -r 3 -o md 2>/dev/null | grep "pp"
echo "== 1 GPU"
timeout 300 oracle/_ref/bin/llama-bench -m $M -p 512,2048,4096 -n 0 -fa 1 -ngl 99 -dev B200:0 -r 3 -o md 2>/dev/null | grep "pp"

#!/bin/bash
set -u
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
N=$(nvidia-smi -L | wc -l)
echo "== $N GPUs, -sm layer"
GGML_B200_HOST_TIMING=1 timeout 300 oracle/_ref/bin/llama-bench -m $M -p 512,2048,4096 -n 32 -fa 1 -ngl 99 -sm layer -r 3 -o md 2> gpurun_out/sm5.err | grep "pp\|tg" | tee gpurun_out/llama_bench_r02_sm_layer_${N}gpu_pipelined.md; grep "synchronize" gpurun_out/sm5.err
echo "== 1 GPU"
timeout 300 oracle/_ref/bin/llama-bench -m $M -p 512,2048,4096 -n 32 -fa 1 -ngl 99 -dev B200:0 -r 3 -o md 2>/dev/null | grep "pp\|tg" | tee -a gpurun_out/llama_bench_r02_sm_layer_${N}gpu_pipelined.md
F=/tmp/b200_bench_qwen3_8b_f16.gguf
python tools/make_gguf.py $F --ftype f16 --reuse-layers 2>&1 | tail -1
echo "== F16 model, $N GPUs -sm layer"
timeout 600 oracle/_ref/bin/llama-bench -m $F -p 2048 -n 64 -d 0,4096 -fa 1 -ngl 99 -sm layer -r 2 -o md 2>/dev/null | grep "pp\|tg" | tee gpurun_out/llama_bench_r02_f16_${N}gpu_pipelined.md

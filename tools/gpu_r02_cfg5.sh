#!/bin/bash
# BASELINE.json configs[4]: Qwen3-8B-shape F16 model, ctx 8192, layer-split pipeline over N GPUs through the UNMODIFIED reference llama-bench (-sm layer) + the plugin.
# gpurun --gpus N --timeout 900 -- 'bash tools/gpu_r02_cfg5.sh N'
set -u
N=${1:-2}
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
F=/tmp/b200_bench_qwen3_8b_f16.gguf
python tools/make_gguf.py $F --ftype f16 --reuse-layers 2>&1 | tail -1
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -$N
timeout 600 oracle/_ref/bin/llama-bench -m $F -p 2048 -n 64 -d 0,4096,8000 -fa 1 -ngl 99 -sm layer -r 2 -o md 2>gpurun_out/lb_cfg5.err | grep -v "^$" | tee gpurun_out/llama_bench_r02_f16_${N}gpu_layer_split.md
tail -3 gpurun_out/lb_cfg5.err
if [ "$N" = "2" ]; then
  M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
  python tools/make_gguf.py $M 2>&1 | tail -1
  timeout 300 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 64 -d 0,2048 -fa 1 -ngl 99 -sm layer -r 2 -o md 2>>gpurun_out/lb_cfg5.err | grep "pp\|tg" | tee gpurun_out/llama_bench_r02_q4km_${N}gpu_layer_split.md
fi

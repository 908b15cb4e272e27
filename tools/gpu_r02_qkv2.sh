#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "flash_attn" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_plugin_gpu.py -m gpu -x -q --tb=short -k "FLASH_ATTN_EXT" 2>&1 | tail -4
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
timeout 300 oracle/_ref/bin/llama-bench -m $M -p 512 -n 64 -d 2048 -fa 1 -ngl 99 -ctk q8_0 -ctv q8_0 -r 2 -o md 2>/dev/null | grep "tg\|pp" | tee gpurun_out/llama_bench_r02_qkv2.md
timeout 300 oracle/_ref/bin/llama-bench -m $M -p 0 -n 64 -d 2048 -fa 1 -ngl 99 -ctk q4_0 -ctv q4_0 -r 2 -o md 2>/dev/null | grep "tg\|pp" | tee -a gpurun_out/llama_bench_r02_qkv2.md
GGML_B200_DISABLE_ENGINE=1 timeout 300 oracle/_ref/bin/llama-bench -m $M -p 0 -n 64 -d 2048 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | grep "tg\|pp" | tee -a gpurun_out/llama_bench_r02_qkv2.md

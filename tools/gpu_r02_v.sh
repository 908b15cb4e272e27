#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "mul_mat_multi" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_plugin_gpu.py -m gpu -x -q --tb=short -k "tile_fusion or engine_route or decode_only or MUL_MAT" 2>&1 | tail -6
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
for nf in 0 1; do echo "== GGML_B200_NO_TILE_FUSION=$nf"; GGML_B200_NO_TILE_FUSION=$nf timeout 300 oracle/_ref/bin/llama-bench -m $M -p 512 -n 0 -fa 1 -ngl 99 -r 3 -o md 2>/dev/null | grep pp
GGML_B200_NO_TILE_FUSION=$nf timeout 300 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 3 -o md 2>/dev/null | grep pp; done | tee gpurun_out/llama_bench_r02_v.md

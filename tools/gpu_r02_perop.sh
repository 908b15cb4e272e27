#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_plugin_gpu.py -m gpu -x -q --tb=short -k "engine_route or decode_only or three_host or fusions_are or tts_llama or MUL_MAT or RMS_NORM" 2>&1 | tail -8
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
for nf in 0 1; do echo "== GGML_B200_NO_TILE_FUSION=$nf (per-op route: engine disabled)"
GGML_B200_NO_TILE_FUSION=$nf GGML_B200_DISABLE_ENGINE=1 timeout 300 oracle/_ref/bin/llama-bench -m $M -p 0 -n 64 -d 0,2048 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | grep "tg"
GGML_B200_NO_TILE_FUSION=$nf timeout 300 oracle/_ref/bin/llama-bench -m $M -p 0 -n 64 -d 2048 -fa 1 -ngl 99 -ctk q8_0 -ctv q8_0 -r 2 -o md 2>/dev/null | grep "tg"
done | tee gpurun_out/llama_bench_r02_perop.md
echo "== engine route (default)"
timeout 300 oracle/_ref/bin/llama-bench -m $M -p 0 -n 64 -d 0,2048 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | grep "tg" | tee -a gpurun_out/llama_bench_r02_perop.md

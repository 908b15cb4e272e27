#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "decode_matvec_epilogues or mul_mat_add or f16_weights" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_plugin_gpu.py -m gpu -x -q --tb=short -k "tts_llama or fusions_are or decode_only or MUL_MAT or three_host" 2>&1 | tail -8
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
T=/tmp/tts_llama_f16.gguf
python tools/make_gguf.py $T --arch llama --ftype f16 --embd 768 --ff 3072 --heads 12 --kv-heads 12 --head-dim 64 --layers 20 --vocab 6562 2>&1 | tail -1
F=/tmp/b200_bench_qwen3_8b_f16.gguf
python tools/make_gguf.py $F --ftype f16 --reuse-layers 2>&1 | tail -1
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
for nf in 0 1; do echo "== GGML_B200_NO_TILE_FUSION=$nf"
GGML_B200_NO_TILE_FUSION=$nf timeout 300 oracle/_ref/bin/llama-bench -m $T -p 0 -n 128 -d 0,512 -fa 1 -ngl 99 -r 3 -o md 2>/dev/null | grep "tg"
GGML_B200_NO_TILE_FUSION=$nf timeout 300 oracle/_ref/bin/llama-bench -m $F -p 0 -n 64 -d 0,4096 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | grep "tg"
GGML_B200_NO_TILE_FUSION=$nf timeout 300 oracle/_ref/bin/llama-bench -m $M -p 0 -n 64 -d 2048 -fa 1 -ngl 99 -ctk q8_0 -ctv q8_0 -r 2 -o md 2>/dev/null | grep "tg"
done | tee gpurun_out/llama_bench_r02_decode_fusions.md

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_plugin_gpu.py -m gpu -x -q --tb=short -s -k "tts_llama" 2>&1 | tail -12
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
T=/tmp/tts_llama_f16.gguf
python tools/make_gguf.py $T --arch llama --ftype f16 --embd 768 --ff 3072 --heads 12 --kv-heads 12 --head-dim 64 --layers 20 --vocab 6562 2>&1 | tail -1
echo "== TTS-shaped llama (768 wide, 20 layers, F16): B200 plugin"
timeout 300 oracle/_ref/bin/llama-bench -m $T -p 128 -n 128 -d 0,512 -fa 1 -ngl 99 -r 3 -o md 2>/dev/null | grep "pp\|tg" | tee gpurun_out/llama_bench_r02_tts_llama.md
echo "== same binary, reference CPU backend"
GGML_BACKEND_PATH= timeout 300 oracle/_ref/bin/llama-bench -m $T -p 128 -n 128 -d 0 -ngl 0 -r 2 -o md 2>/dev/null | grep "pp\|tg" | tee -a gpurun_out/llama_bench_r02_tts_llama.md

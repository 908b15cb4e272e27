#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r02_p.log
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
Q=/tmp/b200_q8_0_2048.gguf
python tools/make_gguf.py $Q --ftype q8_0 --embd 2048 --ff 6144 --heads 16 --kv-heads 4 --layers 24 --vocab 32000 2>&1 | tail -1
echo "== Q8_0 model (n_embd 2048, 24 layers): engine route vs per-op route"
timeout 300 oracle/_ref/bin/llama-bench -m $Q -p 0 -n 64 -d 512 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | grep "tg" | tee gpurun_out/llama_bench_r02_q8_0_engine.md
GGML_B200_DISABLE_ENGINE=1 timeout 300 oracle/_ref/bin/llama-bench -m $Q -p 0 -n 64 -d 512 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | grep "tg" | tee -a gpurun_out/llama_bench_r02_q8_0_engine.md

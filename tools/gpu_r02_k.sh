#!/bin/bash
# residual-ADD-in-the-GEMM-epilogue pass: every GPU test, then bench.py (prefill leg) and the reference llama-bench pp numbers with the fusion on and off
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r02_k.log
timeout 600 python bench.py --steps 32 --warmup 8 --no-plugin-e2e > gpurun_out/bench_r02_k.json 2> gpurun_out/bench.err; tail -c 1800 gpurun_out/bench_r02_k.json; tail -3 gpurun_out/bench.err
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
for nf in 0 1; do echo "== GGML_B200_NO_TILE_FUSION=$nf"; GGML_B200_NO_TILE_FUSION=$nf timeout 300 oracle/_ref/bin/llama-bench -m $M -p 512,2048 -n 0 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | grep pp
GGML_B200_NO_TILE_FUSION=$nf timeout 300 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | grep pp; done | tee gpurun_out/llama_bench_r02_k.md

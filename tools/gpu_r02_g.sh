#!/bin/bash
# round 2, pass G: bench.py through the boundary (llama-bench + plugin as the e2e leg, llama-bench CPU as the reference arm), host-overhead probe
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "whole_token or engine_matches" 2>&1 | tail -3
timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/bench_r02_g.json 2> gpurun_out/bench.err; tail -c 3500 gpurun_out/bench_r02_g.json; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02_g_reference.json 2>> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench_r02_g_reference.json
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
echo "== llama-bench tg128 d0/d2048 (plugin)"; timeout 600 oracle/_ref/bin/llama-bench -m $M -p 0 -n 128 -d 0,2048 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | tee gpurun_out/llama_bench_r02_g.md
echo "== host-overhead probe: GGML_B200_NULL_COMPUTE=1 (no launches)"; GGML_B200_NULL_COMPUTE=1 timeout 600 oracle/_ref/bin/llama-bench -m $M -p 0 -n 128 -d 0 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | tee gpurun_out/llama_bench_r02_g_null.md
echo "== pp"; timeout 600 oracle/_ref/bin/llama-bench -m $M -p 512,2048 -n 0 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | tee -a gpurun_out/llama_bench_r02_g.md
timeout 600 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | tee -a gpurun_out/llama_bench_r02_g.md

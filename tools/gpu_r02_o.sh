#!/bin/bash
# F16 model (BASELINE.json configs[4]) on 1 GPU through the reference llama-bench + plugin; quantised KV cache end to end; F16 matvec GB/s
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "f16_weights or float_weights" 2>&1 | tail -5
timeout 300 python tools/f16_matvec_bench.py 2>&1 | grep -v Warn | tee gpurun_out/f16_matvec_bench.txt
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
echo "== Q4_K_M, q8_0 / q4_0 KV cache"
timeout 300 oracle/_ref/bin/llama-bench -m $M -p 512 -n 64 -d 2048 -fa 1 -ngl 99 -ctk q8_0 -ctv q8_0 -r 2 -o md 2>gpurun_out/lb_qkv.err | grep -v "^$" | tee gpurun_out/llama_bench_r02_qkv.md
timeout 300 oracle/_ref/bin/llama-bench -m $M -p 0 -n 64 -d 2048 -fa 1 -ngl 99 -ctk q4_0 -ctv q4_0 -r 2 -o md 2>>gpurun_out/lb_qkv.err | grep "tg\|pp" | tee -a gpurun_out/llama_bench_r02_qkv.md
tail -3 gpurun_out/lb_qkv.err
F=/tmp/b200_bench_qwen3_8b_f16.gguf
( time python tools/make_gguf.py $F --ftype f16 ) 2>&1 | tail -4
echo "== F16 model, 1 GPU"
timeout 600 oracle/_ref/bin/llama-bench -m $F -p 512,2048 -n 64 -d 0,4096 -fa 1 -ngl 99 -r 2 -o md 2>gpurun_out/lb_f16.err | grep -v "^$" | tee gpurun_out/llama_bench_r02_f16_1gpu.md
tail -3 gpurun_out/lb_f16.err

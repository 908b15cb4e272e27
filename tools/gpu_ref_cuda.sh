#!/bin/bash
# the reference's OWN CUDA backend (oracle/_ref/lib/libggml-cuda.so, built for sm_100 by `make -C oracle ref_cuda`) beside ours, same llama-bench binary,
# same GGUF, same box.  gpurun --timeout 900 -- 'bash tools/gpu_ref_cuda.sh'
set -u
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
for be in cuda b200; do
  if [ $be = cuda ]; then export GGML_BACKEND_PATH=$PWD/oracle/_ref/lib/libggml-cuda.so; else export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so; fi
  echo "== backend: $be"
  timeout 300 oracle/_ref/bin/llama-bench -m $M -p 512,2048 -n 64 -d 0,2048 -fa 1 -ngl 99 -r 2 -o md 2> gpurun_out/llama_bench_$be.err | tee gpurun_out/llama_bench_ref_${be}.md
  timeout 300 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 2 -o md 2>> gpurun_out/llama_bench_$be.err | tee -a gpurun_out/llama_bench_ref_${be}.md
  tail -2 gpurun_out/llama_bench_$be.err
done

#!/bin/bash
# End-to-end through the UNMODIFIED reference (libllama + llama-bench from oracle/_ref) with libggml-b200.so loaded via GGML_BACKEND_PATH:
#   1. greedy-token / logit parity CPU backend vs B200 backend on a 4-layer synthetic Qwen3 Q4_K_M GGUF (tests/native/llama_parity.cpp)
#   2. llama-bench tg/pp on the full 36-layer Qwen3-8B-shaped Q4_K_M GGUF (BASELINE.json configs[1], configs[2]) + the CPU backend beside it
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_llama.sh [parity|bench|all]'
set -u
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
what=${1:-all}
if [ "$what" = parity ] || [ "$what" = all ]; then
  # (a) Gaussian F32 weights quantised by the reference's own llama-quantize -> Q4_K_M with Q6_K attn_v / ffn_down / output, as in a real file
  python tools/make_gguf.py /tmp/f32.gguf --layers 4 --vocab 8192 --ftype f32 2>&1 | tail -1
  GGML_BACKEND_PATH= timeout 600 oracle/_ref/bin/llama-quantize /tmp/f32.gguf /tmp/q4l.gguf q4_k_m $(nproc) > gpurun_out/quantize.log 2>&1; tail -2 gpurun_out/quantize.log
  rm -f /tmp/f32.gguf
  : > gpurun_out/llama_parity.json
  timeout 600 oracle/_ref/bin/llama_parity /tmp/q4l.gguf 48 48 $(nproc) 1 2> gpurun_out/llama_parity.err | tee -a gpurun_out/llama_parity.json
  timeout 600 oracle/_ref/bin/llama_parity /tmp/q4l.gguf 48 48 $(nproc) 0 2>> gpurun_out/llama_parity.err | tee -a gpurun_out/llama_parity.json
  # the reference CPU backend against itself (plain vs repacked weights): the yardstick for what "parity" can mean on this model
  GGML_BACKEND_PATH= timeout 600 oracle/_ref/bin/llama_parity /tmp/q4l.gguf 48 48 $(nproc) 1 1 2>> gpurun_out/llama_parity.err | tee -a gpurun_out/llama_parity.json
  # (b) F16 weights: no activation quantisation on either side
  python tools/make_gguf.py /tmp/f16.gguf --layers 4 --vocab 8192 --ftype f16 2>&1 | tail -1
  timeout 600 oracle/_ref/bin/llama_parity /tmp/f16.gguf 48 48 $(nproc) 1 2>> gpurun_out/llama_parity.err | tee -a gpurun_out/llama_parity.json
  tail -3 gpurun_out/llama_parity.err
fi
if [ "$what" = bench ] || [ "$what" = all ]; then
  python tools/make_gguf.py /tmp/q8b.gguf 2>&1 | tail -1
  timeout 900 oracle/_ref/bin/llama-bench -m /tmp/q8b.gguf -p ${PP:-512} -n ${TG:-64} -d ${DEPTH:-0,2048} -fa 1 -ngl 99 -r 2 -o md 2> gpurun_out/llama_bench.err | tee gpurun_out/llama_bench_b200.md
  tail -3 gpurun_out/llama_bench.err
  if [ "${CPU:-1}" = 1 ]; then
    GGML_BACKEND_PATH= timeout 600 oracle/_ref/bin/llama-bench -m /tmp/q8b.gguf -p 0 -n 8 -fa 1 -ngl 0 -t $(nproc) -r 1 -o md 2>> gpurun_out/llama_bench.err | tee gpurun_out/llama_bench_cpu.md
  fi
fi

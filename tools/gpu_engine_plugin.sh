#!/bin/bash
# plugin -> decode engine: parity of the engine route against the per-op route and the CPU, then llama-bench tg with and without it
set -u
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
python tools/make_gguf.py /tmp/f32.gguf --layers 4 --vocab 8192 --ftype f32 2>&1 | tail -1
GGML_BACKEND_PATH= timeout 600 oracle/_ref/bin/llama-quantize /tmp/f32.gguf /tmp/q4l.gguf q4_k_m $(nproc) > gpurun_out/quantize.log 2>&1; rm -f /tmp/f32.gguf
echo "== engine vs per-op (mode 5)"; timeout 300 oracle/_ref/bin/llama_parity /tmp/q4l.gguf 48 64 $(nproc) 1 5 2>&1 | tail -3 | tee gpurun_out/llama_parity_engine.json
echo "== engine vs CPU";             timeout 300 oracle/_ref/bin/llama_parity /tmp/q4l.gguf 48 64 $(nproc) 1 2>&1 | tail -2 | tee -a gpurun_out/llama_parity_engine.json
echo "== per-op vs CPU";             GGML_B200_DISABLE_ENGINE=1 timeout 300 oracle/_ref/bin/llama_parity /tmp/q4l.gguf 48 64 $(nproc) 1 2>&1 | tail -2 | tee -a gpurun_out/llama_parity_engine.json
if [ "${1:-}" != "parity" ]; then
  python tools/make_gguf.py /tmp/q8b.gguf 2>&1 | tail -1
  timeout 600 oracle/_ref/bin/llama-bench -m /tmp/q8b.gguf -p 0 -n 64 -d 0,2048 -fa 1 -ngl 99 -r 2 -o md 2> gpurun_out/llama_bench.err | tee gpurun_out/llama_bench_engine.md
  GGML_B200_DISABLE_ENGINE=1 timeout 600 oracle/_ref/bin/llama-bench -m /tmp/q8b.gguf -p 0 -n 64 -d 0,2048 -fa 1 -ngl 99 -r 2 -o md 2>> gpurun_out/llama_bench.err | tee gpurun_out/llama_bench_perop.md
  tail -3 gpurun_out/llama_bench.err
fi

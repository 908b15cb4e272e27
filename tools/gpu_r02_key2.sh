#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_plugin_gpu.py -m gpu -x -q --tb=short -k "engine_route or decode_only or three_host or fusions_are or kshift" 2>&1 | tail -4
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
GGML_B200_HOST_TIMING=1 timeout 300 oracle/_ref/bin/llama-bench -m $M -p 0 -n 128 -d 0,2048 -fa 1 -ngl 99 -r 3 -o md 2> gpurun_out/key.err | grep "tg"; grep "host time" gpurun_out/key.err | tail -1

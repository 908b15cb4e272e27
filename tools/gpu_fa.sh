#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "flash_attn" 2>&1 | tail -25 | tee gpurun_out/pytest_fa.log
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
timeout 600 oracle/_ref/bin/test-backend-ops test -b B200:0 -o FLASH_ATTN_EXT > gpurun_out/tbo_fa.log 2>&1; tail -3 gpurun_out/tbo_fa.log | sed 's/\x1b\[[0-9;]*m//g'; grep -c "FAIL" gpurun_out/tbo_fa.log
if [ "${1:-}" = bench ]; then
  python tools/make_gguf.py /tmp/q8b.gguf 2>&1 | tail -1
  timeout 900 oracle/_ref/bin/llama-bench -m /tmp/q8b.gguf -p 512,2048 -n 0 -d 0,2048 -fa 1 -ngl 99 -r 2 -o md 2> gpurun_out/llama_bench.err | tee gpurun_out/llama_bench_pp.md
  timeout 900 oracle/_ref/bin/llama-bench -m /tmp/q8b.gguf -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 2 -o md 2>> gpurun_out/llama_bench.err | tee -a gpurun_out/llama_bench_pp.md
fi

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
timeout 600 oracle/_ref/bin/test-backend-ops test -b B200:0 -o FLASH_ATTN_EXT,GLU,MUL_MAT > gpurun_out/tbo_sel.log 2>&1; tail -3 gpurun_out/tbo_sel.log | sed 's/\x1b\[[0-9;]*m//g'; grep -c "FAIL" gpurun_out/tbo_sel.log
python tools/make_gguf.py /tmp/q8b.gguf 2>&1 | tail -1
timeout 900 oracle/_ref/bin/llama-bench -m /tmp/q8b.gguf -p 512,2048 -n 64 -d 0,2048 -fa 1 -ngl 99 -r 2 -o md 2> gpurun_out/llama_bench.err | tee gpurun_out/llama_bench_r01b.md
timeout 900 oracle/_ref/bin/llama-bench -m /tmp/q8b.gguf -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 2 -o md 2>> gpurun_out/llama_bench.err | tee -a gpurun_out/llama_bench_r01b.md
tail -2 gpurun_out/llama_bench.err
bash tools/gpu_pp_shares.sh 2>&1 | tail -22

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
timeout 600 oracle/_ref/bin/test-backend-ops test -b B200:0 -o FLASH_ATTN_EXT,MUL_MAT > gpurun_out/tbo_sel.log 2>&1; tail -3 gpurun_out/tbo_sel.log | sed 's/\x1b\[[0-9;]*m//g'; grep -c "FAIL" gpurun_out/tbo_sel.log
python tools/make_gguf.py /tmp/f32.gguf --layers 4 --vocab 8192 --ftype f32 2>&1 | tail -1
GGML_BACKEND_PATH= timeout 600 oracle/_ref/bin/llama-quantize /tmp/f32.gguf /tmp/q4l.gguf q4_k_m $(nproc) > gpurun_out/quantize.log 2>&1; rm -f /tmp/f32.gguf
timeout 300 oracle/_ref/bin/llama_parity /tmp/q4l.gguf 600 16 $(nproc) 1 2>&1 | tail -1
python tools/make_gguf.py /tmp/q8b.gguf 2>&1 | tail -1
timeout 900 oracle/_ref/bin/llama-bench -m /tmp/q8b.gguf -p 512,2048 -n 0 -fa 1 -ngl 99 -r 2 -o md 2> gpurun_out/llama_bench.err | tee gpurun_out/llama_bench_pp2.md
timeout 900 oracle/_ref/bin/llama-bench -m /tmp/q8b.gguf -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 2 -o md 2>> gpurun_out/llama_bench.err | tee -a gpurun_out/llama_bench_pp2.md
unset GGML_BACKEND_PATH
python bench.py --steps 32 --warmup 4 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('decode', j['value'], 'prefill', j['prefill']['value'], j['prefill']['roofline'])"

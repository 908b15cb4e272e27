#!/bin/bash
# does `llama-bench -sm layer` really spread the model over the devices, and does the reference scheduler pipeline the ubatches of a prompt over them?
set -u
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
timeout 300 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 32 -fa 1 -ngl 99 -sm layer -r 2 -o md -v 2> gpurun_out/lb_sm_verbose.err | grep "pp\|tg"
grep -i "buffer size\|pipeline\|offloaded\|using device\|n_copies\|graph splits" gpurun_out/lb_sm_verbose.err | sort | uniq -c | sort -rn | head -20
echo "== one device only (-dev B200:0)"
timeout 300 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 32 -fa 1 -ngl 99 -dev B200:0 -r 2 -o md 2>/dev/null | grep "pp\|tg"

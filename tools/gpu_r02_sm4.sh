#!/bin/bash
set -u
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
python tools/make_gguf.py /tmp/q2l.gguf --layers 2 --vocab 8192 2>&1 | tail -1
GGML_SCHED_DEBUG=2 timeout 120 oracle/_ref/bin/llama-bench -m /tmp/q2l.gguf -p 1024 -n 0 -fa 1 -ngl 99 -r 1 --no-warmup -v -o md > gpurun_out/sched_debug.txt 2>&1
grep -c "SPLIT" gpurun_out/sched_debug.txt
grep -n "## SPLIT" gpurun_out/sched_debug.txt | head -40
# the nodes of the small splits
awk '/## SPLIT/{s=$0; n=0} {n++; if (n<=6) print}' gpurun_out/sched_debug.txt | grep -v "^$" | tail -80 | cut -c1-200

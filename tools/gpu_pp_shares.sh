#!/bin/bash
set -u
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
python tools/make_gguf.py /tmp/q8b.gguf --layers 6 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/launches_pp.csv \
   oracle/_ref/bin/llama-bench -m /tmp/q8b.gguf -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 1 --no-warmup -o md > gpurun_out/ncu_pp.log 2>&1
tail -4 gpurun_out/ncu_pp.log
python tools/launch_shares.py gpurun_out/launches_pp.csv | tee gpurun_out/launch_shares_pp.txt

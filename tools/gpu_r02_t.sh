#!/bin/bash
# where does llama-bench pp2048 lose against bench.py's prefill leg?  (a) host-only time per prompt (no launches at all), (b) kernel shares of the plugin path
set -u
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
echo "== host only (GGML_B200_NULL_COMPUTE=2)"
GGML_B200_NULL_COMPUTE=2 timeout 300 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 3 -o md 2>/dev/null | grep pp | tee gpurun_out/llama_bench_r02_pp_host_only.md
GGML_B200_NULL_COMPUTE=2 timeout 300 oracle/_ref/bin/llama-bench -m $M -p 512 -n 0 -fa 1 -ngl 99 -r 3 -o md 2>/dev/null | grep pp | tee -a gpurun_out/llama_bench_r02_pp_host_only.md
echo "== full"
timeout 300 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 3 -o md 2>/dev/null | grep pp
python tools/make_gguf.py /tmp/q8b.gguf --layers 6 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/launches_pp.csv \
   oracle/_ref/bin/llama-bench -m /tmp/q8b.gguf -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 1 --no-warmup -o md > gpurun_out/ncu_pp.log 2>&1
python tools/launch_shares.py gpurun_out/launches_pp.csv | tee gpurun_out/launch_shares_pp_r02.txt

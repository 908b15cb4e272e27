#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chaos_yardstick.py -m gpu -x -q --tb=short 2>&1 | tail -40 > gpurun_out/pytest_chaos.log; tail -25 gpurun_out/pytest_chaos.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mul_mat_add or fused_activation" 2>&1 | tail -4
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-plugin-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['prefill']['value'], d['prefill']['ms'], d['prefill']['roofline']['frac'])"
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
for nf in 0 1; do echo "== GGML_B200_NO_TILE_FUSION=$nf"; GGML_B200_NO_TILE_FUSION=$nf timeout 300 oracle/_ref/bin/llama-bench -m $M -p 512 -n 0 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | grep pp
GGML_B200_NO_TILE_FUSION=$nf timeout 300 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 2 -o md 2>/dev/null | grep pp; done | tee gpurun_out/llama_bench_r02_l.md

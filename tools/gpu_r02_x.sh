#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "mul_mat_multi or fused_activation" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_plugin_gpu.py -m gpu -x -q --tb=short -k "tile_fusion or MUL_MAT" 2>&1 | tail -6
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
S=/tmp/small_f32.gguf; SQ=/tmp/small_q4.gguf
python tools/make_gguf.py $S --layers 4 --vocab 8192 --ftype f32 2>&1 | tail -1
env -u GGML_BACKEND_PATH oracle/_ref/bin/llama-quantize $S $SQ q4_k_m 16 > /dev/null 2>&1
echo "== llama_parity mode 8 (fusions on vs off), 600-token prompt: merged q/k/v + gate/up launches are active"
oracle/_ref/bin/llama_parity $SQ 600 8 16 1 8 2>/dev/null | grep "^{" | tee gpurun_out/llama_parity_mode8_600.json
echo "== llama_parity mode 0 (CPU vs B200, batched 600-token prompt)"
oracle/_ref/bin/llama_parity $SQ 600 8 16 1 0 2>/dev/null | grep "^{" | tee gpurun_out/llama_parity_mode0_600.json
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
timeout 300 oracle/_ref/bin/llama-bench -m $M -p 512 -n 0 -fa 1 -ngl 99 -r 3 -o md 2>/dev/null | grep pp | tee gpurun_out/llama_bench_r02_x.md
timeout 300 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 3 -o md 2>/dev/null | grep pp | tee -a gpurun_out/llama_bench_r02_x.md
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-plugin-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['prefill']['value'], d['prefill']['ms'], d['prefill']['roofline']['frac'])"

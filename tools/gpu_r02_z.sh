#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "flash_attn or fused_activation" 2>&1 | tail -4
timeout 600 python -m pytest tests/test_plugin_gpu.py -m gpu -x -q --tb=short -k "FLASH_ATTN_EXT" 2>&1 | tail -3
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-plugin-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['prefill']['value'], d['prefill']['ms'], d['prefill']['roofline']['frac'])"
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
S=/tmp/small_f32.gguf; SQ=/tmp/small_q4.gguf
python tools/make_gguf.py $S --layers 4 --vocab 8192 --ftype f32 2>&1 | tail -1
oracle/_ref/bin/llama-quantize $S $SQ q4_k_m 16 2>&1 | tail -1
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
echo "== llama_parity mode 8 (fusions on vs off), 600-token prompt: merged q/k/v + gate/up launches are active"
oracle/_ref/bin/llama_parity $SQ 600 8 16 1 8 2>gpurun_out/parity8.err | grep "^{" | tee gpurun_out/llama_parity_mode8_600.json; tail -2 gpurun_out/parity8.err
echo "== llama_parity mode 0 (CPU vs B200, batched 600-token prompt)"
oracle/_ref/bin/llama_parity $SQ 600 8 16 1 0 2>gpurun_out/parity0.err | grep "^{" | tee gpurun_out/llama_parity_mode0_600.json; tail -2 gpurun_out/parity0.err
echo "== llama_parity mode 1 (CPU vs CPU repacked: the yardstick)"
GGML_BACKEND_PATH= oracle/_ref/bin/llama_parity $SQ 600 8 16 1 1 2>gpurun_out/parity1.err | grep "^{" | tee gpurun_out/llama_parity_mode1_600.json

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "mul_mat_multi or fused_activation" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_plugin_gpu.py -m gpu -x -q --tb=short -k "preload or tile_fusion" 2>&1 | tail -6
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
S=/tmp/small_f32.gguf; SQ=/tmp/small_q4.gguf
python tools/make_gguf.py $S --layers 4 --vocab 8192 --ftype f32 2>&1 | tail -1
oracle/_ref/bin/llama-quantize $S $SQ q4_k_m 16 2>&1 | tail -2
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
echo "== llama_parity mode 8 (fusions on vs off), 600-token prompt: merged q/k/v + gate/up launches are active"
oracle/_ref/bin/llama_parity $SQ 600 8 16 1 8 2>gpurun_out/parity8.err | grep "^{" | tee gpurun_out/llama_parity_mode8_600.json; tail -3 gpurun_out/parity8.err
echo "== llama_parity mode 0 (CPU vs B200, batched 600-token prompt)"
oracle/_ref/bin/llama_parity $SQ 600 8 16 1 0 2>gpurun_out/parity0.err | grep "^{" | tee gpurun_out/llama_parity_mode0_600.json; tail -3 gpurun_out/parity0.err

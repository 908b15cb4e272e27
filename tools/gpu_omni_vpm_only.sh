#!/bin/bash
# the VPM frame after the F16 tensor-core GEMM took ragged K (ffn_down k = 4304): 2 frames, reference CPU backend vs plugin
set -u
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
T=$(nproc); [ "$T" -gt 16 ] && T=16
python tools/make_omni_gguf.py vpm /tmp/vpm.gguf 2> /dev/null
timeout 14 oracle/_ref/bin/omni_encoders vpm /tmp/vpm.gguf 2 $T 2> gpurun_out/omni_vpm_ragged_k.err | tee gpurun_out/r02_omni_vpm_ragged_k.json
tail -2 gpurun_out/omni_vpm_ragged_k.err

#!/bin/bash
# steady-state timings of the omni encoders on the plugin (several chunks / frames), the A/B switch of k_mm_simt, and one ncu --set full capture of its two attention launches
# gpurun --timeout 85 -- 'bash tools/gpu_omni_final.sh'
set -u
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
T=$(nproc); [ "$T" -gt 16 ] && T=16
python tools/make_omni_gguf.py apm /tmp/apm.gguf 2> /dev/null &
python tools/make_omni_gguf.py vpm /tmp/vpm.gguf 2> /dev/null &
wait
B=oracle/_ref/bin/omni_encoders
timeout 30 $B apm /tmp/apm.gguf 12 $T 2> /dev/null | tee gpurun_out/r02_omni_apm_final.json
timeout 30 $B vpm /tmp/vpm.gguf 3 $T 2> /dev/null | tee gpurun_out/r02_omni_vpm_final.json
B200_NO_SIMT_GEMM=1 timeout 30 $B vpm /tmp/vpm.gguf 2 $T 2> /dev/null | tee gpurun_out/r02_omni_vpm_no_simt.json
timeout 32 ncu --set full --clock-control none --import-source on -k regex:k_mm_simt -s 1 -c 2 -f -o gpurun_out/k_mm_simt_r02 $B vpm /tmp/vpm.gguf 1 $T > gpurun_out/ncu_simt.log 2>&1
tail -2 gpurun_out/ncu_simt.log | cut -c1-160

#!/bin/bash
# BASELINE.json configs[3], the encoder half ("APM 1 s audio chunk + VPM frame"): the reference's UNMODIFIED tools/omni/audition.cpp and vision.cpp (oracle/_ref/bin/omni_encoders,
# tests/native/omni_encoders.cpp) on full-size synthetic GGUFs, reference CPU backend vs libggml-b200.so in one process, plus the scheduler's node placement.
# gpurun --timeout 400 -- 'bash tools/gpu_omni_encoders.sh'
set -u
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
P=$PWD/llama.cpp-omni_b200/lib
T=$(nproc); [ "$T" -gt 16 ] && T=16
python tools/make_omni_gguf.py apm /tmp/apm.gguf 2> /dev/null &
python tools/make_omni_gguf.py vpm /tmp/vpm.gguf 2> /dev/null &
wait
B=oracle/_ref/bin/omni_encoders
GGML_BACKEND_PATH=$P/libggml-b200.so timeout 150 $B apm /tmp/apm.gguf 10 $T 2> gpurun_out/omni_apm.err | tee gpurun_out/r02_omni_apm.json
GGML_BACKEND_PATH=$P/libggml-b200.so timeout 150 $B vpm /tmp/vpm.gguf 3 $T 2> gpurun_out/omni_vpm.err | tee gpurun_out/r02_omni_vpm.json
# the same through the LD_PRELOAD shim, no ggml_backend_load_all() in the host program (what llama-omni-cli does)
OMNI_NO_LOAD_ALL=1 LD_PRELOAD=$P/libggml-b200-preload.so timeout 100 $B apm /tmp/apm.gguf 2 $T 2> gpurun_out/omni_apm_preload.err | tee gpurun_out/r02_omni_apm_preload.json
# node placement: which nodes the scheduler left on the CPU backend
GGML_SCHED_DEBUG=2 GGML_BACKEND_PATH=$P/libggml-b200.so timeout 100 $B apm /tmp/apm.gguf 2 $T 2> gpurun_out/omni_apm_sched.txt > /dev/null
GGML_SCHED_DEBUG=2 GGML_BACKEND_PATH=$P/libggml-b200.so timeout 100 $B vpm /tmp/vpm.gguf 1 $T 2> gpurun_out/omni_vpm_sched.txt > /dev/null
python tools/sched_census.py gpurun_out/omni_apm_sched.txt | tee gpurun_out/r02_omni_apm_sched_census.txt
python tools/sched_census.py gpurun_out/omni_vpm_sched.txt | tee gpurun_out/r02_omni_vpm_sched_census.txt
tail -3 gpurun_out/omni_apm.err gpurun_out/omni_vpm.err gpurun_out/omni_apm_preload.err

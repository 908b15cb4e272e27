#!/bin/bash
# NOT RUN YET (the round's GPU minutes ended): split the 48.9 ms VPM frame and the 6 ms APM chunk into kernels — the launch list the write-up
# profiles/r02_omni_encoders.md asks for — and time the frame without the reference's host-side work (graph build, sincos position embedding) by differencing.
# gpurun --timeout 300 -- 'bash tools/gpu_omni_launch_list.sh'
set -u
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
T=$(nproc); [ "$T" -gt 16 ] && T=16
python tools/make_omni_gguf.py apm /tmp/apm.gguf 2> /dev/null &
python tools/make_omni_gguf.py vpm /tmp/vpm.gguf 2> /dev/null &
wait
B=oracle/_ref/bin/omni_encoders
for what in vpm apm; do
    timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 4000 --csv --log-file gpurun_out/r02_omni_${what}_launches.csv $B $what /tmp/$what.gguf 1 $T > /dev/null 2>&1
    python tools/launch_shares.py gpurun_out/r02_omni_${what}_launches.csv | tee gpurun_out/r02_omni_${what}_launch_shares.txt | head -25
done
# GGML_B200_NULL_COMPUTE=2: the plugin launches nothing — what is left is the reference's host work per frame / chunk
GGML_B200_NULL_COMPUTE=2 timeout 60 $B vpm /tmp/vpm.gguf 3 $T 2> /dev/null | tee gpurun_out/r02_omni_vpm_host_only.json
GGML_B200_NULL_COMPUTE=2 timeout 60 $B apm /tmp/apm.gguf 6 $T 2> /dev/null | tee gpurun_out/r02_omni_apm_host_only.json

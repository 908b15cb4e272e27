#!/bin/bash
# ncu --set full capture of one production k_stream launch (source-level samples).  gpurun --timeout 900 -- 'bash tools/gpu_ncu_stream.sh tag'
set -u
tag=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 3 -c 1 -f -o gpurun_out/k_stream_${tag} \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-prefill > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-120
ls -la gpurun_out/k_stream_${tag}.ncu-rep

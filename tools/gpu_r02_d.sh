#!/bin/bash
# round 2, pass D: engine v2 timing details (no nanosleep in the polls; finer attention stamps)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "whole_token" 2>&1 | tail -3
for f in "" 12; do
  echo "== B200_SD_FLAGS=$f"
  B200_SD_FLAGS=$f timeout 300 python bench.py --steps 48 --warmup 8 --no-cpu-baseline --no-prefill 2>> gpurun_out/bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
done | tee gpurun_out/sd_flags_r02_v4.txt
for f in 4608 524 ; do
B200_SD_FLAGS=$f timeout 200 python tools/engine_profile.py --depth 2048 2>&1 | grep -v Warning > gpurun_out/engine_profile_r02_v4_$f.txt
head -20 gpurun_out/engine_profile_r02_v4_$f.txt
done

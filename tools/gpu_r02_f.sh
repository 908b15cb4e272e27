#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "whole_token or engine_matches" 2>&1 | tail -3
for f in "" 12; do
  echo "== B200_SD_FLAGS=$f"
  B200_SD_FLAGS=$f timeout 300 python bench.py --steps 48 --warmup 8 --no-cpu-baseline --no-prefill 2>> gpurun_out/bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
done | tee gpurun_out/sd_flags_r02_v6.txt
B200_SD_FLAGS=4096 timeout 200 python tools/engine_profile.py --depth 2048 2>&1 | grep -v Warning > gpurun_out/engine_profile_r02_v6.txt
head -20 gpurun_out/engine_profile_r02_v6.txt

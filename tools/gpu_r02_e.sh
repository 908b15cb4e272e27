#!/bin/bash
# pacing experiment: bulk copies in flight per ring (B200_SD_FLAGS bits 4-7) — does a shallower weight queue shorten the hand-off round trips?
set -u
mkdir -p gpurun_out
for f in 4096 4112 4128 4144; do
  echo "== B200_SD_FLAGS=$f"
  B200_SD_FLAGS=$f timeout 300 python bench.py --steps 48 --warmup 8 --no-cpu-baseline --no-prefill 2>> gpurun_out/bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
done | tee gpurun_out/sd_flags_r02_v5.txt
B200_SD_FLAGS=4112 timeout 200 python tools/engine_profile.py --depth 2048 2>&1 | grep -v Warning > gpurun_out/engine_profile_r02_v5_inflight1.txt
head -9 gpurun_out/engine_profile_r02_v5_inflight1.txt

#!/bin/bash
# round 2, pass C: engine v2 + tensor-core attention phase: correctness (oracle whole token, per-op route, FA cases) + timing
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "whole_token or engine_matches" 2>&1 | tail -15 | tee gpurun_out/pytest_engine_v2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for f in "" 4096 12; do
  echo "== B200_SD_FLAGS=$f"
  B200_SD_FLAGS=$f timeout 300 python bench.py --steps 48 --warmup 8 --no-cpu-baseline --no-prefill 2>> gpurun_out/bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
done | tee gpurun_out/sd_flags_r02_v3.txt
for d in 0 3900; do
  echo "== depth $d"
  timeout 300 python bench.py --depth $d --steps 48 --warmup 8 --no-cpu-baseline --no-prefill 2>> gpurun_out/bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
done | tee -a gpurun_out/sd_flags_r02_v3.txt
tail -3 gpurun_out/bench.err
B200_SD_FLAGS=4096 timeout 200 python tools/engine_profile.py --depth 2048 2>&1 | grep -v Warning > gpurun_out/engine_profile_r02_v3.txt
head -20 gpurun_out/engine_profile_r02_v3.txt

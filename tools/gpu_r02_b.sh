#!/bin/bash
# round 2, pass B: engine v2 (flag-in-data hand-off, no grid barriers) — correctness vs the oracle + timing; F16 parity bisect by op
set -u
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
PLUG=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
( python tools/make_gguf.py /tmp/f16.gguf --layers 4 --vocab 8192 --ftype f16 2>&1 | tail -1 ) &
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "whole_token or engine_matches" 2>&1 | tail -15 | tee gpurun_out/pytest_engine_v2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for f in "" 4096 4 12; do
  echo "== B200_SD_FLAGS=$f"
  B200_SD_FLAGS=$f timeout 300 python bench.py --steps 48 --warmup 8 --no-cpu-baseline --no-prefill 2>> gpurun_out/bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
done | tee gpurun_out/sd_flags_r02_v2.txt
tail -3 gpurun_out/bench.err
B200_SD_FLAGS=4096 timeout 200 python tools/engine_profile.py --depth 2048 2>&1 | grep -v Warning > gpurun_out/engine_profile_r02_v2.txt
head -20 gpurun_out/engine_profile_r02_v2.txt
wait
: > gpurun_out/f16_bisect.txt
for ops in "" MUL_MAT RMS_NORM ROPE SOFT_MAX GLU ADD MUL SET_ROWS CPY GET_ROWS "SOFT_MAX,MUL_MAT"; do
  echo "== DISABLE_OPS=$ops" >> gpurun_out/f16_bisect.txt
  GGML_B200_DISABLE_OPS=$ops GGML_BACKEND_PATH=$PLUG timeout 300 oracle/_ref/bin/llama_parity /tmp/f16.gguf 48 16 $(nproc) 0 6 2>> gpurun_out/llama_parity.err | cut -c1-200 >> gpurun_out/f16_bisect.txt
done
cat gpurun_out/f16_bisect.txt

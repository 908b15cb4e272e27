#!/bin/bash
# round 2, pass A: decode-only parity through libllama (mode 6), whole-token engine vs oracle, and decode-engine experiment switches
# (B200_SD_FLAGS: 4 = no weight traffic, 8 = no arithmetic, 1024 = L2 prefetch of staged phases, ...).  gpurun --timeout 1500 -- 'bash tools/gpu_r02_a.sh'
set -u
mkdir -p gpurun_out
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
PLUG=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
( python tools/make_gguf.py /tmp/f32.gguf --layers 4 --vocab 8192 --ftype f32 2>&1 | tail -1
  timeout 600 oracle/_ref/bin/llama-quantize /tmp/f32.gguf /tmp/q4l.gguf q4_k_m $(nproc) > gpurun_out/quantize.log 2>&1; rm -f /tmp/f32.gguf
  python tools/make_gguf.py /tmp/f16.gguf --layers 4 --vocab 8192 --ftype f16 2>&1 | tail -1 ) &
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "whole_token or engine_matches" 2>&1 | tail -15 | tee gpurun_out/pytest_oracle_token.log
wait
: > gpurun_out/llama_parity_r02.json
for model in q4l f16; do for fa in 0 1; do for mode in 6 0; do
  echo "== $model fa=$fa mode=$mode" | tee -a gpurun_out/llama_parity_r02.json
  GGML_BACKEND_PATH=$PLUG timeout 600 oracle/_ref/bin/llama_parity /tmp/$model.gguf 48 32 $(nproc) $fa $mode 2>> gpurun_out/llama_parity.err | tee -a gpurun_out/llama_parity_r02.json
done; done; done
echo "== q4l fa=1 mode=6 ENGINE OFF" | tee -a gpurun_out/llama_parity_r02.json
GGML_B200_DISABLE_ENGINE=1 GGML_BACKEND_PATH=$PLUG timeout 600 oracle/_ref/bin/llama_parity /tmp/q4l.gguf 48 32 $(nproc) 1 6 2>> gpurun_out/llama_parity.err | tee -a gpurun_out/llama_parity_r02.json
tail -3 gpurun_out/llama_parity.err
: > gpurun_out/sd_flags_r02.txt
for f in "" 4096 4 8 12 1024 9216 25600 3072 33792 1025; do
  echo "== B200_SD_FLAGS=$f" >> gpurun_out/sd_flags_r02.txt
  B200_SD_FLAGS=$f timeout 300 python bench.py --steps 48 --warmup 8 --no-cpu-baseline --no-prefill 2>> gpurun_out/bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])" >> gpurun_out/sd_flags_r02.txt
done
cat gpurun_out/sd_flags_r02.txt
B200_SD_FLAGS=4096 timeout 200 python tools/engine_profile.py --depth 2048 2>&1 | grep -v Warning > gpurun_out/engine_profile_r02_base.txt
B200_SD_FLAGS=1024 timeout 200 python tools/engine_profile.py --depth 2048 2>&1 | grep -v Warning > gpurun_out/engine_profile_r02_l2pf.txt
B200_SD_FLAGS=4 timeout 200 python tools/engine_profile.py --depth 2048 2>&1 | grep -v Warning > gpurun_out/engine_profile_r02_dry.txt
head -9 gpurun_out/engine_profile_r02_base.txt gpurun_out/engine_profile_r02_l2pf.txt gpurun_out/engine_profile_r02_dry.txt

#!/bin/bash
# engine iteration pass: parity tests, per-phase profile, short bench.   gpurun --timeout 900 -- 'bash tools/gpu_engine.sh'
set -u
mkdir -p gpurun_out; rm -f gpurun_out/engine_profile.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for f in ${SD_FLAGS_LIST:-0}; do
  B200_SD_FLAGS=$f timeout 200 python tools/engine_profile.py --depth 2048 2>&1 | grep -v Warning | tee -a gpurun_out/engine_profile.txt
done
timeout 300 python bench.py --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench.err
python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
    print("bench:", j["value"], "tok/s  frac", j["roofline"]["frac"], " e2e", j["e2e"]["value"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/bench.err").read()[-2000:])
PY

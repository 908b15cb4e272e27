#!/bin/bash
# scaling pass on an 8-GPU box: N = 2, 4, 8 with the device-side hop, N = 8 also with host-issued NCCL.  gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_r02_scale.sh'
set -u
mkdir -p gpurun_out
run() { N=$1; hop=$2; echo "== N=$N hop=$hop"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 64 --warmup 8 --hop $hop \
      > gpurun_out/bench_r02_gpus${N}_${hop}.json 2> gpurun_out/bench_multi_${N}_${hop}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_r02_gpus${N}_${hop}.json").read().strip().splitlines()[-1]); print(d["n_gpus"], "value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], d["config"]["parallelism"])
except Exception as e: print("FAILED", e)
PY
  grep -i "error\|Traceback" gpurun_out/bench_multi_${N}_${hop}.err | head -3; }
run 8 peer; run 8 nccl; run 4 peer; run 2 peer; run 2 nccl

#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench, and the ncu launch list.  Usage (from the container):
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh [quick]'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 64 --warmup 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${1:-}" != "quick" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 1300 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  tail -3 gpurun_out/ncu_bench.log
fi

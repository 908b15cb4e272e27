#!/bin/bash
# One complete GPU-box pass for a round: parity tests, smoke, bench (+reference arm), per-phase engine timeline, the ncu launch list of the
# bench command and ONE `--set full` capture of the dominant kernel.   gpurun --timeout 1700 -- 'bash tools/gpu_round.sh [tag]'
set -u
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 64 --warmup 8 > gpurun_out/bench_$tag.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench_$tag.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --depth 0 --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/bench_${tag}_d0.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --depth 3900 --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/bench_${tag}_d3900.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --per-op --steps 64 --warmup 8 --no-cpu-baseline > gpurun_out/bench_${tag}_perop.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2>> gpurun_out/bench.err
tail -c 1200 gpurun_out/bench_${tag}_reference.json
timeout 200 python tools/engine_profile.py --depth 2048 2>&1 | grep -v Warning > gpurun_out/engine_profile_$tag.txt
head -12 gpurun_out/engine_profile_$tag.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 3 -c 1 -f -o gpurun_out/k_stream_$tag \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out

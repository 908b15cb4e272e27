#!/bin/bash
# final pass of round 2: every GPU test, smoke, the bench lines (both arms), the ncu evidence of the dominant kernel, the reference's llama-bench on the plugin.
# gpurun --timeout 2400 -- 'bash tools/gpu_final_r02.sh'
set -u
tag=r02
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_${tag}_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 64 --warmup 8 > gpurun_out/bench_${tag}_final.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench_${tag}_final.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --depth 0 --steps 64 --warmup 8 --no-cpu-baseline --no-prefill --no-plugin-e2e > gpurun_out/bench_${tag}_final_d0.json 2>> gpurun_out/bench.err
timeout 300 python bench.py --depth 3900 --steps 64 --warmup 8 --no-cpu-baseline --no-prefill --no-plugin-e2e > gpurun_out/bench_${tag}_final_d3900.json 2>> gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_${tag}_final_reference.json 2>> gpurun_out/bench.err; tail -c 700 gpurun_out/bench_${tag}_final_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 60 --csv --log-file gpurun_out/launches_${tag}_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prefill --no-plugin-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 3 -c 1 -f -o gpurun_out/k_stream_${tag}_final \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-prefill --no-plugin-e2e > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-80
python tools/ncu_summary.py gpurun_out/k_stream_${tag}_final.ncu-rep > gpurun_out/k_stream_${tag}_final_ncu_summary.txt 2>&1; head -30 gpurun_out/k_stream_${tag}_final_ncu_summary.txt
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
M=/tmp/b200_bench_qwen3_8b_q4_k_m.gguf
python tools/make_gguf.py $M 2>&1 | tail -1
timeout 900 oracle/_ref/bin/llama-bench -m $M -p 512,2048 -n 128 -d 0,2048 -fa 1 -ngl 99 -r 2 -o md 2> gpurun_out/llama_bench.err | tee gpurun_out/llama_bench_${tag}_final.md
timeout 900 oracle/_ref/bin/llama-bench -m $M -p 2048 -n 0 -ub 2048 -b 2048 -fa 1 -ngl 99 -r 2 -o md 2>> gpurun_out/llama_bench.err | tee -a gpurun_out/llama_bench_${tag}_final.md

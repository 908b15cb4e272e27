#!/bin/bash
set -u
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
timeout 1500 oracle/_ref/bin/test-backend-ops test -b B200:0 > gpurun_out/tbo.log 2>&1
echo "tbo exit $?"
grep -c "OK$\|\[1;32mOK" gpurun_out/tbo.log | sed 's/^/ok lines: /'
grep -c "not supported" gpurun_out/tbo.log | sed 's/^/not supported: /'
grep "FAIL\|ERR\|abort\|Abort" gpurun_out/tbo.log | sed 's/\x1b\[[0-9;]*m//g' | cut -c1-220 | head -30
tail -4 gpurun_out/tbo.log | sed 's/\x1b\[[0-9;]*m//g'
unset GGML_BACKEND_PATH
SKIP=104 bash tools/gpu_tc_ncu.sh r01_v2
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_fa_prefill -s 2 -c 1 -f -o gpurun_out/k_fa_prefill_r01 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fa.log 2>&1
tail -2 gpurun_out/ncu_fa.log

#!/bin/bash
# plugin pass on the GPU box: the reference's own per-op parity harness (tests/test-backend-ops.cpp, built by oracle/Makefile)
# against libggml-b200.so loaded through GGML_BACKEND_PATH.   gpurun --timeout 1500 -- 'bash tools/gpu_plugin.sh [ops...]'
set -u
mkdir -p gpurun_out
export GGML_BACKEND_PATH=$PWD/llama.cpp-omni_b200/lib/libggml-b200.so
export LD_LIBRARY_PATH=$PWD/oracle/_ref/lib:$PWD/llama.cpp-omni_b200/lib:${LD_LIBRARY_PATH:-}
OPS="${*:-ALL}"
if [ "$OPS" = "ALL" ]; then
  timeout 1200 oracle/_ref/bin/test-backend-ops test -b B200:0 > gpurun_out/tbo.log 2>&1
else
  : > gpurun_out/tbo.log
  for o in $OPS; do timeout 600 oracle/_ref/bin/test-backend-ops test -b B200:0 -o $o >> gpurun_out/tbo.log 2>&1; done
fi
echo "exit $?"
grep -c "OK$\|\[1;32mOK" gpurun_out/tbo.log | sed 's/^/ok lines: /'
grep -c "not supported" gpurun_out/tbo.log | sed 's/^/not supported: /'
grep "FAIL\|ERR\|error\|abort\|Abort" gpurun_out/tbo.log | sed 's/\x1b\[[0-9;]*m//g' | cut -c1-220 | head -60
tail -5 gpurun_out/tbo.log | sed 's/\x1b\[[0-9;]*m//g'

#!/bin/bash
# tcgen05 prefill GEMM pass: parity tests of the tensor-core path, then the shape bench.   gpurun --timeout 600 -- 'bash tools/gpu_tc.sh'
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "prefill_tensor_core or mul_mat_vs_oracle" 2>&1 | tail -25 | tee gpurun_out/pytest_tc.log
timeout 200 python tools/prefill_bench.py ${1:-2048} 2>&1 | grep -v Warning | tee gpurun_out/prefill_bench.txt
timeout 100 python tools/prefill_bench.py 512 2>&1 | grep -v Warning | tee -a gpurun_out/prefill_bench.txt

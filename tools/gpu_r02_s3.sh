#!/bin/bash
# round 2, third session: the whole GPU suite (incl. the omni-encoder tests, k_mm_simt parity, the hidden-state pattern) with the encoder timings written to gpurun_out/
# gpurun --timeout 345 -- 'bash tools/gpu_r02_s3.sh'
set -u
mkdir -p gpurun_out
export OMNI_RESULTS_DIR=$PWD/gpurun_out
timeout 335 python -m pytest tests -m gpu -q --durations=25 -p no:cacheprovider 2>&1 | tail -60 > gpurun_out/r02_pytest_gpu_session3.log
tail -45 gpurun_out/r02_pytest_gpu_session3.log

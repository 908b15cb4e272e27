#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_r02_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3

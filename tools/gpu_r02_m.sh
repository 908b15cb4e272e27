#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chaos_yardstick.py -m gpu -x -q --tb=short 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "mul_mat_add or fused_activation or wave" 2>&1 | tail -30
timeout 900 python -m pytest tests/test_plugin_gpu.py -m gpu -q --tb=short -k "CONCAT or SIN or tile_fusion" 2>&1 | tail -30

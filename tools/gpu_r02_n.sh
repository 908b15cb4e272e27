#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "quantised or rope or set_rows or flash_attn" 2>&1 | tail -30
timeout 900 python -m pytest tests/test_plugin_gpu.py -m gpu -q --tb=short -k "ROPE or SET_ROWS or GET_ROWS or CPY or CONT or FLASH_ATTN_EXT" 2>&1 | tail -30

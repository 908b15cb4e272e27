#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "mul_mat_multi" 2>&1 | grep -v "^$" | tail -60

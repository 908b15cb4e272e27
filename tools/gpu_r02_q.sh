#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -s -k "whole_token" 2>&1 | grep -v Warn | tail -15

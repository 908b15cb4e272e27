#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_plugin_gpu.py -m gpu -x -q --tb=short -s -k "kshift" 2>&1 | tail -12

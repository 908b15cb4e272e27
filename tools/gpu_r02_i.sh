#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_activation or im2col or pool_1d or norm_matches or flash_attn" 2>&1 | tail -4
timeout 600 python -m pytest tests/test_plugin_gpu.py -q -x -k "IM2COL or NORM" 2>&1 | tail -3
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-plugin-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['prefill']['value'], d['prefill']['ms'], d['prefill']['roofline']['frac'])"

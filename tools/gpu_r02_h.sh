#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "rope or rms_norm or qkv_post or norm or engine_matches" 2>&1 | tail -4
timeout 600 python -m pytest tests/test_plugin_gpu.py -q -x -k "ROPE or RMS_NORM or NORM or IM2COL" 2>&1 | tail -4
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-plugin-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['prefill']['value'], d['prefill']['ms'], d['prefill']['roofline']['frac'])"

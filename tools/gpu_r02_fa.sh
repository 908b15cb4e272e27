#!/bin/bash
# tcgen05 FA prefill: parity tests, the reference's harness for the op, timing vs the mma.sync kernel
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "flash_attn" 2>&1 | tail -12 | tee gpurun_out/pytest_fa_tc.log
timeout 600 python -m pytest tests/test_plugin_gpu.py -q -x -k "FLASH_ATTN_EXT" 2>&1 | tail -5
cat > /tmp/fa_time.py <<'PY'
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
ops = load_package().ops
for n in (512, 2048):
    H, HK, D = 32, 8, 128
    q = torch.randn(n, H, D, device='cuda'); k = (torch.randn(n, HK, D, device='cuda') * 0.5).half(); v = torch.randn(n, HK, D, device='cuda').half()
    mask = torch.full((n, n), float('-inf'), device='cuda').triu(1).half()
    out = torch.empty(n, H, D, device='cuda'); scratch = torch.empty(1 << 20, dtype=torch.uint8, device='cuda')
    f = lambda: ops.flash_attn(q.permute(1, 0, 2), k.permute(1, 0, 2), v.permute(1, 0, 2), mask, D ** -0.5, out=out, scratch=scratch)
    for _ in range(3): f()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    fl = 4.0 * H * D * n * n / 2
    print(f"n={n}: {us:.1f} us per call, {fl / us / 1e6:.1f} TFLOP/s (causal)")
PY
echo "== tcgen05"; timeout 120 python /tmp/fa_time.py
echo "== mma.sync"; B200_DISABLE_FA_TC=1 timeout 120 python /tmp/fa_time.py

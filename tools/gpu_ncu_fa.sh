#!/bin/bash
set -u
mkdir -p gpurun_out
cat > /tmp/fa_one.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from __graft_entry__ import load_package
ops = load_package().ops
n, H, HK, D = 2048, 32, 8, 128
q = torch.randn(n, H, D, device='cuda'); k = (torch.randn(n, HK, D, device='cuda') * 0.5).half(); v = torch.randn(n, HK, D, device='cuda').half()
mask = torch.full((n, n), float('-inf'), device='cuda').triu(1).half()
out = torch.empty(n, H, D, device='cuda'); scratch = torch.empty(1 << 20, dtype=torch.uint8, device='cuda')
for _ in range(3): ops.flash_attn(q.permute(1, 0, 2), k.permute(1, 0, 2), v.permute(1, 0, 2), mask, D ** -0.5, out=out, scratch=scratch)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fa_tc -s 2 -c 1 -f -o gpurun_out/k_fa_tc_${1:-v1} python /tmp/fa_one.py > gpurun_out/ncu_fa.log 2>&1
tail -2 gpurun_out/ncu_fa.log | cut -c1-100

#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --tb=short -k "wave" 2>&1 | tail -4
timeout 600 python -m pytest tests/test_plugin_gpu.py -m gpu -x -q --tb=short -k "CONCAT" 2>&1 | tail -3
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from __graft_entry__ import load_package
ops = load_package().ops
w = (torch.randn(512, 256, 16, device="cuda") * 0.02).half(); x = torch.randn(512, 200, device="cuda")
for _ in range(3): ops.conv_transpose_1d(w, x, 8)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.conv_transpose_1d(w, x, 8)
e1.record(); torch.cuda.synchronize()
print("conv_transpose_1d [K16, Cout256, Cin512] x [L200, Cin512], s0 = 8: %.1f us per call" % (e0.elapsed_time(e1) / 20 * 1e3))
PY

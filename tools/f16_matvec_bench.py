#!/usr/bin/env python
"""GB/s of the single-column F16 matvec (k_mmvf16_stream) at the Qwen3-8B F16 shapes (BASELINE.json configs[4]); weights rotate over copies > L2."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402

ops = load_package().ops
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
dev = torch.device("cuda:0")
for name, m, k in [("wq/wo", 4096, 4096), ("wk/wv", 1024, 4096), ("gate/up", 12288, 4096), ("down", 4096, 12288), ("lm_head", 151748, 4096)]:
    ncopy = max(2, int(400e6 // (m * k * 2)) + 1)
    ws = [torch.randn(m, k, device=dev).half() * 0.05 for _ in range(ncopy)]
    x = torch.randn(1, k, device=dev)
    out = torch.empty(1, m, device=dev)
    iters = 40
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(3):
            ops.mul_mat(ws[i % ncopy], ops.F16, m, k, x, w_ne=[k, m], out=out)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for i in range(iters):
                ops.mul_mat(ws[i % ncopy], ops.F16, m, k, x, w_ne=[k, m], out=out)
        best = 1e30
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st); st.synchronize()
            best = min(best, e0.elapsed_time(e1))
    us = best / iters * 1e3
    gbs = m * k * 2 / us / 1e3
    print(f"{name:8s} {m:6d} x {k:5d}  {us:8.2f} us  {gbs:7.1f} GB/s  {gbs / PEAK:.3f} of measured HBM peak")

#!/usr/bin/env python
"""Launch ONE decode matvec shape a few times (for ncu): python tools/prof_one.py {lm_head|gateup|qkv|down}"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package
pkg = load_package(); ops, dec = pkg.ops, pkg.decode
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "lm_head"
gen = torch.Generator(device=dev); gen.manual_seed(0)
E, F = 4096, 12288
if which == "lm_head":
    wt, shapes, k, sw = ops.Q6_K, [(151748, E)], E, False
elif which == "gateup":
    wt, shapes, k, sw = ops.Q4_K, [(F, E), (F, E)], E, True
elif which == "qkv":
    wt, shapes, k, sw = ops.Q4_K, [(4096, E), (1024, E), (1024, E)], E, False
else:
    wt, shapes, k, sw = ops.Q4_K, [(E, F)], F, False
ws = [dec._rand_weight(wt, m, kk, gen, dev) for m, kk in shapes]
x = torch.randn(1, k, device=dev)
act = ops.quantize_act(wt, x)
ys = [torch.zeros(m, device=dev) for m, _ in shapes]
layout = ops.LAYOUT_PLANAR if wt == ops.Q6_K else ops.LAYOUT_NATIVE
for _ in range(5):
    jobs = [ops.make_job(w, wt, m, k, y, None, layout) for w, (m, _), y in zip(ws, shapes, ys)]
    if sw:
        ops.matvec_q_swiglu(jobs[0], jobs[1], ys[0], act, k)
    else:
        ops.matvec_q(jobs, act, k)
torch.cuda.synchronize()
print("done", which)

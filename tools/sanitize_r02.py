#!/usr/bin/env python
"""Small invocations of the round-2 kernels for compute-sanitizer (ragged shapes on purpose):  compute-sanitizer --tool memcheck python tools/sanitize_r02.py
k_mmq_tc with three weight segments, the residual epilogue, k_mmvf16_stream (plain / residual / GLU), k_quant_rows / k_dequant_rows (native + planar), k_fa_decode over a
q8_0 / q4_0 cache, k_fa_tc, the Token2Wav op set, ROPE on F16."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from __graft_entry__ import load_package
pkg = load_package(); ops, dec = pkg.ops, pkg.decode
dev = torch.device("cuda:0")
torch.manual_seed(0)
gen = torch.Generator(device=dev); gen.manual_seed(0)
acc = 0.0
def mark(tag):
    torch.cuda.synchronize(); print("  %-28s %.6f" % (tag, acc))
# merged q/k/v launch (ragged m, n chosen so that the merged launch needs no split-K) and the residual epilogue
k, n = 1024, 600
ws = [(dec._rand_weight(t, m, k, gen, dev), t, m, ops.LAYOUT_PLANAR if t == ops.Q6_K else ops.LAYOUT_NATIVE) for t, m in ((ops.Q4_K, 1100), (ops.Q4_K, 300), (ops.Q6_K, 260))]
x = torch.randn(n, k, device=dev)
outs = [torch.empty(n, m, device=dev) for (_, _, m, _) in ws]
ops.mul_mat_multi(ws, x, outs); acc += float(sum(o.abs().sum() for o in outs))
r = torch.randn(n, 1100, device=dev)
acc += float(ops.mul_mat_add(ws[0][0], ops.Q4_K, 1100, k, x, r, torch.empty_like(r)).abs().sum())
mark("multi + resid")
# F16 streaming matvec: plain, residual, GLU (m odd, k = 3 x 256)
wf = (torch.randn(131, 768, device=dev) * 0.05).half(); wf2 = (torch.randn(131, 768, device=dev) * 0.05).half()
xv = torch.randn(1, 768, device=dev); rv = torch.randn(1, 131, device=dev)
acc += float(ops.mul_mat(wf, ops.F16, 131, 768, xv, w_ne=[768, 131]).abs().sum())
acc += float(ops.mul_mat_add(wf, ops.F16, 131, 768, xv, rv, torch.empty_like(rv)).abs().sum())
acc += float(ops.mul_mat_glu(ops.GLU_SWIGLU, wf, wf2, ops.F16, 131, 768, xv).abs().sum())
mark("f16 matvec")
# quantised matvec epilogues
wq = dec._rand_weight(ops.Q4_K, 300, 512, gen, dev); wq2 = dec._rand_weight(ops.Q4_K, 300, 512, gen, dev)
xq = torch.randn(1, 512, device=dev); rq = torch.randn(1, 300, device=dev)
acc += float(ops.mul_mat_add(wq, ops.Q4_K, 300, 512, xq, rq, torch.empty_like(rq)).abs().sum())
acc += float(ops.mul_mat_glu(ops.GLU_SWIGLU, wq, wq2, ops.Q4_K, 300, 512, xq).abs().sum())
mark("quant matvec epilogues")
# rows in and out of the block formats
for t, bs in ((ops.Q8_0, 34), (ops.Q4_0, 18)):
    src = torch.randn(37, 1024, device=dev); idx = torch.randperm(96, device=dev)[:37].to(torch.int64)
    cache = torch.zeros(96 * 1024 // 32 * bs, dtype=torch.uint8, device=dev)
    ops.set_rows(src, idx, ops.T(cache, t, ne=[1024, 96]))
    back = torch.empty(96, 1024, device=dev); ops.cpy(ops.T(cache, t, ne=[1024, 96]), back); acc += float(back.abs().sum())
    planar = ops.to_planar(t, cache)
    acc += float(ops.get_rows(ops.T(planar, t, ne=[1024, 96], layout=ops.LAYOUT_PLANAR), torch.arange(0, 96, 5, device=dev, dtype=torch.int32)).abs().sum())
    # decode attention over the quantised cache: [D = 128, n_kv = 96, 8 kv-head slices of one 1024-wide row]
    D, HK, n_kv = 128, 8, 96
    desc = ops.T(cache, t, ne=[D, n_kv, HK], nb=[bs, HK * D // 32 * bs, D // 32 * bs, n_kv * HK * D // 32 * bs])
    q = torch.randn(32, 3, D, device=dev); mask = torch.zeros(64, n_kv, device=dev).half(); mask[:, 90:] = float("-inf")
    acc += float(ops.flash_attn(q, desc, desc, mask, 0.088).abs().sum())
    q2 = torch.randn(8, 40, D, device=dev)                         # >= 16 query tokens: F16 staging + the tensor-core kernel
    acc += float(ops.flash_attn(q2, desc, desc, mask, 0.088).abs().sum())
mark("quant rows + FA")
wk = dec._rand_weight(ops.Q4_K, 64, 512, gen, dev)
acc += float(ops.get_rows(ops.T(wk, ops.Q4_K, ne=[512, 64]), torch.tensor([0, 63, 7], device=dev, dtype=torch.int32)).abs().sum())
mark("get_rows q4_K")
# ROPE on F16, Token2Wav ops
h = torch.randn(70, 4, 128, device=dev).half(); pos = torch.arange(70, device=dev, dtype=torch.int32)
acc += float(ops.rope(h, pos, 128, 2).float().abs().sum()) + float(ops.rope(h[:3, :1].contiguous(), pos[:3].contiguous(), 64, 0).float().abs().sum())
a = torch.randn(3, 5, 77, device=dev)
acc += float(ops.concat(a, a, 1).abs().sum()) + float(ops.repeat(a, [6, 5, 154]).abs().sum()) + float(ops.sum_rows(a).abs().sum()) + float(ops.pad_reflect_1d(a, 4, 2).abs().sum())
acc += float(ops.pad(a, [1, 2, 0, 3, 2, 0, 0, 0]).abs().sum()) + float(ops.arange(0.0, 100.0, 0.5, dev).abs().sum()) + float(ops.unary_param(ops.LEAKY_RELU, a, 0.1).abs().sum())
acc += float(ops.conv_transpose_1d((torch.randn(33, 17, 16, device=dev) * 0.1).half(), torch.randn(33, 50, device=dev), 8).abs().sum())
acc += float(ops.conv_transpose_1d(torch.randn(5, 3, 7, device=dev) * 0.1, torch.randn(5, 9, device=dev), 2).abs().sum())
mark("rope + wave")
print("sanitize_r02 run done", acc)

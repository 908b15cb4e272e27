#!/usr/bin/env python
"""Run-to-run reproducibility of the tcgen05 GEMM variants on ragged shapes (split-K with two commuting addends, merged launches, residual epilogue): every call is
made three times on the same inputs and the outputs must be bit-identical; also prints input checksums so that two PROCESSES can be compared."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package
pkg = load_package(); ops, dec = pkg.ops, pkg.decode
dev = torch.device("cuda:0")
torch.manual_seed(0)
gen = torch.Generator(device=dev); gen.manual_seed(0)
k, n = 1024, 600
ws = [(dec._rand_weight(t, m, k, gen, dev), t, m, ops.LAYOUT_PLANAR if t == ops.Q6_K else ops.LAYOUT_NATIVE) for t, m in ((ops.Q4_K, 1100), (ops.Q4_K, 300), (ops.Q6_K, 260))]
x = torch.randn(n, k, device=dev, generator=gen)
print("inputs:", [int(w[0].to(torch.int64).sum()) for w in ws], float(x.double().sum()))
bad = 0
for trial, (kk, nn) in enumerate(((1024, 600), (1024, 2048))):
    xx = torch.randn(nn, kk, device=dev, generator=gen)
    ref = None
    for rep in range(3):
        outs = [torch.full((nn, m), float("nan"), device=dev) for (_, _, m, _) in ws]
        ops.mul_mat_multi(ws, xx, outs)
        torch.cuda.synchronize()
        assert all(torch.isfinite(o).all() for o in outs), "unwritten output elements"
        if ref is None: ref = [o.clone() for o in outs]
        else:
            for i, (a, b) in enumerate(zip(ref, outs)):
                if not torch.equal(a, b): bad += 1; print(f"mul_mat_multi n={nn} segment {i}: rep {rep} differs, max |d| = {float((a - b).abs().max()):.3g}")
    print(f"mul_mat_multi n={nn}: sum |y| = {float(sum(o.double().abs().sum() for o in ref)):.6f}")
    for (w, t, m, lay) in ws:
        r = torch.randn(nn, m, device=dev, generator=gen)
        ref1 = None
        for rep in range(3):
            o = torch.full((nn, m), float("nan"), device=dev)
            ops.mul_mat_add(w, t, m, kk, xx, r, o, layout=lay)
            single = ops.mul_mat(w, t, m, kk, xx, layout=lay)
            torch.cuda.synchronize()
            assert torch.isfinite(o).all() and torch.isfinite(single).all()
            if ref1 is None: ref1 = (o.clone(), single.clone())
            elif not (torch.equal(ref1[0], o) and torch.equal(ref1[1], single)): bad += 1; print(f"mul_mat(_add) m={m} n={nn}: rep {rep} differs")
print("non-reproducible calls:", bad)

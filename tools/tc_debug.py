import sys, numpy as np, torch
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from __graft_entry__ import load_package
import oracle_c as O
ops = load_package().ops
name, m, k, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
t = {"q4_K": O.Q4_K, "q6_K": O.Q6_K}[name]
rng = np.random.default_rng(1)
bs = O.BLOCK[t][1]
b = rng.integers(0, 256, (m * k // 256, bs), dtype=np.uint8)
def h(nn, lo=1e-3, hi=1e-2): return rng.uniform(lo, hi, nn).astype(np.float16).view(np.uint8).reshape(nn, 2)
if t == O.Q4_K: b[:, 0:2] = h(len(b)); b[:, 2:4] = h(len(b))
else: b[:, 208:210] = h(len(b), 1e-4, 1e-3)
x = rng.standard_normal((n, k)).astype(np.float32)
wd = torch.from_numpy(b).cuda()
planar = t == O.Q6_K
if planar: wd = ops.to_planar(t, wd)
wf = torch.from_numpy(O.dequant(t, b, k)).cuda().double()
exact = (torch.from_numpy(x).cuda().double() @ wf.T)
for rep in range(3):
    got = ops.mul_mat(wd, t, m, k, torch.from_numpy(x).cuda(), layout=ops.LAYOUT_PLANAR if planar else ops.LAYOUT_NATIVE).double()
    torch.cuda.synchronize()
    err = (got - exact).abs()
    scale = exact.abs().mean().item()
    bad = err > 0.02 * scale + 1e-9
    print(f"rep {rep}: nmse {((got-exact)**2).sum().item()/(exact**2).sum().item():.3e}  bad elems {int(bad.sum())} of {bad.numel()}  max err/scale {err.max().item()/scale:.3f}")
    if bad.any():
        cols = bad.any(1).nonzero().flatten().cpu().numpy(); rows = bad.any(0).nonzero().flatten().cpu().numpy()
        print("  bad cols (tokens):", len(cols), cols[:20], "...", cols[-5:])
        print("  bad rows (weights):", len(rows), rows[:20], "...", rows[-5:])
        for r_ in rows[:6]:
            bc = bad[:, r_].nonzero().flatten().cpu().numpy()
            print(f"   row {r_}: {len(bc)} bad cols, first {bc[:4]}, last {bc[-3:]}, in nt0 {int((bc < 256).sum())} nt1 {int((bc >= 256).sum())}")
        mt = np.unique(rows // 128); print("  m-tiles with errors:", mt[:40], " rows within tile:", np.unique(rows % 128)[:40])

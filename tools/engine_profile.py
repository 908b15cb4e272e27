#!/usr/bin/env python
"""Per-phase timeline of the persistent decode engine (device timestamps of CTA 0): python tools/engine_profile.py [--depth D]"""
import argparse, ctypes as C, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package
pkg = load_package(); ops, dec = pkg.ops, pkg.decode
ap = argparse.ArgumentParser(); ap.add_argument("--depth", type=int, default=2048); ap.add_argument("--layers", type=int, default=36)
a = ap.parse_args()
cfg = dec.LLMConfig(n_layer=a.layers)
D = dec.Qwen3Decoder(cfg, "cuda:0")
n_kv = (a.depth + 1 + 255) // 256 * 256
for lw in D.L:
    lw["k_cache"][:a.depth].normal_(0, 0.5); lw["v_cache"][:a.depth].normal_(0, 1.0)
hi = dec.Qwen3Decoder.host_inputs(cfg, a.depth, n_kv, pinned=False)
D.x_in.normal_(0, 0.05); D.pos.copy_(hi["pos"]); D.kv_idx.copy_(hi["kv_idx"]); D.mask_f32[:, :n_kv].copy_(hi["mask"])
D.build_engine()
for _ in range(3):
    D.step_engine(n_kv)
torch.cuda.synchronize()
L = ops.lib()
n = L.b200_decoder_n_phases(D._engine)
buf = np.zeros((n, 8), np.uint64)
ops.check(L.b200_decoder_profile(D._engine, n_kv, buf.ctypes.data_as(C.c_void_p), ops.stream()))
t = buf.astype(np.int64)
names = ["A qkv", "B attn", "C wo", "D gate/up", "E down"]
tot = (t[-1, 2] - t[0, 0]) / 1e3
print(f"token: {tot:.1f} us over {n} phases (n_kv={n_kv})")
rows = {}
for p in range(n):
    nm = names[p % 5] if p < n - 1 else "Z lm_head"
    pro, work, bar = (t[p, 1] - t[p, 0]) / 1e3, (t[p, 2] - t[p, 1]) / 1e3, (t[p, 3] - t[p, 2]) / 1e3
    rows.setdefault(nm, []).append((pro, work, bar))
print(f"{'phase':12s} {'n':>4s} {'prologue':>9s} {'work':>9s} {'barrier':>9s} {'total/phase':>12s} {'sum':>9s}  (us, CTA 0)")
for nm, v in rows.items():
    v = np.array(v)
    m = v.mean(0)
    print(f"{nm:12s} {len(v):4d} {m[0]:9.2f} {m[1]:9.2f} {m[2]:9.2f} {m.sum():12.2f} {v.sum():9.1f}")
# finer stamps (slots 4..7), relative to the phase start, averaged over layers
for k, nm in enumerate(names):
    sel = [p for p in range(n - 1) if p % 5 == k]
    d = np.array([[(t[p, s] - t[p, 0]) / 1e3 if t[p, s] else np.nan for s in (4, 5, 6, 7, 1, 2)] for p in sel])
    print(f"  {nm:10s} stamps +us: s4={np.nanmean(d[:,0]):.2f} s5={np.nanmean(d[:,1]):.2f} s6={np.nanmean(d[:,2]):.2f} s7={np.nanmean(d[:,3]):.2f} | prologue_end={np.nanmean(d[:,4]):.2f} work_end={np.nanmean(d[:,5]):.2f}")

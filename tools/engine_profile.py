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
nsm = L.b200_device_sm_count(0)
buf = np.zeros((n, nsm, 8), np.uint64)
ops.check(L.b200_decoder_profile(D._engine, n_kv, buf.ctypes.data_as(C.c_void_p), ops.stream()))
t = buf.astype(np.int64)
names = ["A qkv", "B attn", "C wo", "D gate/up", "E down"]
print(f"token: {(t[-1, :, 2].max() - t[0, :, 0].min()) / 1e3:.1f} us over {n} phases (n_kv={n_kv}), B200_SD_FLAGS={__import__('os').environ.get('B200_SD_FLAGS', '0')}")
print(f"{'phase':10s} {'n':>3s} | {'pro mean':>8s} {'pro max':>8s} | {'work mean':>9s} {'work max':>8s} | {'max-med arr':>11s} {'bar exit lat':>12s} | {'phase total':>11s}   (us, over all CTAs)")
rows = {}
for p in range(n):
    nm = names[p % 5] if p < n - 1 else "Z lm_head"
    act = t[p, :, 0] > 0
    pro, work = (t[p, act, 1] - t[p, act, 0]) / 1e3, (t[p, act, 2] - t[p, act, 1]) / 1e3
    skew = (t[p, act, 2].max() - np.median(t[p, act, 2])) / 1e3
    exitlat = (t[p, act, 3].min() - t[p, act, 2].max()) / 1e3
    total = (t[p, act, 3].max() - t[p, act, 0].min()) / 1e3
    rows.setdefault(nm, []).append((pro.mean(), pro.max(), work.mean(), work.max(), skew, exitlat, total))
for nm, v in rows.items():
    m = np.array(v).mean(0)
    print(f"{nm:10s} {len(v):3d} | {m[0]:8.2f} {m[1]:8.2f} | {m[2]:9.2f} {m[3]:8.2f} | {m[4]:11.2f} {m[5]:12.2f} | {m[6]:11.2f}")
for k, nm in enumerate(names):
    sel = [p for p in range(n - 1) if p % 5 == k]
    d = np.array([[np.mean([(t[p, c, s] - t[p, c, 0]) / 1e3 for c in range(nsm) if t[p, c, s]] or [np.nan]) for s in (4, 5, 6, 7, 1, 2)] for p in sel])
    if k != 1:
        wt = np.mean([t[p, :, 7].mean() for p in sel]) / 1965.0
        print(f"  {nm:10s} warp 0 waited for weights {wt:.2f} us of its work time")
        d[:, 3] = np.nan
    print(f"  {nm:10s} stamps +us (mean over CTAs): s4={np.nanmean(d[:,0]):.2f} s5={np.nanmean(d[:,1]):.2f} s6={np.nanmean(d[:,2]):.2f} s7={np.nanmean(d[:,3]):.2f} | prologue_end={np.nanmean(d[:,4]):.2f} work_end={np.nanmean(d[:,5]):.2f}")

#!/usr/bin/env python
"""Small invocations of every new kernel for compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize.py"""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from __graft_entry__ import load_package
pkg = load_package(); ops, dec = pkg.ops, pkg.decode
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev); gen.manual_seed(0)
rng = np.random.default_rng(0)
# tcgen05 GEMMs: ragged m, split-K and 128-column variants, every type
for wt, m, k, n in ((ops.Q4_K, 200, 512, 17), (ops.Q6_K, 129, 1024, 300), (ops.Q4_K, 1024, 2048, 130)):
    w = dec._rand_weight(wt, m, k, gen, dev)
    y = ops.mul_mat(w, wt, m, k, torch.randn(n, k, device=dev), layout=ops.LAYOUT_PLANAR if wt == ops.Q6_K else ops.LAYOUT_NATIVE)
for wt, bs in ((ops.Q5_K, 176), (ops.Q8_0, 34), (ops.Q4_0, 18)):
    m, k, n = 130, 512, 40
    blk = ops.BLOCK[wt][0]
    raw = torch.randint(0, 256, (m * k // blk, bs), dtype=torch.uint8, generator=gen, device=dev)
    raw[:, 0:2] = torch.tensor([0x66, 0x1e], dtype=torch.uint8, device=dev)      # f16 d ~ 6e-3
    if wt == ops.Q5_K: raw[:, 2:4] = torch.tensor([0x66, 0x1e], dtype=torch.uint8, device=dev)
    w = raw.reshape(-1)
    layout = ops.LAYOUT_NATIVE
    if wt in ops.PAYLOAD: w, layout = ops.to_planar(wt, w), ops.LAYOUT_PLANAR
    y = ops.mul_mat(w, wt, m, k, torch.randn(n, k, device=dev), layout=layout)
wf = (torch.randn(300, 256, device=dev) * 0.05).half()
y = ops.mul_mat(wf, ops.F16, 300, 256, torch.randn(77, 256, device=dev), w_ne=[256, 300])
# attention prefill, ragged
n_q, n_kv, D, H, HK = 70, 200, 128, 4, 1
q = torch.randn(n_q, H, D, device=dev); kk = torch.randn(HK, n_kv, D, device=dev).half(); vv = torch.randn(HK, n_kv, D, device=dev).half()
mask = torch.zeros(128, n_kv, device=dev).half(); mask[:, 150:] = float("-inf")
o = ops.flash_attn(q.permute(1, 0, 2), kk, vv, mask, 0.088)
# decode engine, two tokens
cfg = dec.LLMConfig(name="san", n_embd=2048, n_layer=2, n_head=8, n_head_kv=2, n_ff=4096, n_vocab=4096, n_ctx=512)
B = dec.Qwen3Decoder(cfg, dev, seed=1)
B.build_engine()
for step in range(2):
    hi = dec.Qwen3Decoder.host_inputs(cfg, 100 + step, 256, pinned=False)
    B.x_in.normal_(0, 0.05); B.pos.copy_(hi["pos"]); B.kv_idx.copy_(hi["kv_idx"]); B.mask_f32[:, :256].copy_(hi["mask"])
    B.step_engine(256)
torch.cuda.synchronize()
print("sanitize run done", float(y.abs().sum()), float(o.abs().sum()), float(B.logits.abs().sum()))

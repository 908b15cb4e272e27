#!/usr/bin/env python
"""Prefill GEMM (k_mmq_tc, tcgen05) timing through the C-ABI at the Qwen3-8B layer shapes: python tools/prefill_bench.py [n_tokens]
Prints TFLOP/s per shape and the linear-layer time of a whole 36-layer prompt vs MEASURED_PEAKS.json bf16_tflops (sustained)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
ops, dec = pkg.ops, pkg.decode
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2048
pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
PEAK = pk.get("bf16_tflops_sustained", 1400.0)
E, F = 4096, 12288
shapes = [("wq/wo q4_K", ops.Q4_K, 4096, E, 2), ("wk q4_K", ops.Q4_K, 1024, E, 1), ("wv q6_K", ops.Q6_K, 1024, E, 1),
          ("gate/up q4_K", ops.Q4_K, F, E, 2), ("down q6_K", ops.Q6_K, E, F, 0.5), ("down q4_K", ops.Q4_K, E, F, 0.5)]
if "--f16" in sys.argv:                                   # the F16 model of BASELINE.json configs[4]: k_mm_f16_tc
    shapes = [("wq/wo f16", ops.F16, 4096, E, 2), ("wk/wv f16", ops.F16, 1024, E, 2), ("gate/up f16", ops.F16, F, E, 2), ("down f16", ops.F16, E, F, 1)]
gen = torch.Generator(device=dev)
gen.manual_seed(0)
total_us, total_flop = 0.0, 0.0
out = {}
for name, wt, m, k, per_layer in shapes:
    w = dec._rand_weight(wt, m, k, gen, dev) if wt != ops.F16 else (torch.randn(m, k, device=dev) * 0.02).half()
    x = torch.randn(n, k, device=dev)
    y = torch.empty(n, m, device=dev)
    layout = ops.LAYOUT_PLANAR if wt == ops.Q6_K else ops.LAYOUT_NATIVE
    import ctypes as C
    L = ops.lib()
    wd, xd, od = ops.T(w, wt, ne=[k, m], layout=layout), ops.T(x), ops.T(y)
    sb = L.b200_mul_mat_scratch_bytes(C.byref(wd), C.byref(xd))
    scratch = torch.empty(max(sb, 16), dtype=torch.uint8, device=dev)

    def call():
        ops.check(L.b200_mul_mat(C.byref(wd), C.byref(xd), C.byref(od), C.c_void_p(scratch.data_ptr()), C.c_size_t(sb), ops.stream()))
    iters = 10
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            call()
        st.synchronize()
        g = torch.cuda.CUDAGraph()                       # the launches of `iters` calls replayed as ONE graph: no host time in the measurement
        with torch.cuda.graph(g, stream=st):
            for _ in range(iters):
                call()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            g.replay()
            e1.record(st)
            st.synchronize()
            best = min(best, e0.elapsed_time(e1))
    us = best / iters * 1e3
    fl = 2.0 * m * k * n
    out[name] = {"us": round(us, 1), "tflops": round(fl / us / 1e6, 1)}
    print(f"{name:14s} m={m:6d} k={k:6d} n={n:5d}: {us:9.1f} us  {fl / us / 1e6:8.1f} TFLOP/s  ({fl / us / 1e6 / PEAK:.1%} of measured sustained {PEAK})")
    total_us += us * per_layer
    total_flop += fl * per_layer
layer_ms = total_us / 1e3
print(json.dumps({"n_tokens": n, "linear_ms_per_layer": round(layer_ms, 3), "linear_tok_per_s_36_layers": round(n / (36 * layer_ms / 1e3), 1),
                  "linear_tflops": round(total_flop / total_us / 1e6, 1), "frac_of_measured_sustained": round(total_flop / total_us / 1e6 / PEAK, 4),
                  "shapes": out}))

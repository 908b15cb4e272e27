#!/usr/bin/env python
"""Per-kernel timing through the C-ABI: each decode-path launch at the Qwen3-8B shapes, CUDA-event timed over rotating weight
copies larger than L2 (126 MB) so every launch streams from HBM.  Prints GB/s vs MEASURED_PEAKS.json."""
import ctypes as C
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
ops, dec = pkg.ops, pkg.decode
P = C.c_void_p
dev = torch.device("cuda:0")
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def timeit(fn, iters=40, warm=3, reps=5):
    """GPU time per call: `iters` calls captured into ONE CUDA graph (no host launch overhead), best of `reps` replays."""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(warm):
            fn(i)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for i in range(iters):
                fn(i)
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            g.replay()
            e1.record(st)
            st.synchronize()
            best = min(best, e0.elapsed_time(e1))
    return best / iters * 1e3          # us


def bench_matvec(name, wtype, shapes, swiglu=False, residual=False):
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    k = shapes[0][1]
    nbytes = sum(m * ops.row_size(wtype, k) for m, _ in shapes)
    ncopies = max(2, int(300e6 // nbytes) + 1)
    copies = [[dec._rand_weight(wtype, m, k, gen, dev) for m, _ in shapes] for _ in range(ncopies)]
    x = torch.randn(1, k, device=dev)
    act = ops.quantize_act(wtype, x)
    ys = [torch.zeros(m, device=dev) for m, _ in shapes]
    res = torch.zeros(shapes[0][0], device=dev)
    layout = ops.LAYOUT_PLANAR if wtype == ops.Q6_K else ops.LAYOUT_NATIVE

    def fn(i):
        ws = copies[i % ncopies]
        jobs = [ops.make_job(w, wtype, m, k, y, res if residual else None, layout) for w, (m, _), y in zip(ws, shapes, ys)]
        if swiglu:
            ops.matvec_q_swiglu(jobs[0], jobs[1], ys[0], act, k)
        else:
            ops.matvec_q(jobs, act, k)
    us = timeit(fn)
    print(f"{name:34s} {nbytes / 1e6:8.2f} MB  {us:8.2f} us  {nbytes / us / 1e3:8.1f} GB/s  {nbytes / us / 1e3 / PEAK:6.1%}")


def main():
    E, F, Q, KV = 4096, 12288, 4096, 1024
    bench_matvec("qkv q4_K (3 jobs)", ops.Q4_K, [(Q, E), (KV, E), (KV, E)])
    bench_matvec("q,k q4_K (2 jobs)", ops.Q4_K, [(Q, E), (KV, E)])
    bench_matvec("v q6_K", ops.Q6_K, [(KV, E)])
    bench_matvec("wo q4_K + residual", ops.Q4_K, [(E, Q)], residual=True)
    bench_matvec("gate/up q4_K + swiglu", ops.Q4_K, [(F, E), (F, E)], swiglu=True)
    bench_matvec("down q4_K + residual", ops.Q4_K, [(E, F)], residual=True)
    bench_matvec("down q6_K + residual", ops.Q6_K, [(E, F)], residual=True)
    bench_matvec("lm_head q6_K", ops.Q6_K, [(151748, E)])
    # small ops
    x = torch.randn(1, E, device=dev)
    w = torch.ones(E, device=dev)
    act = torch.zeros(ops.lib().b200_act_bytes(ops.Q4_K, F), dtype=torch.uint8, device=dev)
    L = ops.lib()
    st = ops.stream()
    us = timeit(lambda i: L.b200_rms_norm_quantize(P(x.data_ptr()), C.c_int64(E), P(w.data_ptr()), None, C.c_int64(E), P(act.data_ptr()),
                                                   ops.Q4_K, C.c_int64(E), C.c_int64(1), C.c_float(1e-6), st))
    print(f"{'rms_norm_quantize 4096':34s} {us:8.2f} us")
    h = torch.randn(1, F, device=dev)
    us = timeit(lambda i: L.b200_quantize_act(ops.Q4_K, P(h.data_ptr()), C.c_int64(F), P(act.data_ptr()), C.c_int64(F), C.c_int64(1), st))
    print(f"{'quantize_act 12288':34s} {us:8.2f} us")
    for n_kv in (256, 2304, 4096):
        cfg = dec.LLMConfig()
        D = 128
        nl = 8
        kc = [torch.randn(4096, 1024, device=dev).half() for _ in range(nl)]
        vc = [torch.randn(4096, 1024, device=dev).half() for _ in range(nl)]
        q = torch.randn(1, 32, D, device=dev)
        mask = torch.zeros(64, n_kv, dtype=torch.float16, device=dev)
        out = torch.zeros(1, 32, D, device=dev)
        scratch = torch.zeros(8 << 20, dtype=torch.uint8, device=dev)

        def fa(i):
            kv = kc[i % nl][:n_kv].view(n_kv, 8, D).permute(1, 0, 2)
            vv = vc[i % nl][:n_kv].view(n_kv, 8, D).permute(1, 0, 2)
            ops.flash_attn(q.permute(1, 0, 2), kv, vv, mask, 0.088, out=out, scratch=scratch)
        us = timeit(fa)
        nbytes = 2 * n_kv * 1024 * 2
        print(f"{'flash_attn n_kv=%d' % n_kv:34s} {nbytes / 1e6:8.2f} MB  {us:8.2f} us  {nbytes / us / 1e3:8.1f} GB/s  {nbytes / us / 1e3 / PEAK:6.1%}")


if __name__ == "__main__":
    main()

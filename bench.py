#!/usr/bin/env python
"""bench.py — BASELINE.json metric: Qwen3-8B (MiniCPM-o-4.5 LLM) Q4_K_M batch-1 decode tok/s on B200, with the HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--depth D] [--impl b200|reference]

A "step" = one decoded token = one pass of the hot path (36 layers + lm_head) over synthetic weights of the named shape
(BASELINE.json configs[1]; ctx = 4096, KV depth D, default 2048 = the mean depth of a 4096-token generation).
  value : tok/s, inputs resident in HBM, the step replayed as ONE CUDA graph, CUDA-event timed on the launching stream.
  e2e   : tok/s through the drop-in boundary: the UNMODIFIED reference llama-bench (oracle/_ref/bin) with libggml-b200.so loaded through
          GGML_BACKEND_PATH on a synthetic 36-layer GGUF at the same KV depth — graph build, scheduler, host input copies (ggml-backend.cpp:1435-1442),
          our backend, logits read-back (llama-context.cpp:1144), all inside its timed region.  `e2e_cabi` is the same token through the C-ABI
          with pinned host buffers from ctypes (--no-plugin-e2e reports only that one).
  roofline : algorithmic bytes/token (weights + KV read/write) / measured time vs MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference : the reference's own ggml CPU backend (oracle/_ref, built from /root/reference by oracle/Makefile)
          on a bounded sample of the same workload (whole layers of the same shapes/types), all host threads.
N > 1: layer-split pipeline (LLAMA_SPLIT_MODE_LAYER, src/llama-model.cpp:2130-2185): rank r owns a contiguous layer range and its KV
cache; the only exchange is one NCCL send/recv of the 16 KiB hidden state per boundary.  N independent decode streams are kept in
flight (one per stage) so every GPU works every step; value = tokens of all streams / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


REF_BIN = ROOT / "oracle" / "_ref" / "bin" / "llama-bench"
PLUGIN = ROOT / "llama.cpp-omni_b200" / "lib" / "libggml-b200.so"
GGUF = Path(os.environ.get("B200_BENCH_GGUF", "/tmp/b200_bench_qwen3_8b_q4_k_m.gguf"))


def bench_gguf() -> Path:
    """The synthetic Qwen3-8B-shaped Q4_K_M GGUF both arms decode (tools/make_gguf.py: random valid quant blocks, the tensor names / types of a
    llama-quantize file; 5.0 GB, ~20 s to write).  Generated once per box."""
    if not GGUF.exists() or GGUF.stat().st_size < 5_000_000_000:
        tmp = GGUF.with_suffix(".tmp")
        subprocess.check_call([sys.executable, str(ROOT / "tools" / "make_gguf.py"), str(tmp)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        tmp.rename(GGUF)
    return GGUF


def llama_bench(plugin: bool, n_gen: int, depth: int, threads: int, reps: int = 1, timeout: int = 1500) -> dict:
    """The reference's OWN llama-bench (oracle/_ref, built unmodified from /root/reference by oracle/Makefile), either on its ggml CPU backend
    (plugin=False, -ngl 0) or with libggml-b200.so loaded through GGML_BACKEND_PATH exactly as a llama.cpp-omni user would (-ngl 99): the whole
    decode path — graph build, scheduler, input copies from host memory, our backend, logits back to the host — on the same GGUF."""
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = f"{ROOT / 'oracle' / '_ref' / 'lib'}:{PLUGIN.parent}:" + env.get("LD_LIBRARY_PATH", "")
    if plugin:
        env["GGML_BACKEND_PATH"] = str(PLUGIN)
    else:
        env.pop("GGML_BACKEND_PATH", None)
    cmd = [str(REF_BIN), "-m", str(bench_gguf()), "-p", "0", "-n", str(n_gen), "-d", str(depth), "-fa", "1", "-ngl", "99" if plugin else "0",
           "-t", str(threads), "-r", str(reps), "-o", "json"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    if out.returncode != 0:
        raise RuntimeError(f"llama-bench failed ({out.returncode}): {out.stderr[-600:]}")
    rows = json.loads(out.stdout)
    r = [x for x in rows if x.get("n_gen", 0) == n_gen][-1]
    return {"tok_s": float(r["avg_ts"]), "ms": 1e3 / float(r["avg_ts"]), "backends": r.get("backends"), "n_gen": n_gen, "depth": depth, "reps": reps}


OMNI_BIN = ROOT / "oracle" / "_ref" / "bin" / "omni_encoders"


def omni_encoders_leg(threads: int) -> dict:
    """BASELINE.json configs[3], the encoder half ("APM 1 s audio chunk + VPM frame"): the reference's UNMODIFIED tools/omni/audition.cpp and vision.cpp
    (oracle/_ref/bin/omni_encoders = tests/native/omni_encoders.cpp around them) on full-size synthetic Whisper / SigLip GGUFs (tools/make_omni_gguf.py), the reference CPU
    backend and the plugin in ONE process: milliseconds per 1 s chunk / per 448 x 448 frame on both, and how far the plugin's embeddings are from the CPU backend's.  The first
    chunk / frame (first-launch costs) is left out of the medians.  An extra object of the JSON line; never the headline metric."""
    if not OMNI_BIN.exists() or not PLUGIN.exists():
        return {"unavailable": "oracle/_ref/bin/omni_encoders or the plugin is not built"}
    tmpdir = Path(os.environ.get("TMPDIR", "/tmp"))
    # B200_BENCH_OMNI_SMALL=1: 2-layer encoders (plumbing checks of this leg on a box without a GPU: tests/test_cabi_exports.py); never a bench number
    small = os.environ.get("B200_BENCH_OMNI_SMALL") == "1"
    shrink = {"apm": ["--layers", "2", "--d-model", "256", "--heads", "4", "--proj", "512"], "vpm": ["--layers", "2", "--embd", "288", "--heads", "4", "--ff", "512", "--proj", "512"]}
    files, makers = {}, []
    for what, size in (("apm", 600_000_000), ("vpm", 1_000_000_000)):
        f = files[what] = tmpdir / f"b200_bench_omni_{what}{'_small' if small else ''}.gguf"
        if not f.exists() or f.stat().st_size < (1_000_000 if small else size):
            tmp = f.with_suffix(".tmp")
            makers.append((subprocess.Popen([sys.executable, str(ROOT / "tools" / "make_omni_gguf.py"), what, str(tmp)] + (shrink[what] if small else []),
                                            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL), tmp, f))
    for proc, tmp, f in makers:
        if proc.wait(timeout=240) != 0:
            return {"error": f"tools/make_omni_gguf.py failed for {f.name}"}
        tmp.rename(f)
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = f"{ROOT / 'oracle' / '_ref' / 'lib'}:{PLUGIN.parent}:" + env.get("LD_LIBRARY_PATH", "")
    env["GGML_BACKEND_PATH"] = str(PLUGIN)
    med = lambda v: sorted(v)[len(v) // 2]
    out = {"path": "oracle/_ref/bin/omni_encoders: audition_audio_encode / vision_image_encode of the unmodified reference, use_gpu = false (ggml CPU backend, "
                   f"{threads} threads) vs use_gpu = true (libggml-b200.so through GGML_BACKEND_PATH) in one process; synthetic F16 GGUFs; medians without the first call"}
    for key, what, n in (("apm_1s_audio_chunk", "apm", 9), ("vpm_448x448_frame", "vpm", 3)):
        r = subprocess.run([str(OMNI_BIN), what, str(files[what]), str(n), str(threads)], env=env, capture_output=True, text=True, timeout=180)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        if r.returncode != 0 or not lines:
            out[key] = {"error": (r.stderr or r.stdout)[-300:]}
            continue
        res = json.loads(lines[-1])
        if "error" in res:
            out[key] = res
            continue
        on_plugin = "devices:" in r.stderr and "B200" in r.stderr.split("devices:")[-1].splitlines()[0]
        if not on_plugin:                                    # (no GPU-type device: the harness's second side fell back to the CPU backend — not a plugin number)
            out[key] = {"error": "no B200 device was registered in the harness process", "reference_cpu_ms": round(med(res["ms_cpu"][1:]), 3)}
            continue
        out[key] = {"ms": round(med(res["ms_gpu"][1:]), 3), "reference_cpu_ms": round(med(res["ms_cpu"][1:]), 3), "calls": n, "max_rel_err_vs_cpu": res["max_rel_err"],
                    "nmse_vs_cpu": res["nmse"], "non_finite": res["non_finite"]}
    return out


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int = 0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                                          str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------ reference arm
def reference_layer_sample(cfg, n_layers: int, threads: int, steps: int, warmup: int) -> dict:
    """The reference CPU backend (oracle/_ref libggml-cpu.so, unmodified) on `n_layers` whole layers' worth of decode matvecs +
    swiglu/norms at the Qwen3-8B shapes; tok/s is extrapolated to the 36-layer + lm_head token by BYTES (the CPU path is as
    bandwidth-bound as ours)."""
    import numpy as np
    import refggml as R
    from __graft_entry__ import load_package
    dec = load_package().decode
    assert R.available(), "oracle/_ref is missing: run `make -C oracle ref` where /root/reference exists"
    rng = np.random.default_rng(0)
    E, F, q, kv = cfg.n_embd, cfg.n_ff, cfg.n_head * cfg.head_dim, cfg.n_head_kv * cfg.head_dim
    shapes = [("wq", q, E), ("wk", kv, E), ("wv", kv, E), ("wo", E, q), ("gate", F, E), ("up", F, E), ("down", E, F)]
    tmap = {12: R.Q4_K, 14: R.Q6_K}
    sample_bytes = 0
    with R.Graph(mem_mb=2048) as g:
        x = g.tensor(R.F32, [E, 1], rng.standard_normal((1, E)).astype(np.float32))
        outs = []
        for il in range(n_layers):
            ty = dec.layer_types(cfg, il)
            for name, m, k in shapes:
                t = tmap[ty[name]]
                nbytes = m * R.row_size(t, k)
                raw = rng.integers(0, 256, nbytes, dtype=np.uint8)
                blk = raw.reshape(-1, R.BLOCK[t][1])
                sc = np.float16(rng.uniform(2e-5, 2e-4, blk.shape[0])).view(np.uint8).reshape(-1, 2)
                if t == R.Q4_K:
                    blk[:, 0:2], blk[:, 2:4] = sc, sc
                else:
                    blk[:, 208:210] = sc
                w = g.tensor(t, [k, m], raw)
                xin = x if k == E else g.tensor(R.F32, [k, 1], rng.standard_normal((1, k)).astype(np.float32))
                outs.append(g.op("ggml_mul_mat", w, xin))
                sample_bytes += nbytes
        gr = g.base.ggml_new_graph(g.ctx)
        for o in outs:
            g.base.ggml_build_forward_expand(gr, o)
        for _ in range(warmup):
            assert g.cpu.ggml_graph_compute_with_ctx(g.ctx, gr, threads) == 0
        t0 = time.perf_counter()
        for _ in range(steps):
            assert g.cpu.ggml_graph_compute_with_ctx(g.ctx, gr, threads) == 0
        dt = (time.perf_counter() - t0) / steps
    full = dec.weight_bytes_per_token(cfg)
    ms_token = dt * 1e3 * full / sample_bytes
    return {"ms_per_step": ms_token, "value": 1e3 / ms_token, "sample_ms": dt * 1e3, "sample_bytes": sample_bytes,
            "sample": f"{n_layers} of {cfg.n_layer} layers (7 weight matvecs each, {sample_bytes / 1e6:.0f} MB of Q4_K/Q6_K rows) per step on the "
                      f"reference ggml CPU backend, {threads} threads; extrapolated to one token by weight bytes ({full / 1e9:.3f} GB)"}


def run_reference(args) -> None:
    """--impl reference: the reference's own CPU implementation of the path — its llama-bench on its ggml CPU backend, all host threads, WHOLE decoded
    tokens of the same GGUF at the same KV depth as the B200 arm (same config: nothing is extrapolated).  Bounded: `steps` tokens per run."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from __graft_entry__ import load_package
    cfg = load_package().decode.LLMConfig()
    threads = os.cpu_count() or 1
    depth = max(0, min(args.depth, cfg.n_ctx - 2))
    n_gen = max(4, min(args.steps, 64))
    if REF_BIN.exists():
        r = llama_bench(False, n_gen, depth, threads)
        value, ms = r["tok_s"], r["ms"]
        sample = (f"{n_gen} whole decoded tokens (36 layers + lm_head, attention at KV depth {depth}) by the reference's llama-bench on its ggml CPU backend "
                  f"(oracle/_ref, AVX2 build), {threads} threads, after its own 1-token warm-up and a {depth}-token prompt")
    else:                                                    # reference binaries not built on this box: matvec-only layer sample, extrapolated by bytes
        r = reference_layer_sample(cfg, 2, threads, max(1, args.steps), max(1, args.warmup))
        value, ms, sample = r["value"], r["ms_per_step"], r["sample"]
    line = {"impl": "reference", "metric": "decode_tok_per_s", "value": round(value, 3), "unit": "tok/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int8 x int4/int6 -> int32 -> f32 (q8_K activations), f16 KV", "data": "synthetic",
            "config": {"workload": f"{cfg.name} batch=1 decode, ctx={cfg.n_ctx}, KV depth {depth}", "path": "oracle/_ref/bin/llama-bench -ngl 0 (reference CPU backend)",
                       "timing": "llama-bench's own clock around each llama_decode + synchronize; weights (5 GB) exceed the CPU caches"},
            "cpu_baseline": {"value": round(value, 3), "unit": "tok/s", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": round(value, 3), "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    pkg = load_package()
    ops, dec = pkg.ops, pkg.decode
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.lib()
    cfg = dec.LLMConfig.tiny() if args.tiny else dec.LLMConfig()
    # contiguous layer ranges balanced by the bytes a token streams per stage, lm_head with the last stage (llama.cpp-omni_b200/pipeline.py)
    # stage cost = streamed bytes + the fixed latency of the layer's 5 dependent phase transitions, expressed in bytes of streaming time
    # (measured, profiles/README.md: a layer takes ~54 us of which ~18 us is its bytes at the HBM peak; lm_head 87 us of which 79 us is bytes)
    peak_bps = peaks()[0] * 1e9
    layer_cost = [dec.weight_bytes_per_token(cfg, range(i, i + 1), with_head=False) + int(36e-6 * peak_bps) for i in range(cfg.n_layer)]
    head_cost = dec.weight_bytes_per_token(cfg, range(0), with_head=True) + int(8e-6 * peak_bps)
    ranges = pkg.pipeline.partition_layers(layer_cost, head_cost, world)
    pipe = pkg.pipeline.Pipeline(rank, world, ranges, dist)
    layers = ranges[rank]
    last = rank == pipe.last
    D = dec.Qwen3Decoder(cfg, dev, layers=layers, has_head=last, seed=rank)
    # decode positions depth, depth+1, ...; a run longer than the context wraps back to `depth` (same n_kv, same bytes per step)
    depth = max(0, min(args.depth, cfg.n_ctx - 2))
    span = max(1, min(args.steps + args.warmup + 1, cfg.n_ctx - depth - 1))
    n_kv = min(cfg.n_ctx, (depth + span + 255) // 256 * 256)
    # pre-fill the KV cache up to `depth` so attention reads real (finite) rows
    for lw in D.L:
        lw["k_cache"][:depth].normal_(0, 0.5)
        lw["v_cache"][:depth].normal_(0, 1.0)
    E = cfg.n_embd
    n_rows = min(args.steps + args.warmup + 1, 4096)
    host_embd = (torch.randn(n_rows, E) * 0.05).pin_memory()     # rows of token_embd gathered on the host
    host_logits = torch.empty(cfg.n_vocab).pin_memory()
    stream = torch.cuda.Stream(device=dev)
    n_streams = world                                      # decode streams in flight across the pipeline

    # what llama's set_inputs writes per token (pos, KV cell index, KQ mask), prepared ONCE in pinned memory for every position of the run: building
    # and pinning them per token cost ~0.2 ms of host time per stream-item and was what bounded the N > 1 numbers of round 1
    host_in = [dec.Qwen3Decoder.host_inputs(cfg, depth + j, n_kv) for j in range(span)]

    def set_inputs(i: int):
        hi = host_in[i % span]
        D.pos.copy_(hi["pos"], non_blocking=True)
        D.kv_idx.copy_(hi["kv_idx"], non_blocking=True)
        D.mask_f32[:, :n_kv].copy_(hi["mask"], non_blocking=True)
        return hi["pos"].numel() * 4 + 8 + hi["mask"].numel() * 4

    engine = not args.per_op
    with torch.cuda.stream(stream):
        set_inputs(0)
        D.x_in.copy_(host_embd[0], non_blocking=True)
        if engine:
            D.build_engine()
        D.step(n_kv, engine)                               # eager once (module load, attribute setup)
        stream.synchronize()
        # N > 1: the stage hop runs ON THE DEVICE (llama.cpp-omni_b200/pipeline.py PeerHop, csrc/hop.cu): wait-for-slot -> copy-in -> stage -> ack -> peer
        # write + release are kernels of the SAME captured graph, one graph per slot; `--hop nccl` keeps the host-issued NCCL send/recv for comparison
        hop = None
        if world > 1 and args.hop == "peer":
            try:
                hop = pkg.pipeline.PeerHop(pipe, E, ops, torch, dist, dev, n_slots=2)
            except Exception as ex:                                  # no IPC / peer access on this box: host-issued NCCL send/recv
                print(f"[bench] peer hop unavailable ({ex}); using NCCL send/recv", file=sys.stderr)
                hop = None
        graphs = []
        for slot in range(2 if hop is not None else 1):
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_, stream=stream):
                launches = D.step(n_kv, engine) if hop is None else hop.enqueue_recv(slot, D.x_in) + D.step(n_kv, engine)
                if hop is not None:
                    launches += hop.enqueue_send(slot, D.x_out)
            graphs.append(g_)
        graph = graphs[0]
    hidden = D.x_in
    item = [0]                                              # stream-items this rank has enqueued (slot = item % 2, identical on every rank)

    def pipeline_step(i: int, e2e: bool) -> tuple[int, int]:
        """One step of every in-flight stream through this rank's stage.  Returns (h2d, d2h) bytes when e2e."""
        h2d = d2h = 0
        for _s in range(n_streams):
            if e2e:                                         # (the device-resident leg keeps its inputs in HBM at every N)
                h2d += set_inputs(i)
            if rank == pipe.first and e2e:
                D.x_in.copy_(host_embd[i % n_rows], non_blocking=True)
                h2d += E * 4
            if hop is not None:
                if pipe.is_active:
                    graphs[item[0] % 2].replay()                # wait -> copy-in -> stage -> ack -> send, all on the device
                item[0] += 1
            else:
                pipe.stage_step(hidden, D.x_out, graph.replay)      # recv from the previous stage -> one graph replay -> send to the next
            if last and e2e:
                host_logits.copy_(D.logits, non_blocking=True)
                d2h += cfg.n_vocab * 4
        if e2e and last:
            torch.cuda.current_stream().synchronize()       # the sampler needs the logits before the next token
        return h2d, d2h

    def timed(e2e: bool, steps: int, warmup: int):
        with torch.cuda.stream(stream):
            for i in range(warmup):
                pipeline_step(i, e2e)
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ev0.record(stream)
            hb = db = 0
            for i in range(steps):
                a, b = pipeline_step(warmup + i, e2e)
                hb, db = hb + a, db + b
            ev1.record(stream)
            stream.synchronize()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            if world > 1:
                dist.barrier()
            ms = ev0.elapsed_time(ev1)
        if e2e:
            ms = max(ms, wall)                              # host copies/syncs are part of the end-to-end time
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item(), hb // max(steps, 1), db // max(steps, 1)

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_total, _, _ = timed(False, args.steps, args.warmup)
    clk = clocks.stop() if rank == 0 else None
    ms_e2e, h2d, d2h = timed(True, args.steps, args.warmup)
    if hop is not None:
        if hop.errors():
            raise RuntimeError("peer hop timed out (see csrc/hop.cu error words)")
    e2e_cabi = args.steps * n_streams / (ms_e2e / 1e3)
    e2e_path = "C-ABI with host buffers (pinned H2D inputs, logits D2H per token)"
    plug = None
    if world == 1 and not args.tiny and not args.per_op and not args.no_plugin_e2e and REF_BIN.exists() and PLUGIN.exists():
        # the headline end-to-end number goes THROUGH THE BOUNDARY: the unmodified reference llama-bench with libggml-b200.so as its backend
        torch.cuda.synchronize()
        try:
            plug = llama_bench(True, max(args.steps, 128), depth, os.cpu_count() or 1, reps=2)
            e2e_path = ("oracle/_ref/bin/llama-bench -ngl 99 with GGML_BACKEND_PATH=libggml-b200.so (the reference's graph build, scheduler, host input copies and "
                        f"logits read-back around our backend; {plug['n_gen']} tokens x {plug['reps']} runs at KV depth {depth})")
        except Exception as e:                               # the line must still be printed: e2e then is the C-ABI leg, and says so
            plug = None
            e2e_path += f" [llama-bench + plugin leg failed: {str(e)[:200]}]"
    if world > 1:
        hb = torch.tensor([h2d, d2h], device=dev)
        dist.all_reduce(hb)
        h2d, d2h = int(hb[0]), int(hb[1])

    tokens = args.steps * n_streams
    ms_step = ms_total / args.steps
    value = tokens / (ms_total / 1e3)
    e2e_value = plug["tok_s"] if plug else tokens / (ms_e2e / 1e3)
    if plug:                                                # what the reference's scheduler moves per token (llama-graph.cpp inputs, llama-context.cpp:1144 logits)
        h2d = cfg.n_embd * 4 + 4 + 64 * n_kv * 4 + 8 + 8 + 4
        d2h = cfg.n_vocab * 4
    peak, peak_src = peaks()
    # roofline of the dominant kernel class (the weight matvecs + KV reads = the whole step's algorithmic bytes; one launch = one
    # graph replay = one token through this rank's stage).  At N > 1 every rank moves 1/N of the bytes per stream-step.
    w_bytes = dec.weight_bytes_per_token(cfg, layers, with_head=last)
    kv_bytes = dec.kv_bytes_per_token(cfg, n_kv, layers)
    step_bytes = (w_bytes + kv_bytes) * n_streams
    achieved = step_bytes / (ms_step / 1e3) / 1e9
    traffic = None                                          # dram bytes of one k_stream launch from the committed ncu capture of this workload
    tf = ROOT / "profiles" / "k_stream_traffic.json"
    if engine and world == 1 and tf.exists():
        t = json.loads(tf.read_text())
        if t.get("n_kv") == n_kv and not args.tiny:
            traffic = int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
    line = {"metric": "decode_tok_per_s", "value": round(value, 2), "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int8 x int4/int6 -> int32 -> f32 (q8_K activations), f16 KV", "data": "synthetic",
            "config": {"workload": f"{cfg.name} batch=1 decode, ctx={cfg.n_ctx}, KV depth {depth}..{depth + span - 1} (n_kv={n_kv})",
                       "streams_in_flight": n_streams, "parallelism": "single GPU" if world == 1 else f"layer-split pipeline pp{world}, layers per stage {[len(r) for r in ranges]} (+lm_head on the last)",
                       "hop": None if world == 1 else ("device-side peer write + release flag over NVLink (csrc/hop.cu), captured in the stage's CUDA graph" if hop is not None else "NCCL send/recv issued from the host between graph replays"),
                       "engine": "persistent (1 kernel/token)" if engine else "per-op launches",
                       "timing": f"one CUDA graph per token; weights {w_bytes / 1e9:.2f} GB/rank exceed the 126 MB L2, so no flush is needed"},
            "gpu_launches": launches * args.steps * n_streams,
            "e2e": {"value": round(e2e_value, 2), "unit": "tok/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "path": e2e_path},
            "e2e_cabi": {"value": round(e2e_cabi, 2), "unit": "tok/s", "path": "b200_decoder_step through ctypes with pinned host inputs and logits read-back per token"},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_step": int(step_bytes),
                         "kernel": ("k_stream (persistent decode engine: all weight matvecs + attention of the token in one launch)" if engine else
                                    "k_stream / k_mmvq per-op launches + k_fa_decode") + ": algorithmic weight+KV bytes of one token / graph-replay time"},
            "clocks": clk}
    if rank == 0 and world == 1 and not args.no_prefill and not args.tiny:
        line["prefill"] = prefill_leg(D, cfg, dec, torch, dev, args.prefill_tokens)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            r = None
            if REF_BIN.exists() and not args.tiny:
                try:
                    r = llama_bench(False, 16, 0, threads)
                except Exception as e:
                    print(f"bench.py: reference llama-bench failed, falling back to the layer sample: {str(e)[:200]}", file=sys.stderr)
            if r is not None:
                line["cpu_baseline"] = {"value": round(r["tok_s"], 3), "unit": "tok/s", "cores": threads, "kind": "reference",
                                        "sample": "16 whole decoded tokens (36 layers + lm_head) of the same GGUF by the reference's llama-bench on its ggml CPU backend "
                                                  f"(oracle/_ref), {threads} threads, KV depth 0..15 (bounded sample: `--impl reference` times the full depth)"}
            else:
                r = reference_layer_sample(cfg, 1, threads, 3, 1)
                line["cpu_baseline"] = {"value": round(r["value"], 3), "unit": "tok/s", "cores": threads, "kind": "reference", "sample": r["sample"]}
        if world == 1 and not args.tiny and not args.no_encoders:
            try:
                line["omni_encoders"] = omni_encoders_leg(min(os.cpu_count() or 1, 16))
            except Exception as e:                           # the line must still be printed
                line["omni_encoders"] = {"error": str(e)[:300]}
        print(json.dumps(line))
    if hop is not None:
        hop.close()
    if world > 1:
        dist.destroy_process_group()


def prefill_leg(D, cfg, dec, torch, dev, n_tokens: int) -> dict:
    """Second half of BASELINE.json's metric ("prefill tok/s"): ONE ubatch of n_tokens (configs[2]: 2048) through every layer + the last token's
    lm_head, per-op C-ABI launches exactly as the ggml plugin issues them for an n-token graph; CUDA-event timed, 3 warm-up passes.  The tensor
    roofline: linear-layer FLOPs (2 x tokens x weight elements) + causal attention FLOPs over the whole pass vs MEASURED_PEAKS bf16_tflops_sustained."""
    p = ROOT / "MEASURED_PEAKS.json"
    peak = float(json.loads(p.read_text()).get("bf16_tflops_sustained", 1400.0)) if p.exists() else 1400.0
    n = min(n_tokens, cfg.n_ctx)
    n_kv = (n + 255) // 256 * 256
    x = torch.randn(n, cfg.n_embd, device=dev) * 0.05
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        for _ in range(3):
            _, launches = D.prefill(x, 0, n_kv)
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record(st)
        for _ in range(reps):
            D.prefill(x, 0, n_kv)
        e1.record(st)
        st.synchronize()
    ms = e0.elapsed_time(e1) / reps
    per_layer = 2 * cfg.n_embd * (cfg.n_head + cfg.n_head_kv) * cfg.head_dim + 3 * cfg.n_embd * cfg.n_ff
    flops = 2.0 * n * cfg.n_layer * per_layer + cfg.n_layer * 4.0 * cfg.n_head * cfg.head_dim * n * n / 2
    for lw in D.L:                                   # leave the KV cache as the decode legs expect it
        lw["k_cache"][:n].normal_(0, 0.5)
        lw["v_cache"][:n].normal_(0, 1.0)
    return {"metric": "prefill_tok_per_s", "value": round(n / (ms / 1e3), 1), "unit": "tok/s", "n_tokens": n, "ms": round(ms, 2), "gpu_launches_per_pass": launches,
            "roofline": {"bound": "tensor", "achieved": round(flops / (ms / 1e3) / 1e12, 1), "peak": peak, "unit": "TFLOP/s", "frac": round(flops / (ms / 1e3) / 1e12 / peak, 4),
                         "kernel": "whole pass: k_mmq_tc (tcgen05 dequant-GEMM; q/k/v and gate/up merged launches) + k_fa_tc (tcgen05 attention) + elementwise; algorithmic FLOPs = linear + causal attention (SURVEY.md 8d)"},
            "config": {"workload": f"{cfg.name} prefill n_tokens={n} (one ubatch), empty KV cache"}}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--depth", type=int, default=2048)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tiny", action="store_true", help="tiny model (plumbing checks only; not a bench line)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-prefill", action="store_true", help="skip the prefill leg (the extra `prefill` object of the JSON line)")
    ap.add_argument("--prefill-tokens", type=int, default=2048)
    ap.add_argument("--per-op", action="store_true", help="one launch per (fused) op instead of the persistent decode engine")
    ap.add_argument("--hop", default="peer", choices=["peer", "nccl"], help="N > 1: how the hidden state crosses a stage boundary")
    ap.add_argument("--no-plugin-e2e", action="store_true", help="e2e through the C-ABI only (skip the llama-bench + plugin leg)")
    ap.add_argument("--no-encoders", action="store_true", help="skip the omni-encoder leg (the extra `omni_encoders` object: APM 1 s chunk / VPM frame, BASELINE.json configs[3])")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

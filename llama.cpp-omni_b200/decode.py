"""Host-side mirror of the reference's Qwen3 decode graph, issued as C-ABI calls (include/b200_ops.h).

Mirrors llm_build_qwen3 (src/llama-model.cpp:9287-9406): per layer RMS_NORM*w -> wq/wk/wv -> q/k-norm + RoPE(neox) -> KV write ->
FLASH_ATTN_EXT -> wo + residual -> RMS_NORM*w -> gate/up -> swiglu -> down + residual; then output_norm and lm_head.  The tensor
types follow llama-quantize's Q4_K_M recipe (src/llama-quant.cpp:185-187, 225-227, 302-303, 358-364): everything Q4_K except attn_v
and ffn_down in the `use_more_bits` layers and output.weight, which are Q6_K (held in the planar layout).

This is bench / test plumbing over the product's C-ABI: it launches exactly the kernels the ggml plugin launches for the same graph
(csrc/ggml_b200/), on torch's current stream so that a step can be captured into a CUDA graph.  Weights are synthetic: random but
valid quantised blocks (no checkpoint can be fetched here), norm weights 1 + 0.1*N(0,1).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import ops

P = C.c_void_p


@dataclass
class LLMConfig:
    name: str = "MiniCPM-o-4.5-LLM(Qwen3-8B)-Q4_K_M"
    n_embd: int = 4096
    n_layer: int = 36
    n_head: int = 32
    n_head_kv: int = 8
    head_dim: int = 128
    n_ff: int = 12288
    n_vocab: int = 151748
    n_ctx: int = 4096
    rms_eps: float = 1e-6
    rope_base: float = 1e6
    n_ctx_orig: int = 40960

    @staticmethod
    def tiny() -> "LLMConfig":
        return LLMConfig(name="qwen3-tiny", n_embd=1024, n_layer=2, n_head=8, n_head_kv=2, head_dim=128, n_ff=3072, n_vocab=4096, n_ctx=512)


def use_more_bits(i: int, n: int) -> bool:          # src/llama-quant.cpp:120-122
    return i < n // 8 or i >= 7 * n // 8 or (i - n // 8) % 3 == 2


def layer_types(cfg: LLMConfig, il: int) -> dict:
    hi = ops.Q6_K if use_more_bits(il, cfg.n_layer) else ops.Q4_K
    return {"wq": ops.Q4_K, "wk": ops.Q4_K, "wv": hi, "wo": ops.Q4_K, "gate": ops.Q4_K, "up": ops.Q4_K, "down": hi}


def weight_bytes_per_token(cfg: LLMConfig, layers: range | None = None, with_head: bool = True) -> int:
    """Algorithmic HBM bytes of weights one decoded token must read."""
    layers = range(cfg.n_layer) if layers is None else layers
    q, kv = cfg.n_head * cfg.head_dim, cfg.n_head_kv * cfg.head_dim
    shapes = {"wq": (q, cfg.n_embd), "wk": (kv, cfg.n_embd), "wv": (kv, cfg.n_embd), "wo": (cfg.n_embd, q),
              "gate": (cfg.n_ff, cfg.n_embd), "up": (cfg.n_ff, cfg.n_embd), "down": (cfg.n_embd, cfg.n_ff)}
    total = 0
    for il in layers:
        ty = layer_types(cfg, il)
        for n, (m, k) in shapes.items():
            total += m * ops.row_size(ty[n], k)
        total += (2 * cfg.n_embd + 2 * cfg.head_dim) * 4
    if with_head:
        total += cfg.n_vocab * ops.row_size(ops.Q6_K, cfg.n_embd) + cfg.n_embd * 4
    return total


def kv_bytes_per_token(cfg: LLMConfig, n_kv: int, layers: range | None = None) -> int:
    n = cfg.n_layer if layers is None else len(layers)
    return 2 * n * cfg.n_head_kv * cfg.head_dim * 2 * (n_kv + 1)


def _rand_weight(wtype: int, m: int, k: int, gen: torch.Generator, dev) -> torch.Tensor:
    """Random valid blocks, generated on the device.  Q6_K is produced directly in the planar layout (payload plane | d plane)."""
    nblk = m * k // ops.BLOCK[wtype][0]

    def halfs(n, lo, hi):
        return (torch.rand(n, generator=gen, device=dev) * (hi - lo) + lo).to(torch.float16).view(torch.uint8).reshape(n, 2)
    if wtype == ops.Q4_K:
        w = torch.randint(0, 256, (nblk, 144), dtype=torch.uint8, generator=gen, device=dev)
        w[:, 0:2] = halfs(nblk, 2e-5, 2e-4)
        w[:, 2:4] = halfs(nblk, 2e-5, 2e-4)
        return w.reshape(-1)
    if wtype == ops.Q6_K:
        pay = torch.randint(0, 256, (nblk, 208), dtype=torch.uint8, generator=gen, device=dev)
        d = halfs(nblk, 2e-5, 2e-4)
        return torch.cat([pay.reshape(-1), d.reshape(-1)])
    raise ValueError(wtype)


class Qwen3Decoder:
    """Weights + KV cache of a contiguous layer range on one GPU, and the launch sequence of one batch-1 decode step."""

    def __init__(self, cfg: LLMConfig, device="cuda:0", layers: range | None = None, has_head: bool = True, seed: int = 0):
        self.cfg, self.dev = cfg, torch.device(device)
        self.layers = range(cfg.n_layer) if layers is None else layers
        self.has_head = has_head
        gen = torch.Generator(device=self.dev)
        gen.manual_seed(seed)
        q, kv, E, F = cfg.n_head * cfg.head_dim, cfg.n_head_kv * cfg.head_dim, cfg.n_embd, cfg.n_ff
        shapes = {"wq": (q, E), "wk": (kv, E), "wv": (kv, E), "wo": (E, q), "gate": (F, E), "up": (F, E), "down": (E, F)}
        self.L = []
        for il in self.layers:
            ty = layer_types(cfg, il)
            lw = {"types": ty}
            for n, (m, k) in shapes.items():
                lw[n] = _rand_weight(ty[n], m, k, gen, self.dev)
            for n, sz in (("attn_norm", E), ("ffn_norm", E), ("q_norm", cfg.head_dim), ("k_norm", cfg.head_dim)):
                lw[n] = 1 + 0.1 * torch.randn(sz, generator=gen, device=self.dev)
            lw["k_cache"] = torch.zeros((cfg.n_ctx, kv), dtype=torch.float16, device=self.dev)
            lw["v_cache"] = torch.zeros((cfg.n_ctx, kv), dtype=torch.float16, device=self.dev)
            self.L.append(lw)
        if has_head:
            self.out_norm = 1 + 0.1 * torch.randn(E, generator=gen, device=self.dev)
            self.lm_head = _rand_weight(ops.Q6_K, cfg.n_vocab, E, gen, self.dev)
            self.logits = torch.zeros(cfg.n_vocab, device=self.dev)
        # step inputs (device side; the host writes them before every step exactly like llama's set_inputs)
        self.x_in = torch.zeros(E, device=self.dev)                 # embedding row of the token (GET_ROWS runs on the host side)
        self.pos = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.kv_idx = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.mask_f32 = torch.zeros((64, cfg.n_ctx), device=self.dev)    # [GGML_PAD(n_tokens, 64), n_kv] F32 (llama-graph.cpp:1532 casts)
        self.mask_f16 = torch.zeros((64, cfg.n_ctx), dtype=torch.float16, device=self.dev)
        # activations
        self.xa, self.xb = torch.zeros(E, device=self.dev), torch.zeros(E, device=self.dev)
        self.q, self.k, self.v = torch.zeros(q, device=self.dev), torch.zeros(kv, device=self.dev), torch.zeros(kv, device=self.dev)
        self.attn = torch.zeros(q, device=self.dev)
        self.h = torch.zeros(F, device=self.dev)
        L = ops.lib()
        self.act_e = torch.zeros(L.b200_act_bytes(ops.Q4_K, E), dtype=torch.uint8, device=self.dev)
        self.act_q = torch.zeros(L.b200_act_bytes(ops.Q4_K, q), dtype=torch.uint8, device=self.dev)
        self.act_f = torch.zeros(L.b200_act_bytes(ops.Q4_K, F), dtype=torch.uint8, device=self.dev)
        self.fa_scratch = torch.zeros(8 << 20, dtype=torch.uint8, device=self.dev)
        self.rope = ops.RopeParams(cfg.head_dim, 2, cfg.n_ctx_orig, cfg.rope_base, 1.0, 0.0, 1.0, 32.0, 1.0)
        self.launches_per_step = 0

    # ---- host-side input preparation (what llama_context::set_inputs writes; src/llama-kv-cache.cpp:1142-1204) ---------------
    @staticmethod
    def host_inputs(cfg: LLMConfig, pos: int, n_kv: int, pinned: bool = True) -> dict:
        mask = torch.full((64, n_kv), float("-inf"))
        mask[0, :pos + 1] = 0.0
        d = {"pos": torch.tensor([pos], dtype=torch.int32), "kv_idx": torch.tensor([pos], dtype=torch.int64), "mask": mask}
        return {k: v.pin_memory() for k, v in d.items()} if pinned else d

    def _job(self, w, wtype, m, k, y, residual=None):
        layout = ops.LAYOUT_PLANAR if wtype == ops.Q6_K else ops.LAYOUT_NATIVE
        return ops.make_job(w, wtype, m, k, y, residual, layout)

    def step(self, n_kv: int, engine: bool = False) -> int:
        """Enqueue one decode step on the current stream (capturable).  x_in/pos/kv_idx/mask_f32[:, :n_kv] must be set.  Returns #launches."""
        if engine:
            return self.step_engine(n_kv)
        cfg, L = self.cfg, ops.lib()
        E, F, D = cfg.n_embd, cfg.n_ff, cfg.head_dim
        q, kv = cfg.n_head * D, cfg.n_head_kv * D
        st = ops.stream()
        n = 0
        mask16 = self.mask_f16[:, :n_kv]
        ops.cpy(self.mask_f32[:, :n_kv], mask16)
        n += 1
        x, x_alt = self.x_in, self.xa
        for lw in self.L:
            ty = lw["types"]
            ops.check(L.b200_rms_norm_quantize(P(x.data_ptr()), C.c_int64(E), P(lw["attn_norm"].data_ptr()), None, C.c_int64(E),
                                               P(self.act_e.data_ptr()), ops.Q4_K, C.c_int64(E), C.c_int64(1), C.c_float(cfg.rms_eps), st))
            jobs = [self._job(lw["wq"], ty["wq"], q, E, self.q), self._job(lw["wk"], ty["wk"], kv, E, self.k),
                    self._job(lw["wv"], ty["wv"], kv, E, self.v)]
            n += 1 + self._matvec(jobs, self.act_e, E)
            ops.check(L.b200_qkv_post(P(self.q.data_ptr()), P(self.k.data_ptr()), P(self.v.data_ptr()), P(lw["q_norm"].data_ptr()),
                                      P(lw["k_norm"].data_ptr()), P(self.pos.data_ptr()), P(self.kv_idx.data_ptr()), ops.I64,
                                      P(lw["k_cache"].data_ptr()), P(lw["v_cache"].data_ptr()), C.c_int64(kv * 2), C.c_int64(kv * 2),
                                      D, cfg.n_head, cfg.n_head_kv, C.c_int64(1), C.c_int64(q), C.c_int64(kv), C.c_int64(kv),
                                      C.byref(self.rope), C.c_float(cfg.rms_eps), st))
            kview = lw["k_cache"][:n_kv].view(n_kv, cfg.n_head_kv, D).permute(1, 0, 2)
            vview = lw["v_cache"][:n_kv].view(n_kv, cfg.n_head_kv, D).permute(1, 0, 2)
            ops.flash_attn(self.q.view(1, cfg.n_head, D).permute(1, 0, 2), kview, vview, mask16, 1.0 / D ** 0.5,
                           out=self.attn.view(1, cfg.n_head, D), scratch=self.fa_scratch)
            ops.check(L.b200_quantize_act(ops.Q4_K, P(self.attn.data_ptr()), C.c_int64(q), P(self.act_q.data_ptr()), C.c_int64(q), C.c_int64(1), st))
            x1 = x_alt
            n += 4 + self._matvec([self._job(lw["wo"], ty["wo"], E, q, x1, residual=x)], self.act_q, q)
            ops.check(L.b200_rms_norm_quantize(P(x1.data_ptr()), C.c_int64(E), P(lw["ffn_norm"].data_ptr()), None, C.c_int64(E),
                                               P(self.act_e.data_ptr()), ops.Q4_K, C.c_int64(E), C.c_int64(1), C.c_float(cfg.rms_eps), st))
            ops.matvec_q_swiglu(self._job(lw["gate"], ty["gate"], F, E, self.h), self._job(lw["up"], ty["up"], F, E, self.h), self.h, self.act_e, E)
            ops.check(L.b200_quantize_act(ops.Q4_K, P(self.h.data_ptr()), C.c_int64(F), P(self.act_f.data_ptr()), C.c_int64(F), C.c_int64(1), st))
            x2 = self.xb if x1 is self.xa else self.xa
            n += 3 + self._matvec([self._job(lw["down"], ty["down"], E, F, x2, residual=x1)], self.act_f, F)
            x, x_alt = x2, (self.xb if x2 is self.xa else self.xa)
        self.x_out = x
        if self.has_head:
            ops.check(L.b200_rms_norm_quantize(P(x.data_ptr()), C.c_int64(E), P(self.out_norm.data_ptr()), None, C.c_int64(E),
                                               P(self.act_e.data_ptr()), ops.Q6_K, C.c_int64(E), C.c_int64(1), C.c_float(cfg.rms_eps), st))
            n += 1 + self._matvec([self._job(self.lm_head, ops.Q6_K, cfg.n_vocab, E, self.logits)], self.act_e, E)
        self.launches_per_step = n
        return n

    # ---- the persistent decode engine: the whole step as ONE cooperative kernel (include/b200_ops.h b200_decoder_*) ---------------
    def build_engine(self, x_out: torch.Tensor | None = None, hidden_out: torch.Tensor | None = None):
        cfg = self.cfg
        kvb = cfg.n_head_kv * cfg.head_dim * 2

        def W(t, wtype):
            return ops.Weight(t.data_ptr(), wtype, ops.LAYOUT_PLANAR if wtype in ops.PAYLOAD else ops.LAYOUT_NATIVE)     # q6_K / q8_0 / q4_0 stream from the planar layout
        layers = (ops.DecodeLayer * len(self.L))()
        for i, lw in enumerate(self.L):
            ty = lw["types"]
            for n in ("wq", "wk", "wv", "wo", "gate", "up", "down"):
                setattr(layers[i], n, W(lw[n], ty[n]))
            for n in ("attn_norm", "ffn_norm", "q_norm", "k_norm", "k_cache", "v_cache"):
                setattr(layers[i], n, lw[n].data_ptr() if lw[n] is not None else None)      # q_norm / k_norm None: llama arch
            layers[i].k_row_bytes = layers[i].v_row_bytes = kvb
        d = ops.DecodeDesc()
        d.n_layer, d.n_embd, d.n_head, d.n_head_kv, d.head_dim, d.n_ff, d.n_vocab = (len(self.L), cfg.n_embd, cfg.n_head, cfg.n_head_kv,
                                                                                   cfg.head_dim, cfg.n_ff, cfg.n_vocab)
        d.rms_eps, d.attn_scale, d.rope, d.layers = cfg.rms_eps, 1.0 / cfg.head_dim ** 0.5, self.rope, layers
        if self.has_head:
            d.out_norm, d.lm_head, d.logits = self.out_norm.data_ptr(), W(self.lm_head, ops.Q6_K), self.logits.data_ptr()
        self.engine_x_out = x_out if x_out is not None else torch.zeros(cfg.n_embd, device=self.dev)
        d.x_out = self.engine_x_out.data_ptr()
        d.hidden_out = hidden_out.data_ptr() if hidden_out is not None else None
        d.x_in, d.pos, d.kv_idx = self.x_in.data_ptr(), self.pos.data_ptr(), self.kv_idx.data_ptr()
        d.mask = self.mask_f16.data_ptr()                                 # row 0 of the [64, n_ctx] F16 mask
        h = C.c_void_p()
        ops.check(ops.lib().b200_decoder_create(C.byref(d), C.byref(h)))
        self._engine, self._engine_keep = h, (layers, d)
        return h

    def step_engine(self, n_kv: int) -> int:
        """mask cast (F32 -> F16, as llama-graph.cpp:1532 does) + the one-kernel decode step.  Returns #launches."""
        ops.cpy(self.mask_f32[:1, :n_kv], self.mask_f16[:1, :n_kv])
        ops.check(ops.lib().b200_decoder_step(self._engine, n_kv, ops.stream()))
        self.x_out = self.engine_x_out
        return 2

    # ---- prefill: one ubatch of n tokens through the same C-ABI entry points the ggml plugin calls node by node for an n-token graph -------
    def prefill(self, x: torch.Tensor, pos0: int, n_kv: int, fused_tiles: bool = True) -> tuple[torch.Tensor, int]:
        """x [n, n_embd] F32 (embedding rows) at positions pos0 .. pos0+n-1 -> (logits of the LAST token, #launches).  Quantised MUL_MATs with
        n > 8 columns run on the tcgen05 dequant-GEMM (csrc/mmq_tc.cu), attention on k_fa_prefill (csrc/fa_prefill.cu); KV rows are written."""
        cfg, L = self.cfg, ops.lib()
        n, E, F, D = x.shape[0], cfg.n_embd, cfg.n_ff, cfg.head_dim
        q, kv = cfg.n_head * D, cfg.n_head_kv * D
        st = ops.stream()
        dev = self.dev
        pos = torch.arange(pos0, pos0 + n, dtype=torch.int32, device=dev)
        idx = pos.to(torch.int64)
        n_pad = (n + 63) // 64 * 64
        ar = torch.arange(n_kv, device=dev)[None, :]
        mask = torch.full((n_pad, n_kv), float("-inf"), device=dev)
        mask[:n] = torch.where(ar <= pos[:, None].to(torch.int64), 0.0, float("-inf"))
        mask16 = torch.empty((n_pad, n_kv), dtype=torch.float16, device=dev)
        ops.cpy(mask, mask16)
        nl = 1
        bufs = {k: torch.empty((n, m), device=dev) for k, m in (("a", E), ("q", q), ("k", kv), ("v", kv), ("attn", q), ("x1", E), ("g", F), ("u", F), ("h", F), ("x2", E))}
        fa_scratch = torch.empty(1 << 20, dtype=torch.uint8, device=dev)

        mm_scratch = torch.empty((n + 255) // 256 * 256 * max(E, F, q) * 2, dtype=torch.uint8, device=dev)      # F16 activation tiles of one MUL_MAT

        def mm(w, wtype, m, k, xin, out, reuse=False):
            ops.mul_mat(w, wtype, m, k, xin, layout=ops.LAYOUT_PLANAR if wtype == ops.Q6_K else ops.LAYOUT_NATIVE, out=out, scratch=mm_scratch, reuse_act=reuse)
        # fused_tiles: RMS_NORM, FLASH_ATTN_EXT and SWIGLU write the NEXT MUL_MAT's F16 activation tiles directly (b200_*_tiles) and every tensor-core MUL_MAT runs
        # with B200_MM_REUSE_ACT: no F32 round trip and no k_x_to_f16_tiles pass per MUL_MAT (the ggml plugin fuses the same pairs when the use counts allow it)
        ft = fused_tiles and n >= 64 and 1024 < E <= 4096 and D == 128 and E % 64 == 0 and F % 64 == 0
        for lw in self.L:
            ty = lw["types"]
            if ft:
                ops.rms_norm_tiles(x, cfg.rms_eps, lw["attn_norm"], mm_scratch)
            else:
                ops.rms_norm(x, cfg.rms_eps, w=lw["attn_norm"], out=bufs["a"])
            if ft:                                    # q / k / v as ONE launch over the concatenated m-tiles (b200_mul_mat_multi)
                lay = lambda t: ops.LAYOUT_PLANAR if t == ops.Q6_K else ops.LAYOUT_NATIVE
                ops.mul_mat_multi([(lw["wq"], ty["wq"], q, lay(ty["wq"])), (lw["wk"], ty["wk"], kv, lay(ty["wk"])), (lw["wv"], ty["wv"], kv, lay(ty["wv"]))],
                                  bufs["a"], [bufs["q"], bufs["k"], bufs["v"]], scratch=mm_scratch, reuse_act=True)
            else:
                mm(lw["wq"], ty["wq"], q, E, bufs["a"], bufs["q"], ft); mm(lw["wk"], ty["wk"], kv, E, bufs["a"], bufs["k"], n > 8); mm(lw["wv"], ty["wv"], kv, E, bufs["a"], bufs["v"], n > 8)
            ops.check(L.b200_qkv_post(P(bufs["q"].data_ptr()), P(bufs["k"].data_ptr()), P(bufs["v"].data_ptr()), P(lw["q_norm"].data_ptr()),
                                      P(lw["k_norm"].data_ptr()), P(pos.data_ptr()), P(idx.data_ptr()), ops.I64,
                                      P(lw["k_cache"].data_ptr()), P(lw["v_cache"].data_ptr()), C.c_int64(kv * 2), C.c_int64(kv * 2),
                                      D, cfg.n_head, cfg.n_head_kv, C.c_int64(n), C.c_int64(q), C.c_int64(kv), C.c_int64(kv),
                                      C.byref(self.rope), C.c_float(cfg.rms_eps), st))
            kview = lw["k_cache"][:n_kv].view(n_kv, cfg.n_head_kv, D).permute(1, 0, 2)
            vview = lw["v_cache"][:n_kv].view(n_kv, cfg.n_head_kv, D).permute(1, 0, 2)
            if ft:
                ops.flash_attn_tiles(bufs["q"].view(n, cfg.n_head, D).permute(1, 0, 2), kview, vview, mask16, 1.0 / D ** 0.5, bufs["attn"].view(n, cfg.n_head, D), fa_scratch, mm_scratch)
            else:
                ops.flash_attn(bufs["q"].view(n, cfg.n_head, D).permute(1, 0, 2), kview, vview, mask16, 1.0 / D ** 0.5,
                               out=bufs["attn"].view(n, cfg.n_head, D), scratch=fa_scratch)
            if ft:                                    # residual ADD in the GEMM epilogue (b200_mul_mat_add)
                ops.mul_mat_add(lw["wo"], ty["wo"], E, q, bufs["attn"], x, bufs["x1"], layout=ops.LAYOUT_PLANAR if ty["wo"] == ops.Q6_K else ops.LAYOUT_NATIVE, scratch=mm_scratch, reuse_act=True)
            else:
                mm(lw["wo"], ty["wo"], E, q, bufs["attn"], bufs["x1"], ft)
                ops.binary(ops.ADD, bufs["x1"], x, out=bufs["x1"])
            if ft:
                ops.rms_norm_tiles(bufs["x1"], cfg.rms_eps, lw["ffn_norm"], mm_scratch)
            else:
                ops.rms_norm(bufs["x1"], cfg.rms_eps, w=lw["ffn_norm"], out=bufs["a"])
            if ft:                                    # gate / up as one launch: 2 x 96 m-tiles share the waves (one round of tiles less per layer)
                lay = lambda t: ops.LAYOUT_PLANAR if t == ops.Q6_K else ops.LAYOUT_NATIVE
                ops.mul_mat_multi([(lw["gate"], ty["gate"], F, lay(ty["gate"])), (lw["up"], ty["up"], F, lay(ty["up"]))], bufs["a"], [bufs["g"], bufs["u"]],
                                  scratch=mm_scratch, reuse_act=True)
            else:
                mm(lw["gate"], ty["gate"], F, E, bufs["a"], bufs["g"], ft); mm(lw["up"], ty["up"], F, E, bufs["a"], bufs["u"], n > 8)
            if ft:
                ops.glu_tiles(ops.GLU_SWIGLU, bufs["g"], bufs["u"], mm_scratch)
            else:
                ops.check(L.b200_glu(ops.GLU_SWIGLU, ops._ref(ops.T(bufs["g"])), ops._ref(ops.T(bufs["u"])), ops._ref(ops.T(bufs["h"])), 0, st))
            if ft:
                x = ops.mul_mat_add(lw["down"], ty["down"], E, F, bufs["h"], bufs["x1"], bufs["x2"], layout=ops.LAYOUT_PLANAR if ty["down"] == ops.Q6_K else ops.LAYOUT_NATIVE, scratch=mm_scratch, reuse_act=True)
            else:
                mm(lw["down"], ty["down"], E, F, bufs["h"], bufs["x2"], ft)      # the layer input (x) is dead once x1 exists: x2 may be the same buffer
                x = ops.binary(ops.ADD, bufs["x2"], bufs["x1"], out=bufs["x2"])
            nl += 2 * 7 + 8                           # 7 GEMMs (+ their activation tiling), norm x2, qkv_post, kvmax + fa, add x2, glu
        logits = None
        if self.has_head:
            last = ops.rms_norm(x[n - 1:n], cfg.rms_eps, w=self.out_norm)
            logits = ops.mul_mat(self.lm_head, ops.Q6_K, cfg.n_vocab, E, last, layout=ops.LAYOUT_PLANAR)
            nl += 3
        return logits, nl

    def _matvec(self, jobs, act, k) -> int:
        """One launch per run of equal weight type (a Q4_K_M layer mixes Q4_K and Q6_K in q/k/v)."""
        n, i = 0, 0
        while i < len(jobs):
            j = i
            while j < len(jobs) and jobs[j].type == jobs[i].type:
                j += 1
            ops.matvec_q(jobs[i:j], act, k)
            n += 1
            i = j
        return n

"""One whole decoded token on the CPU ORACLE: the chain of oracle/*.c restatements that the reference's ggml CPU backend executes for
the batch-1 graph of llm_build_qwen3 / llm_build_llama (src/llama-model.cpp:9287-9406, build_attn src/llama-graph.cpp:1546-1597,
build_ffn :713-819):

    a = RMS_NORM(x)*attn_norm -> q8_K -> wq/wk/wv -> [q/k RMS_NORM*w] -> ROPE -> SET_ROWS(F16 cache) -> FLASH_ATTN_EXT
      -> q8_K -> wo (+x) -> RMS_NORM*ffn_norm -> q8_K -> gate/up -> SWIGLU -> q8_K -> down (+) -> ... -> RMS_NORM*output_norm -> lm_head

TEST INFRASTRUCTURE ONLY (imported by tests/ and __graft_entry__.smoke()): the product never sees it.  Weights are NATIVE ggml blocks
(numpy uint8); K/V caches are numpy F16 [n_ctx, n_head_kv*D] exactly as llama-kv-cache lays them out (llama-kv-cache.cpp:981-1009).
"""
from __future__ import annotations

import numpy as np

import oracle_c as O


def _mm(wtype, w, x, m, k):
    return O.mul_mat(wtype, w, x.reshape(1, k), m, k)[0]


def oracle_token(cfg, layers, x, pos: int, n_kv: int, head=None, rope_mode: int = 2, f16_acc: bool = False):
    """layers: list of dicts {wq, wk, wv, wo, gate, up, down: native uint8 blocks; types: {name: ggml type}; attn_norm, ffn_norm,
    q_norm, k_norm (None = llama arch): f32; k_cache, v_cache: F16 [n_ctx, n_head_kv*D], UPDATED IN PLACE like SET_ROWS does}.
    head: None (pipeline stage: returns the residual stream) or {out_norm, lm_head, type}.
    Returns (logits or None, residual stream after the last layer, RMS_NORM(x)*output_norm or None)."""
    E, F, D, H, HK = cfg.n_embd, cfg.n_ff, cfg.head_dim, cfg.n_head, cfg.n_head_kv
    q_n, kv_n = H * D, HK * D
    x = np.asarray(x, np.float32).copy()
    mask = np.full((1, n_kv), -np.inf, np.float16)
    mask[0, :pos + 1] = 0
    for L in layers:
        ty = L["types"]
        a = O.rms_norm(x.reshape(1, E), cfg.rms_eps)[0] * L["attn_norm"]
        q = _mm(ty["wq"], L["wq"], a, q_n, E).reshape(1, H, D)
        k = _mm(ty["wk"], L["wk"], a, kv_n, E).reshape(1, HK, D)
        v = _mm(ty["wv"], L["wv"], a, kv_n, E)
        if L.get("q_norm") is not None:
            q = (O.rms_norm(q.reshape(H, D), cfg.rms_eps) * L["q_norm"]).reshape(1, H, D).astype(np.float32)
            k = (O.rms_norm(k.reshape(HK, D), cfg.rms_eps) * L["k_norm"]).reshape(1, HK, D).astype(np.float32)
        p = np.array([pos], np.int32)
        q = O.rope(q, p, D, rope_mode, cfg.n_ctx_orig, cfg.rope_base)
        k = O.rope(k, p, D, rope_mode, cfg.n_ctx_orig, cfg.rope_base)
        L["k_cache"][pos] = k.reshape(kv_n).astype(np.float16)          # SET_ROWS F32 -> F16 (round to nearest even, set-rows / cpy)
        L["v_cache"][pos] = v.astype(np.float16)
        kc = np.ascontiguousarray(L["k_cache"][:n_kv].reshape(n_kv, HK, D).transpose(1, 0, 2))
        vc = np.ascontiguousarray(L["v_cache"][:n_kv].reshape(n_kv, HK, D).transpose(1, 0, 2))
        attn = O.flash_attn(q, kc, vc, mask, 1.0 / np.sqrt(D), f16_acc=f16_acc).reshape(q_n)
        x1 = _mm(ty["wo"], L["wo"], attn, E, q_n) + x
        f = O.rms_norm(x1.reshape(1, E), cfg.rms_eps)[0] * L["ffn_norm"]
        h = O.swiglu(_mm(ty["gate"], L["gate"], f, F, E), _mm(ty["up"], L["up"], f, F, E))
        x = _mm(ty["down"], L["down"], h, E, F) + x1
    if head is None:
        return None, x, None
    hn = O.rms_norm(x.reshape(1, E), cfg.rms_eps)[0] * head["out_norm"]
    return _mm(head["type"], head["lm_head"], hn.astype(np.float32), head["n_vocab"], E), x, hn.astype(np.float32)

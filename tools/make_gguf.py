#!/usr/bin/env python
"""Write a synthetic Qwen3-architecture GGUF (v3) with the tensor names / types of a llama-quantize Q4_K_M file, without the
reference's gguf-py (which does not travel to the GPU box) and without a 32 GB F32 intermediate (random quant blocks
with sane scales).  Quant blocks are random bytes whose f16 scale fields are set so that the dequantised weights are ~N(0, 0.02^2)
and zero-mean (q4_K: dmin = 7.5 d balances E[sc q] against E[m]; q6_K is symmetric by construction).

  python tools/make_gguf.py out.gguf [--layers 36] [--vocab 151748] [--ftype q4_k_m|f16] [--seed 0]

Layout follows the GGUF spec as read by ggml/src/gguf.cpp (magic, version 3, counts, KV pairs, tensor infos, 32-byte aligned data);
per-tensor types follow src/llama-quant.cpp:185-187,225-227,302-303,358-364 (attn_v / ffn_down -> Q6_K on use_more_bits layers,
output -> Q6_K, the rest Q4_K, norms F32)."""
from __future__ import annotations

import argparse
import struct
import sys

import numpy as np

F32, F16, Q4_0, Q8_0, Q4_K, Q5_K, Q6_K = 0, 1, 2, 8, 12, 13, 14
BLOCK = {F32: (1, 4), F16: (1, 2), Q4_K: (256, 144), Q6_K: (256, 210), Q8_0: (32, 34), Q4_0: (32, 18)}
ALIGN = 32


def use_more_bits(i: int, n: int) -> bool:          # src/llama-quant.cpp:120-122
    return i < n // 8 or i >= 7 * n // 8 or (i - n // 8) % 3 == 2


def gstr(s: str) -> bytes:
    b = s.encode()
    return struct.pack("<Q", len(b)) + b


def kv(key: str, vtype: int, payload: bytes) -> bytes:
    return gstr(key) + struct.pack("<I", vtype) + payload


def kv_u32(k, v): return kv(k, 4, struct.pack("<I", v))
def kv_f32(k, v): return kv(k, 6, struct.pack("<f", v))
def kv_str(k, v): return kv(k, 8, gstr(v))


def tensor_bytes(rng: np.random.Generator, ttype: int, ne0: int, ne1: int, kind: str) -> np.ndarray:
    n = ne0 * ne1
    if ttype == F32:
        if kind == "norm":
            return (1.0 + 0.1 * rng.standard_normal(n)).astype(np.float32).view(np.uint8)
        out = np.empty(n, np.float32)
        step = 1 << 24
        for i in range(0, n, step):
            out[i:i + step] = 0.02 * rng.standard_normal(min(step, n - i), dtype=np.float32)
        return out.view(np.uint8)
    if ttype == F16:
        out = np.empty(n, np.float16)
        step = 1 << 24
        for i in range(0, n, step):
            out[i:i + step] = (0.02 * rng.standard_normal(min(step, n - i), dtype=np.float32)).astype(np.float16)
        return out.view(np.uint8)
    qk, bs = BLOCK[ttype]
    nb = n // qk
    raw = rng.integers(0, 256, nb * bs, dtype=np.uint8).reshape(nb, bs)
    if ttype == Q4_K:
        d = np.float16(rng.uniform(0.8e-4, 1.4e-4, nb))
        raw[:, 0:2] = d.view(np.uint8).reshape(-1, 2)
        raw[:, 2:4] = np.float16(7.5 * d.astype(np.float32)).view(np.uint8).reshape(-1, 2)
    elif ttype == Q6_K:
        raw[:, 208:210] = np.float16(rng.uniform(1.0e-5, 2.0e-5, nb)).view(np.uint8).reshape(-1, 2)
    elif ttype in (Q8_0, Q4_0):
        raw[:, 0:2] = np.float16(rng.uniform(1e-4, 3e-4, nb) * (1 if ttype == Q8_0 else 16)).view(np.uint8).reshape(-1, 2)
    return raw.reshape(-1)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--layers", type=int, default=36)
    ap.add_argument("--vocab", type=int, default=151748)
    ap.add_argument("--embd", type=int, default=4096)
    ap.add_argument("--ff", type=int, default=12288)
    ap.add_argument("--heads", type=int, default=32)
    ap.add_argument("--kv-heads", type=int, default=8)
    ap.add_argument("--head-dim", type=int, default=128)
    ap.add_argument("--ctx", type=int, default=40960)
    ap.add_argument("--ftype", default="q4_k_m", choices=["q4_k_m", "f16", "q8_0", "f32"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--arch", default="qwen3", choices=["qwen3", "llama"], help="llama: no q/k-norm tensors, ROPE norm mode (the MiniCPM-o TTS llama: --embd 768 --ff 3072 --heads 12 --kv-heads 12 --head-dim 64 --layers 20)")
    ap.add_argument("--reuse-layers", action="store_true", help="speed-only files: every layer reuses layer 0's random bytes (minutes -> seconds for a 15 GB F16 file)")
    a = ap.parse_args()
    E, F, Q, KV, D, L = a.embd, a.ff, a.heads * a.head_dim, a.kv_heads * a.head_dim, a.head_dim, a.layers
    rng = np.random.default_rng(a.seed)

    def wtype(name: str, il: int) -> int:
        if a.ftype == "f32":                             # input for the reference's own llama-quantize (realistic quantised weights)
            return F32
        if a.ftype == "f16":
            return F16
        if a.ftype == "q8_0":
            return Q8_0
        if name == "output":
            return Q6_K
        if name in ("attn_v", "ffn_down") and use_more_bits(il, L):
            return Q6_K
        return Q4_K

    tensors = [("token_embd.weight", wtype("token_embd", 0), E, a.vocab, "w")]
    for il in range(L):
        p = f"blk.{il}."
        tensors += [(p + "attn_norm.weight", F32, E, 1, "norm"), (p + "attn_q.weight", wtype("attn_q", il), E, Q, "w"),
                    (p + "attn_k.weight", wtype("attn_k", il), E, KV, "w"), (p + "attn_v.weight", wtype("attn_v", il), E, KV, "w"),
                    (p + "attn_output.weight", wtype("attn_output", il), Q, E, "w"), (p + "attn_q_norm.weight", F32, D, 1, "norm"),
                    (p + "attn_k_norm.weight", F32, D, 1, "norm"), (p + "ffn_norm.weight", F32, E, 1, "norm"),
                    (p + "ffn_gate.weight", wtype("ffn_gate", il), E, F, "w"), (p + "ffn_up.weight", wtype("ffn_up", il), E, F, "w"),
                    (p + "ffn_down.weight", wtype("ffn_down", il), F, E, "w")]
    tensors += [("output_norm.weight", F32, E, 1, "norm"), ("output.weight", wtype("output", 0), E, a.vocab, "w")]
    if a.arch == "llama":
        tensors = [t for t in tensors if not t[0].endswith(("attn_q_norm.weight", "attn_k_norm.weight"))]

    ftype_id = {"q4_k_m": 15, "f16": 1, "q8_0": 7, "f32": 0}[a.ftype]
    A = a.arch
    kvs = [kv_str("general.architecture", A), kv_str("general.name", f"synthetic-{A}-{L}L-{a.ftype}"), kv_u32("general.file_type", ftype_id),
           kv_u32(f"{A}.block_count", L), kv_u32(f"{A}.context_length", a.ctx), kv_u32(f"{A}.embedding_length", E),
           kv_u32(f"{A}.feed_forward_length", F), kv_u32(f"{A}.attention.head_count", a.heads), kv_u32(f"{A}.attention.head_count_kv", a.kv_heads),
           kv_u32(f"{A}.attention.key_length", D), kv_u32(f"{A}.attention.value_length", D), kv_f32(f"{A}.attention.layer_norm_rms_epsilon", 1e-6),
           kv_f32(f"{A}.rope.freq_base", 1e6), kv_u32(f"{A}.vocab_size", a.vocab), kv_str("tokenizer.ggml.model", "no_vocab"),
           kv_u32("general.alignment", ALIGN)]
    if A == "llama":
        kvs.append(kv_u32("llama.rope.dimension_count", D))

    infos, off = [], 0
    for name, t, ne0, ne1, _ in tensors:
        qk, bs = BLOCK[t]
        nbytes = ne0 * ne1 // qk * bs
        dims = [ne0] if ne1 == 1 else [ne0, ne1]
        infos.append(gstr(name) + struct.pack("<I", len(dims)) + b"".join(struct.pack("<Q", d) for d in dims) + struct.pack("<IQ", t, off))
        off += (nbytes + ALIGN - 1) // ALIGN * ALIGN
    header = struct.pack("<IIQQ", 0x46554747, 3, len(tensors), len(kvs)) + b"".join(kvs) + b"".join(infos)
    with open(a.out, "wb") as f:
        f.write(header)
        f.write(b"\0" * (-len(header) % ALIGN))
        cache = {}
        for name, t, ne0, ne1, kind in tensors:
            key = (name.split(".", 2)[2] if name.startswith("blk.") else name, t, ne0, ne1)
            if a.reuse_layers and key in cache:
                b = cache[key]
            else:
                b = tensor_bytes(rng, t, ne0, ne1, kind)
                if a.reuse_layers and name.startswith("blk."):
                    cache[key] = b
            f.write(b.tobytes())
            f.write(b"\0" * (-b.size % ALIGN))
    print(f"wrote {a.out}: {len(tensors)} tensors, {off / 1e9:.3f} GB of tensor data, ftype {a.ftype}", file=sys.stderr)


if __name__ == "__main__":
    main()

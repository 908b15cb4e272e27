#!/usr/bin/env python
"""Write synthetic GGUFs for the omni encoders that sit before the LLM on BASELINE.json configs[3], with exactly the KV keys and tensor
names the reference's own loaders ask for, so that the UNMODIFIED tools/omni/audition.cpp / vision.cpp load and run them (tests/native/omni_encoders.cpp):

  python tools/make_omni_gguf.py apm out.gguf [--layers 24 --d-model 1024 --heads 16 --proj 4096]     Whisper-medium-shaped audio encoder + 2-layer audio projector
  python tools/make_omni_gguf.py vpm out.gguf [--layers 27 --embd 1152 --heads 16 --ff 4304 --proj 4096] SigLip-so400m-shaped ViT + MiniCPM-V resampler

APM: keys / names from audition.cpp:817-862 (d_model, encoder_attention_heads, encoder_layers, n_mel, n_fft, max_source_positions), :1066-1115 (encoder.* and
audio_projector.* tensors), :1118-1137 (the `filters` array).  Matrices F16 (ggml_conv_1d needs an F16 kernel: ggml.c im2col dst type), biases / norms / positions F32.
VPM: keys from vision.cpp load_hparams, tensor names from omni-impl.h TN_* (v.blk.%d.*, resampler.*).
No gguf-py (it does not travel to the GPU box): the container format is written by hand as tools/make_gguf.py does (ggml/src/gguf.cpp)."""
from __future__ import annotations

import argparse
import struct
import sys

import numpy as np

F32, F16 = 0, 1
ALIGN = 32


def gstr(s: str) -> bytes:
    b = s.encode()
    return struct.pack("<Q", len(b)) + b


def kv(key: str, vtype: int, payload: bytes) -> bytes:
    return gstr(key) + struct.pack("<I", vtype) + payload


def kv_u32(k, v): return kv(k, 4, struct.pack("<I", v))
def kv_i32(k, v): return kv(k, 5, struct.pack("<i", v))
def kv_f32(k, v): return kv(k, 6, struct.pack("<f", v))
def kv_bool(k, v): return kv(k, 7, struct.pack("<?", v))
def kv_str(k, v): return kv(k, 8, gstr(v))
def kv_arr_f32(k, a): return kv(k, 9, struct.pack("<IQ", 6, len(a)) + np.asarray(a, np.float32).tobytes())


def write_gguf(path: str, kvs: list[bytes], tensors: list[tuple[str, list[int], np.ndarray]]) -> None:
    """tensors: (name, ggml ne (dim 0 first), array of F32 or F16 values with prod(ne) elements)"""
    infos, off = [], 0
    for name, ne, arr in tensors:
        assert arr.dtype in (np.float32, np.float16) and arr.size == int(np.prod(ne)), name
        t = F32 if arr.dtype == np.float32 else F16
        infos.append(gstr(name) + struct.pack("<I", len(ne)) + b"".join(struct.pack("<Q", d) for d in ne) + struct.pack("<IQ", t, off))
        off += (arr.nbytes + ALIGN - 1) // ALIGN * ALIGN
    header = struct.pack("<IIQQ", 0x46554747, 3, len(tensors), len(kvs)) + b"".join(kvs) + b"".join(infos)
    with open(path, "wb") as f:
        f.write(header)
        f.write(b"\0" * (-len(header) % ALIGN))
        for _, _, arr in tensors:
            f.write(arr.tobytes())
            f.write(b"\0" * (-arr.nbytes % ALIGN))
    print(f"wrote {path}: {len(tensors)} tensors, {off / 1e6:.1f} MB of tensor data", file=sys.stderr)


class Gen:
    def __init__(self, seed: int):
        self.rng = np.random.default_rng(seed)

    def w(self, *ne, std=None, dtype=np.float16):
        """a matrix with ggml ne = (in, out, ...): std 1/sqrt(in) keeps activations O(1) through the stack"""
        s = std if std is not None else 1.0 / np.sqrt(ne[0])
        return (self.rng.standard_normal(int(np.prod(ne)), dtype=np.float32) * s).astype(dtype)

    def b(self, n, std=0.02):
        return (self.rng.standard_normal(n, dtype=np.float32) * std).astype(np.float32)

    def g(self, n):
        return (1.0 + 0.1 * self.rng.standard_normal(n, dtype=np.float32)).astype(np.float32)


def make_apm(a) -> None:
    S, L, H, P, M, NFFT, CTX = a.d_model, a.layers, a.heads, a.proj, 80, 400, 1500
    g = Gen(a.seed)
    kvs = [kv_str("general.architecture", "clip"), kv_str("general.name", f"synthetic-whisper-encoder-{L}L-{S}"), kv_str("general.model_type", "minicpmo"),
           kv_u32("d_model", S), kv_u32("encoder_attention_heads", H), kv_u32("encoder_layers", L), kv_u32("n_mel", M), kv_u32("n_fft", NFFT),
           kv_u32("max_source_positions", CTX), kv_arr_f32("filters", g.rng.uniform(0, 0.05, M * NFFT)), kv_u32("general.alignment", ALIGN)]
    # sinusoidal positions as Whisper's (values in [-1, 1])
    pos = np.arange(CTX, dtype=np.float32)[:, None] * np.exp(-np.log(10000.0) / (S // 2 - 1) * np.arange(S // 2, dtype=np.float32))[None, :]
    pe = np.concatenate([np.sin(pos), np.cos(pos)], axis=1).astype(np.float32)                    # [CTX, S] rows = positions -> ggml ne [S, CTX]
    T = [("encoder.positional_embedding", [S, CTX], pe.reshape(-1)),
         ("encoder.conv1.weight", [3, M, S], g.w(3, M, S, std=1.0 / np.sqrt(3 * M))), ("encoder.conv1.bias", [1, S], g.b(S)),
         ("encoder.conv2.weight", [3, S, S], g.w(3, S, S, std=1.0 / np.sqrt(3 * S))), ("encoder.conv2.bias", [1, S], g.b(S)),
         ("encoder.ln_post.weight", [S], g.g(S)), ("encoder.ln_post.bias", [S], g.b(S))]
    for i in range(L):
        p = f"encoder.blocks.{i}."
        T += [(p + "attn_ln.weight", [S], g.g(S)), (p + "attn_ln.bias", [S], g.b(S)),
              (p + "attn.query.weight", [S, S], g.w(S, S)), (p + "attn.query.bias", [S], g.b(S)),
              (p + "attn.key.weight", [S, S], g.w(S, S)),
              (p + "attn.value.weight", [S, S], g.w(S, S)), (p + "attn.value.bias", [S], g.b(S)),
              (p + "attn.out.weight", [S, S], g.w(S, S, std=0.5 / np.sqrt(S))), (p + "attn.out.bias", [S], g.b(S)),
              (p + "mlp_ln.weight", [S], g.g(S)), (p + "mlp_ln.bias", [S], g.b(S)),
              (p + "mlp.0.weight", [S, 4 * S], g.w(S, 4 * S)), (p + "mlp.0.bias", [4 * S], g.b(4 * S)),
              (p + "mlp.2.weight", [4 * S, S], g.w(4 * S, S, std=0.5 / np.sqrt(4 * S))), (p + "mlp.2.bias", [S], g.b(S))]
    T += [("audio_projector.linear1.weight", [S, P], g.w(S, P)), ("audio_projector.linear1.bias", [P], g.b(P)),
          ("audio_projector.linear2.weight", [P, P], g.w(P, P)), ("audio_projector.linear2.bias", [P], g.b(P))]
    write_gguf(a.out, kvs, T)


def make_vpm(a) -> None:
    E, L, H, FF, P, PS, IMG, NQ = a.embd, a.layers, a.heads, a.ff, a.proj, 14, 448, 64
    g = Gen(a.seed)
    kvs = [kv_str("general.architecture", "clip"), kv_str("general.name", f"synthetic-siglip-{L}L-{E}-resampler"), kv_str("general.model_type", "minicpmo"),
           kv_u32("clip.vision.embedding_length", E), kv_u32("clip.vision.attention.head_count", H), kv_u32("clip.vision.feed_forward_length", FF),
           kv_u32("clip.vision.block_count", L), kv_u32("clip.vision.projection_dim", 0), kv_f32("clip.vision.attention.layer_norm_epsilon", 1e-6),
           kv_u32("clip.vision.image_size", IMG), kv_u32("clip.vision.patch_size", PS), kv_i32("clip.minicpmv_version", 100045), kv_u32("clip.minicpmv_query_num", NQ),
           kv_bool("clip.use_gelu", True), kv_arr_f32("clip.vision.image_mean", [0.5, 0.5, 0.5]), kv_arr_f32("clip.vision.image_std", [0.5, 0.5, 0.5]),
           kv_u32("general.alignment", ALIGN)]
    T = [("v.patch_embd.weight", [PS, PS, 3, E], g.w(PS, PS, 3, E, std=1.0 / np.sqrt(3 * PS * PS))), ("v.patch_embd.bias", [E], g.b(E)),
         ("v.position_embd.weight", [E, 70 * 70], g.w(E, 70 * 70, std=0.1, dtype=np.float32)),
         ("v.post_ln.weight", [E], g.g(E)), ("v.post_ln.bias", [E], g.b(E))]
    for i in range(L):
        p = f"v.blk.{i}."
        T += [(p + "ln1.weight", [E], g.g(E)), (p + "ln1.bias", [E], g.b(E)), (p + "ln2.weight", [E], g.g(E)), (p + "ln2.bias", [E], g.b(E))]
        for n in ("attn_q", "attn_k", "attn_v"):
            T += [(p + n + ".weight", [E, E], g.w(E, E)), (p + n + ".bias", [E], g.b(E))]
        T += [(p + "attn_out.weight", [E, E], g.w(E, E, std=0.5 / np.sqrt(E))), (p + "attn_out.bias", [E], g.b(E)),
              (p + "ffn_up.weight", [E, FF], g.w(E, FF)), (p + "ffn_up.bias", [FF], g.b(FF)),
              (p + "ffn_down.weight", [FF, E], g.w(FF, E, std=0.5 / np.sqrt(FF))), (p + "ffn_down.bias", [E], g.b(E))]
    T += [("resampler.pos_embed_k", [P, 70 * 70], g.w(P, 70 * 70, std=0.1, dtype=np.float32)), ("resampler.query", [P, NQ], g.w(P, NQ, std=1.0, dtype=np.float32)),
          ("resampler.proj.weight", [P, P], g.w(P, P)), ("resampler.kv.weight", [E, P], g.w(E, P))]
    for n in ("q", "k", "v", "out"):
        T += [(f"resampler.attn.{n}.weight", [P, P], g.w(P, P)), (f"resampler.attn.{n}.bias", [P], g.b(P))]
    for n in ("q", "kv", "post"):
        T += [(f"resampler.ln_{n}.weight", [P], g.g(P)), (f"resampler.ln_{n}.bias", [P], g.b(P))]
    write_gguf(a.out, kvs, T)


def main() -> None:
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="what", required=True)
    p = sub.add_parser("apm")
    p.add_argument("out"); p.add_argument("--layers", type=int, default=24); p.add_argument("--d-model", type=int, default=1024)
    p.add_argument("--heads", type=int, default=16); p.add_argument("--proj", type=int, default=4096); p.add_argument("--seed", type=int, default=0)
    p = sub.add_parser("vpm")
    p.add_argument("out"); p.add_argument("--layers", type=int, default=27); p.add_argument("--embd", type=int, default=1152); p.add_argument("--heads", type=int, default=16)
    p.add_argument("--ff", type=int, default=4304); p.add_argument("--proj", type=int, default=4096); p.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    {"apm": make_apm, "vpm": make_vpm}[a.what](a)


if __name__ == "__main__":
    main()

"""ctypes view of the UNMODIFIED reference compiled into oracle/_ref (libggml-base.so + libggml-cpu.so).

TEST INFRASTRUCTURE ONLY.  Used (a) to pin oracle/*.c against the real reference, (b) to mint the golden
fixtures under tests/golden/ (tests/golden/make_golden.py), (c) as the "reference" CPU arm of bench.py.
Nothing here reads /root/reference at run time: oracle/_ref travels to the GPU box prebuilt.

API mirrored: ggml/include/ggml.h (ggml_init :~2370, ggml_new_tensor, ggml_mul_mat, ggml_rms_norm, ggml_rope_ext,
ggml_swiglu_split, ggml_flash_attn_ext :2179-2202, ggml_set_rows, ggml_soft_max_ext, ggml_quantize_chunk) and
ggml/include/ggml-cpu.h (ggml_graph_compute_with_ctx).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_LIB = ROOT / "oracle" / "_ref" / "lib"

# ggml_type ids (ggml.h:379-421)
F32, F16, Q4_0, Q8_0, Q4_K, Q5_K, Q6_K, Q8_K, I32, I64, BF16 = 0, 1, 2, 8, 12, 13, 14, 15, 26, 27, 30
TYPE_NAMES = {F32: "f32", F16: "f16", Q4_0: "q4_0", Q8_0: "q8_0", Q4_K: "q4_K", Q5_K: "q5_K", Q6_K: "q6_K", BF16: "bf16"}
BLOCK = {F32: (1, 4), F16: (1, 2), BF16: (1, 2), Q4_0: (32, 18), Q8_0: (32, 34), Q4_K: (256, 144), Q5_K: (256, 176),
         Q6_K: (256, 210), Q8_K: (256, 292), I32: (1, 4), I64: (1, 8)}


def row_size(t: int, k: int) -> int:
    b, s = BLOCK[t]
    assert k % b == 0
    return k // b * s


def available() -> bool:
    return (REF_LIB / "libggml-base.so").exists() and (REF_LIB / "libggml-cpu.so").exists()


class _InitParams(C.Structure):
    _fields_ = [("mem_size", C.c_size_t), ("mem_buffer", C.c_void_p), ("no_alloc", C.c_bool)]


class _Tensor(C.Structure):  # prefix of struct ggml_tensor (ggml.h:626-658), enough to read ne/nb/data
    _fields_ = [("type", C.c_int), ("buffer", C.c_void_p), ("ne", C.c_int64 * 4), ("nb", C.c_size_t * 4),
                ("op", C.c_int), ("op_params", C.c_int32 * 16), ("flags", C.c_int32), ("src", C.c_void_p * 10),
                ("view_src", C.c_void_p), ("view_offs", C.c_size_t), ("data", C.c_void_p)]


_libs = None


def libs():
    global _libs
    if _libs is None:
        base = C.CDLL(str(REF_LIB / "libggml-base.so"), mode=C.RTLD_GLOBAL)
        cpu = C.CDLL(str(REF_LIB / "libggml-cpu.so"), mode=C.RTLD_GLOBAL)
        P, I, F, S, L = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64
        sig = {
            "ggml_init": (P, [_InitParams]), "ggml_free": (None, [P]),
            "ggml_new_tensor": (P, [P, I, I, C.POINTER(L)]),
            "ggml_nbytes": (S, [P]), "ggml_new_graph": (P, [P]), "ggml_build_forward_expand": (None, [P, P]),
            "ggml_mul_mat": (P, [P, P, P]), "ggml_rms_norm": (P, [P, P, F]), "ggml_mul": (P, [P, P, P]), "ggml_add": (P, [P, P, P]),
            "ggml_rope_ext": (P, [P, P, P, P, I, I, I, F, F, F, F, F, F]),
            "ggml_swiglu_split": (P, [P, P, P]),
            "ggml_flash_attn_ext": (P, [P, P, P, P, P, F, F, F][0:5] + [F, F, F]), "ggml_flash_attn_ext_set_prec": (None, [P, I]),
            "ggml_set_rows": (P, [P, P, P, P]), "ggml_get_rows": (P, [P, P, P]),
            "ggml_soft_max_ext": (P, [P, P, P, F, F]),
            "ggml_norm": (P, [P, P, F]), "ggml_im2col": (P, [P, P, P, I, I, I, I, I, I, C.c_bool, I]), "ggml_pool_1d": (P, [P, P, I, I, I, I]),
            "ggml_permute": (P, [P, P, I, I, I, I]), "ggml_cont": (P, [P, P]),
            "ggml_view_3d": (P, [P, P, L, L, L, S, S, S]),
            "ggml_quantize_chunk": (S, [I, P, P, L, L, L, P]),
            # Token2Wav op set
            "ggml_sin": (P, [P, P]), "ggml_cos": (P, [P, P]), "ggml_log": (P, [P, P]), "ggml_elu": (P, [P, P]), "ggml_step": (P, [P, P]), "ggml_sgn": (P, [P, P]),
            "ggml_hardswish": (P, [P, P]), "ggml_hardsigmoid": (P, [P, P]), "ggml_leaky_relu": (P, [P, P, F, C.c_bool]), "ggml_clamp": (P, [P, P, F, F]),
            "ggml_concat": (P, [P, P, P, I]), "ggml_repeat": (P, [P, P, P]), "ggml_arange": (P, [P, F, F, F]), "ggml_sum_rows": (P, [P, P]),
            "ggml_pad_ext": (P, [P, P] + [I] * 8), "ggml_pad_reflect_1d": (P, [P, P, I, I]), "ggml_conv_transpose_1d": (P, [P, P, P, I, I, I]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(base, name)
            fn.restype, fn.argtypes = res, args
        cpu.ggml_graph_compute_with_ctx.restype = I
        cpu.ggml_graph_compute_with_ctx.argtypes = [P, P, I]
        _libs = (base, cpu)
    return _libs


# ---- raw (de)quantisers / dot products --------------------------------------------------------------------------
def quantize(t: int, x: np.ndarray) -> np.ndarray:
    """ggml_quantize_chunk (ggml.c) on an [nrows, k] float32 array -> uint8 [nrows, row_size]."""
    base, _ = libs()
    x = np.ascontiguousarray(x, dtype=np.float32)
    nrows, k = x.shape
    out = np.empty((nrows, row_size(t, k)), dtype=np.uint8)
    n = base.ggml_quantize_chunk(t, x.ctypes.data, out.ctypes.data, 0, nrows, k, None)
    assert n == out.nbytes
    return out


def dequantize(t: int, q: np.ndarray, k: int) -> np.ndarray:
    base, _ = libs()
    q = np.ascontiguousarray(q, dtype=np.uint8)
    nrows = q.size // row_size(t, k)
    out = np.empty((nrows, k), dtype=np.float32)
    fn = getattr(base, f"dequantize_row_{TYPE_NAMES[t]}")
    fn.restype, fn.argtypes = None, [C.c_void_p, C.c_void_p, C.c_int64]
    rs = row_size(t, k)
    for r in range(nrows):
        fn(q.ctypes.data + r * rs, out.ctypes.data + r * k * 4, k)
    return out


def quantize_act(t: int, x: np.ndarray, simd: bool = False) -> np.ndarray:
    """Activation quantiser: q8_0 / q8_K.  simd=True calls the CPU backend's (AVX2) flavour, else the *_ref scalar one."""
    base, cpu = libs()
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    out = np.empty(row_size(t, x.size), dtype=np.uint8)
    name = {Q8_0: "quantize_row_q8_0", Q8_K: "quantize_row_q8_K"}[t]
    fn = getattr(cpu, name) if simd else getattr(base, name + "_ref")
    fn.restype, fn.argtypes = None, [C.c_void_p, C.c_void_p, C.c_int64]
    fn(x.ctypes.data, out.ctypes.data, x.size)
    return out


def vec_dot(wt: int, w: np.ndarray, act: np.ndarray, k: int, generic: bool = False) -> float:
    _, cpu = libs()
    at = "q8_0" if wt in (Q4_0, Q8_0) else "q8_K"
    fn = getattr(cpu, f"ggml_vec_dot_{TYPE_NAMES[wt]}_{at}" + ("_generic" if generic else ""))
    fn.restype = None
    fn.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    s = C.c_float(0)
    fn(k, C.byref(s), 0, w.ctypes.data, 0, act.ctypes.data, 0, 1)
    return s.value


# ---- single-op graphs on the reference CPU backend -------------------------------------------------------------
class Graph:
    """Tiny builder: g = Graph(); a = g.tensor(F32, [k, n], data); out = g.op('ggml_rms_norm', a, eps); g.run(out)."""

    def __init__(self, mem_mb: int = 256):
        self.base, self.cpu = libs()
        self.ctx = self.base.ggml_init(_InitParams(mem_mb << 20, None, False))
        assert self.ctx

    def tensor(self, t: int, ne, data: np.ndarray | None = None):
        ne = list(ne)
        arr = (C.c_int64 * len(ne))(*ne)
        p = self.base.ggml_new_tensor(self.ctx, t, len(ne), arr)
        if data is not None:
            data = np.ascontiguousarray(data)
            nbytes = self.base.ggml_nbytes(p)
            assert data.nbytes == nbytes, (data.nbytes, nbytes)
            C.memmove(_Tensor.from_address(p).data, data.ctypes.data, nbytes)
        return p

    def op(self, name: str, *args):
        fn = getattr(self.base, name)
        return fn(self.ctx, *args)

    def run(self, out, n_threads: int = 1, dtype=np.float32) -> np.ndarray:
        g = self.base.ggml_new_graph(self.ctx)
        self.base.ggml_build_forward_expand(g, out)
        st = self.cpu.ggml_graph_compute_with_ctx(self.ctx, g, n_threads)
        assert st == 0
        t = _Tensor.from_address(out)
        shape = [int(t.ne[i]) for i in range(4)][::-1]
        nbytes = self.base.ggml_nbytes(out)
        buf = (C.c_uint8 * nbytes).from_address(t.data)
        return np.frombuffer(buf, dtype=dtype).reshape(shape).copy()

    def close(self):
        if self.ctx:
            self.base.ggml_free(self.ctx)
            self.ctx = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def mul_mat(wt: int, wq: np.ndarray, x: np.ndarray, m: int, k: int, n_threads: int = 1) -> np.ndarray:
    """dst[n, m] = x[n, k] @ W[m, k]^T with W of ggml type wt given as raw bytes (reference CPU backend, full MUL_MAT path)."""
    n = x.shape[0]
    with Graph() as g:
        a = g.tensor(wt, [k, m], wq)
        b = g.tensor(F32, [k, n], x.astype(np.float32))
        return g.run(g.op("ggml_mul_mat", a, b), n_threads).reshape(n, m)


def rms_norm(x: np.ndarray, eps: float) -> np.ndarray:
    with Graph() as g:
        a = g.tensor(F32, [x.shape[1], x.shape[0]], x)
        return g.run(g.op("ggml_rms_norm", a, C.c_float(eps))).reshape(x.shape)


def rope(x: np.ndarray, pos: np.ndarray, n_dims: int, mode: int, n_ctx_orig: int, freq_base: float, freq_scale=1.0,
         ext_factor=0.0, attn_factor=1.0, beta_fast=32.0, beta_slow=1.0) -> np.ndarray:
    """x: [n_tok, n_head, head_dim] float32."""
    n_tok, n_head, hd = x.shape
    with Graph() as g:
        a = g.tensor(F32, [hd, n_head, n_tok], x)
        p = g.tensor(I32, [n_tok], pos.astype(np.int32))
        out = g.op("ggml_rope_ext", a, p, None, n_dims, mode, n_ctx_orig, C.c_float(freq_base), C.c_float(freq_scale),
                   C.c_float(ext_factor), C.c_float(attn_factor), C.c_float(beta_fast), C.c_float(beta_slow))
        return g.run(out).reshape(x.shape)


def swiglu(gate: np.ndarray, up: np.ndarray) -> np.ndarray:
    with Graph() as g:
        a = g.tensor(F32, [gate.size], gate)
        b = g.tensor(F32, [up.size], up)
        return g.run(g.op("ggml_swiglu_split", a, b)).reshape(gate.shape)


def set_rows_f16(src: np.ndarray, idx: np.ndarray, n_dst_rows: int) -> np.ndarray:
    nrows, ncols = src.shape
    with Graph() as g:
        dst = g.tensor(F16, [ncols, n_dst_rows], np.zeros((n_dst_rows, ncols), np.float16))
        s = g.tensor(F32, [ncols, nrows], src)
        i = g.tensor(I64, [nrows], idx.astype(np.int64))
        return g.run(g.op("ggml_set_rows", dst, s, i), dtype=np.float16).reshape(n_dst_rows, ncols)


def flash_attn(q: np.ndarray, k: np.ndarray, v: np.ndarray, mask: np.ndarray | None, scale: float, n_threads: int = 1) -> np.ndarray:
    """q: F32 [n_q, n_head, D] (llama layout before permute); k, v: F16 [n_head_kv, n_kv, D]; mask: F16 [n_q_pad, n_kv].
    Returns F32 [n_q, n_head, D] — the same call shape llama-graph.cpp:1303-1435 builds (q permuted to [D, n_q, n_head])."""
    n_q, n_head, D = q.shape
    n_head_kv, n_kv, _ = k.shape
    with Graph(512) as g:
        qt = g.tensor(F32, [D, n_head, n_q], q)
        qp = g.op("ggml_permute", qt, 0, 2, 1, 3)                 # -> [D, n_q, n_head]
        kt = g.tensor(F16, [D, n_kv, n_head_kv], k.astype(np.float16))
        vt = g.tensor(F16, [D, n_kv, n_head_kv], v.astype(np.float16))
        mt = g.tensor(F16, [n_kv, mask.shape[0]], mask.astype(np.float16)) if mask is not None else None
        out = g.op("ggml_flash_attn_ext", qp, kt, vt, mt, C.c_float(scale), C.c_float(0.0), C.c_float(0.0))
        g.base.ggml_flash_attn_ext_set_prec(out, 10)              # GGML_PREC_F32, as llama-graph.cpp:1347 does
        return g.run(out, n_threads).reshape(n_q, n_head, D)


# ---- APM / VPM encoder ops ----------------------------------------------------------------------------------------
def norm(x: np.ndarray, eps: float) -> np.ndarray:
    with Graph() as g:
        a = g.tensor(F32, [x.shape[-1], x.size // x.shape[-1]], x)
        return g.run(g.op("ggml_norm", a, C.c_float(eps))).reshape(x.shape)


def im2col(x: np.ndarray, KH: int, KW: int, OC: int, s0, s1, p0, p1, d0, d1, is_2d: bool, f16: bool = True) -> np.ndarray:
    """x F32 [N, IC, IH, IW] (IH = 1 for 1-D) -> [N, OH, OW, IC*KH*KW] through ggml_im2col on the reference CPU backend."""
    N, IC, IH, IW = x.shape
    with Graph() as g:
        if is_2d:
            k = g.tensor(F16, [KW, KH, IC, OC], np.zeros((OC, IC, KH, KW), np.float16))
            b = g.tensor(F32, [IW, IH, IC, N], x)
        else:
            k = g.tensor(F16, [KW, IC, OC], np.zeros((OC, IC, KW), np.float16))
            b = g.tensor(F32, [IW, IC, N], x.reshape(N, IC, IW))
        out = g.op("ggml_im2col", k, b, s0, s1, p0, p1, d0, d1, is_2d, F16 if f16 else F32)
        r = g.run(out, dtype=np.float16 if f16 else np.float32)
    return r.reshape(N, -1, r.shape[-2], r.shape[-1]) if is_2d else r.reshape(N, 1, r.shape[-2], r.shape[-1])


def pool_1d(x: np.ndarray, op: int, k: int) -> np.ndarray:
    with Graph() as g:
        a = g.tensor(F32, [x.shape[-1], x.size // x.shape[-1]], x)
        return g.run(g.op("ggml_pool_1d", a, op, k, k, 0)).reshape(list(x.shape[:-1]) + [x.shape[-1] // k])


# ---- Token2Wav op set: single-op graphs on the reference CPU backend; arrays in numpy order (last axis = ggml dim 0) ------------------------
def _t(g, a: np.ndarray, t: int = F32):
    return g.tensor(t, list(a.shape)[::-1], a)


def wave_unary(name: str, x: np.ndarray, p0: float = 0.0, p1: float = 0.0) -> np.ndarray:
    with Graph() as g:
        a = _t(g, x.astype(np.float32))
        if name == "leaky_relu":
            out = g.op("ggml_leaky_relu", a, C.c_float(p0), False)
        elif name == "clamp":
            out = g.op("ggml_clamp", a, C.c_float(p0), C.c_float(p1))
        else:
            out = g.op("ggml_" + name, a)
        return g.run(out).reshape(x.shape)


def wave_concat(a: np.ndarray, b: np.ndarray, dim: int) -> np.ndarray:
    with Graph() as g:
        r = g.run(g.op("ggml_concat", _t(g, a), _t(g, b), dim))
    return r.reshape(np.concatenate([a, b], axis=a.ndim - 1 - dim).shape)


def wave_repeat(a: np.ndarray, reps) -> np.ndarray:
    shape = [s * r for s, r in zip(a.shape, reps)]
    with Graph() as g:
        like = g.tensor(F32, shape[::-1])
        return g.run(g.op("ggml_repeat", _t(g, a), like)).reshape(shape)


def wave_pad(a: np.ndarray, lp_rp8) -> np.ndarray:
    with Graph() as g:
        r = g.run(g.op("ggml_pad_ext", _t(g, a), *[int(v) for v in lp_rp8]))
    ne = list(a.shape)[::-1] + [1] * (4 - a.ndim)
    out_ne = [ne[d] + lp_rp8[2 * d] + lp_rp8[2 * d + 1] for d in range(4)]
    return r.reshape(out_ne[::-1][4 - a.ndim:])


def wave_pad_reflect_1d(a: np.ndarray, p0: int, p1: int) -> np.ndarray:
    with Graph() as g:
        return g.run(g.op("ggml_pad_reflect_1d", _t(g, a), p0, p1)).reshape(list(a.shape[:-1]) + [a.shape[-1] + p0 + p1])


def wave_arange(start: float, stop: float, step: float) -> np.ndarray:
    with Graph() as g:
        return g.run(g.op("ggml_arange", C.c_float(start), C.c_float(stop), C.c_float(step))).reshape(-1)


def wave_sum_rows(a: np.ndarray) -> np.ndarray:
    with Graph() as g:
        return g.run(g.op("ggml_sum_rows", _t(g, a))).reshape(list(a.shape[:-1]) + [1])


def wave_conv_transpose_1d(w: np.ndarray, x: np.ndarray, s0: int, f16: bool = False) -> np.ndarray:
    """w [Cin, Cout, K], x [Cin, L] -> [Cout, (L - 1) * s0 + K]"""
    with Graph() as g:
        wt = _t(g, w.astype(np.float16), F16) if f16 else _t(g, w.astype(np.float32))
        r = g.run(g.op("ggml_conv_transpose_1d", wt, _t(g, x.astype(np.float32)), s0, 0, 1))
    return r.reshape(w.shape[1], (x.shape[-1] - 1) * s0 + w.shape[2])

"""numpy wrapper over oracle/*.c (liboracle_c.so) — the CPU restatement of the reference's hot-path arithmetic.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; never by the
product package.  Function-by-function reference citations live in oracle/quants_port.c and oracle/ops_port.c.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
SO = ROOT / "oracle" / "_ref" / "lib" / "liboracle_c.so"

F32, F16, Q4_0, Q8_0, Q4_K, Q5_K, Q6_K, Q8_K, BF16 = 0, 1, 2, 8, 12, 13, 14, 15, 30
QUANT_TYPES = (Q4_0, Q8_0, Q4_K, Q5_K, Q6_K)
NAMES = {F32: "f32", F16: "f16", BF16: "bf16", Q4_0: "q4_0", Q8_0: "q8_0", Q4_K: "q4_K", Q5_K: "q5_K", Q6_K: "q6_K"}
BLOCK = {F32: (1, 4), F16: (1, 2), BF16: (1, 2), Q4_0: (32, 18), Q8_0: (32, 34), Q4_K: (256, 144), Q5_K: (256, 176),
         Q6_K: (256, 210), Q8_K: (256, 292)}

_lib = None


def build() -> None:
    """Compile the C restatement (gcc only; no reference sources needed)."""
    subprocess.check_call(["make", "-s", "-C", str(ROOT / "oracle"), "port"])


def lib():
    global _lib
    if _lib is None:
        if not SO.exists():
            build()
        L = C.CDLL(str(SO))
        P, I, F, L64 = C.c_void_p, C.c_int, C.c_float, C.c_int64
        L.or_row_size.restype, L.or_row_size.argtypes = C.c_size_t, [I, L64]
        L.or_dequant_row.restype, L.or_dequant_row.argtypes = I, [I, P, P, L64]
        L.or_quantize_q8_0.restype, L.or_quantize_q8_0.argtypes = None, [P, P, L64, I]
        L.or_quantize_q8_K.restype, L.or_quantize_q8_K.argtypes = None, [P, P, L64]
        L.or_quantize_q4_0.restype, L.or_quantize_q4_0.argtypes = None, [P, P, L64]
        L.or_mul_mat.restype, L.or_mul_mat.argtypes = I, [I, P, P, P, L64, L64, L64, I]
        L.or_rms_norm.restype, L.or_rms_norm.argtypes = None, [P, P, L64, L64, F]
        L.or_rope.restype = None
        L.or_rope.argtypes = [P, P, P, P, L64, L64, L64, I, I, I, F, F, F, F, F, F]
        L.or_swiglu.restype, L.or_swiglu.argtypes = None, [P, P, P, L64]
        L.or_set_rows_f16.restype, L.or_set_rows_f16.argtypes = None, [P, P, P, L64, L64]
        L.or_soft_max.restype, L.or_soft_max.argtypes = None, [P, P, P, L64, L64, F]
        L.or_norm.restype, L.or_norm.argtypes = None, [P, P, L64, L64, F]
        L.or_im2col.restype, L.or_im2col.argtypes = None, [P, P, I] + [L64] * 8 + [I] * 6
        L.or_pool_1d.restype, L.or_pool_1d.argtypes = None, [P, P, L64, L64, I, I]
        L.or_flash_attn_f16.restype = None
        L.or_flash_attn_f16.argtypes = [P, P, P, P, P] + [L64] * 10 + [F, I]
        _lib = L
    return _lib


def row_size(t: int, k: int) -> int:
    b, s = BLOCK[t]
    assert k % b == 0, (t, k)
    return k // b * s


def _p(a: np.ndarray):
    return a.ctypes.data


def dequant(t: int, q: np.ndarray, k: int) -> np.ndarray:
    q = np.ascontiguousarray(q, dtype=np.uint8)
    n = q.size // row_size(t, k) * k
    y = np.empty(n, np.float32)
    assert lib().or_dequant_row(t, _p(q), _p(y), n) == 0
    return y.reshape(-1, k)


def quantize_q8_0(x: np.ndarray, variant: int = 1) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    y = np.empty(row_size(Q8_0, x.size), np.uint8)
    lib().or_quantize_q8_0(_p(x), _p(y), x.size, variant)
    return y


def quantize_q4_0(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    y = np.empty(row_size(Q4_0, x.size), np.uint8)
    lib().or_quantize_q4_0(_p(x), _p(y), x.size)
    return y


def quantize_q8_K(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32).reshape(-1)
    y = np.empty(row_size(Q8_K, x.size), np.uint8)
    lib().or_quantize_q8_K(_p(x), _p(y), x.size)
    return y


def mul_mat(t: int, w: np.ndarray, x: np.ndarray, m: int, k: int, q8_0_variant: int = 1) -> np.ndarray:
    """dst[n, m] = x[n, k] . W[m, k]^T; W raw bytes of ggml type t."""
    w = np.ascontiguousarray(w).view(np.uint8)
    x = np.ascontiguousarray(x, np.float32)
    n = x.shape[0]
    y = np.empty((n, m), np.float32)
    assert lib().or_mul_mat(t, _p(w), _p(x), _p(y), m, k, n, q8_0_variant) == 0
    return y


def rms_norm(x: np.ndarray, eps: float) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    lib().or_rms_norm(_p(x), _p(y), x.shape[-1], x.size // x.shape[-1], eps)
    return y


def rope(x: np.ndarray, pos: np.ndarray, n_dims: int, mode: int, n_ctx_orig: int = 40960, freq_base: float = 1e6,
         freq_scale=1.0, ext_factor=0.0, attn_factor=1.0, beta_fast=32.0, beta_slow=1.0, freq_factors=None) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    n_tok, n_head, hd = x.shape
    pos = np.ascontiguousarray(pos, np.int32)
    y = np.empty_like(x)
    ff = None if freq_factors is None else _p(np.ascontiguousarray(freq_factors, np.float32))
    lib().or_rope(_p(x), _p(y), _p(pos), ff, hd, n_head, n_tok, n_dims, mode, n_ctx_orig, freq_base, freq_scale,
                  ext_factor, attn_factor, beta_fast, beta_slow)
    return y


def swiglu(gate: np.ndarray, up: np.ndarray) -> np.ndarray:
    gate = np.ascontiguousarray(gate, np.float32)
    up = np.ascontiguousarray(up, np.float32)
    y = np.empty_like(gate)
    lib().or_swiglu(_p(gate), _p(up), _p(y), gate.size)
    return y


def set_rows_f16(src: np.ndarray, idx: np.ndarray, dst: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(src, np.float32)
    idx = np.ascontiguousarray(idx, np.int64)
    dst = np.ascontiguousarray(dst, np.float16).copy()
    lib().or_set_rows_f16(_p(src), _p(idx), _p(dst), src.shape[1], src.shape[0])
    return dst


def soft_max(x: np.ndarray, mask: np.ndarray | None, scale: float) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    m = None if mask is None else _p(np.ascontiguousarray(mask, np.float32))
    lib().or_soft_max(_p(x), m, _p(y), x.shape[-1], x.size // x.shape[-1], scale)
    return y


def flash_attn(q: np.ndarray, k: np.ndarray, v: np.ndarray, mask: np.ndarray | None, scale: float, f16_acc: bool = True) -> np.ndarray:
    """q F32 [n_q, n_head, D]; k, v F16 [n_head_kv, n_kv, D]; mask F16 [>=n_q, n_kv] -> F32 [n_q, n_head, D]."""
    q = np.ascontiguousarray(q, np.float32)
    k = np.ascontiguousarray(k, np.float16)
    v = np.ascontiguousarray(v, np.float16)
    n_q, n_head, D = q.shape
    n_head_kv, n_kv, _ = k.shape
    out = np.empty_like(q)
    mp, ms = None, 0
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.float16)
        mp, ms = _p(mask), mask.shape[1]
    lib().or_flash_attn_f16(_p(q), _p(k), _p(v), mp, _p(out), D, n_q, n_head, n_head_kv, n_kv,
                            n_head * D, D, n_kv * D, D, ms, scale, int(f16_acc))
    return out


def norm(x: np.ndarray, eps: float) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    y = np.empty_like(x)
    lib().or_norm(_p(x), _p(y), x.shape[-1], x.size // x.shape[-1], eps)
    return y


def conv_out(i: int, k: int, s: int, p: int, d: int) -> int:       # ggml_calc_conv_output_size
    return (i + 2 * p - d * (k - 1) - 1) // s + 1


def im2col(x: np.ndarray, KH: int, KW: int, s0, s1, p0, p1, d0, d1, f16: bool = True) -> np.ndarray:
    """x F32 [N, IC, IH, IW] -> [N, OH, OW, IC*KH*KW] (F16 by default, as ggml_conv_1d / ggml_conv_2d ask for)."""
    x = np.ascontiguousarray(x, np.float32)
    N, IC, IH, IW = x.shape
    OH, OW = (1 if (IH == 1 and KH == 1) else conv_out(IH, KH, s1, p1, d1)), conv_out(IW, KW, s0, p0, d0)      # 1-D: s1 = p1 = d1 = 0
    y = np.empty((N, OH, OW, IC * KH * KW), np.float16 if f16 else np.float32)
    lib().or_im2col(_p(x), _p(y), int(f16), N, IC, IH, IW, KH, KW, OH, OW, s0, s1, p0, p1, d0, d1)
    return y


def pool_1d(x: np.ndarray, op: int, k: int) -> np.ndarray:
    x = np.ascontiguousarray(x, np.float32)
    n = x.shape[-1]
    y = np.empty(list(x.shape[:-1]) + [n // k], np.float32)
    lib().or_pool_1d(_p(x), _p(y), n, x.size // n, op, k)
    return y

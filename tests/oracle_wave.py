"""TEST INFRASTRUCTURE — numpy restatements of the reference CPU implementations of the Token2Wav op set, each citing the reference
lines it follows (ggml/src/ggml-cpu/ops.cpp, unary-ops.cpp).  Pinned against the live reference by tests/test_oracle_pin.py::test_wave_ops_match_reference (CPU) and
used as the oracle of the GPU tests (tests/test_gpu_parity.py::test_wave_*).  Arrays are in torch / numpy order: the LAST axis is ggml dim 0."""
import numpy as np

F = np.float32


def unary(name: str, x: np.ndarray, p0: float = 0.0, p1: float = 0.0) -> np.ndarray:
    """ggml-cpu/unary-ops.cpp op_* (sin, cos, log, elu, step, sgn, hardswish, hardsigmoid), ops.cpp:2453-2480 (leaky_relu), :5309-5340 (clamp); all in f32"""
    x = x.astype(F)
    if name == "sin": return np.sin(x)
    if name == "cos": return np.cos(x)
    if name == "log": return np.log(x)
    if name == "elu": return np.where(x > 0, x, np.expm1(x)).astype(F)
    if name == "step": return (x > 0).astype(F)
    if name == "sgn": return np.sign(x).astype(F)
    if name == "hardswish": return (x * np.minimum(F(1), np.maximum(F(0), (x + F(3)) / F(6)))).astype(F)
    if name == "hardsigmoid": return np.minimum(F(1), np.maximum(F(0), (x + F(3)) / F(6))).astype(F)
    if name == "leaky_relu": return (np.maximum(x, 0) + F(p0) * np.minimum(x, 0)).astype(F)       # ((x > 0) ? x : 0) + slope * ((x < 0) ? x : 0)
    if name == "clamp": return np.clip(x, F(p0), F(p1))
    raise KeyError(name)


def concat(a: np.ndarray, b: np.ndarray, ggml_dim: int) -> np.ndarray:
    """ops.cpp:1968-2010: dst = [a ; b] along ggml dim"""
    return np.concatenate([a, b], axis=a.ndim - 1 - ggml_dim)


def repeat(a: np.ndarray, reps) -> np.ndarray:
    """ops.cpp:1637-1700: dst[i] = src[i mod ne]"""
    return np.tile(a, reps)


def pad(a: np.ndarray, lp_rp8) -> np.ndarray:
    """ops.cpp:7592-7640: zeros left (lp) and right (rp) of every ggml dim; lp_rp8 = [lp0, rp0, ..., lp3, rp3]"""
    pads = [(lp_rp8[2 * d], lp_rp8[2 * d + 1]) for d in range(4)][:a.ndim][::-1]
    return np.pad(a, pads)


def pad_reflect_1d(a: np.ndarray, p0: int, p1: int) -> np.ndarray:
    """ops.cpp:7664-7692: left[-i] = left[i], right[i] = right[-i]"""
    return np.pad(a, [(0, 0)] * (a.ndim - 1) + [(p0, p1)], mode="reflect")


def arange(start: float, stop: float, step: float) -> np.ndarray:
    """ops.cpp:7762-7785: n = ceil((stop - start) / step) in f32, value = start + step * i in f32"""
    n = int(np.ceil((F(stop) - F(start)) / F(step)))
    return (F(start) + F(step) * np.arange(n, dtype=F)).astype(F)


def sum_rows(a: np.ndarray) -> np.ndarray:
    """ops.cpp:1399-1430: ggml_vec_sum_f32 accumulates the row in ggml_float (double), rounded once"""
    return a.astype(np.float64).sum(-1, keepdims=True).astype(F)


def conv_transpose_1d(w: np.ndarray, x: np.ndarray, s0: int) -> np.ndarray:
    """ops.cpp:6040-6130 (p0 = 0, d0 = 1): w [Cin, Cout, K] (ggml ne [K, Cout, Cin]), x [Cin, L] -> [Cout, (L - 1) * s0 + K]; dst[co, l * s0 + k] += sum_ci x[ci, l] * w[ci, co, k]
    (f64 here: the reference's f32 order is one of many valid ones)"""
    Cin, Cout, K = w.shape
    L = x.shape[-1]
    out = np.zeros((Cout, (L - 1) * s0 + K))
    w64, x64 = w.astype(np.float64), x.astype(np.float64)
    for k in range(K):
        out[:, k:k + (L - 1) * s0 + 1:s0] += np.einsum("ic,il->cl", w64[:, :, k], x64)
    return out

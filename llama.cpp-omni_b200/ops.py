"""ctypes mirror of include/b200_ops.h (libb200ops.so).  Arguments are torch CUDA tensors described the ggml way:
ne[0] is the contiguous dimension (= the LAST torch dim), nb[] are byte strides.  No CPU fallback exists."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
SO = HERE / "lib" / "libb200ops.so"

F32, F16, Q4_0, Q8_0, Q4_K, Q5_K, Q6_K, I32, I64, BF16 = 0, 1, 2, 8, 12, 13, 14, 26, 27, 30
QUANT = (Q4_0, Q8_0, Q4_K, Q5_K, Q6_K)
BLOCK = {F32: (1, 4), F16: (1, 2), BF16: (1, 2), I32: (1, 4), I64: (1, 8), Q4_0: (32, 18), Q8_0: (32, 34), Q4_K: (256, 144),
         Q5_K: (256, 176), Q6_K: (256, 210)}
PAYLOAD = {Q4_0: 16, Q8_0: 32, Q6_K: 208}
LAYOUT_NATIVE, LAYOUT_PLANAR = 0, 1
_TORCH2B = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16, torch.int32: I32, torch.int64: I64}

EXPORTS = ["b200_abi_version", "b200_error_string", "b200_device_sm_count", "b200_repack_supported", "b200_repack_scatter",
           "b200_repack_gather", "b200_act_bytes", "b200_quantize_act", "b200_mul_mat_supported", "b200_mul_mat_scratch_bytes",
           "b200_mul_mat", "b200_mul_mat_ex", "b200_matvec_q", "b200_matvec_q_swiglu", "b200_rms_norm", "b200_rms_norm_quantize", "b200_rope",
           "b200_set_rows", "b200_get_rows", "b200_cpy", "b200_binary", "b200_unary", "b200_glu", "b200_scale", "b200_soft_max",
           "b200_flash_attn_supported", "b200_flash_attn_scratch_bytes", "b200_flash_attn", "b200_qkv_post",
           "b200_decoder_create", "b200_decoder_step", "b200_decoder_n_phases", "b200_decoder_profile", "b200_decoder_destroy",
           "b200_norm", "b200_im2col", "b200_pool_1d", "b200_rms_norm_tiles", "b200_glu_tiles", "b200_flash_attn_tiles", "b200_mul_mat_add", "b200_mul_mat_multi", "b200_mul_mat_multi_merges", "b200_mul_mat_glu", "b200_unary_param", "b200_concat", "b200_repeat", "b200_arange", "b200_sum_rows", "b200_pad", "b200_pad_reflect_1d", "b200_conv_transpose_1d", "b200_ipc_alloc", "b200_ipc_open", "b200_ipc_close", "b200_ipc_free", "b200_hop_send", "b200_hop_wait", "b200_hop_ack"]


class Tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("type", C.c_int32), ("layout", C.c_int32), ("ne", C.c_int64 * 4), ("nb", C.c_int64 * 4)]


class MatvecJob(C.Structure):
    _fields_ = [("w", C.c_void_p), ("type", C.c_int32), ("layout", C.c_int32), ("m", C.c_int64), ("row_stride_bytes", C.c_int64),
                ("y", C.c_void_p), ("residual", C.c_void_p)]


class RopeParams(C.Structure):
    _fields_ = [("n_dims", C.c_int32), ("mode", C.c_int32), ("n_ctx_orig", C.c_int32), ("freq_base", C.c_float),
                ("freq_scale", C.c_float), ("ext_factor", C.c_float), ("attn_factor", C.c_float), ("beta_fast", C.c_float),
                ("beta_slow", C.c_float)]


class Weight(C.Structure):
    _fields_ = [("data", C.c_void_p), ("type", C.c_int32), ("layout", C.c_int32)]


class DecodeLayer(C.Structure):
    _fields_ = [(n, Weight) for n in ("wq", "wk", "wv", "wo", "gate", "up", "down")] + \
               [(n, C.c_void_p) for n in ("attn_norm", "ffn_norm", "q_norm", "k_norm", "k_cache", "v_cache")] + \
               [("k_row_bytes", C.c_int64), ("v_row_bytes", C.c_int64)]


class DecodeDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_layer", "n_embd", "n_head", "n_head_kv", "head_dim", "n_ff", "n_vocab")] + \
               [("rms_eps", C.c_float), ("attn_scale", C.c_float), ("rope", RopeParams), ("layers", C.POINTER(DecodeLayer)),
                ("out_norm", C.c_void_p), ("lm_head", Weight)] + \
               [(n, C.c_void_p) for n in ("x_in", "pos", "kv_idx", "mask", "logits", "hidden_out", "x_out")]


class B200Error(RuntimeError):
    pass


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is None:
        if not SO.exists():
            raise B200Error(f"{SO} is missing: build it with `make -C {HERE / 'csrc'}` — there is no fallback path")
        L = C.CDLL(str(SO))
        L.b200_error_string.restype = C.c_char_p
        L.b200_act_bytes.restype = C.c_size_t
        L.b200_act_bytes.argtypes = [C.c_int, C.c_int64]
        L.b200_mul_mat_scratch_bytes.restype = C.c_size_t
        L.b200_flash_attn_scratch_bytes.restype = C.c_size_t
        L.b200_decoder_destroy.restype = None
        L.b200_decoder_destroy.argtypes = [C.c_void_p]
        L.b200_decoder_step.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.b200_decoder_n_phases.argtypes = [C.c_void_p]
        L.b200_decoder_profile.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        L.b200_ipc_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]
        L.b200_ipc_open.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        L.b200_ipc_close.argtypes = [C.c_void_p]
        L.b200_ipc_free.argtypes = [C.c_void_p]
        L.b200_hop_send.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200_hop_wait.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.b200_hop_ack.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise B200Error(f"b200 call failed ({rc}): {lib().b200_error_string(rc).decode()}")


def stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def T(t: torch.Tensor | None, type: int | None = None, ne=None, nb=None, layout: int = LAYOUT_NATIVE):
    """b200_tensor view of a torch tensor.  Default: ne = reversed shape, nb from torch strides.  For quantised weights pass
    a uint8 tensor plus type and ne=[k, m(, ...)]; nb is then the packed ggml stride."""
    if t is None:
        return None
    if isinstance(t, Tensor):                         # a ready descriptor (quantised tensors: T(u8, Q8_0, ne=[...], nb=[...]))
        return t
    d = Tensor()
    d.data = t.data_ptr()
    d.layout = layout
    if type is None:
        type = _TORCH2B[t.dtype]
    d.type = type
    if ne is None:
        shape = list(t.shape)[::-1]
        strides = [s * t.element_size() for s in t.stride()][::-1]
        assert len(shape) <= 4
        while len(shape) < 4:
            shape.append(1)
            strides.append(strides[-1] * shape[-2] if strides else t.element_size())
        ne, nb = shape, strides
    else:
        ne = list(ne) + [1] * (4 - len(ne))
        if nb is None:
            blk, bs = BLOCK[type]
            nb = [bs, ne[0] // blk * bs]
            nb.append(nb[1] * ne[1])
            nb.append(nb[2] * ne[2])
    for i in range(4):
        d.ne[i] = int(ne[i])
        d.nb[i] = int(nb[i])
    return d


def _ref(d):
    return C.byref(d) if d is not None else None


def row_size(t: int, k: int) -> int:
    blk, bs = BLOCK[t]
    assert k % blk == 0
    return k // blk * bs


# ---- ops (each returns the output tensor) ---------------------------------------------------------------------------------
def to_planar(wtype: int, w_native: torch.Tensor) -> torch.Tensor:
    """Re-lay a native quantised tensor (uint8, any shape) into the planar layout, through the chunked scatter entry point."""
    L = lib()
    n = w_native.numel()
    nblocks = n // BLOCK[wtype][1]
    out = torch.empty(n, dtype=torch.uint8, device=w_native.device)
    flat = w_native.reshape(-1)
    step = 1 << 20                                   # 1 MiB chunks like llama-model-loader.cpp:1077-1093
    for off in range(0, n, step):
        sz = min(step, n - off)
        check(L.b200_repack_scatter(wtype, C.c_void_p(flat.data_ptr() + off), C.c_void_p(out.data_ptr()), C.c_int64(nblocks),
                                    C.c_int64(off), C.c_int64(sz), stream()))
    return out


def from_planar(wtype: int, w_planar: torch.Tensor) -> torch.Tensor:
    L = lib()
    n = w_planar.numel()
    nblocks = n // BLOCK[wtype][1]
    out = torch.empty(n, dtype=torch.uint8, device=w_planar.device)
    step = (1 << 20) + 6
    for off in range(0, n, step):
        sz = min(step, n - off)
        check(L.b200_repack_gather(wtype, C.c_void_p(w_planar.data_ptr()), C.c_void_p(out.data_ptr() + off), C.c_int64(nblocks),
                                   C.c_int64(off), C.c_int64(sz), stream()))
    return out


def quantize_act(wtype: int, x: torch.Tensor) -> torch.Tensor:
    """x F32 [n, k] -> activation records uint8 [n, act_bytes]."""
    L = lib()
    n, k = x.shape
    ab = L.b200_act_bytes(wtype, k)
    act = torch.empty((n, ab), dtype=torch.uint8, device=x.device)
    check(L.b200_quantize_act(wtype, C.c_void_p(x.data_ptr()), C.c_int64(x.stride(0)), C.c_void_p(act.data_ptr()), C.c_int64(k),
                              C.c_int64(n), stream()))
    return act


def mul_mat(w: torch.Tensor, wtype: int, m: int, k: int, x: torch.Tensor, layout: int = LAYOUT_NATIVE, w_ne=None, w_nb=None,
            out: torch.Tensor | None = None, scratch: torch.Tensor | None = None, reuse_act: bool = False) -> torch.Tensor:
    """dst[..., n, m] = x[..., n, k] . W[m, k]^T   (GGML_OP_MUL_MAT).  `scratch` + `reuse_act`: b200_mul_mat_ex with B200_MM_REUSE_ACT — the caller
    passes the SAME scratch as for the previous call on the same x (q/k/v, gate/up), whose prepared activations are then not rebuilt."""
    L = lib()
    wd = T(w, wtype, ne=w_ne or [k, m], nb=w_nb, layout=layout)
    xd = T(x)
    if out is None:
        out = torch.empty(list(x.shape[:-1]) + [m], dtype=torch.float32, device=x.device)
    od = T(out)
    if not L.b200_mul_mat_supported(C.byref(wd), C.byref(xd), C.byref(od)):
        raise B200Error("mul_mat: unsupported")
    sb = L.b200_mul_mat_scratch_bytes(C.byref(wd), C.byref(xd))
    if scratch is None or scratch.numel() < sb:
        if reuse_act:
            raise B200Error("mul_mat: reuse_act needs the scratch of the previous call")
        scratch = torch.empty(max(sb, 16), dtype=torch.uint8, device=x.device)
    check(L.b200_mul_mat_ex(C.byref(wd), C.byref(xd), C.byref(od), C.c_void_p(scratch.data_ptr()), C.c_size_t(scratch.numel()), int(reuse_act), stream()))
    return out


def mul_mat_add(w: torch.Tensor, wtype: int, m: int, k: int, x: torch.Tensor, residual: torch.Tensor, out: torch.Tensor, layout: int = LAYOUT_NATIVE,
                scratch: torch.Tensor | None = None, reuse_act: bool = False) -> torch.Tensor:
    """out = x . W^T + residual (b200_mul_mat_add: the ADD rides in the tensor-core GEMM's epilogue when it can)."""
    L = lib()
    wd, xd, rd, od = T(w, wtype, ne=[k, m], layout=layout), T(x), T(residual), T(out)
    sb = L.b200_mul_mat_scratch_bytes(C.byref(wd), C.byref(xd))
    if scratch is None or scratch.numel() < sb:
        if reuse_act:
            raise B200Error("mul_mat_add: reuse_act needs the scratch that holds the tiles")
        scratch = torch.empty(max(sb, 16), dtype=torch.uint8, device=x.device)
    check(L.b200_mul_mat_add(C.byref(wd), C.byref(xd), C.byref(rd), C.byref(od), C.c_void_p(scratch.data_ptr()), C.c_size_t(scratch.numel()), int(reuse_act), stream()))
    return out


def mul_mat_glu(op: int, w_gate: torch.Tensor, w_up: torch.Tensor, wtype: int, m: int, k: int, x: torch.Tensor, layout: int = LAYOUT_NATIVE) -> torch.Tensor:
    """out[m] = glu(op)(W_gate . x) * (W_up . x) for ONE activation column (b200_mul_mat_glu: the decode graph's gate / up / SWIGLU triple in one launch)."""
    L = lib()
    gd, ud, xd = T(w_gate, wtype, ne=[k, m], layout=layout), T(w_up, wtype, ne=[k, m], layout=layout), T(x)
    out = torch.empty(list(x.shape[:-1]) + [m], dtype=torch.float32, device=x.device)
    sb = L.b200_mul_mat_scratch_bytes(C.byref(gd), C.byref(xd))
    scratch = torch.empty(max(sb, 16), dtype=torch.uint8, device=x.device)
    check(L.b200_mul_mat_glu(op, C.byref(gd), C.byref(ud), C.byref(xd), _ref(T(out)), C.c_void_p(scratch.data_ptr()), C.c_size_t(scratch.numel()), 0, stream()))
    return out


def mul_mat_multi(ws, x: torch.Tensor, outs, scratch: torch.Tensor | None = None, reuse_act: bool = False):
    """outs[i] = x . W_i^T for 2-3 weights over the same activations (b200_mul_mat_multi).  ws: list of (tensor, wtype, m, layout)."""
    L = lib()
    k = x.shape[-1]
    wds = [T(w, t, ne=[k, m], layout=lay) for (w, t, m, lay) in ws]
    xd = T(x)
    ods = [T(o) for o in outs]
    sb = max(L.b200_mul_mat_scratch_bytes(C.byref(wd), C.byref(xd)) for wd in wds)
    if scratch is None or scratch.numel() < sb:
        if reuse_act:
            raise B200Error("mul_mat_multi: reuse_act needs the scratch that holds the tiles")
        scratch = torch.empty(max(sb, 16), dtype=torch.uint8, device=x.device)
    WP = (C.POINTER(Tensor) * len(wds))(*[C.pointer(wd) for wd in wds])
    DP = (C.POINTER(Tensor) * len(ods))(*[C.pointer(od) for od in ods])
    check(L.b200_mul_mat_multi(len(wds), WP, C.byref(xd), DP, C.c_void_p(scratch.data_ptr()), C.c_size_t(scratch.numel()), int(reuse_act), stream()))
    return outs


def make_job(w: torch.Tensor, wtype: int, m: int, k: int, y: torch.Tensor, residual: torch.Tensor | None = None,
             layout: int = LAYOUT_NATIVE) -> MatvecJob:
    j = MatvecJob()
    j.w, j.type, j.layout, j.m, j.row_stride_bytes = w.data_ptr(), wtype, layout, m, row_size(wtype, k)
    j.y = y.data_ptr()
    j.residual = residual.data_ptr() if residual is not None else None
    return j


def matvec_q(jobs: list[MatvecJob], act: torch.Tensor, k: int) -> None:
    arr = (MatvecJob * len(jobs))(*jobs)
    check(lib().b200_matvec_q(arr, len(jobs), C.c_void_p(act.data_ptr()), C.c_int64(k), stream()))


def matvec_q_swiglu(gate: MatvecJob, up: MatvecJob, y: torch.Tensor, act: torch.Tensor, k: int) -> None:
    check(lib().b200_matvec_q_swiglu(C.byref(gate), C.byref(up), C.c_void_p(y.data_ptr()), C.c_void_p(act.data_ptr()), C.c_int64(k), stream()))


def rms_norm(x: torch.Tensor, eps: float, w: torch.Tensor | None = None, add: torch.Tensor | None = None,
             out: torch.Tensor | None = None) -> torch.Tensor:
    out = torch.empty_like(x) if out is None else out
    check(lib().b200_rms_norm(_ref(T(x)), _ref(T(w)), _ref(T(add)), _ref(T(out)), C.c_float(eps), stream()))
    return out


def rms_norm_quantize(x: torch.Tensor, w: torch.Tensor, wtype: int, eps: float, want_f32: bool = False):
    """x F32 [n, k] -> (act records [n, act_bytes], y F32 or None)."""
    L = lib()
    n, k = x.shape
    ab = L.b200_act_bytes(wtype, k)
    act = torch.empty((n, ab), dtype=torch.uint8, device=x.device)
    y = torch.empty_like(x) if want_f32 else None
    check(L.b200_rms_norm_quantize(C.c_void_p(x.data_ptr()), C.c_int64(x.stride(0)), C.c_void_p(w.data_ptr()),
                                   C.c_void_p(y.data_ptr() if want_f32 else None), C.c_int64(k), C.c_void_p(act.data_ptr()), wtype,
                                   C.c_int64(k), C.c_int64(n), C.c_float(eps), stream()))
    return act, y


def rope(x: torch.Tensor, pos: torch.Tensor, n_dims: int, mode: int, n_ctx_orig: int = 40960, freq_base: float = 1e6,
         freq_scale: float = 1.0, ext_factor: float = 0.0, attn_factor: float = 1.0, beta_fast: float = 32.0, beta_slow: float = 1.0,
         freq_factors: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """x F32 [n_tok, n_head, head_dim], pos I32 [n_tok]."""
    out = torch.empty_like(x) if out is None else out
    p = RopeParams(n_dims, mode, n_ctx_orig, freq_base, freq_scale, ext_factor, attn_factor, beta_fast, beta_slow)
    ff = C.c_void_p(freq_factors.data_ptr()) if freq_factors is not None else None
    check(lib().b200_rope(_ref(T(x)), C.c_void_p(pos.data_ptr()), ff, _ref(T(out)), C.byref(p), stream()))
    return out


def set_rows(src: torch.Tensor, idx: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    check(lib().b200_set_rows(_ref(T(src)), _ref(T(idx)), _ref(T(dst)), stream()))
    return dst


def get_rows(src, idx: torch.Tensor) -> torch.Tensor:
    ne0 = int(src.ne[0]) if isinstance(src, Tensor) else src.shape[-1]          # a Tensor descriptor: quantised rows (native or planar)
    out = torch.empty(list(idx.shape) + [ne0], dtype=torch.float32, device=idx.device)
    check(lib().b200_get_rows(_ref(T(src)), _ref(T(idx)), _ref(T(out)), stream()))
    return out


def cpy(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    check(lib().b200_cpy(_ref(T(src)), _ref(T(dst)), stream()))
    return dst


ADD, SUB, MUL, DIV = 0, 1, 2, 3


def binary(op: int, a: torch.Tensor, b: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    out = torch.empty_like(a) if out is None else out
    check(lib().b200_binary(op, _ref(T(a)), _ref(T(b)), _ref(T(out)), stream()))
    return out


SILU, GELU, RELU, GELU_QUICK, TANH, SIGMOID, GELU_ERF, NEG, EXP, SQR, SQRT, ABS = range(12)


def unary(op: int, x: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(x)
    check(lib().b200_unary(op, _ref(T(x)), _ref(T(out)), stream()))
    return out


GLU_REGLU, GLU_GEGLU, GLU_SWIGLU, GLU_GEGLU_ERF, GLU_GEGLU_QUICK = 0, 1, 2, 4, 5


def glu(op: int, gate: torch.Tensor, up: torch.Tensor | None = None, swapped: bool = False) -> torch.Tensor:
    shape = list(gate.shape)
    if up is None:
        shape[-1] //= 2
    out = torch.empty(shape, dtype=gate.dtype, device=gate.device)
    check(lib().b200_glu(op, _ref(T(gate)), _ref(T(up)), _ref(T(out)), int(swapped), stream()))
    return out


def scale(x: torch.Tensor, s: float, b: float = 0.0) -> torch.Tensor:
    out = torch.empty_like(x)
    check(lib().b200_scale(_ref(T(x)), _ref(T(out)), C.c_float(s), C.c_float(b), stream()))
    return out


def soft_max(x: torch.Tensor, mask: torch.Tensor | None, s: float) -> torch.Tensor:
    out = torch.empty_like(x)
    check(lib().b200_soft_max(_ref(T(x)), _ref(T(mask)), _ref(T(out)), C.c_float(s), C.c_float(0.0), stream()))
    return out


def flash_attn(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: torch.Tensor | None, s: float,
               out: torch.Tensor | None = None, scratch: torch.Tensor | None = None) -> torch.Tensor:
    """q F32 [n_head, n_q, D] (any strides: pass the permuted view llama builds), k/v F16 [n_head_kv, n_kv, D] (any row strides),
    mask F16 [n_q_pad, n_kv] or None -> F32 [n_q, n_head, D]."""
    L = lib()
    n_head, n_q, D = q.shape[-3:]                      # k / v may be Tensor descriptors of a quantised cache (q8_0 / q4_0)
    if out is None:
        out = torch.empty(list(q.shape[:-3]) + [n_q, n_head, D], dtype=torch.float32, device=q.device)
    qd, kd, vd, md, od = T(q), T(k), T(v), T(mask), T(out)
    if not L.b200_flash_attn_supported(_ref(qd), _ref(kd), _ref(vd), _ref(md), _ref(od)):
        raise B200Error("flash_attn: unsupported")
    sb = L.b200_flash_attn_scratch_bytes(_ref(qd), _ref(kd))
    if scratch is None:
        scratch = torch.empty(max(sb, 16), dtype=torch.uint8, device=q.device)
    check(L.b200_flash_attn(_ref(qd), _ref(kd), _ref(vd), _ref(md), _ref(od), C.c_float(s), C.c_float(0.0), C.c_float(0.0),
                            C.c_void_p(scratch.data_ptr()), C.c_size_t(scratch.numel()), stream()))
    return out


def norm(x: torch.Tensor, eps: float) -> torch.Tensor:
    """GGML_OP_NORM over the last dim."""
    out = torch.empty_like(x)
    check(lib().b200_norm(_ref(T(x)), _ref(T(out)), C.c_float(eps), stream()))
    return out


def im2col(kernel: torch.Tensor, x: torch.Tensor, s0: int, s1: int, p0: int, p1: int, d0: int, d1: int, is_2d: bool, dtype=torch.float16) -> torch.Tensor:
    """GGML_OP_IM2COL.  2-D: kernel [OC, IC, KH, KW], x [N, IC, IH, IW] -> [N, OH, OW, IC*KH*KW]; 1-D: kernel [OC, IC, KW], x [N, IC, IW] -> [N, OW, IC*KW]."""
    def osz(i, k, s, p, d):
        return (i + 2 * p - d * (k - 1) - 1) // s + 1
    if is_2d:
        N, IC, IH, IW = x.shape
        KH, KW = kernel.shape[-2:]
        out = torch.empty((N, osz(IH, KH, s1, p1, d1), osz(IW, KW, s0, p0, d0), IC * KH * KW), dtype=dtype, device=x.device)
    else:
        N, IC, IW = x.shape
        KW = kernel.shape[-1]
        out = torch.empty((N, osz(IW, KW, s0, p0, d0), IC * KW), dtype=dtype, device=x.device)
    check(lib().b200_im2col(_ref(T(kernel)), _ref(T(x)), _ref(T(out)), s0, s1, p0, p1, d0, d1, int(is_2d), stream()))
    return out


def pool_1d(x: torch.Tensor, op: int, k: int) -> torch.Tensor:
    """GGML_OP_POOL_1D along the last dim, kernel == stride, no padding (op 0 = max, 1 = avg)."""
    out = torch.empty(list(x.shape[:-1]) + [x.shape[-1] // k], dtype=torch.float32, device=x.device)
    check(lib().b200_pool_1d(_ref(T(x)), _ref(T(out)), op, k, k, 0, stream()))
    return out


# ---- Token2Wav op set (csrc/ops_wave.cu; torch shapes are the reversed ggml ne) ---------------------------------------------------------------------------
SIN, COS, LOG, ELU, STEP, SGN, HARDSWISH, HARDSIGMOID, LEAKY_RELU, CLAMP = range(12, 22)


def unary_param(op: int, x: torch.Tensor, p0: float = 0.0, p1: float = 0.0) -> torch.Tensor:
    out = torch.empty_like(x)
    check(lib().b200_unary_param(op, _ref(T(x)), _ref(T(out)), C.c_float(p0), C.c_float(p1), stream()))
    return out


def concat(a: torch.Tensor, b: torch.Tensor, ggml_dim: int) -> torch.Tensor:
    out = torch.empty(torch.cat([a, b], dim=a.dim() - 1 - ggml_dim).shape, dtype=a.dtype, device=a.device)
    check(lib().b200_concat(_ref(T(a)), _ref(T(b)), _ref(T(out)), ggml_dim, stream()))
    return out


def repeat(x: torch.Tensor, shape) -> torch.Tensor:
    out = torch.empty(shape, dtype=x.dtype, device=x.device)
    check(lib().b200_repeat(_ref(T(x)), _ref(T(out)), stream()))
    return out


def arange(start: float, stop: float, step: float, device) -> torch.Tensor:
    import math
    out = torch.empty(int(math.ceil((stop - start) / step)), dtype=torch.float32, device=device)
    check(lib().b200_arange(_ref(T(out)), C.c_float(start), C.c_float(step), stream()))
    return out


def sum_rows(x: torch.Tensor) -> torch.Tensor:
    out = torch.empty(list(x.shape[:-1]) + [1], dtype=torch.float32, device=x.device)
    check(lib().b200_sum_rows(_ref(T(x)), _ref(T(out)), stream()))
    return out


def pad(x: torch.Tensor, lp_rp8) -> torch.Tensor:
    """lp_rp8 = [lp0, rp0, ..., lp3, rp3] in ggml dimension order (dim 0 = the last torch dim)."""
    ne = list(x.shape)[::-1] + [1] * (4 - x.dim())
    out_ne = [ne[d] + lp_rp8[2 * d] + lp_rp8[2 * d + 1] for d in range(4)]
    out = torch.empty(out_ne[::-1][4 - x.dim():], dtype=torch.float32, device=x.device)
    arr = (C.c_int32 * 8)(*lp_rp8)
    check(lib().b200_pad(_ref(T(x)), _ref(T(out)), arr, stream()))
    return out


def pad_reflect_1d(x: torch.Tensor, p0: int, p1: int) -> torch.Tensor:
    out = torch.empty(list(x.shape[:-1]) + [x.shape[-1] + p0 + p1], dtype=torch.float32, device=x.device)
    check(lib().b200_pad_reflect_1d(_ref(T(x)), _ref(T(out)), p0, p1, stream()))
    return out


def conv_transpose_1d(kernel: torch.Tensor, x: torch.Tensor, s0: int) -> torch.Tensor:
    """kernel [Cin, Cout, K] (ggml ne [K, Cout, Cin]) F32 / F16, x [Cin, L] F32 -> [Cout, (L - 1) * s0 + K]."""
    Cin, Cout, K = kernel.shape
    out = torch.empty((Cout, (x.shape[-1] - 1) * s0 + K), dtype=torch.float32, device=x.device)
    check(lib().b200_conv_transpose_1d(_ref(T(kernel)), _ref(T(x)), _ref(T(out)), s0, stream()))
    return out


# ---- producers that write the following tensor-core MUL_MAT's activation tiles directly (include/b200_ops.h): pass the MUL_MAT's `scratch`, then call
# mul_mat(..., scratch=scratch, reuse_act=True)
def rms_norm_tiles(x: torch.Tensor, eps: float, w: torch.Tensor | None, tiles: torch.Tensor, out: torch.Tensor | None = None) -> None:
    dst = T(out) if out is not None else T(x)
    if out is None:
        dst.data = None
    check(lib().b200_rms_norm_tiles(_ref(T(x)), _ref(T(w)), _ref(dst), C.c_void_p(tiles.data_ptr()), C.c_float(eps), stream()))


def glu_tiles(op: int, gate: torch.Tensor, up: torch.Tensor, tiles: torch.Tensor) -> None:
    check(lib().b200_glu_tiles(op, _ref(T(gate)), _ref(T(up)), C.c_void_p(tiles.data_ptr()), stream()))


def flash_attn_tiles(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, mask: torch.Tensor | None, s: float, out_shape_like: torch.Tensor, scratch: torch.Tensor,
                     tiles: torch.Tensor) -> None:
    """as flash_attn, but the [n_q, n_head * D] result goes to `tiles` (the wo MUL_MAT's scratch) as F16 activation tiles; out_shape_like only describes the shape."""
    qd, kd, vd, md, od = T(q), T(k), T(v), T(mask), T(out_shape_like)
    check(lib().b200_flash_attn_tiles(_ref(qd), _ref(kd), _ref(vd), _ref(md), _ref(od), C.c_float(s), C.c_void_p(scratch.data_ptr()), C.c_size_t(scratch.numel()),
                                      C.c_void_p(tiles.data_ptr()), stream()))

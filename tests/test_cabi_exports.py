"""CPU: the C-ABI library loads and exports every symbol include/b200_ops.h declares (no compute calls without a GPU),
and the host-side accounting (Q4_K_M type recipe, algorithmic bytes/token) matches SURVEY.md §8d."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from __graft_entry__ import load_package, ROOT

HDR = ROOT / "include" / "b200_ops.h"
SO = ROOT / "llama.cpp-omni_b200" / "lib" / "libb200ops.so"


def declared():
    txt = HDR.read_text()
    return sorted(set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", txt)))


def test_header_declares_what_python_binds():
    ops = load_package().ops
    assert set(ops.EXPORTS) == set(declared())


@pytest.mark.skipif(not SO.exists(), reason="libb200ops.so not built (run __graft_entry__.build())")
def test_library_exports_every_declared_symbol():
    out = subprocess.check_output(["nm", "-D", "--defined-only", str(SO)], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = [s for s in declared() if s not in exported]
    assert not missing, missing
    extra = [s for s in exported if not s.startswith("b200_")]
    assert not extra, f"non-API symbols leak from the library: {extra[:5]}"
    lib = C.CDLL(str(SO))
    assert lib.b200_abi_version() == 3
    lib.b200_error_string.restype = C.c_char_p
    assert b"unsupported" in lib.b200_error_string(-1000)
    lib.b200_act_bytes.restype, lib.b200_act_bytes.argtypes = C.c_size_t, [C.c_int, C.c_int64]
    assert lib.b200_act_bytes(12, 4096) == 4096 + 4 * 16 + 2 * 256            # q8_K record: qs | f32 d per 256 | i16 bsum per 16
    assert lib.b200_act_bytes(8, 4096) == 4096 + 2 * 128 + 2 * 128             # q8_0 record: qs | f16 d per 32 | i16 sum per 32


def test_no_fallback_when_library_missing(monkeypatch, tmp_path):
    ops = load_package().ops
    monkeypatch.setattr(ops, "_lib", None)
    monkeypatch.setattr(ops, "SO", tmp_path / "nope.so")
    with pytest.raises(ops.B200Error):
        ops.lib()


def test_q4_k_m_recipe_and_bytes_per_token():
    dec = load_package().decode
    cfg = dec.LLMConfig()
    hi = [il for il in range(cfg.n_layer) if dec.use_more_bits(il, cfg.n_layer)]
    assert len(hi) == 18                                                       # 18 of 36 layers carry Q6_K wv / ffn_down
    assert dec.weight_bytes_per_token(cfg) == 4_671_134_848
    assert dec.kv_bytes_per_token(cfg, 4096) == 147_456 * 4097
    # layer-split sharding covers every layer exactly once
    for world in (1, 2, 4, 8):
        per = (cfg.n_layer + world - 1) // world
        cover = [il for r in range(world) for il in range(r * per, min(cfg.n_layer, (r + 1) * per))]
        assert cover == list(range(cfg.n_layer))
        total = sum(dec.weight_bytes_per_token(cfg, range(r * per, min(cfg.n_layer, (r + 1) * per)), with_head=(r == world - 1)) for r in range(world))
        assert total == 4_671_134_848


@pytest.mark.skipif(not SO.exists(), reason="libb200ops.so not built (run __graft_entry__.build())")
def test_mul_mat_routing_predicates_for_float_operands():
    """b200_mul_mat_supported / b200_mul_mat_scratch_bytes are host-only predicates (the plugin's supports_op calls them on tensors without buffers,
    llama-model.cpp:286-291), so they can be pinned without a GPU: F16 ACTIVATIONS only together with F16 weights (ggml_conv_1d / conv_2d: im2col x kernel — the
    one F16 x F16 product the reference CPU backend has, ggml-cpu.c type_traits_cpu[F16].vec_dot_type), and the F16 tensor-core GEMM's tile scratch rounds K up to its 64-wide
    tile (SigLip ffn_down: k = 4304)."""
    ops = load_package().ops
    L = ops.lib()

    def desc(type_, ne, elem, nb0=None):
        d = ops.Tensor()
        d.data, d.type, d.layout = 0x10000000, type_, ops.LAYOUT_NATIVE
        ne = list(ne) + [1] * (4 - len(ne))
        nb = [nb0 or elem, (nb0 or elem) * ne[0]]
        nb += [nb[1] * ne[1], nb[1] * ne[1] * ne[2]]
        for i in range(4):
            d.ne[i], d.nb[i] = ne[i], nb[i]
        return d

    sup = lambda w, x, y: L.b200_mul_mat_supported(C.byref(w), C.byref(x), C.byref(y))
    L.b200_mul_mat_scratch_bytes.restype = C.c_size_t
    # conv1 of the Whisper encoder: im2col [240, 100] F16 x kernel [240, 1024] F16 -> [100, 1024] F32
    w16, x16, y = desc(ops.F16, [240, 100], 2), desc(ops.F16, [240, 1024], 2), desc(ops.F32, [100, 1024], 4)
    assert sup(w16, x16, y)
    assert L.b200_mul_mat_scratch_bytes(C.byref(w16), C.byref(x16)) == 0                              # k_mm_simt reads F16 activations in place
    assert not sup(desc(ops.F32, [240, 100], 4), x16, y)                                             # F32 x F16: not a product the CPU backend has
    assert not sup(desc(ops.BF16, [240, 100], 2), x16, y)
    assert not sup(w16, desc(ops.F16, [240, 1024], 2, nb0=4), y)                                     # activation rows must be contiguous
    assert not sup(w16, desc(ops.F16, [248, 1024], 2), y)                                            # k mismatch
    # the F16 tensor-core GEMM: k = 4304 (67.25 K tiles) -> tiles for 68, n padded to 256 columns
    w, x = desc(ops.F16, [4304, 1152], 2), desc(ops.F32, [4304, 1000], 4)
    assert sup(w, x, desc(ops.F32, [1152, 1000], 4))
    assert L.b200_mul_mat_scratch_bytes(C.byref(w), C.byref(x)) == 1024 * 4352 * 2
    # k = 72 F32 x F32 (SigLip K.Q^T per head): supported, no scratch (k_mm_simt), batch dims broadcast
    wk, xq = desc(ops.F32, [72, 1024, 16], 4), desc(ops.F32, [72, 1024, 16], 4)
    assert sup(wk, xq, desc(ops.F32, [1024, 1024, 16], 4)) and L.b200_mul_mat_scratch_bytes(C.byref(wk), C.byref(xq)) == 0
    assert not sup(wk, desc(ops.F32, [72, 1024, 24], 4), desc(ops.F32, [1024, 1024, 24], 4))        # 24 % 16 != 0: not a ggml broadcast


@pytest.mark.skipif(not (ROOT / "oracle" / "_ref" / "bin" / "omni_encoders").exists() or not SO.exists(), reason="oracle/_ref/bin/omni_encoders is not built")
def test_bench_encoder_leg_never_reports_a_cpu_number_as_the_plugin(monkeypatch, tmp_path):
    """bench.py's `omni_encoders` object (BASELINE.json configs[3], encoder half): on a box without a GPU the harness's second side falls back to the reference CPU backend —
    the leg must then say so instead of printing that time as the plugin's (plumbing check with 2-layer encoders)."""
    import bench
    monkeypatch.setenv("B200_BENCH_OMNI_SMALL", "1")
    monkeypatch.setenv("TMPDIR", str(tmp_path))
    r = bench.omni_encoders_leg(2)
    for key in ("apm_1s_audio_chunk", "vpm_448x448_frame"):
        assert "ms" not in r[key] and "no B200 device" in r[key]["error"] and r[key]["reference_cpu_ms"] > 0, r

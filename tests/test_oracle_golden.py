"""CPU: the C restatement (oracle/*.c) against the committed golden vectors minted from the reference (tests/golden).

Bars: block decode and activation quantisers BIT-EXACT; MUL_MAT within a few ulp of the sum of |terms| (float
accumulation order differs between the reference's generic / AVX2 / sgemm code paths, the integer parts are exact).
"""
import numpy as np
import pytest

import oracle_c as O


def _canon_q8K(b):
    """quantize_row_q8_K_ref leaves bsums of an all-zero block unwritten (ggml-quants.c:2569-2574): ignore them."""
    b = np.array(b, dtype=np.uint8).reshape(-1, 292).copy()
    zero = (b[:, :4].view(np.float32)[:, 0] == 0)
    b[zero, 260:] = 0
    return b

QT = {"q4_0": O.Q4_0, "q8_0": O.Q8_0, "q4_K": O.Q4_K, "q5_K": O.Q5_K, "q6_K": O.Q6_K}


@pytest.mark.parametrize("name", QT)
def test_dequant_bit_exact(golden, name):
    g = golden["mul_mat"]
    got = O.dequant(QT[name], g[f"w_{name}"][:2], 1024)
    assert np.array_equal(got, g[f"deq_{name}"])


def test_activation_quantisers_bit_exact(golden):
    g = golden["mul_mat"]
    x = g["x"]
    assert np.array_equal(O.quantize_q8_0(x[0], variant=0), g["act_q8_0_ref"])
    assert np.array_equal(O.quantize_q8_0(x[0], variant=1), g["act_q8_0_simd"])
    for i in range(3):
        assert np.array_equal(_canon_q8K(O.quantize_q8_K(x[i])), _canon_q8K(g["act_q8_K"][i]))


@pytest.mark.parametrize("name", list(QT) + ["f16"])
def test_mul_mat(golden, name):
    g = golden["mul_mat"]
    t = QT.get(name, O.F16)
    w, x, ref = g[f"w_{name}"], g["x"], g[f"y_{name}"]
    got = O.mul_mat(t, w, x, 48, 1024)
    # scale of the summands: |W| . |x|
    wf = O.dequant(t, w, 1024) if name != "f16" else w.view(np.float16).astype(np.float32).reshape(48, 1024)
    mag = np.abs(x) @ np.abs(wf).T
    assert np.all(np.abs(got - ref) <= 2e-6 * mag + 1e-12), np.abs(got - ref).max()


def test_rms_norm_rope_glu_setrows(golden):
    g = golden["ops"]
    assert np.allclose(O.rms_norm(g["rms_x"], 1e-6), g["rms_y"], rtol=2e-7, atol=0)
    for key, nd, mode, ctx, base in (("rope_neox", 128, 2, 40960, 1e6), ("rope_norm", 128, 0, 4096, 1e4),
                                     ("rope_neox_partial", 64, 2, 40960, 1e6)):
        got = O.rope(g["rope_x"], g["rope_pos"], nd, mode, ctx, base)
        assert np.abs(got - g[key]).max() <= 2e-6, key          # glibc sinf/cosf both sides; theta chain identical
    # the reference's SIMD expf differs from libm by ~1 ulp
    assert np.allclose(O.swiglu(g["glu_gate"], g["glu_up"]), g["glu_y"], rtol=1e-6, atol=1e-7)
    dst0 = np.zeros_like(g["sr_dst"])
    assert np.array_equal(O.set_rows_f16(g["sr_src"], g["sr_idx"], dst0).view(np.uint16), g["sr_dst"].view(np.uint16))


@pytest.mark.parametrize("tag", ["dec", "pre"])
def test_flash_attn(golden, tag):
    g = golden["flash_attn"]
    q = g[f"{tag}_q"]
    got = O.flash_attn(q, g[f"{tag}_k"], g[f"{tag}_v"], g[f"{tag}_mask"], 1.0 / np.sqrt(128), f16_acc=True)
    ref = g[f"{tag}_out"]
    # the reference accumulates V in F16 (ops.cpp:8040-8060): the port reproduces that, differing only in the f32 dot order
    assert np.abs(got - ref).max() <= 3e-3 * np.abs(ref).max()
    exact = O.flash_attn(q, g[f"{tag}_k"], g[f"{tag}_v"], g[f"{tag}_mask"], 1.0 / np.sqrt(128), f16_acc=False)
    assert np.abs(exact - ref).max() <= 1e-2 * np.abs(ref).max()

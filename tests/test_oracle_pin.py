"""CPU: pin oracle/*.c against the reference itself (oracle/_ref, live) on fresh seeded inputs and edge cases.

Skipped when oracle/_ref has not been built (it is built by `make -C oracle ref` / __graft_entry__.build() where
/root/reference exists, and travels prebuilt to the GPU box).
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
import refggml as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
QT = [O.Q4_0, O.Q8_0, O.Q4_K, O.Q5_K, O.Q6_K]


@pytest.mark.parametrize("t", QT)
@pytest.mark.parametrize("seed", [1, 2])
def test_block_decode_matches_reference(t, seed):
    rng = np.random.default_rng(seed)
    k = 2048
    x = (rng.standard_normal((6, k)) * rng.choice([1e-3, 1.0, 30.0], (6, 1))).astype(np.float32)
    x[0, :256] = 0.0
    q = R.quantize(t, x)
    assert np.array_equal(O.dequant(t, q, k), R.dequantize(t, q, k))


def test_random_bytes_decode():
    """decoders must agree on arbitrary bit patterns, not only on quantiser output (all 6-bit scale/min codes, all nibbles)."""
    rng = np.random.default_rng(3)
    for t in QT:
        bs = O.BLOCK[t]
        raw = rng.integers(0, 256, (4, 8 * bs[1]), dtype=np.uint8)
        k = 8 * bs[0]
        blk = raw.reshape(4, 8, bs[1])
        # keep the f16 scale fields finite
        for off in {O.Q4_0: [0], O.Q8_0: [0], O.Q4_K: [0, 2], O.Q5_K: [0, 2], O.Q6_K: [208]}[t]:
            blk[:, :, off + 1] &= 0x3B
        assert np.array_equal(O.dequant(t, raw, k), R.dequantize(t, raw, k))


@pytest.mark.parametrize("seed", [0, 7])
def test_activation_quantisers(seed):
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal(4096) * 2.5).astype(np.float32)
    x[256:512] = 0.0
    x[700] = -40.0            # negative max: q8_K keeps the sign in iscale
    assert np.array_equal(_canon_q8K(O.quantize_q8_K(x)), _canon_q8K(R.quantize_act(R.Q8_K, x)))
    assert np.array_equal(_canon_q8K(O.quantize_q8_K(x)), _canon_q8K(R.quantize_act(R.Q8_K, x, simd=True)))
    assert np.array_equal(O.quantize_q8_0(x, 0), R.quantize_act(R.Q8_0, x))
    assert np.array_equal(O.quantize_q8_0(x, 1), R.quantize_act(R.Q8_0, x, simd=True))
    x[32:64] = -x[0:32]                                       # equal magnitudes, opposite signs: the first largest element wins
    x[96] = 7.25; x[97] = -7.25
    assert np.array_equal(O.quantize_q4_0(x), R.quantize(R.Q4_0, x.reshape(1, -1)).reshape(-1))      # ggml_quantize_chunk -> quantize_row_q4_0_ref


@pytest.mark.parametrize("t", QT)
def test_vec_dot_and_mul_mat(t):
    rng = np.random.default_rng(11 + t)
    m, k, n = 32, 4096, 2
    w = (rng.standard_normal((m, k)) * 0.02).astype(np.float32)
    x = rng.standard_normal((n, k)).astype(np.float32)
    wq = R.quantize(t, w)
    ref = R.mul_mat(t, wq, x, m, k)
    got = O.mul_mat(t, wq, x, m, k)
    mag = np.abs(x) @ np.abs(O.dequant(t, wq, k)).T
    assert np.all(np.abs(got - ref) <= 2e-6 * mag)
    # the generic C vec_dot and the SIMD one of the reference agree with the port to the same bar
    act = R.quantize_act(R.Q8_0 if t in (O.Q4_0, O.Q8_0) else R.Q8_K, x[0], simd=True)
    for r in range(4):
        a = R.vec_dot(t, wq[r], act, k, generic=True)
        b = R.vec_dot(t, wq[r], act, k, generic=False)
        assert abs(a - got[0, r]) <= 2e-6 * mag[0, r] and abs(b - got[0, r]) <= 2e-6 * mag[0, r]


def test_ops_match_reference():
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((5, 4096)) * 2).astype(np.float32)
    assert np.allclose(O.rms_norm(x, 1e-6), R.rms_norm(x, 1e-6), rtol=2e-7, atol=0)
    qk = rng.standard_normal((3, 8, 128)).astype(np.float32)
    pos = np.array([3, 2047, 40000], np.int32)
    assert np.abs(O.rope(qk, pos, 128, 2) - R.rope(qk, pos, 128, 2, 40960, 1e6)).max() <= 2e-6
    g, u = (rng.standard_normal(4096) * 5).astype(np.float32), rng.standard_normal(4096).astype(np.float32)
    assert np.allclose(O.swiglu(g, u), R.swiglu(g, u), rtol=1e-6, atol=1e-7)


def test_flash_attn_matches_reference():
    rng = np.random.default_rng(9)
    D, n_head, n_head_kv, n_kv = 128, 8, 2, 512
    q = rng.standard_normal((1, n_head, D)).astype(np.float32)
    k = (rng.standard_normal((n_head_kv, n_kv, D)) * 0.3).astype(np.float16)
    v = rng.standard_normal((n_head_kv, n_kv, D)).astype(np.float16)
    mask = np.zeros((64, n_kv), np.float16)
    mask[:, 400:] = -np.inf
    ref = R.flash_attn(q, k, v, mask, 1 / np.sqrt(D))
    got = O.flash_attn(q, k, v, mask, 1 / np.sqrt(D), f16_acc=True)
    assert np.abs(got - ref).max() <= 3e-3 * np.abs(ref).max()


def test_encoder_ops_match_reference():
    """NORM / IM2COL / POOL_1D (the APM / VPM encoder graphs): oracle/ops_port.c against single-op graphs on the live reference CPU backend."""
    rng = np.random.default_rng(21)
    x = (rng.standard_normal((7, 1280)) * 3 + 0.5).astype(np.float32)
    assert np.allclose(O.norm(x, 1e-5), R.norm(x, 1e-5), rtol=2e-6, atol=2e-6)
    # Whisper conv_1d (k = 3, stride 1 and 2, "same" padding) and SigLip patch embedding (14 x 14, stride 14)
    a = rng.standard_normal((1, 16, 1, 100)).astype(np.float32)
    for s in (1, 2):
        assert np.array_equal(O.im2col(a, 1, 3, s, 0, 1, 0, 1, 0).view(np.uint16), R.im2col(a, 1, 3, 8, s, 0, 1, 0, 1, 0, False).view(np.uint16))
    img = rng.standard_normal((2, 3, 56, 56)).astype(np.float32)
    assert np.array_equal(O.im2col(img, 14, 14, 14, 14, 0, 0, 1, 1).view(np.uint16), R.im2col(img, 14, 14, 8, 14, 14, 0, 0, 1, 1, True).view(np.uint16))
    assert np.array_equal(O.im2col(img, 3, 3, 2, 2, 1, 1, 1, 1, f16=False), R.im2col(img, 3, 3, 8, 2, 2, 1, 1, 1, 1, True, f16=False))
    t = rng.standard_normal((5, 250)).astype(np.float32)
    for op in (0, 1):
        assert np.array_equal(O.pool_1d(t, op, 5), R.pool_1d(t, op, 5))


def test_make_gguf_llama_arch_loads_in_the_reference(tmp_path):
    """tools/make_gguf.py --arch llama (the TTS-transformer shape of MiniCPM-o: no q/k-norm, ROPE norm mode) must be a file the unmodified reference loads and decodes:
    llama_parity mode 3 = the reference CPU backend against itself (batched vs token-by-token prompt), also with the K-shift knob (llama_memory_seq_rm / seq_add)."""
    import json, os, subprocess, sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    exe = root / "oracle" / "_ref" / "bin" / "llama_parity"
    if not exe.exists():
        pytest.skip("oracle/_ref is not built")
    f = tmp_path / "tiny_llama.gguf"
    subprocess.check_call([sys.executable, str(root / "tools" / "make_gguf.py"), str(f), "--arch", "llama", "--ftype", "f16", "--embd", "256", "--ff", "512", "--heads", "4",
                           "--kv-heads", "4", "--head-dim", "64", "--layers", "2", "--vocab", "512"], stderr=subprocess.DEVNULL)
    env = dict(os.environ, LD_LIBRARY_PATH=f"{root / 'oracle' / '_ref' / 'lib'}:" + os.environ.get("LD_LIBRARY_PATH", ""))
    env.pop("GGML_BACKEND_PATH", None)
    for extra in ({}, {"PARITY_KSHIFT": "1"}):
        r = subprocess.run([str(exe), str(f), "16", "6", "2", "0", "3"], env=dict(env, **extra), capture_output=True, text=True, timeout=300)
        res = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        assert "error" not in res and res["max_rel_logit_err"] <= 5e-3, res


def test_wave_ops_match_reference():
    """The numpy restatements of the Token2Wav op set (tests/oracle_wave.py: the oracle of tests/test_gpu_parity.py::test_wave_*) against the live reference CPU backend."""
    import oracle_wave as W
    rng = np.random.default_rng(61)
    x = (rng.standard_normal((3, 5, 77)) * 3).astype(np.float32)
    for name in ("sin", "cos", "elu", "step", "sgn", "hardswish", "hardsigmoid"):
        np.testing.assert_allclose(W.unary(name, x), R.wave_unary(name, x), rtol=2e-6, atol=2e-6, err_msg=name)
    np.testing.assert_allclose(W.unary("log", np.abs(x) + 0.1), R.wave_unary("log", np.abs(x) + 0.1), rtol=2e-6, atol=2e-6)
    assert np.array_equal(W.unary("leaky_relu", x, 0.1), R.wave_unary("leaky_relu", x, 0.1))
    assert np.array_equal(W.unary("clamp", x, -1.5, 2.0), R.wave_unary("clamp", x, -1.5, 2.0))
    a = rng.standard_normal((3, 4, 5, 11)).astype(np.float32)
    for dim in range(4):
        shp = list(a.shape); shp[3 - dim] = 7
        b = rng.standard_normal(shp).astype(np.float32)
        assert np.array_equal(W.concat(a, b, dim), R.wave_concat(a, b, dim)), dim
    for reps in [(1, 1, 1, 2), (2, 1, 3, 1), (2, 2, 2, 2)]:
        assert np.array_equal(W.repeat(a, reps), R.wave_repeat(a, reps)), reps
    lr = [1, 2, 3, 4, 0, 1, 2, 0]
    assert np.array_equal(W.pad(a, lr), R.wave_pad(a, lr))
    y = rng.standard_normal((2, 80, 300)).astype(np.float32)
    assert np.array_equal(W.pad_reflect_1d(y, 7, 3), R.wave_pad_reflect_1d(y, 7, 3))
    assert np.array_equal(W.arange(0.5, 100.25, 0.75), R.wave_arange(0.5, 100.25, 0.75))
    for shape in [(3, 5, 1000), (2, 33), (1, 4097)]:
        z = rng.standard_normal(shape).astype(np.float32)
        assert np.array_equal(W.sum_rows(z), R.wave_sum_rows(z)), shape
    for (L, Cin, Cout, K, s0, f16) in [(197, 32, 16, 16, 1, False), (3, 2, 3, 2, 3, False), (50, 64, 32, 16, 8, True), (121, 16, 8, 11, 5, False)]:
        w = rng.standard_normal((Cin, Cout, K)).astype(np.float32)
        if f16:
            w = w.astype(np.float16).astype(np.float32)
        xx = rng.standard_normal((Cin, L)).astype(np.float32)
        if f16:
            xx = xx.astype(np.float16).astype(np.float32)                 # the CPU's f16-kernel variant rounds x to f16 too (ops.cpp:5993-5999); F16-exact inputs make both sides agree
        ref = R.wave_conv_transpose_1d(w, xx, s0, f16)
        got = W.conv_transpose_1d(w, xx, s0)
        mag = W.conv_transpose_1d(np.abs(w), np.abs(xx), s0)
        assert np.all(np.abs(got - ref) <= 2e-6 * mag + 1e-9), (L, Cin, Cout, K, s0)

"""GPU parity: the sm_100a kernels, called through the C-ABI (include/b200_ops.h via ctypes), against the CPU oracle
(oracle/*.c restating the reference's ggml CPU backend) and the committed golden vectors minted from the reference itself.

Bars (BASELINE.json north_star): integer / byte / index work BIT-EXACT (activation quantisers, repack, KV-cache writes, gathers);
MUL_MAT: the integer sub-block dot products are exact, only the order of the final f32 accumulation differs -> |err| <=
2e-6 * sum|w||x| (a few ulp of the summand scale; far inside the reference's own NMSE 5e-4 bar, tests/test-backend-ops.cpp:3300);
float ops within the tolerance written in each test.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import oracle_c as O

pytestmark = pytest.mark.gpu

from __graft_entry__ import load_package  # noqa: E402

QT = {"q4_0": O.Q4_0, "q8_0": O.Q8_0, "q4_K": O.Q4_K, "q5_K": O.Q5_K, "q6_K": O.Q6_K}


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    return load_package().ops


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rand_blocks(rng, t, nblocks):
    """Random but VALID quantised blocks (no quantiser needed): random payload bytes, f16 scales in a sane range."""
    bs = O.BLOCK[t][1]
    b = rng.integers(0, 256, (nblocks, bs), dtype=np.uint8)

    def h(n, lo=1e-3, hi=1e-2, signed=False):
        v = rng.uniform(lo, hi, n).astype(np.float16)
        if signed:
            v *= rng.choice([-1, 1], n).astype(np.float16)
        return v.view(np.uint8).reshape(n, 2)
    if t in (O.Q4_0, O.Q8_0):
        b[:, 0:2] = h(nblocks, signed=True)
    elif t in (O.Q4_K, O.Q5_K):
        b[:, 0:2] = h(nblocks)
        b[:, 2:4] = h(nblocks)
    elif t == O.Q6_K:
        b[:, 208:210] = h(nblocks, 1e-4, 1e-3, signed=True)
    return b


def mm_check(got, ref, w_deq, x, factor=2e-6):
    mag = np.abs(x) @ np.abs(w_deq).T
    err = np.abs(got - ref)
    assert np.all(err <= factor * mag + 1e-12), (err.max(), (err / (mag + 1e-30)).max())


def nmse(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300))


def tc_routed(ops, t, planar, n, k):
    """n > 8 columns of q4_K / q5_K (native) and q6_K / q8_0 / q4_0 (planar) go to the tcgen05 dequant-GEMM (csrc/mmq_tc.cu): F16 operands, F32 accumulation."""
    return n > 8 and k % 256 == 0 and (t in (ops.Q4_K, ops.Q5_K) or (t in (ops.Q6_K, ops.Q8_0, ops.Q4_0) and planar))


def tc_check(got, oracle_ref, w_deq, x):
    """Bars of the tensor-core path: the reference's own MUL_MAT bar against the CPU oracle (NMSE <= 5e-4, tests/test-backend-ops.cpp:3300)
    and, since the oracle's q8_K activations are the larger error, a much tighter one against the exact product of the dequantised weights."""
    exact = x.astype(np.float64) @ w_deq.astype(np.float64).T
    assert nmse(got, exact) <= 2e-6, nmse(got, exact)
    if oracle_ref is not None:
        assert nmse(got, oracle_ref) <= 5e-4, nmse(got, oracle_ref)


# ---------------------------------------------------------------------------------------------------------------- quantisers
def _canon_q8K(rec, k):
    return rec


@pytest.mark.parametrize("k", [256, 1024, 4096, 12288])
def test_quantize_q8K_bit_exact(ops, k):
    rng = np.random.default_rng(k)
    x = rng.standard_normal((3, k)).astype(np.float32) * 2
    x[1] *= 1e-3
    x[2, :256] = 0.0
    x[0, 5] = -x[0, 7]                     # tie in |x|: the reference keeps the FIRST maximum
    act = ops.quantize_act(ops.Q4_K, dev(x)).cpu().numpy()
    nb = k // 256
    for i in range(3):
        ref = O.quantize_q8_K(x[i]).reshape(nb, 292)
        d_ref = ref[:, :4].copy().view(np.float32)[:, 0]
        qs_ref = ref[:, 4:260].reshape(-1)
        bs_ref = ref[:, 260:292].copy().view(np.int16).reshape(-1)
        zero = np.repeat(d_ref == 0, 16)
        rec = act[i]
        assert np.array_equal(rec[:k], qs_ref)
        assert np.array_equal(rec[k:k + 4 * nb].view(np.float32), d_ref)
        got_bs = rec[k + 4 * nb:k + 4 * nb + 2 * (k // 16)].view(np.int16)
        assert np.array_equal(got_bs[~zero], bs_ref[~zero])      # ref leaves bsums of an all-zero block unwritten
        assert np.all(got_bs[zero] == 0)


@pytest.mark.parametrize("k", [32, 1024, 4096])
def test_quantize_q8_0_bit_exact(ops, k):
    rng = np.random.default_rng(k + 1)
    x = rng.standard_normal((2, k)).astype(np.float32)
    x[1, :32] = 0
    act = ops.quantize_act(ops.Q8_0, dev(x)).cpu().numpy()
    nb = k // 32
    for i in range(2):
        ref = O.quantize_q8_0(x[i], variant=1).reshape(nb, 34)
        assert np.array_equal(act[i][:k], ref[:, 2:].reshape(-1))
        assert np.array_equal(act[i][k:k + 2 * nb].view(np.uint16), ref[:, :2].copy().view(np.uint16).reshape(-1))
        bs = act[i][k + 2 * nb:k + 4 * nb].view(np.int16)
        assert np.array_equal(bs, ref[:, 2:].view(np.int8).astype(np.int32).sum(1).astype(np.int16))


# ---------------------------------------------------------------------------------------------------------------- MUL_MAT
@pytest.mark.parametrize("name", list(QT) + ["f16"])
@pytest.mark.parametrize("planar", [False, True])
def test_mul_mat_golden(ops, golden, name, planar):
    g = golden["mul_mat"]
    t = QT.get(name, O.F16)
    if planar and t not in ops.PAYLOAD:
        pytest.skip("type is never repacked")
    w, x, ref = g[f"w_{name}"], g["x"], g[f"y_{name}"]
    wd = dev(w.view(np.uint8))
    layout = ops.LAYOUT_NATIVE
    if planar:
        wd, layout = ops.to_planar(t, wd), ops.LAYOUT_PLANAR
    wf = O.dequant(t, w, 1024) if name != "f16" else w.view(np.float16).astype(np.float32).reshape(48, 1024)
    for n in (1, 3):
        got = ops.mul_mat(wd, t, 48, 1024, dev(x[:n]), layout=layout).cpu().numpy()
        mm_check(got, ref[:n], wf, x[:n], 3e-6 if name != "f16" else 1e-5)


SHAPES = [(4096, 4096, 1), (1024, 4096, 1), (12288, 4096, 1), (4096, 12288, 1), (40, 256, 1), (7, 512, 5), (33, 256, 8), (64, 1024, 11)]


@pytest.mark.parametrize("name", QT)
@pytest.mark.parametrize("m,k,n", SHAPES)
def test_mul_mat_vs_oracle(ops, name, m, k, n):
    """Seeded random blocks at the BASELINE.json layer shapes (Qwen3-8B: 4096x4096, 1024x4096, 12288x4096, 4096x12288) and ragged
    small ones; n > 8 exercises the column chunking."""
    t = QT[name]
    rng = np.random.default_rng(hash((name, m, k, n)) & 0xffff)
    blocks = rand_blocks(rng, t, m * k // O.BLOCK[t][0])
    x = rng.standard_normal((n, k)).astype(np.float32)
    ref = O.mul_mat(t, blocks, x, m, k)
    wf = O.dequant(t, blocks, k)
    wd = dev(blocks)
    got = ops.mul_mat(wd, t, m, k, dev(x)).cpu().numpy()
    (tc_check(got, ref, wf, x) if tc_routed(ops, t, False, n, k) else mm_check(got, ref, wf, x))
    if t in ops.PAYLOAD:
        got = ops.mul_mat(ops.to_planar(t, wd), t, m, k, dev(x), layout=ops.LAYOUT_PLANAR).cpu().numpy()
        (tc_check(got, ref, wf, x) if tc_routed(ops, t, True, n, k) else mm_check(got, ref, wf, x))


@pytest.mark.parametrize("name,m,k,n", [("q4_K", 4096, 4096, 512), ("q6_K", 1024, 4096, 300), ("q4_K", 200, 512, 17), ("q6_K", 129, 256, 9),
                                        ("q4_K", 12288, 4096, 2048), ("q4_K", 4096, 12288, 257), ("q6_K", 4096, 12288, 64),
                                        ("q5_K", 1024, 4096, 300), ("q8_0", 1024, 4096, 300), ("q4_0", 1024, 4096, 300), ("q5_K", 130, 512, 40),
                                        ("q8_0", 200, 768, 33), ("q4_0", 129, 256, 9)])
def test_mul_mat_prefill_tensor_core(ops, name, m, k, n):
    """Prefill GEMM (BASELINE.json configs[2] shapes and ragged ones) through b200_mul_mat -> k_mmq_tc.  Exact product of the dequantised
    weights in f64 on the GPU for the whole output; the CPU oracle (q8_K activations) on a sample of rows and columns."""
    t = QT[name]
    rng = np.random.default_rng(hash((name, m, k, n)) & 0xffff)
    blocks = rand_blocks(rng, t, m * k // O.BLOCK[t][0])
    x = rng.standard_normal((n, k)).astype(np.float32)
    wd = dev(blocks)
    planar = t in ops.PAYLOAD
    if planar:
        wd = ops.to_planar(t, wd)
    got = ops.mul_mat(wd, t, m, k, dev(x), layout=ops.LAYOUT_PLANAR if planar else ops.LAYOUT_NATIVE)
    torch.cuda.synchronize()
    wf = O.dequant(t, blocks, k)
    exact = (dev(x).double() @ dev(wf).double().T)
    err = nmse(got.cpu().numpy(), exact.cpu().numpy())
    assert err <= 2e-6, err
    rows = np.unique(np.concatenate([np.arange(min(m, 40)), np.arange(max(0, m - 40), m), rng.integers(0, m, 48)]))
    cols = np.unique(np.concatenate([np.arange(min(n, 4)), np.arange(max(0, n - 4), n), rng.integers(0, n, 4)]))
    sub = blocks.reshape(m, -1)[rows]
    ref = O.mul_mat(t, sub, x[cols], len(rows), k)
    assert nmse(got.cpu().numpy()[np.ix_(cols, rows)], ref) <= 5e-4


@pytest.mark.parametrize("name,m,k,n", [("q4_K", 4096, 4096, 2048), ("q6_K", 4096, 12288, 512), ("q4_K", 1024, 4096, 512), ("q4_K", 200, 512, 17), ("q6_K", 4096, 4096, 3)])
def test_mul_mat_add_is_mul_mat_then_add(ops, name, m, k, n):
    """b200_mul_mat_add (the residual ADD behind wo / ffn_down riding in the tcgen05 GEMM's epilogue; split-K launches and matvecs fall back to MUL_MAT + ADD
    inside the call): bit for bit the two separate ops, also when the residual IS the destination."""
    t = QT[name]
    rng = np.random.default_rng(hash((name, m, k, n, "add")) & 0xffff)
    blocks = rand_blocks(rng, t, m * k // O.BLOCK[t][0])
    planar = t in ops.PAYLOAD
    wd = ops.to_planar(t, dev(blocks)) if planar else dev(blocks)
    lay = ops.LAYOUT_PLANAR if planar else ops.LAYOUT_NATIVE
    x = dev(rng.standard_normal((n, k)).astype(np.float32))
    r = dev(rng.standard_normal((n, m)).astype(np.float32))
    want = ops.binary(ops.ADD, ops.mul_mat(wd, t, m, k, x, layout=lay), r)
    got = ops.mul_mat_add(wd, t, m, k, x, r, torch.empty_like(r), layout=lay)
    torch.cuda.synchronize()
    assert torch.isfinite(want).all()
    assert torch.equal(got, want)
    inplace = r.clone()                                # residual == dst: fine in the epilogue (each thread reads its element before writing it); the
    try:                                               # unfused fall-back would overwrite the residual first and must refuse instead
        ops.mul_mat_add(wd, t, m, k, x, inplace, inplace, layout=lay)
        torch.cuda.synchronize()
        assert torch.equal(inplace, want)
    except ops.B200Error as e:
        assert "-1000" in str(e) and torch.equal(inplace, r)


@pytest.mark.parametrize("name,m,k", [("q4_K", 4096, 4096), ("q6_K", 4096, 12288), ("q8_0", 1024, 2048), ("q4_0", 4096, 4096), ("q5_K", 300, 512), ("f16", 4096, 4096), ("f16", 768, 3072),
                                      ("f16", 70, 256)])
def test_decode_matvec_epilogues_are_the_separate_ops(ops, name, m, k):
    """One activation column (the decode graph): b200_mul_mat_add (residual in the matvec's epilogue, also in place) and b200_mul_mat_glu (gate / up / SWIGLU in one launch)
    against MUL_MAT, ADD and MUL_MAT, MUL_MAT, GLU — bit for bit (the same dot products, then the same single F32 operations)."""
    rng = np.random.default_rng(hash((name, m, k, "dec")) & 0xffff)
    x = dev(rng.standard_normal((1, k)).astype(np.float32))
    r = dev(rng.standard_normal((1, m)).astype(np.float32))
    if name == "f16":
        t, lay = ops.F16, ops.LAYOUT_NATIVE
        wg = (dev(rng.standard_normal((m, k)).astype(np.float32)) * 0.05).half(); wu = (dev(rng.standard_normal((m, k)).astype(np.float32)) * 0.05).half()
    else:
        t = QT[name]
        planar = t in ops.PAYLOAD
        lay = ops.LAYOUT_PLANAR if planar else ops.LAYOUT_NATIVE
        mk = lambda: (lambda b: ops.to_planar(t, dev(b)) if planar else dev(b))(rand_blocks(rng, t, m * k // O.BLOCK[t][0]))
        wg, wu = mk(), mk()
    g = ops.mul_mat(wg, t, m, k, x, layout=lay, w_ne=[k, m]); u = ops.mul_mat(wu, t, m, k, x, layout=lay, w_ne=[k, m])
    want_add = ops.binary(ops.ADD, g, r)
    got_add = ops.mul_mat_add(wg, t, m, k, x, r, torch.empty_like(r), layout=lay)
    inplace = r.clone()
    ops.mul_mat_add(wg, t, m, k, x, inplace, inplace, layout=lay)
    want_glu = torch.empty_like(g)
    ops.check(ops.lib().b200_glu(ops.GLU_SWIGLU, ops._ref(ops.T(g)), ops._ref(ops.T(u)), ops._ref(ops.T(want_glu)), 0, ops.stream()))
    got_glu = ops.mul_mat_glu(ops.GLU_SWIGLU, wg, wu, t, m, k, x, layout=lay)
    torch.cuda.synchronize()
    assert torch.isfinite(want_add).all() and torch.isfinite(want_glu).all()
    assert torch.equal(got_add, want_add) and torch.equal(inplace, want_add)
    assert torch.equal(got_glu, want_glu)


@pytest.mark.parametrize("names,ms,k,n", [(("q4_K", "q4_K", "q6_K"), (4096, 1024, 1024), 4096, 2048), (("q4_K", "q4_K", "q4_K"), (4096, 1024, 1024), 4096, 512),
                                           (("q5_K", "q6_K"), (200, 130), 512, 300), (("q4_K", "q8_0"), (256, 256), 512, 64), (("q4_K", "q4_K", "q6_K"), (512, 128, 128), 1024, 40),
                                           (("q4_K", "q6_K"), (512, 128), 1024, 3)])
def test_mul_mat_multi_is_the_separate_mul_mats(ops, names, ms, k, n):
    """b200_mul_mat_multi (q / k / v in one tcgen05 launch over the concatenated m-tiles; groups it cannot merge — split-K shapes, non-K-quants, matvecs — run
    one after the other inside the call): the separate MUL_MATs — bit for bit, except where the separate launch of a small matrix splits K over two CTAs (two partial
    sums added in F32) while the merged launch accumulates the whole K in one TMEM accumulator: there the two differ by F32 summation order only."""
    rng = np.random.default_rng(hash((names, ms, k, n)) & 0xffff)
    x = dev(rng.standard_normal((n, k)).astype(np.float32))
    ws, want = [], []
    for name, m in zip(names, ms):
        t = QT[name]
        blocks = rand_blocks(rng, t, m * k // O.BLOCK[t][0])
        planar = t in ops.PAYLOAD
        wd = ops.to_planar(t, dev(blocks)) if planar else dev(blocks)
        lay = ops.LAYOUT_PLANAR if planar else ops.LAYOUT_NATIVE
        ws.append((wd, t, m, lay))
        want.append(ops.mul_mat(wd, t, m, k, x, layout=lay))
    outs = [torch.empty_like(w) for w in want]
    ops.mul_mat_multi(ws, x, outs)
    torch.cuda.synchronize()
    for o, w in zip(outs, want):
        assert torch.isfinite(w).all()
        if not torch.equal(o, w):
            err = float(((o.double() - w.double()) ** 2).sum() / (w.double() ** 2).sum())
            assert err <= 1e-10, err                    # (tensor-core F32 accumulation of 2 x K/2 vs 1 x K: ~3e-6 relative rms)


def test_mul_mat_lm_head_shape(ops):
    """output.weight of MiniCPM-o-4.5: Q6_K [4096, 151748]; oracle on a row sample, linearity on the full output."""
    m, k = 151748, 4096
    rng = np.random.default_rng(7)
    blocks = rand_blocks(rng, O.Q6_K, m * k // 256)
    x = rng.standard_normal((1, k)).astype(np.float32)
    wd = ops.to_planar(ops.Q6_K, dev(blocks))
    got = ops.mul_mat(wd, ops.Q6_K, m, k, dev(x), layout=ops.LAYOUT_PLANAR).cpu().numpy()
    rows = np.concatenate([np.arange(0, 64), np.arange(m - 64, m), rng.integers(0, m, 128)])
    sub = blocks.reshape(m, -1)[rows]
    ref = O.mul_mat(O.Q6_K, sub, x, len(rows), k)
    mm_check(got[:, rows], ref, O.dequant(O.Q6_K, sub, k), x)
    # homogeneity: q8_K quantisation is scale-equivariant for power-of-two scales -> y(2x) == 2 y(x) exactly
    got2 = ops.mul_mat(wd, ops.Q6_K, m, k, dev(2 * x), layout=ops.LAYOUT_PLANAR).cpu().numpy()
    assert np.array_equal(got2, 2 * got)


def test_mul_mat_empty_and_batched(ops):
    rng = np.random.default_rng(3)
    blocks = rand_blocks(rng, O.Q4_K, 16 * 512 // 256)
    wd = dev(blocks)
    y = ops.mul_mat(wd, ops.Q4_K, 16, 512, torch.empty((0, 512), device="cuda"))
    assert y.shape == (0, 16)
    # batch dims with broadcast: x [2, 3, n=2, k], w 2-D
    x = rng.standard_normal((2, 3, 2, 512)).astype(np.float32)
    got = ops.mul_mat(wd, ops.Q4_K, 16, 512, dev(x)).cpu().numpy()
    ref = O.mul_mat(O.Q4_K, blocks, x.reshape(-1, 512), 16, 512).reshape(2, 3, 2, 16)
    mm_check(got.reshape(-1, 16), ref.reshape(-1, 16), O.dequant(O.Q4_K, blocks, 512), x.reshape(-1, 512))


@pytest.mark.parametrize("dt", ["f32", "f16", "bf16"])
def test_mul_mat_float_weights(ops, dt):
    rng = np.random.default_rng(11)
    m, k, n = 96, 768, 3
    w = (rng.standard_normal((m, k)) * 0.05).astype(np.float32)
    x = rng.standard_normal((n, k)).astype(np.float32)
    tdt = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}[dt]
    wt = torch.from_numpy(w).cuda().to(tdt)
    got = ops.mul_mat(wt, {"f32": ops.F32, "f16": ops.F16, "bf16": ops.BF16}[dt], m, k, dev(x), w_ne=[k, m]).cpu().numpy()
    xr = torch.from_numpy(x).to(tdt).double()             # the CPU backend rounds activations to the weight type (vec_dot_type)
    ref = (xr @ wt.cpu().double().T).numpy()
    assert np.abs(got - ref).max() <= 1e-5 * (np.abs(x) @ np.abs(w).T).max()


@pytest.mark.parametrize("m,k", [(4096, 4096), (4096, 12288), (151748, 4096), (65, 768), (64, 256), (3072, 768)])
def test_mul_mat_f16_weights_decode_stream(ops, m, k):
    """The single-column F16 matvec (k_mmvf16_stream: the decode path of an F16 model and of the TTS llama): products of F16-rounded operands, F32 accumulation."""
    rng = np.random.default_rng(m + k)
    w = dev((rng.standard_normal((m, k)) * 0.05).astype(np.float32)).half()
    x = rng.standard_normal((1, k)).astype(np.float32)
    got = ops.mul_mat(w, ops.F16, m, k, dev(x), w_ne=[k, m])
    xr = dev(x).half().double()
    ref = xr @ w.double().T
    mag = xr.abs() @ w.double().abs().T
    torch.cuda.synchronize()
    assert bool(((got.double() - ref).abs() <= 2e-6 * mag + 1e-12).all())


@pytest.mark.parametrize("m,k,n,batch", [(4096, 4096, 512, None), (300, 1024, 77, None), (129, 64, 9, None), (1152, 4304 // 16 * 16 - 4304 % 64, 200, None),
                                          (256, 256, 33, (2, 3))])
def test_mul_mat_f16_weights_tensor_core(ops, m, k, n, batch):
    """F16 weights with more than 8 columns -> k_mm_f16_tc (2-D TMA with the 128-byte swizzle straight into the UMMA operand layout).  Same
    arithmetic as the CPU oracle: activations rounded to F16 (vec_dot_type of F16 weights), products accumulated in F32.
    (K that is not a multiple of the 64-wide tile: tests/test_zz_ragged_k.py)"""
    k = k // 64 * 64
    rng = np.random.default_rng(m + k + n)
    w = (rng.standard_normal((m, k)) * 0.05).astype(np.float16)
    xs = (n, k) if batch is None else batch + (n, k)
    x = rng.standard_normal(xs).astype(np.float32)
    wt = torch.from_numpy(w).cuda()
    got = ops.mul_mat(wt, ops.F16, m, k, dev(x), w_ne=[k, m]).cpu().numpy()
    xr = torch.from_numpy(x).cuda().half().double()
    ref = (xr @ wt.double().T).cpu().numpy()
    mag = (np.abs(x.reshape(-1, k)) @ np.abs(w.astype(np.float32)).T).reshape(ref.shape)
    assert np.all(np.abs(got - ref) <= 2e-6 * mag + 1e-9), (np.abs(got - ref) / (mag + 1e-30)).max()


@pytest.mark.parametrize("wdt,xdt,m,k,n,wb,xb", [
    ("f32", "f32", 1024, 72, 1024, (1, 16), (1, 16)),      # SigLip K.Q^T: F32 x F32, k = d_head = 72, one slice per head (vision.cpp:662)
    ("f32", "f32", 72, 1000, 333, (1, 4), (1, 4)),         # V.softmax: m = d_head, k = n_pos
    ("f16", "f32", 64, 150, 50, (1, 16), (1, 16)),         # Whisper V.softmax over the F16 encoder cache, k = 50 . (iter + 1) (audition.cpp:620)
    ("f16", "f32", 300, 64, 50, (1, 16), (1, 16)),         # Whisper K.Q^T: tensor-core eligible but 16 small slices -> the one-launch route
    ("f16", "f16", 100, 240, 1024, (1, 1), (1, 1)),        # ggml_conv_1d: im2col [IC . K, OL] F16 x kernel [IC . K, OC] F16 (conv1 of the Whisper encoder)
    ("f16", "f16", 1024, 588, 1152, (1, 1), (1, 1)),       # ggml_conv_2d patch embedding: 3 . 14 . 14 = 588
    ("f16", "f16", 3, 33, 5, (2, 1), (2, 3)),              # F16 activations with <= 8 columns, broadcast over dim 3
    ("bf16", "f32", 65, 100, 9, (1, 1), (2, 2)),           # ragged everything, 2-D weight broadcast over both batch dims
    ("f32", "f32", 129, 31, 65, (3, 2), (3, 2))])
def test_mul_mat_float_weights_many_columns_simt(ops, wdt, xdt, m, k, n, wb, xb):
    """Float weights, more than 8 columns, off the tcgen05 route (k % 64 != 0, F32 / BF16 weights, F16 activations, many small slices) -> k_mm_simt: ONE launch over all
    batch slices.  Arithmetic = the CPU oracle's: activations rounded to the weight type (vec_dot_type), products accumulated in F32.  `wb` / `xb` = (ne3, ne2) of W / x."""
    rng = np.random.default_rng(m * 7 + k * 3 + n)
    tdt = {"f32": torch.float32, "f16": torch.float16, "bf16": torch.bfloat16}
    w = torch.from_numpy((rng.standard_normal(wb + (m, k)) * 0.05).astype(np.float32)).cuda().to(tdt[wdt])
    x = torch.from_numpy(rng.standard_normal(xb + (n, k)).astype(np.float32)).cuda().to(tdt[xdt])
    got = ops.mul_mat(w, {"f32": ops.F32, "f16": ops.F16, "bf16": ops.BF16}[wdt], m, k, x, w_ne=[k, m, wb[1], wb[0]],
                      w_nb=[w.element_size(), k * w.element_size(), m * k * w.element_size(), wb[1] * m * k * w.element_size()]).cpu().numpy()
    xr = x.to(tdt[wdt]).double()
    wd = w.double()
    ref = torch.empty(xb + (n, m), dtype=torch.float64)
    mag = torch.empty_like(ref)
    for i3 in range(xb[0]):
        for i2 in range(xb[1]):
            ws = wd[i3 // (xb[0] // wb[0]), i2 // (xb[1] // wb[1])]
            ref[i3, i2] = (xr[i3, i2] @ ws.T).cpu()
            mag[i3, i2] = (xr[i3, i2].abs() @ ws.abs().T).cpu()
    assert got.shape == tuple(ref.shape)
    assert np.all(np.abs(got - ref.numpy()) <= 2e-6 * mag.numpy() + 1e-9), float((np.abs(got - ref.numpy()) / (mag.numpy() + 1e-30)).max())
    # the A/B switch: one k_mmvf launch per 8 columns gives the same numbers up to the summation order (F32 activations only: k_mmvf reads F32)
    if xdt == "f32":
        os.environ["B200_NO_SIMT_GEMM"] = "1"
        try:
            old = ops.mul_mat(w, {"f32": ops.F32, "f16": ops.F16, "bf16": ops.BF16}[wdt], m, k, x, w_ne=[k, m, wb[1], wb[0]],
                              w_nb=[w.element_size(), k * w.element_size(), m * k * w.element_size(), wb[1] * m * k * w.element_size()]).cpu().numpy()
        finally:
            del os.environ["B200_NO_SIMT_GEMM"]
        assert np.all(np.abs(got - old) <= 4e-6 * mag.numpy() + 1e-9)


def test_matvec_jobs_residual_swiglu(ops):
    rng = np.random.default_rng(5)
    k = 1024
    x = rng.standard_normal((1, k)).astype(np.float32)
    act = ops.quantize_act(ops.Q4_K, dev(x))
    ms = [256, 64, 64]
    ws = [rand_blocks(rng, O.Q4_K, m * k // 256) for m in ms]
    ys = [torch.zeros(m, device="cuda") for m in ms]
    res = rng.standard_normal(ms[0]).astype(np.float32)
    jobs = [ops.make_job(dev(w), ops.Q4_K, m, k, y) for w, m, y in zip(ws, ms, ys)]
    keep = [dev(w) for w in ws]
    jobs = [ops.make_job(kw, ops.Q4_K, m, k, y) for kw, m, y in zip(keep, ms, ys)]
    rd = dev(res)
    jobs[0].residual = rd.data_ptr()
    ops.matvec_q(jobs, act, k)
    for i, (w, m) in enumerate(zip(ws, ms)):
        ref = O.mul_mat(O.Q4_K, w, x, m, k)[0] + (res if i == 0 else 0)
        mm_check(ys[i].cpu().numpy()[None], ref[None], O.dequant(O.Q4_K, w, k), x)
    # gate/up + swiglu epilogue
    m = 384
    wg, wu = rand_blocks(rng, O.Q4_K, m * k // 256), rand_blocks(rng, O.Q4_K, m * k // 256)
    dg, du = dev(wg), dev(wu)
    y = torch.zeros(m, device="cuda")
    ops.matvec_q_swiglu(ops.make_job(dg, ops.Q4_K, m, k, y), ops.make_job(du, ops.Q4_K, m, k, y), y, act, k)
    g, u = O.mul_mat(O.Q4_K, wg, x, m, k)[0], O.mul_mat(O.Q4_K, wu, x, m, k)[0]
    # silu amplifies a relative error eps of g to ~|g|*eps for very negative g: check the dots through the unfused launch (the
    # same warp arithmetic, bit-identical) and the epilogue against the oracle's swiglu of those dots
    yg, yu = torch.zeros(m, device="cuda"), torch.zeros(m, device="cuda")
    ops.matvec_q([ops.make_job(dg, ops.Q4_K, m, k, yg), ops.make_job(du, ops.Q4_K, m, k, yu)], act, k)
    mm_check(yg.cpu().numpy()[None], g[None], O.dequant(O.Q4_K, wg, k), x)
    mm_check(yu.cpu().numpy()[None], u[None], O.dequant(O.Q4_K, wu, k), x)
    assert np.allclose(y.cpu().numpy(), O.swiglu(yg.cpu().numpy(), yu.cpu().numpy()), rtol=3e-6, atol=1e-30)


@pytest.mark.parametrize("name", ["q4_0", "q8_0", "q6_K"])
def test_repack_round_trip(ops, name):
    t = QT[name]
    rng = np.random.default_rng(9)
    nblocks = 70001                     # odd count: chunk boundaries fall inside blocks
    raw = rng.integers(0, 256, nblocks * O.BLOCK[t][1], dtype=np.uint8)
    wd = dev(raw)
    planar = ops.to_planar(t, wd)
    back = ops.from_planar(t, planar)
    assert torch.equal(back, wd)
    pay = ops.PAYLOAD[t]
    p = planar.cpu().numpy()
    nat = raw.reshape(nblocks, -1)
    d_off = 208 if t == O.Q6_K else 0
    p_off = 0 if t == O.Q6_K else 2
    assert np.array_equal(p[:nblocks * pay].reshape(nblocks, pay), nat[:, p_off:p_off + pay])
    assert np.array_equal(p[nblocks * pay:].reshape(nblocks, 2), nat[:, d_off:d_off + 2])


# ---------------------------------------------------------------------------------------------------------------- norm / rope / kv
def test_rms_norm_golden_and_fused_forms(ops, golden):
    g = golden["ops"]
    got = ops.rms_norm(dev(g["rms_x"]), 1e-6).cpu().numpy()
    assert np.allclose(got, g["rms_y"], rtol=1e-6, atol=0)
    rng = np.random.default_rng(2)
    for rows, n in ((1, 4096), (40, 128), (5, 1000), (3, 12288)):
        x = rng.standard_normal((rows, n)).astype(np.float32) * 3
        w = (1 + 0.1 * rng.standard_normal(n)).astype(np.float32)
        a = rng.standard_normal((rows, n)).astype(np.float32)
        ref = O.rms_norm(x, 1e-6)
        assert np.allclose(ops.rms_norm(dev(x), 1e-6).cpu().numpy(), ref, rtol=1e-6, atol=0)
        assert np.allclose(ops.rms_norm(dev(x), 1e-6, w=dev(w)).cpu().numpy(), ref * w, rtol=1e-6, atol=0)
        assert np.allclose(ops.rms_norm(dev(x), 1e-6, w=dev(w), add=dev(a)).cpu().numpy(), ref * w + a, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("wtype", ["q4_K", "q8_0"])
def test_rms_norm_quantize_fused(ops, wtype):
    """The fused producer must equal rms_norm -> mul -> quantize run as separate C-ABI calls (same f32 op sequence)."""
    t = QT[wtype]
    rng = np.random.default_rng(4)
    for n, k in ((1, 4096), (3, 1024), (2, 12288)):
        x = rng.standard_normal((n, k)).astype(np.float32) * 2
        w = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
        xd, wd = dev(x), dev(w)
        act, y = ops.rms_norm_quantize(xd, wd, t, 1e-6, want_f32=True)
        y_sep = ops.rms_norm(xd, 1e-6, w=wd)
        assert np.allclose(y.cpu().numpy(), y_sep.cpu().numpy(), rtol=3e-7, atol=0)
        act_sep = ops.quantize_act(t, y)                      # quantising the fused kernel's own f32 output must be identical
        assert torch.equal(act, act_sep)
        ref = O.rms_norm(x, 1e-6) * w
        assert np.allclose(y.cpu().numpy(), ref, rtol=1e-6, atol=0)


def test_rope_golden(ops, golden):
    g = golden["ops"]
    x, pos = dev(g["rope_x"]), dev(g["rope_pos"])
    for key, nd, mode, ctx, base in (("rope_neox", 128, 2, 40960, 1e6), ("rope_norm", 128, 0, 4096, 1e4),
                                     ("rope_neox_partial", 64, 2, 40960, 1e6)):
        got = ops.rope(x, pos, nd, mode, ctx, base).cpu().numpy()
        assert np.abs(got - g[key]).max() <= 4e-6, key         # theta chain identical; CUDA sincosf vs glibc <= 2 ulp


def test_rope_yarn_and_freq_factors(ops):
    rng = np.random.default_rng(6)
    x = rng.standard_normal((3, 4, 64)).astype(np.float32)
    pos = np.array([3, 500, 9000], np.int32)
    ff = (1 + rng.uniform(0, 1, 32)).astype(np.float32)
    ref = O.rope(x, pos, 64, 0, 4096, 1e4, freq_scale=0.25, ext_factor=1.0, attn_factor=1.1, freq_factors=ff)
    got = ops.rope(dev(x), dev(pos), 64, 0, 4096, 1e4, freq_scale=0.25, ext_factor=1.0, attn_factor=1.1, freq_factors=dev(ff)).cpu().numpy()
    assert np.abs(got - ref).max() <= 1e-5


def test_set_rows_get_rows_cpy(ops, golden):
    g = golden["ops"]
    dst = torch.zeros((12, 256), dtype=torch.float16, device="cuda")
    ops.set_rows(dev(g["sr_src"]), dev(g["sr_idx"]), dst)
    assert np.array_equal(dst.cpu().numpy().view(np.uint16), g["sr_dst"].view(np.uint16))
    # i32 indices, f32 destination, strided destination rows (a KV-cache view)
    rng = np.random.default_rng(8)
    src = rng.standard_normal((5, 1024)).astype(np.float32)
    idx = np.array([7, 0, 3, 9, 1], np.int32)
    big = torch.zeros((10, 2048), dtype=torch.float16, device="cuda")
    ops.set_rows(dev(src), dev(idx), big[:, :1024])
    exp = np.zeros((10, 2048), np.float16)
    exp[idx, :1024] = src.astype(np.float16)
    assert np.array_equal(big.cpu().numpy().view(np.uint16), exp.view(np.uint16))
    # get_rows f32 / f16
    table = rng.standard_normal((50, 96)).astype(np.float32)
    ids = np.array([[49, 0, 7], [7, 7, 1]], np.int32)[0]
    assert np.array_equal(ops.get_rows(dev(table), dev(ids)).cpu().numpy(), table[ids])
    assert np.array_equal(ops.get_rows(dev(table.astype(np.float16)), dev(ids)).cpu().numpy(), table.astype(np.float16)[ids].astype(np.float32))
    # cpy: f32 -> f16 contiguous (the KQ-mask cast), f16 -> f32, transposed source
    m = rng.standard_normal((64, 320)).astype(np.float32)
    m[m > 1] = -np.inf
    out = torch.empty((64, 320), dtype=torch.float16, device="cuda")
    ops.cpy(dev(m), out)
    assert np.array_equal(out.cpu().numpy().view(np.uint16), m.astype(np.float16).view(np.uint16))
    back = torch.empty((64, 320), dtype=torch.float32, device="cuda")
    ops.cpy(out, back)
    assert np.array_equal(back.cpu().numpy(), m.astype(np.float16).astype(np.float32))
    tr = torch.empty((320, 64), dtype=torch.float32, device="cuda")
    ops.cpy(dev(m).t(), tr)
    assert np.array_equal(tr.cpu().numpy(), m.T)


def test_elementwise(ops, golden):
    g = golden["ops"]
    got = ops.glu(ops.GLU_SWIGLU, dev(g["glu_gate"]), dev(g["glu_up"])).cpu().numpy()
    assert np.allclose(got, g["glu_y"], rtol=2e-6, atol=1e-7)
    rng = np.random.default_rng(10)
    a = rng.standard_normal((2, 3, 5, 64)).astype(np.float32)
    for op, f in ((ops.ADD, np.add), (ops.SUB, np.subtract), (ops.MUL, np.multiply), (ops.DIV, np.divide)):
        for bshape in ((2, 3, 5, 64), (64,), (1, 3, 1, 64), (2, 1, 5, 1)):
            b = rng.standard_normal(bshape).astype(np.float32) + 3
            assert np.array_equal(ops.binary(op, dev(a), dev(b)).cpu().numpy(), f(a, b).astype(np.float32)), (op, bshape)
    x = rng.standard_normal(1000).astype(np.float32) * 3
    xt = torch.from_numpy(x)
    F = torch.nn.functional
    for op, ref in ((ops.SILU, F.silu(xt)), (ops.RELU, F.relu(xt)), (ops.GELU, F.gelu(xt, approximate="tanh")), (ops.TANH, torch.tanh(xt)),
                    (ops.SIGMOID, torch.sigmoid(xt)), (ops.GELU_ERF, F.gelu(xt)), (ops.NEG, -xt), (ops.EXP, torch.exp(xt)),
                    (ops.SQR, xt * xt), (ops.ABS, xt.abs()), (ops.GELU_QUICK, xt * torch.sigmoid(1.702 * xt))):
        assert np.allclose(ops.unary(op, dev(x)).cpu().numpy(), ref.numpy(), rtol=1e-5, atol=1e-6), op
    assert np.allclose(ops.scale(dev(x), 0.5, 1.0).cpu().numpy(), x * 0.5 + 1.0)
    # single-tensor GLU (halves of a row), swapped or not
    xx = rng.standard_normal((4, 128)).astype(np.float32)
    assert np.allclose(ops.glu(ops.GLU_SWIGLU, dev(xx)).cpu().numpy(), O.swiglu(xx[:, :64], xx[:, 64:]), rtol=2e-6, atol=1e-7)
    assert np.allclose(ops.glu(ops.GLU_SWIGLU, dev(xx), swapped=True).cpu().numpy(), O.swiglu(xx[:, 64:], xx[:, :64]), rtol=2e-6, atol=1e-7)
    # soft_max with f16 mask
    s = rng.standard_normal((6, 300)).astype(np.float32) * 4
    mask = np.zeros((6, 300), np.float32)
    mask[:, 250:] = -np.inf
    got = ops.soft_max(dev(s), dev(mask.astype(np.float16)), 0.3).cpu().numpy()
    assert np.allclose(got, O.soft_max(s, mask, 0.3), rtol=2e-6, atol=1e-9)


# ---------------------------------------------------------------------------------------------------------------- attention
@pytest.mark.parametrize("tag", ["dec", "pre"])
def test_flash_attn_golden(ops, golden, tag):
    g = golden["flash_attn"]
    q, k, v, mask, ref = g[f"{tag}_q"], g[f"{tag}_k"], g[f"{tag}_v"], g[f"{tag}_mask"], g[f"{tag}_out"]
    qd = dev(q).permute(1, 0, 2)                                         # [n_head, n_q, D] view, as llama-graph.cpp:1303 permutes
    got = ops.flash_attn(qd, dev(k), dev(v), dev(mask), 1.0 / np.sqrt(128)).cpu().numpy()
    # the reference CPU backend accumulates V in f16 (ops.cpp:8040-8060); we accumulate in f32 -> compare at its own noise level
    assert np.abs(got - ref).max() <= 1e-2 * np.abs(ref).max()
    exact = O.flash_attn(q, k, v, mask, 1.0 / np.sqrt(128), f16_acc=False)
    assert np.abs(got - exact).max() <= 2e-5 * np.abs(exact).max() + 1e-6


@pytest.mark.parametrize("n_kv,cur,D,n_head,n_head_kv", [(4096, 4000, 128, 32, 8), (256, 0, 128, 32, 8), (512, 300, 64, 12, 12),
                                                         (1024, 1023, 128, 8, 4), (8192, 5000, 128, 32, 8)])
def test_flash_attn_decode_cache_views(ops, n_kv, cur, D, n_head, n_head_kv):
    """Decode at the BASELINE.json shapes over a strided F16 KV-cache view (row stride = n_head_kv*D halves, head stride = D)."""
    rng = np.random.default_rng(n_kv + cur)
    q = rng.standard_normal((1, n_head, D)).astype(np.float32)
    kc = (rng.standard_normal((n_kv, n_head_kv, D)) * 0.5).astype(np.float16)
    vc = rng.standard_normal((n_kv, n_head_kv, D)).astype(np.float16)
    kc[cur + 1:] = np.float16(np.nan)                                    # never-written cells: must not leak through the mask
    vc[cur + 1:] = np.float16(np.nan)
    mask = np.zeros((64, n_kv), np.float16)
    mask[:, cur + 1:] = -np.inf
    kd, vd = dev(kc).permute(1, 0, 2), dev(vc).permute(1, 0, 2)          # [n_head_kv, n_kv, D] views with cache strides
    got = ops.flash_attn(dev(q).permute(1, 0, 2), kd, vd, dev(mask), 1.0 / np.sqrt(D)).cpu().numpy()
    kk, vv = np.ascontiguousarray(kc.transpose(1, 0, 2)), np.ascontiguousarray(vc.transpose(1, 0, 2))
    kk[:, cur + 1:], vv[:, cur + 1:] = 0, 0
    ref = O.flash_attn(q, kk, vv, mask, 1.0 / np.sqrt(D), f16_acc=False)
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max() + 1e-6


def test_flash_attn_small_batch_causal(ops):
    rng = np.random.default_rng(12)
    n_q, n_kv, D, n_head, n_head_kv = 7, 448, 128, 16, 4
    q = rng.standard_normal((n_q, n_head, D)).astype(np.float32)
    k = (rng.standard_normal((n_head_kv, n_kv, D)) * 0.5).astype(np.float16)
    v = rng.standard_normal((n_head_kv, n_kv, D)).astype(np.float16)
    mask = np.zeros((64, n_kv), np.float16)
    for i in range(n_q):
        mask[i, 400 + i + 1:] = -np.inf
    got = ops.flash_attn(dev(q).permute(1, 0, 2), dev(k), dev(v), dev(mask), 0.088).cpu().numpy()
    ref = O.flash_attn(q, k, v, mask, 0.088, f16_acc=False)
    assert np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max() + 1e-6


@pytest.mark.parametrize("n_q,n_kv,past,D,n_head,n_head_kv", [(512, 512, 0, 128, 8, 2), (300, 768, 411, 128, 8, 2), (64, 256, 150, 64, 6, 6),
                                                             (17, 100, 83, 128, 4, 1), (2048, 2048, 0, 128, 4, 1)])
def test_flash_attn_prefill_tensor_core(ops, n_q, n_kv, past, D, n_head, n_head_kv):
    """>= 16 query tokens -> k_fa_prefill (mma.sync tiles, csrc/fa_prefill.cu): causal mask over a cache with `past` earlier positions, ragged
    sizes, GQA.  P is rounded to f16 for the P.V product (as in the reference's fattn-mma-f16), so the bar is the reference's own
    FLASH_ATTN_EXT bar, NMSE <= 5e-4 (tests/test-backend-ops.cpp:5085), plus a tighter elementwise bound against the f32 oracle."""
    rng = np.random.default_rng(n_q + n_kv + D)
    q = rng.standard_normal((n_q, n_head, D)).astype(np.float32)
    k = (rng.standard_normal((n_head_kv, n_kv, D)) * 0.5).astype(np.float16)
    v = rng.standard_normal((n_head_kv, n_kv, D)).astype(np.float16)
    n_q_pad = (n_q + 63) // 64 * 64
    mask = np.zeros((n_q_pad, n_kv), np.float16)
    for i in range(n_q):
        mask[i, past + i + 1:] = -np.inf
    mask[n_q:] = -np.inf
    scale = 1.0 / np.sqrt(D)
    got = ops.flash_attn(dev(q).permute(1, 0, 2), dev(k), dev(v), dev(mask), scale).cpu().numpy()
    assert np.isfinite(got).all()
    if n_q * n_kv * n_head <= 2 ** 23:
        ref = O.flash_attn(q, k, v, mask, scale, f16_acc=False)
    else:                                     # the C oracle is O(n_q n_kv): use it on a sample of query rows, f64 torch for all of them
        rows = np.unique(np.concatenate([np.arange(4), np.arange(n_q - 4, n_q), rng.integers(0, n_q, 24)]))
        sub = O.flash_attn(q[rows], k, v, mask[rows], scale, f16_acc=False)
        assert nmse(got[rows], sub) <= 5e-4 and np.abs(got[rows] - sub).max() <= 4e-3 * np.abs(sub).max()
        qt = dev(q).half().double().permute(1, 0, 2)                                     # the oracle rounds Q to f16
        kt, vt = dev(k).double().repeat_interleave(n_head // n_head_kv, 0), dev(v).double().repeat_interleave(n_head // n_head_kv, 0)
        sc = qt @ kt.transpose(1, 2) * scale + dev(mask[:n_q]).double()[None]
        ref = (torch.softmax(sc, -1) @ vt).permute(1, 0, 2).cpu().numpy()
    assert nmse(got, ref) <= 5e-4, nmse(got, ref)
    assert np.abs(got - ref).max() <= 4e-3 * np.abs(ref).max(), np.abs(got - ref).max() / np.abs(ref).max()


def test_qkv_post_matches_separate_ops(ops):
    """b200_qkv_post == rms_norm*w -> rope -> set_rows run as separate C-ABI calls (Qwen3 shapes; 2 tokens)."""
    rng = np.random.default_rng(13)
    D, n_head, n_head_kv, n_tok, n_ctx = 128, 32, 8, 2, 64
    q = rng.standard_normal((n_tok, n_head, D)).astype(np.float32)
    k = rng.standard_normal((n_tok, n_head_kv, D)).astype(np.float32)
    v = rng.standard_normal((n_tok, n_head_kv * D)).astype(np.float32)
    qw = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
    kw = (1 + 0.1 * rng.standard_normal(D)).astype(np.float32)
    pos = np.array([17, 4000], np.int32)
    idx = np.array([5, 41], np.int64)
    qd, kd, vd, qwd, kwd, posd, idxd = map(dev, (q, k, v, qw, kw, pos, idx))
    # separate ops
    q_ref = ops.rope(ops.rms_norm(qd, 1e-6, w=qwd), posd, D, 2)
    k_ref = ops.rope(ops.rms_norm(kd, 1e-6, w=kwd), posd, D, 2)
    kc_ref = torch.zeros((n_ctx, n_head_kv * D), dtype=torch.float16, device="cuda")
    vc_ref = torch.zeros_like(kc_ref)
    ops.set_rows(k_ref.reshape(n_tok, -1), idxd, kc_ref)
    ops.set_rows(vd, idxd, vc_ref)
    # fused
    kc, vc = torch.zeros_like(kc_ref), torch.zeros_like(vc_ref)
    p = ops.RopeParams(D, 2, 40960, 1e6, 1.0, 0.0, 1.0, 32.0, 1.0)
    P = C.c_void_p
    ops.check(ops.lib().b200_qkv_post(P(qd.data_ptr()), P(kd.data_ptr()), P(vd.data_ptr()), P(qwd.data_ptr()), P(kwd.data_ptr()),
                                      P(posd.data_ptr()), P(idxd.data_ptr()), ops.I64, P(kc.data_ptr()), P(vc.data_ptr()),
                                      C.c_int64(n_head_kv * D * 2), C.c_int64(n_head_kv * D * 2), D, n_head, n_head_kv,
                                      C.c_int64(n_tok), C.c_int64(n_head * D), C.c_int64(n_head_kv * D), C.c_int64(n_head_kv * D),
                                      C.byref(p), C.c_float(1e-6), ops.stream()))
    assert torch.allclose(qd, q_ref, rtol=2e-6, atol=2e-6)
    assert torch.equal(vc, vc_ref)
    assert (kc.float() - kc_ref.float()).abs().max().item() <= 2e-3        # one f16 ulp at |k| <= 4
    ref_np = O.rope(O.rms_norm(q.reshape(-1, D), 1e-6).reshape(q.shape) * qw, pos, D, 2)
    assert np.abs(qd.cpu().numpy() - ref_np).max() <= 1e-5


# ---------------------------------------------------------------------------------------------------------------- decode engine
@pytest.mark.parametrize("which", ["tiny", "qwen3-8b-4layers"])
def test_decode_engine_matches_per_op_path(which):
    """The one-kernel decode step (b200_decoder_*) must reproduce the per-op launch sequence: same integer dot products, same f32 op
    order inside a block; only the split-KV chunking and the K-split partial sums regroup f32 additions -> tight tolerance on the
    logits and identical KV-cache rows."""
    dec = load_package().decode
    cfg = (dec.LLMConfig(name="small", n_embd=2048, n_layer=3, n_head=8, n_head_kv=2, n_ff=8192, n_vocab=4096, n_ctx=512) if which == "tiny"
           else dec.LLMConfig(n_layer=4, n_vocab=8192, n_ctx=1024))
    n_kv, depth = 512, 300
    A = dec.Qwen3Decoder(cfg, "cuda:0", seed=1)
    B = dec.Qwen3Decoder(cfg, "cuda:0", seed=1)                           # same seed -> identical weights
    for a, b in zip(A.L, B.L):
        a["k_cache"][:depth].normal_(0, 0.5)
        a["v_cache"][:depth].normal_(0, 1.0)
        b["k_cache"].copy_(a["k_cache"])
        b["v_cache"].copy_(a["v_cache"])
    B.build_engine()
    x = torch.randn(cfg.n_embd, device="cuda") * 0.05
    for step in range(3):
        hi = dec.Qwen3Decoder.host_inputs(cfg, depth + step, n_kv, pinned=False)
        for M in (A, B):
            M.x_in.copy_(x * (1 + step))
            M.pos.copy_(hi["pos"])
            M.kv_idx.copy_(hi["kv_idx"])
            M.mask_f32[:, :n_kv].copy_(hi["mask"])
        A.step(n_kv)
        B.step_engine(n_kv)
        torch.cuda.synchronize()
        la, lb = A.logits.cpu().numpy(), B.logits.cpu().numpy()
        assert np.isfinite(lb).all()
        scale = np.abs(la).max()
        # north_star tolerance: logits within 1e-3 relative (an int8 activation step flipped by a last-ulp RMS-norm difference costs ~3e-4)
        assert np.abs(la - lb).max() <= 1e-3 * scale, (step, np.abs(la - lb).max(), scale)
        assert int(la.argmax()) == int(lb.argmax())                       # greedy token id
        for a, b in zip(A.L, B.L):
            row = depth + step
            # the sum of squares of the RMS norm is reduced by 512 threads in one path and 384 in the other: last-ulp differences in the
            # scale can move an int8 activation by one step -> allow a couple of f16 ulps on the freshly written cache rows
            for c in ("k_cache", "v_cache"):
                ra, rb = a[c][row].float(), b[c][row].float()
                assert (ra - rb).abs().max().item() <= 3e-3 * max(1.0, ra.abs().max().item()), c


# ---------------------------------------------------------------------------------------------------------------- whole token vs the ORACLE
def _oracle_model(dec, ops, cfg, rng, variant):
    """A decoder whose weights are known on the host as NATIVE ggml blocks: (engine-side Qwen3Decoder, oracle-side layer dicts, head dict)."""
    import oracle_decode  # noqa: F401  (test infrastructure)
    D = dec.Qwen3Decoder(cfg, "cuda:0", has_head=variant != "stage", seed=3)
    E, F, hd = cfg.n_embd, cfg.n_ff, cfg.head_dim
    q, kv = cfg.n_head * hd, cfg.n_head_kv * hd
    shapes = {"wq": (q, E), "wk": (kv, E), "wv": (kv, E), "wo": (E, q), "gate": (F, E), "up": (F, E), "down": (E, F)}
    o_layers = []
    for il, lw in enumerate(D.L):
        ty = dict(lw["types"])
        if variant == "q5k":                                                  # the q5_K route of the engine (fragments re-read from shared memory)
            ty.update({"wq": ops.Q5_K, "wk": ops.Q5_K, "wo": ops.Q5_K, "down": ops.Q5_K})     # (a q5_K gate/up PAIR exceeds a ring slot: not streamable)
        if variant in ("q4_0", "q8_0"):                                       # whole-model Q4_0 / Q8_0 files: q8_0 activation records (group 32), planar weight rows
            ty = {n: QT[variant] for n in shapes}
        lw["types"] = ty
        ol = {"types": ty}
        for n, (m, k) in shapes.items():
            blocks = rand_blocks(rng, ty[n], m * k // O.BLOCK[ty[n]][0])
            if ty[n] in (O.Q4_K, O.Q5_K):                                     # scales like a real file: d, dmin ~ 1e-4
                blocks[:, 0:4] = np.frombuffer(rng.uniform(2e-5, 2e-4, (blocks.shape[0], 2)).astype(np.float16).tobytes(), np.uint8).reshape(-1, 4)
            elif ty[n] == O.Q6_K:
                blocks[:, 208:210] = np.frombuffer(rng.uniform(2e-5, 2e-4, blocks.shape[0]).astype(np.float16).tobytes(), np.uint8).reshape(-1, 2)
            else:                                                             # q4_0 (codes -8..7) / q8_0 (codes -128..127): d so that the weights are ~1e-2
                lo, hi = (5e-4, 2e-3) if ty[n] == O.Q4_0 else (3e-5, 1.2e-4)
                blocks[:, 0:2] = np.frombuffer(rng.uniform(lo, hi, blocks.shape[0]).astype(np.float16).tobytes(), np.uint8).reshape(-1, 2)
            ol[n] = blocks
            wd = dev(blocks.reshape(-1))
            lw[n] = ops.to_planar(ty[n], wd) if ty[n] in ops.PAYLOAD else wd
        for n in ("attn_norm", "ffn_norm", "q_norm", "k_norm"):
            ol[n] = lw[n].cpu().numpy().astype(np.float32)
        if variant == "llama":                                                # llm_build_llama: no q/k norm, ROPE mode 0 (adjacent pairs)
            lw["q_norm"] = lw["k_norm"] = None
            ol["q_norm"] = ol["k_norm"] = None
        o_layers.append(ol)
    head = None
    if D.has_head:
        blocks = rand_blocks(rng, O.Q6_K, cfg.n_vocab * E // 256)
        blocks[:, 208:210] = np.frombuffer(rng.uniform(2e-5, 2e-4, blocks.shape[0]).astype(np.float16).tobytes(), np.uint8).reshape(-1, 2)
        D.lm_head = ops.to_planar(ops.Q6_K, dev(blocks.reshape(-1)))
        head = {"out_norm": D.out_norm.cpu().numpy().astype(np.float32), "lm_head": blocks, "type": O.Q6_K, "n_vocab": cfg.n_vocab}
    if variant == "llama":
        D.rope = ops.RopeParams(hd, 0, cfg.n_ctx_orig, cfg.rope_base, 1.0, 0.0, 1.0, 32.0, 1.0)
    return D, o_layers, head


@pytest.mark.parametrize("variant", ["qwen3", "llama", "stage", "q5k", "q4_0", "q8_0"])
def test_decode_engine_whole_token_matches_oracle(ops, variant):
    """north_star parity for the dominant kernel: ONE whole token of k_stream (b200_decoder_step) against the CPU ORACLE chain
    (tests/oracle_decode.py: q8_K quantise -> integer-dot matvec -> RMS_NORM -> ROPE -> F16 cache write -> FLASH_ATTN_EXT -> SWIGLU, every
    op an oracle/*.c restatement pinned to the live reference).  Qwen3-8B layer shapes (4096 / 12288 / 32 heads / 8 kv heads), 4-layer slice,
    Q4_K_M type mix.  Bars: logits <= 1e-3 relative, same greedy token, freshly written KV rows within one F16 ulp.  Variants: llama arch
    (no q/k-norm, ROPE mode 0), a head-less pipeline stage (x_out), the q5_K route; hidden_out (omni embeddings=on) is checked on qwen3."""
    import oracle_decode as OD
    dec = load_package().decode
    rng = np.random.default_rng(11)
    cfg = dec.LLMConfig(n_layer=4, n_vocab=8192, n_ctx=1024)
    if variant == "q8_0":                 # a q8_0 gate/up row PAIR must fit one 4608-byte ring slot: n_embd <= 2048 (k + k/16 bytes per row)
        cfg = dec.LLMConfig(name="q8_0-2048", n_embd=2048, n_layer=3, n_head=16, n_head_kv=4, n_ff=6144, n_vocab=8192, n_ctx=1024)
    D, o_layers, head = _oracle_model(dec, ops, cfg, rng, variant)
    n_kv, depth = 512, 300
    kvw = cfg.n_head_kv * cfg.head_dim
    for lw, ol in zip(D.L, o_layers):
        ol["k_cache"] = (rng.standard_normal((cfg.n_ctx, kvw)) * 0.5).astype(np.float16)
        ol["v_cache"] = rng.standard_normal((cfg.n_ctx, kvw)).astype(np.float16)
        lw["k_cache"].copy_(dev(ol["k_cache"]))
        lw["v_cache"].copy_(dev(ol["v_cache"]))
    hidden = torch.zeros(cfg.n_embd, device="cuda") if D.has_head else None
    D.build_engine(hidden_out=hidden)
    x = (rng.standard_normal(cfg.n_embd) * 0.05).astype(np.float32)
    for step in range(2):
        pos = depth + step
        hi = dec.Qwen3Decoder.host_inputs(cfg, pos, n_kv, pinned=False)
        xs = x * (1 + step)
        D.x_in.copy_(dev(xs)); D.pos.copy_(hi["pos"]); D.kv_idx.copy_(hi["kv_idx"]); D.mask_f32[:, :n_kv].copy_(hi["mask"])
        D.step_engine(n_kv)
        torch.cuda.synchronize()
        bar = 1e-3
        if variant in ("q4_0", "q8_0"):
            # q8_0 activation records re-quantise every 32 values with their own scale: on these random weights the ORACLE itself moves by more than 1e-3 when its
            # input is nudged by 2e-6 (what another f32 summation order does; tests/test_chaos_yardstick.py).  The bar is what the oracle differs from its nudged self.
            import copy
            yard = []
            for trial in range(3):
                ol2 = copy.deepcopy(o_layers)
                nudged = (xs.astype(np.float64) * (1.0 + 2e-6 * rng.standard_normal(xs.size))).astype(np.float32)
                pl, _, _ = OD.oracle_token(cfg, ol2, nudged, pos, n_kv, head, rope_mode=2, f16_acc=False)
                bl, _, _ = OD.oracle_token(cfg, copy.deepcopy(o_layers), xs, pos, n_kv, head, rope_mode=2, f16_acc=False)
                yard.append(float(np.abs(pl - bl).max() / np.abs(bl).max()))
            bar = max(1e-3, 2.0 * max(yard))
            print(variant, "oracle vs nudged oracle:", ["%.2e" % y for y in yard], "bar %.2e" % bar)
        ref_logits, ref_x, ref_hn = OD.oracle_token(cfg, o_layers, xs, pos, n_kv, head, rope_mode=0 if variant == "llama" else 2, f16_acc=False)
        # the residual stream (what a pipeline stage hands to the next one)
        gx = D.engine_x_out.cpu().numpy() if not D.has_head else None
        if gx is not None:
            assert np.abs(gx - ref_x).max() <= 1e-3 * np.abs(ref_x).max(), (variant, step, np.abs(gx - ref_x).max(), np.abs(ref_x).max())
        if D.has_head:
            got = D.logits.cpu().numpy()
            assert np.isfinite(got).all()
            rel = np.abs(got - ref_logits).max() / np.abs(ref_logits).max()
            assert rel <= bar, (variant, step, rel, bar)
            top2 = np.sort(ref_logits)[-2:]
            if bar == 1e-3 or top2[1] - top2[0] > 2 * bar * np.abs(ref_logits).max():
                assert int(got.argmax()) == int(ref_logits.argmax()), (variant, step)
            hn = hidden.cpu().numpy()
            assert np.abs(hn - ref_hn).max() <= bar * np.abs(ref_hn).max(), (variant, step)
        for lw, ol in zip(D.L, o_layers):                                     # SET_ROWS parity of this token's cache rows
            for c in ("k_cache", "v_cache"):
                g, r = lw[c][pos].float().cpu().numpy(), ol[c][pos].astype(np.float32)
                assert np.abs(g - r).max() <= 2e-3 * (bar / 1e-3) * max(1.0, np.abs(r).max()), (variant, step, c, np.abs(g - r).max())


# ---------------------------------------------------------------------------------------------------------------- APM / VPM encoder ops
def test_norm_matches_oracle(ops):
    rng = np.random.default_rng(31)
    for shape in ((9, 1280), (3, 5, 1152), (2, 4096), (1, 7)):
        x = (rng.standard_normal(shape) * 3 + 0.5).astype(np.float32)
        got = ops.norm(dev(x), 1e-5).cpu().numpy()
        assert np.allclose(got, O.norm(x, 1e-5), rtol=2e-6, atol=3e-6), shape


@pytest.mark.parametrize("case", ["whisper_conv1_s1", "whisper_conv2_s2", "siglip_patch14", "pad_dilate_f32"])
def test_im2col_matches_oracle(ops, case):
    """bit-exact: a gather plus one F32 -> F16 rounding."""
    rng = np.random.default_rng(32)
    if case.startswith("whisper"):
        s = 1 if case.endswith("s1") else 2
        x = rng.standard_normal((2, 80, 300)).astype(np.float32)                   # [N, IC = mel bins, IW = frames]
        k = torch.zeros((16, 80, 3), dtype=torch.float16, device="cuda")
        got = ops.im2col(k, dev(x), s, 0, 1, 0, 1, 0, False).cpu().numpy()
        ref = O.im2col(x.reshape(2, 80, 1, 300), 1, 3, s, 0, 1, 0, 1, 0)[:, 0]
    elif case == "siglip_patch14":
        x = rng.standard_normal((2, 3, 98, 70)).astype(np.float32)
        k = torch.zeros((8, 3, 14, 14), dtype=torch.float16, device="cuda")
        got = ops.im2col(k, dev(x), 14, 14, 0, 0, 1, 1, True).cpu().numpy()
        ref = O.im2col(x, 14, 14, 14, 14, 0, 0, 1, 1)
    else:
        x = rng.standard_normal((1, 5, 33, 29)).astype(np.float32)
        k = torch.zeros((4, 5, 3, 3), dtype=torch.float16, device="cuda")
        got = ops.im2col(k, dev(x), 2, 1, 1, 2, 2, 1, True, dtype=torch.float32).cpu().numpy()
        ref = O.im2col(x, 3, 3, 2, 1, 1, 2, 2, 1, f16=False)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert np.array_equal(got.view(np.uint16 if got.dtype == np.float16 else np.uint32), ref.view(np.uint16 if ref.dtype == np.float16 else np.uint32))


def test_pool_1d_matches_oracle(ops):
    """bit-exact: the sequential f32 sum of the CPU loop, then one division."""
    rng = np.random.default_rng(33)
    x = rng.standard_normal((3, 1280, 250)).astype(np.float32)                      # audition.cpp:697: k = s = 5 over the token axis
    for op in (0, 1):
        assert np.array_equal(ops.pool_1d(dev(x), op, 5).cpu().numpy(), O.pool_1d(x, op, 5))
    x16 = x[:1].astype(np.float16)
    assert np.array_equal(ops.pool_1d(dev(x16), 1, 2).cpu().numpy(), O.pool_1d(x16.astype(np.float32), 1, 2))


def test_prefill_fused_activation_tiles_are_bit_identical(ops, monkeypatch):
    """RMS_NORM / FLASH_ATTN_EXT / SWIGLU writing the next MUL_MAT's F16 activation tiles directly (b200_*_tiles + B200_MM_REUSE_ACT) and the residual ADD riding in the
    GEMM epilogue must give exactly the logits of the unfused sequence (F32 result + the MUL_MAT's own conversion pass): the tiles hold the same F16 roundings of the
    same F32 values.  The merged q / k / v and gate / up launches (b200_mul_mat_multi) accumulate the whole K in one TMEM accumulator where the separate launch of a
    small matrix splits K over two CTAs: with them on, the two runs agree to F32 summation order (amplified by the F16 roundings of 2 layers), not bit for bit."""
    dec = load_package().decode
    cfg = dec.LLMConfig(name="small", n_embd=2048, n_layer=2, n_head=16, n_head_kv=4, n_ff=6144, n_vocab=4096, n_ctx=512)
    D = dec.Qwen3Decoder(cfg, "cuda:0", seed=2)
    x = torch.randn(200, cfg.n_embd, device="cuda") * 0.05
    a, _ = D.prefill(x, 0, 256, fused_tiles=False)
    ka = [lw["k_cache"][:200].clone() for lw in D.L]
    monkeypatch.setenv("B200_NO_MULTI", "1")
    b, _ = D.prefill(x, 0, 256, fused_tiles=True)
    torch.cuda.synchronize()
    assert torch.isfinite(a).all()
    assert torch.equal(a, b)
    for lw, k0 in zip(D.L, ka):
        assert torch.equal(lw["k_cache"][:200], k0)
    monkeypatch.setenv("B200_NO_MULTI", "0")
    c, _ = D.prefill(x, 0, 256, fused_tiles=True)
    torch.cuda.synchronize()
    assert float((c - a).abs().max() / a.abs().max()) <= 2e-3
    for lw, k0 in zip(D.L, ka):
        assert float((lw["k_cache"][:200].float() - k0.float()).abs().max()) <= 4e-3 * max(1.0, float(k0.float().abs().max()))


# ---- Token2Wav op set (csrc/ops_wave.cu).  Oracle = numpy restatements of the reference CPU loops (file:line in each case); the live reference checks
# the same kernels through its own harness in tests/test_plugin_gpu.py::test_reference_backend_ops_harness.
def test_wave_unary_ops(ops):
    rng = np.random.default_rng(21)
    x = (rng.standard_normal((3, 5, 77)) * 3).astype(np.float32)
    xd = dev(x)
    want = {   # ggml-cpu/unary-ops.cpp op_* and ops.cpp:2453-2480 (leaky_relu), :5309-5340 (clamp)
        ops.SIN: np.sin(x), ops.COS: np.cos(x), ops.ELU: np.where(x > 0, x, np.expm1(x)), ops.STEP: (x > 0).astype(np.float32), ops.SGN: np.sign(x),
        ops.HARDSWISH: x * np.clip((x + 3) / 6, 0, 1), ops.HARDSIGMOID: np.clip((x + 3) / 6, 0, 1),
    }
    for op, w in want.items():
        np.testing.assert_allclose(ops.unary_param(op, xd).cpu().numpy(), w.astype(np.float32), rtol=2e-6, atol=2e-6, err_msg=str(op))
    np.testing.assert_allclose(ops.unary_param(ops.LOG, dev(np.abs(x) + 0.1)).cpu().numpy(), np.log(np.abs(x) + 0.1), rtol=2e-6, atol=2e-6)
    assert np.array_equal(ops.unary_param(ops.LEAKY_RELU, xd, 0.1).cpu().numpy(), np.maximum(x, 0) + np.float32(0.1) * np.minimum(x, 0))
    assert np.array_equal(ops.unary_param(ops.CLAMP, xd, -1.5, 2.0).cpu().numpy(), np.clip(x, -1.5, 2.0))
    h = xd.half()                                                        # F16 in / out: computed in F32, rounded once
    assert torch.equal(ops.unary_param(ops.CLAMP, h, -1.5, 2.0), h.clamp(-1.5, 2.0))
    y = ops.unary_param(ops.ELU, xd.permute(2, 0, 1))                   # strided source
    np.testing.assert_allclose(y.cpu().numpy(), np.where(x > 0, x, np.expm1(x)).transpose(2, 0, 1), rtol=2e-6, atol=2e-6)


def test_wave_concat_repeat_pad_arange_sum_rows(ops):
    rng = np.random.default_rng(22)
    a = rng.standard_normal((3, 4, 5, 11)).astype(np.float32)
    for dim in range(4):                                                 # ops.cpp:1968-2010: dst = [a ; b] along ggml dim
        shp = list(a.shape); shp[3 - dim] = 7
        b = rng.standard_normal(shp).astype(np.float32)
        assert np.array_equal(ops.concat(dev(a), dev(b), dim).cpu().numpy(), np.concatenate([a, b], axis=3 - dim))
    ai = rng.integers(-9, 9, (2, 3, 4), dtype=np.int32); bi = rng.integers(-9, 9, (2, 3, 6), dtype=np.int32)
    assert np.array_equal(ops.concat(dev(ai), dev(bi), 0).cpu().numpy(), np.concatenate([ai, bi], axis=2))
    at = dev(a).permute(0, 1, 3, 2)                                      # non-contiguous first operand
    bt = dev(rng.standard_normal((3, 4, 11, 2)).astype(np.float32))
    assert torch.equal(ops.concat(at, bt, 0), torch.cat([at, bt], dim=3))
    for reps in [(1, 1, 1, 2), (2, 1, 3, 1), (1, 2, 1, 1), (2, 2, 2, 2)]:  # ops.cpp:1637-1700: dst[i] = src[i mod ne]
        out_shape = [s * r for s, r in zip(a.shape, reps)]
        assert np.array_equal(ops.repeat(dev(a), out_shape).cpu().numpy(), np.tile(a, reps))
    h = dev(a).half()
    assert torch.equal(ops.repeat(h, [3, 4, 10, 11]), h.repeat(1, 1, 2, 1))
    # PAD: ops.cpp:7592-7640
    lr = [1, 2, 3, 4, 0, 1, 2, 0]
    want = np.pad(a, [(lr[6], lr[7]), (lr[4], lr[5]), (lr[2], lr[3]), (lr[0], lr[1])])
    assert np.array_equal(ops.pad(dev(a), lr).cpu().numpy(), want)
    # PAD_REFLECT_1D: ops.cpp:7664-7692
    x = rng.standard_normal((2, 80, 300)).astype(np.float32)
    assert np.array_equal(ops.pad_reflect_1d(dev(x), 7, 3).cpu().numpy(), np.pad(x, [(0, 0), (0, 0), (7, 3)], mode="reflect"))
    # ARANGE: ops.cpp:7762-7785, value = start + step * i in f32
    got = ops.arange(0.5, 100.25, 0.75, "cuda").cpu().numpy()
    i = np.arange(len(got), dtype=np.float32)
    assert len(got) == int(np.ceil((100.25 - 0.5) / 0.75)) and np.array_equal(got, np.float32(0.5) + np.float32(0.75) * i)
    # SUM_ROWS: ops.cpp:1399-1430 (f64 accumulator, rounded once)
    for shape in [(3, 5, 1000), (2, 33), (1, 4097), (7, 1)]:
        x = rng.standard_normal(shape).astype(np.float32)
        np.testing.assert_array_equal(ops.sum_rows(dev(x)).cpu().numpy(), x.astype(np.float64).sum(-1, keepdims=True).astype(np.float32))


@pytest.mark.parametrize("L,Cin,Cout,K,s0,wt", [(197, 32, 16, 16, 1, "f32"), (3, 2, 3, 2, 3, "f32"), (3, 2, 1, 3, 1, "f32"), (50, 512, 256, 16, 8, "f16"), (121, 64, 32, 11, 5, "f32"),
                                                 (2, 1, 1, 3, 1, "f32")])
def test_wave_conv_transpose_1d(ops, L, Cin, Cout, K, s0, wt):
    """HiFiGAN upsampling (token2wav-impl.cpp ggml_conv_transpose_1d(w, x, stride, 0, 1)); oracle: the scatter form of ops.cpp:6040-6130 in f64."""
    rng = np.random.default_rng(L * 31 + K)
    w = rng.standard_normal((Cin, Cout, K)).astype(np.float32)
    x = rng.standard_normal((Cin, L)).astype(np.float32)
    wd = dev(w).half() if wt == "f16" else dev(w)
    w64 = wd.float().cpu().numpy().astype(np.float64)
    ref = np.zeros((Cout, (L - 1) * s0 + K))
    for k in range(K):
        ref[:, k:k + (L - 1) * s0 + 1:s0] += np.einsum("ic,il->cl", w64[:, :, k], x.astype(np.float64))
    got = ops.conv_transpose_1d(wd, dev(x), s0).cpu().numpy()
    mag = np.zeros_like(ref)
    for k in range(K):
        mag[:, k:k + (L - 1) * s0 + 1:s0] += np.einsum("ic,il->cl", np.abs(w64[:, :, k]), np.abs(x.astype(np.float64)))
    assert np.all(np.abs(got - ref) <= 2e-6 * mag + 1e-9)


# ---- data formats adjacent to the path (csrc/quant_rows.cu): quantised KV cache, quantised token_embd, K-shift ----------------------------------
def _kv_desc(ops, buf, t, D, n_kv, n_head_kv):
    """the view llama builds of a cache tensor [n_embd_kv, kv_size] (src/llama-kv-cache.cpp:981-1009): [D, n_kv, n_head_kv], heads side by side inside a row"""
    blk, bs = O.BLOCK[t]
    return ops.T(buf, t, ne=[D, n_kv, n_head_kv], nb=[bs, n_head_kv * D // blk * bs, D // blk * bs, n_kv * n_head_kv * D // blk * bs])


@pytest.mark.parametrize("name", ["q8_0", "q4_0"])
def test_set_rows_and_cpy_into_quantised_cache_bit_exact(ops, name):
    """SET_ROWS / CPY F32 -> q8_0 | q4_0 (-ctk / -ctv): the blocks the CPU backend's from_float writes (quantize_row_q8_0 x86 flavour / quantize_row_q4_0_ref)."""
    t = QT[name]
    rng = np.random.default_rng(31)
    n_tok, width, kv_size = 37, 1024, 96
    x = (rng.standard_normal((n_tok, width)) * 2).astype(np.float32)
    x[3, 64:96] = 0.0                                   # an all-zero block
    x[5, 0:16] = 3.0; x[5, 16:32] = -3.0                # tied magnitudes of both signs (q4_0 keeps the first one's sign)
    idx = rng.permutation(kv_size)[:n_tok].astype(np.int64)
    rb = O.row_size(t, width)
    cache = torch.zeros((kv_size, rb), dtype=torch.uint8, device="cuda")
    quant = O.quantize_q8_0 if t == O.Q8_0 else O.quantize_q4_0
    want = np.zeros((kv_size, rb), np.uint8)
    for i, r in enumerate(idx):
        want[r] = quant(x[i])
    ops.set_rows(dev(x), dev(idx), ops.T(cache, t, ne=[width, kv_size]))
    torch.cuda.synchronize()
    assert np.array_equal(cache.cpu().numpy(), want)
    flat = torch.zeros((n_tok, rb), dtype=torch.uint8, device="cuda")
    ops.cpy(dev(x), ops.T(flat, t, ne=[width, n_tok]))
    back = torch.empty((n_tok, width), dtype=torch.float32, device="cuda")
    ops.cpy(ops.T(flat, t, ne=[width, n_tok]), back)
    torch.cuda.synchronize()
    assert np.array_equal(flat.cpu().numpy(), np.stack([quant(r) for r in x]))
    assert np.array_equal(back.cpu().numpy(), O.dequant(t, flat.cpu().numpy(), width))


@pytest.mark.parametrize("name", list(QT))
def test_get_rows_from_quantised_rows(ops, name):
    """GET_ROWS on a quantised token_embd, native blocks and (q4_0 / q8_0 / q6_K) the planar weight layout; oracle = dequantize_row_* port."""
    t = QT[name]
    rng = np.random.default_rng(41 + t)
    n_rows, k = 300, 1024
    blocks = rand_blocks(rng, t, n_rows * k // O.BLOCK[t][0])
    idx = rng.integers(0, n_rows, (2, 17)).astype(np.int32)
    want = O.dequant(t, blocks, k)[idx]
    exact = t in (O.Q4_0, O.Q8_0, O.Q6_K)              # products only; q4_K / q5_K end in d*q - m, which nvcc and gcc may or may not contract to one FMA
    layouts = [(dev(blocks), ops.LAYOUT_NATIVE)] + ([(ops.to_planar(t, dev(blocks)), ops.LAYOUT_PLANAR)] if t in ops.PAYLOAD else [])
    for w, lay in layouts:
        got = ops.get_rows(ops.T(w, t, ne=[k, n_rows], layout=lay), dev(idx[0])).cpu().numpy()
        if exact:
            assert np.array_equal(got, want[0]), lay
        else:
            np.testing.assert_allclose(got, want[0], rtol=1e-6, atol=1e-7)
    wb = dev(blocks)                                   # (keep the device copy alive: a descriptor only holds its address)
    got = ops.get_rows(ops.T(wb, t, ne=[k, n_rows // 2, 2]), dev(idx % (n_rows // 2))).cpu().numpy()      # batched: idx [2, 17] over src [k, 150, 2]
    ref = O.dequant(t, blocks, k).reshape(2, n_rows // 2, k)
    for b in range(2):
        np.testing.assert_allclose(got[b], ref[b][idx[b] % (n_rows // 2)], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name,n_q,n_kv,D,n_head,n_head_kv", [("q8_0", 1, 1024, 128, 32, 8), ("q4_0", 1, 512, 128, 8, 2), ("q8_0", 3, 256, 64, 12, 12), ("q8_0", 200, 512, 128, 8, 2),
                                                               ("q4_0", 64, 256, 128, 8, 2)])
def test_flash_attn_over_quantised_kv(ops, name, n_q, n_kv, D, n_head, n_head_kv):
    """FLASH_ATTN_EXT with a q8_0 / q4_0 KV cache.  >= 16 query tokens: K and V are staged as F16 (the reference's launch_fattn does the same to_fp16 pass), so the result
    must be bit for bit the F16-cache result on the dequantised, F16-rounded values (that F16 run is itself pinned to the oracle by the tests above).  Decode (< 16 query
    tokens): k_fa_decode dequantises in its loads (d * code in F32, no F16 rounding of the products), so it agrees with the F16-staged result to that rounding only."""
    t = QT[name]
    rng = np.random.default_rng(n_q * 7 + n_kv)
    past = n_kv - n_q - 5 if n_q > 1 else n_kv - 9
    kf = (rng.standard_normal((n_kv, n_head_kv * D)) * 0.5).astype(np.float32)
    vf = rng.standard_normal((n_kv, n_head_kv * D)).astype(np.float32)
    quant = O.quantize_q8_0 if t == O.Q8_0 else O.quantize_q4_0
    kq = np.stack([quant(r) for r in kf]); vq = np.stack([quant(r) for r in vf])
    k16 = dev(O.dequant(t, kq, n_head_kv * D)).half().view(n_kv, n_head_kv, D).permute(1, 0, 2)
    v16 = dev(O.dequant(t, vq, n_head_kv * D)).half().view(n_kv, n_head_kv, D).permute(1, 0, 2)
    q = dev(rng.standard_normal((n_head, n_q, D)).astype(np.float32))
    n_pad = (n_q + 63) // 64 * 64
    mask = torch.full((n_pad, n_kv), float("-inf"), dtype=torch.float16, device="cuda")
    for i in range(n_q):
        mask[i, :past + i + 1] = 0
    want = ops.flash_attn(q, k16, v16, mask, 1.0 / D ** 0.5)
    kqd, vqd = dev(kq), dev(vq)
    got = ops.flash_attn(q, _kv_desc(ops, kqd, t, D, n_kv, n_head_kv), _kv_desc(ops, vqd, t, D, n_kv, n_head_kv), mask, 1.0 / D ** 0.5)
    torch.cuda.synchronize()
    assert torch.isfinite(want).all()
    if n_q >= 16:
        assert torch.equal(got, want)
    else:
        assert float((got - want).abs().max()) <= 1e-3 * float(want.abs().max())
        # and against the exact dequantised values in f64 (what the CPU's v_to_float + F32 accumulation computes, up to its q8_0-quantised Q): softmax(q.K^T) V
        kd = dev(O.dequant(t, kq, n_head_kv * D)).double().view(n_kv, n_head_kv, D).permute(1, 0, 2).repeat_interleave(n_head // n_head_kv, 0)
        vd = dev(O.dequant(t, vq, n_head_kv * D)).double().view(n_kv, n_head_kv, D).permute(1, 0, 2).repeat_interleave(n_head // n_head_kv, 0)
        sc = (q.half().double() @ kd.transpose(1, 2)) / D ** 0.5 + mask[:n_q].double()[None]
        ref = (torch.softmax(sc, -1) @ vd).permute(1, 0, 2)
        assert float((got.double() - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


def test_rope_f16_in_place_is_the_f32_rope_rounded_once(ops):
    """K-shift of the F16 cache (ggml_rope_ext_inplace on a cache view; CPU ggml_compute_forward_rope_f16: F16 -> F32, rotate, -> F16)."""
    rng = np.random.default_rng(51)
    n_tok, n_head, D = 300, 8, 128
    x = dev(rng.standard_normal((n_tok, n_head, D)).astype(np.float32)).half()
    pos = dev(rng.integers(-50, 4000, n_tok).astype(np.int32))
    for mode in (0, 2):
        want = ops.rope(x.float(), pos, D, mode).half()
        buf = x.clone()
        ops.rope(buf, pos, D, mode, out=buf)
        torch.cuda.synchronize()
        assert torch.equal(buf, want)
    small = x[:3, :2].contiguous()                       # the generic (non table) kernel
    assert torch.equal(ops.rope(small, pos[:3].contiguous(), 64, 2), ops.rope(small.float(), pos[:3].contiguous(), 64, 2).half())

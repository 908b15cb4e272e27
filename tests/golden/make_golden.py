"""Mint the golden fixtures in tests/golden/ from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile).

Run here (container with /root/reference):   python tests/golden/make_golden.py
The reference has no known-answer files for this path (test-backend-ops seeds from std::random_device),
so we mint fixed-seed vectors from the reference's own code: ggml_quantize_chunk for the weight bytes,
quantize_row_q8_0/q8_K for the activation blocks, and single-op graphs on the reference CPU backend for the outputs.

Determinism knobs (recorded in every file's `meta`): oracle/_ref is the AVX2 build (-mavx2 -mfma -mf16c -mbmi2),
GGML_LLAMAFILE on, no repack (plain ctx tensors), n_threads = 1, seed below.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import refggml as R  # noqa: E402

SEED = 20261017
META = "ref=llama.cpp-omni ggml0.9.4; build=oracle/Makefile AVX2+FMA+F16C, LLAMAFILE=1, no-repack, threads=1; seed=%d" % SEED


def weights(rng, m, k):
    # same flavour as the SURVEY §8d synthetic model: N(0, 0.02^2) weights with a few outliers
    w = rng.standard_normal((m, k)).astype(np.float32) * 0.02
    w[rng.integers(0, m, 8), rng.integers(0, k, 8)] *= 8.0
    return w


def main() -> None:
    assert R.available(), "build oracle/_ref first: make -C oracle ref"
    rng = np.random.default_rng(SEED)

    # ---- quantised MUL_MAT: every north-star weight type, n = 1 (decode) and n = 3 ---------------------------
    mm = {"meta": np.array(META)}
    m, k = 48, 1024
    w = weights(rng, m, k)
    x = rng.standard_normal((3, k)).astype(np.float32)
    x[1] *= 0.05
    x[2, :256] = 0.0                                   # an all-zero q8_K super-block (d = 0 edge case)
    mm["x"] = x
    for t in (R.Q4_0, R.Q8_0, R.Q4_K, R.Q5_K, R.Q6_K, R.F16):
        name = R.TYPE_NAMES[t]
        wq = w.astype(np.float16).view(np.uint8) if t == R.F16 else R.quantize(t, w)
        mm[f"w_{name}"] = wq
        mm[f"y_{name}"] = R.mul_mat(t, wq, x, m, k)
        if t != R.F16:
            mm[f"deq_{name}"] = R.dequantize(t, wq[:2], k)
    mm["act_q8_0_ref"] = R.quantize_act(R.Q8_0, x[0], simd=False)
    mm["act_q8_0_simd"] = R.quantize_act(R.Q8_0, x[0], simd=True)
    mm["act_q8_K"] = np.stack([R.quantize_act(R.Q8_K, x[i]) for i in range(3)])
    np.savez_compressed(HERE / "mul_mat.npz", **mm)

    # ---- elementwise / norm / rope / kv-write ------------------------------------------------------------------
    ops = {"meta": np.array(META)}
    h = rng.standard_normal((3, 512)).astype(np.float32) * 3.0
    ops["rms_x"], ops["rms_y"] = h, R.rms_norm(h, 1e-6)
    qk = rng.standard_normal((4, 6, 128)).astype(np.float32)
    pos = np.array([0, 1, 777, 4095], np.int32)
    ops["rope_x"], ops["rope_pos"] = qk, pos
    ops["rope_neox"] = R.rope(qk, pos, 128, 2, 40960, 1e6)       # Qwen3: neox, theta 1e6
    ops["rope_norm"] = R.rope(qk, pos, 128, 0, 4096, 1e4)        # TTS llama: norm mode
    ops["rope_neox_partial"] = R.rope(qk, pos, 64, 2, 40960, 1e6)
    g, u = rng.standard_normal(1536).astype(np.float32) * 4, rng.standard_normal(1536).astype(np.float32)
    ops["glu_gate"], ops["glu_up"], ops["glu_y"] = g, u, R.swiglu(g, u)
    src = rng.standard_normal((3, 256)).astype(np.float32)
    idx = np.array([5, 0, 9], np.int64)
    ops["sr_src"], ops["sr_idx"], ops["sr_dst"] = src, idx, R.set_rows_f16(src, idx, 12)
    np.savez_compressed(HERE / "ops.npz", **ops)

    # ---- FLASH_ATTN_EXT: decode shape (1 query, GQA 4:1, D = 128) and a small prefill with a causal mask ----------
    fa = {"meta": np.array(META)}
    D, n_head, n_head_kv = 128, 8, 2
    for tag, n_q, n_kv in (("dec", 1, 320), ("pre", 5, 64)):
        q = rng.standard_normal((n_q, n_head, D)).astype(np.float32)
        kk = (rng.standard_normal((n_head_kv, n_kv, D)) * 0.5).astype(np.float16)
        vv = rng.standard_normal((n_head_kv, n_kv, D)).astype(np.float16)
        mask = np.zeros((64, n_kv), np.float16)  # llama pads mask rows to 64 (GGML_KQ_MASK_PAD)
        if n_q == 1:
            mask[0, 300:] = -np.inf                                # padded KV cells beyond the current position
        else:
            for i in range(n_q):
                mask[i, n_kv - n_q + i + 1:] = -np.inf
        out = R.flash_attn(q, kk, vv, mask, 1.0 / np.sqrt(D))
        fa.update({f"{tag}_q": q, f"{tag}_k": kk, f"{tag}_v": vv, f"{tag}_mask": mask, f"{tag}_out": out})
    np.savez_compressed(HERE / "flash_attn.npz", **fa)
    for f in ("mul_mat.npz", "ops.npz", "flash_attn.npz"):
        print(f, (HERE / f).stat().st_size, "bytes")


if __name__ == "__main__":
    main()

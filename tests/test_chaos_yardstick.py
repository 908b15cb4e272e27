"""Why "logits within 1e-3 / identical greedy tokens" cannot be asked of ANY two implementations on a quantised-activation model unless they are
bit-identical — shown with the oracle itself, on weights quantised by the reference's own llama-quantize.

The decode path quantises the activations before every MUL_MAT (q8_K per 256 for K-quants: ggml-quants.c:2555-2592; F16 rounding for F16 weights).
Rounding is discontinuous: a last-bit difference upstream flips a few int8 steps, each flip is a ~1e-4 relative kick to every output of that matmul, the
kicks flip more steps in the next quantiser, and after one transformer layer the difference sits at the quantisation-noise floor (~1e-2 relative for int8
activations, ~1e-3 for F16 ones), independent of how small the first difference was.  The reference shows the same against ITSELF: CPU plain vs CPU repacked
weights 4.6e-2 on the Q4_K_M model, CPU batched-prompt vs token-by-token prompt 1.2e-3 on the F16 model (profiles/r02_llama_parity_*.txt).

  * CPU part: the oracle chain run twice, the second time with its input vector perturbed by 2e-6 relative — the size of the difference a re-ordered f32
    summation leaves in a 4096-term dot product (~sqrt(K) * 2^-24) -> the logits differ by far more than 1e-3.
  * GPU part: the decode engine vs the oracle on the same weights must be no further from the oracle than the oracle is from its perturbed self
    (x2), i.e. the engine's difference IS this noise, not an arithmetic error.  (The tight-tolerance engine-vs-oracle test on well-conditioned weights is
    tests/test_gpu_parity.py::test_decode_engine_whole_token_matches_oracle.)"""
import os
import subprocess
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest

import oracle_c as O
import oracle_decode as OD
from gguf_io import read_gguf

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
pytestmark = pytest.mark.skipif(not (REF / "bin" / "llama-quantize").exists(), reason="oracle/_ref is not built")

# n_ff = 8192: ffn_down is K-split x2 by the engine and every slice of the planar q6_K d plane must stay a 16-byte multiple (16 blocks x f16) for the bulk copies
CFG = SimpleNamespace(n_embd=2048, n_ff=8192, head_dim=128, n_head=16, n_head_kv=4, n_layer=3, n_vocab=4096, n_ctx=256, rms_eps=1e-6, rope_base=1e6,
                      n_ctx_orig=40960)


@pytest.fixture(scope="module")
def quantised_model(tmp_path_factory):
    d = tmp_path_factory.mktemp("chaos")
    f32, q4 = d / "f32.gguf", d / "q4.gguf"
    c = CFG
    subprocess.check_call([sys.executable, str(ROOT / "tools" / "make_gguf.py"), str(f32), "--layers", str(c.n_layer), "--vocab", str(c.n_vocab), "--embd", str(c.n_embd),
                           "--ff", str(c.n_ff), "--heads", str(c.n_head), "--kv-heads", str(c.n_head_kv), "--ftype", "f32"], stderr=subprocess.DEVNULL, timeout=600)
    env = dict(os.environ, LD_LIBRARY_PATH=f"{REF / 'lib'}:" + os.environ.get("LD_LIBRARY_PATH", ""))
    env.pop("GGML_BACKEND_PATH", None)
    subprocess.check_call([str(REF / "bin" / "llama-quantize"), str(f32), str(q4), "q4_k_m", str(os.cpu_count() or 4)], env=env, stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL, timeout=900)
    f32.unlink()
    _, T = read_gguf(str(q4))
    layers = []
    for il in range(c.n_layer):
        p = f"blk.{il}."
        g = lambda n: T[p + n]
        names = {"wq": "attn_q.weight", "wk": "attn_k.weight", "wv": "attn_v.weight", "wo": "attn_output.weight", "gate": "ffn_gate.weight", "up": "ffn_up.weight",
                 "down": "ffn_down.weight"}
        L = {"types": {k: g(v)[0] for k, v in names.items()}}
        for k, v in names.items():
            L[k] = g(v)[2]
        for k, v in (("attn_norm", "attn_norm.weight"), ("ffn_norm", "ffn_norm.weight"), ("q_norm", "attn_q_norm.weight"), ("k_norm", "attn_k_norm.weight")):
            L[k] = g(v)[2].view(np.float32).copy()
        layers.append(L)
    head = {"out_norm": T["output_norm.weight"][2].view(np.float32).copy(), "lm_head": T["output.weight"][2], "type": T["output.weight"][0], "n_vocab": c.n_vocab}
    return layers, head


def _fresh_caches(layers, rng):
    kvw = CFG.n_head_kv * CFG.head_dim
    base = [((rng.standard_normal((CFG.n_ctx, kvw)) * 0.5).astype(np.float16), rng.standard_normal((CFG.n_ctx, kvw)).astype(np.float16)) for _ in layers]

    def make():
        out = []
        for L, (k, v) in zip(layers, base):
            M = dict(L)
            M["k_cache"], M["v_cache"] = k.copy(), v.copy()
            out.append(M)
        return out
    return make


def _rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def _perturb(x, rng, eps=2e-6):
    """what a different (equally valid) f32 summation order does to the vector that feeds a quantiser: ~sqrt(K) ulps on every element"""
    return (x.astype(np.float64) * (1.0 + eps * rng.standard_normal(x.size))).astype(np.float32)


def test_summation_order_noise_reaches_the_quantisation_noise_floor(quantised_model):
    layers, head = quantised_model
    rng = np.random.default_rng(5)
    make = _fresh_caches(layers, rng)
    x = (rng.standard_normal(CFG.n_embd) * 0.05).astype(np.float32)
    pos, n_kv = 100, 128
    a, _, _ = OD.oracle_token(CFG, make(), x, pos, n_kv, head)
    b, _, _ = OD.oracle_token(CFG, make(), x, pos, n_kv, head)
    assert np.array_equal(a, b)                                           # the oracle itself is deterministic
    errs = []
    for trial in range(4):
        c, _, _ = OD.oracle_token(CFG, make(), _perturb(x, rng), pos, n_kv, head)
        errs.append(_rel(c, a))
    # a 2e-6 relative nudge of the input moves the logits by >> 1e-3 (1000x amplification in 3 layers): the map is discontinuous, not ill-conditioned code
    assert max(errs) > 1e-3, errs
    print("2e-6 perturbation -> max relative logit difference per trial:", ["%.2e" % e for e in errs])


@pytest.mark.gpu
def test_engine_is_as_close_to_the_oracle_as_the_oracle_is_to_itself(quantised_model):
    import torch
    from __graft_entry__ import load_package
    pkg = load_package()
    ops, dec = pkg.ops, pkg.decode
    layers, head = quantised_model
    rng = np.random.default_rng(5)
    make = _fresh_caches(layers, rng)
    cfg = dec.LLMConfig(name="chaos", n_embd=CFG.n_embd, n_layer=CFG.n_layer, n_head=CFG.n_head, n_head_kv=CFG.n_head_kv, n_ff=CFG.n_ff, n_vocab=CFG.n_vocab, n_ctx=CFG.n_ctx)
    D = dec.Qwen3Decoder(cfg, "cuda:0", seed=0)
    ol = make()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for lw, L in zip(D.L, ol):
        lw["types"] = dict(L["types"])
        for n in ("wq", "wk", "wv", "wo", "gate", "up", "down"):
            w = dev(L[n].reshape(-1))
            lw[n] = ops.to_planar(L["types"][n], w) if L["types"][n] == ops.Q6_K else w
        for n in ("attn_norm", "ffn_norm", "q_norm", "k_norm"):
            lw[n] = dev(L[n])
        lw["k_cache"].copy_(dev(L["k_cache"])); lw["v_cache"].copy_(dev(L["v_cache"]))
    D.out_norm = dev(head["out_norm"])
    assert head["type"] == ops.Q6_K
    D.lm_head = ops.to_planar(ops.Q6_K, dev(head["lm_head"].reshape(-1)))
    D.build_engine()
    x = (rng.standard_normal(CFG.n_embd) * 0.05).astype(np.float32)
    pos, n_kv = 100, 256
    hi = dec.Qwen3Decoder.host_inputs(cfg, pos, n_kv, pinned=False)
    D.x_in.copy_(dev(x)); D.pos.copy_(hi["pos"]); D.kv_idx.copy_(hi["kv_idx"]); D.mask_f32[:, :n_kv].copy_(hi["mask"])
    D.step_engine(n_kv)
    torch.cuda.synchronize()
    got = D.logits.cpu().numpy()
    ref, _, _ = OD.oracle_token(CFG, make(), x, pos, n_kv, head)
    yard = []
    for trial in range(4):
        c, _, _ = OD.oracle_token(CFG, make(), _perturb(x, rng), pos, n_kv, head)
        yard.append(_rel(c, ref))
    e = _rel(got, ref)
    print("engine vs oracle %.2e; oracle vs 2e-6-perturbed oracle %s" % (e, ["%.2e" % v for v in yard]))
    assert np.isfinite(got).all()
    assert e <= 2.0 * max(max(yard), 1e-3), (e, yard)

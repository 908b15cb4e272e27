#!/usr/bin/env python
"""One launch of every kernel class at the Qwen3-8B shapes inside a cudaProfilerStart/Stop range (for `ncu --profile-from-start off`): the per-op decode sequence of
one layer (k_rms_norm_quantize, k_mmvq / k_stream single phase, k_qkv_post, k_fa_decode, k_fa_combine, k_quantize_q8K), a 2048-token prefill layer (k_mmq_tc, k_fa_tc,
k_rms_norm + tiles, k_glu_tiles, k_binary), the F16 GEMM (k_mm_f16_tc), the legacy mma.sync attention (k_fa_prefill, D = 64) and the encoder ops."""
import sys, torch
sys.path.insert(0, '.')
from __graft_entry__ import load_package
pkg = load_package(); ops, dec = pkg.ops, pkg.decode
cfg = dec.LLMConfig(n_layer=1)
D = dec.Qwen3Decoder(cfg, "cuda:0")
depth, n_kv = 2048, 2304
for lw in D.L:
    lw["k_cache"][:depth].normal_(0, 0.5); lw["v_cache"][:depth].normal_(0, 1.0)
hi = dec.Qwen3Decoder.host_inputs(cfg, depth, n_kv, pinned=False)
D.x_in.normal_(0, 0.05); D.pos.copy_(hi["pos"]); D.kv_idx.copy_(hi["kv_idx"]); D.mask_f32[:, :n_kv].copy_(hi["mask"])
x = torch.randn(2048, cfg.n_embd, device="cuda") * 0.05
wf = (torch.randn(4096, 4096, device="cuda") * 0.02).half(); xf = torch.randn(2048, 4096, device="cuda")
q64 = torch.randn(512, 8, 64, device="cuda"); k64 = torch.randn(512, 8, 64, device="cuda").half(); m64 = torch.zeros(512, 512, device="cuda").half()
enc = torch.randn(1500, 1280, device="cuda"); img = torch.randn(1, 3, 448, 448, device="cuda"); kern = torch.zeros(8, 3, 14, 14, device="cuda").half()
# round-2 additions: the streaming F16 matvec (plain / residual epilogue / gate-up-SWIGLU), quantised-KV rows (SET_ROWS -> q8_0, GET_ROWS from q4_K, FLASH_ATTN_EXT
# decode over a q8_0 cache) and the Token2Wav op set
xv = torch.randn(1, 4096, device="cuda"); rv = torch.randn(1, 4096, device="cuda"); wf2 = (torch.randn(4096, 4096, device="cuda") * 0.02).half()
kvsrc = torch.randn(512, 1024, device="cuda"); kvidx = torch.arange(512, device="cuda", dtype=torch.int64)
kvq = torch.zeros(2304 * 1024 // 32 * 34, dtype=torch.uint8, device="cuda")
emb = torch.randint(0, 255, (32768 * 4096 // 256 * 144,), dtype=torch.uint8, device="cuda"); emb.view(-1, 144)[:, 0:4] = torch.tensor([0, 20, 0, 20], dtype=torch.uint8, device="cuda")
eidx = torch.randint(0, 32768, (512,), device="cuda", dtype=torch.int32)
qd = torch.randn(32, 1, 128, device="cuda"); md = torch.zeros(64, 2304, device="cuda").half()
wav = torch.randn(256, 1500, device="cuda"); ctw = (torch.randn(512, 256, 16, device="cuda") * 0.02).half(); ctx_ = torch.randn(512, 200, device="cuda")
def kvd(buf):
    return ops.T(buf, ops.Q8_0, ne=[128, 2304, 8], nb=[34, 8 * 136, 136, 2304 * 8 * 136])
def run():
    D.step(n_kv)
    D.prefill(x, 0, 2048)
    ops.mul_mat(wf2, ops.F16, 4096, 4096, xv, w_ne=[4096, 4096])
    ops.mul_mat_add(wf2, ops.F16, 4096, 4096, xv, rv, torch.empty_like(rv))
    ops.mul_mat_glu(ops.GLU_SWIGLU, wf2, wf, ops.F16, 4096, 4096, xv)
    ops.set_rows(kvsrc, kvidx, ops.T(kvq, ops.Q8_0, ne=[1024, 2304]))
    ops.get_rows(ops.T(emb, ops.Q4_K, ne=[4096, 32768]), eidx)
    ops.flash_attn(qd, kvd(kvq), kvd(kvq), md, 0.088)
    ops.concat(wav, wav, 1); ops.repeat(wav, [2, 256, 1500]); ops.sum_rows(wav); ops.pad_reflect_1d(wav, 3, 3); ops.unary_param(ops.LEAKY_RELU, wav, 0.1)
    ops.unary_param(ops.SIN, wav); ops.conv_transpose_1d(ctw, ctx_, 8)
    ops.mul_mat(wf.view(torch.uint8).reshape(-1), ops.F16, 4096, 4096, xf)
    ops.flash_attn(q64.permute(1, 0, 2), k64.permute(1, 0, 2), k64.permute(1, 0, 2), m64, 0.125)
    ops.norm(enc, 1e-5); ops.im2col(kern, img, 14, 14, 0, 0, 1, 1, True); ops.pool_1d(enc.t().contiguous(), 1, 5)
run(); torch.cuda.synchronize()
torch.cuda.profiler.start(); run(); torch.cuda.synchronize(); torch.cuda.profiler.stop()

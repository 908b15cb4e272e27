"""GPU: the F16 tcgen05 GEMM (k_mm_f16_tc, csrc/mmq_tc.cu) with K that is NOT a multiple of its 64-wide K tile — SigLip's ffn_down has k = 4304 = 67.25 tiles
(tools/omni/vision.cpp build_ffn).  The weight tile's tail is zero-filled by TMA (the tensor map carries the real k), the activation tiles' tail by k_x_to_f16_tiles.
Arithmetic = the CPU oracle's for F16 weights: activations rounded to F16 (vec_dot_type), F32 accumulation.  The 1152 x 1024 x 4304 shape is what the VPM frame of
tests/test_zz_omni_encoders.py runs 27 times (profiles/r02_omni_encoders.md).  (The file name sorts last on purpose, next to that test.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from __graft_entry__ import load_package  # noqa: E402


@pytest.mark.parametrize("m,k,n", [(1152, 4304, 1024),       # SigLip ffn_down, split-K = 2 over 68 tiles
                                   (300, 4304, 77),          # ragged m, n and k at once
                                   (256, 576, 33),           # 9 K tiles: an odd tile count must not split K (k / 128 used to round it down to "4 units")
                                   (129, 72, 50)])           # one full tile + 8 columns of the second
def test_f16_tensor_core_gemm_with_a_k_tail(m, k, n):
    assert torch.cuda.is_available()
    ops = load_package().ops
    rng = np.random.default_rng(m + k + n)
    w = (rng.standard_normal((m, k)) * 0.05).astype(np.float16)
    x = rng.standard_normal((n, k)).astype(np.float32)
    wt = torch.from_numpy(w).cuda()
    got = ops.mul_mat(wt, ops.F16, m, k, torch.from_numpy(x).cuda(), w_ne=[k, m]).cpu().numpy()
    ref = (torch.from_numpy(x).cuda().half().double() @ wt.double().T).cpu().numpy()
    mag = np.abs(x) @ np.abs(w.astype(np.float32)).T
    assert np.all(np.isfinite(got))
    assert np.all(np.abs(got - ref) <= 2e-6 * mag + 1e-9), float((np.abs(got - ref) / (mag + 1e-30)).max())

#!/usr/bin/env python
"""n = 1 MUL_MAT at the exact Qwen3-8B decode shapes (native and planar layouts) through the C-ABI vs the CPU oracle."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle_c as O
from __graft_entry__ import load_package
from test_gpu_parity import rand_blocks
ops = load_package().ops
rng = np.random.default_rng(0)
QT = {"q4_K": (O.Q4_K, ops.Q4_K), "q6_K": (O.Q6_K, ops.Q6_K)}
for name, m, k in (("q4_K", 4096, 4096), ("q4_K", 1024, 4096), ("q6_K", 1024, 4096), ("q4_K", 512, 12288), ("q6_K", 512, 12288), ("q6_K", 2048, 4096), ("q4_K", 12288, 4096)):
    ot, bt = QT[name]
    w = rand_blocks(rng, ot, m * k // 256)
    for n in (1, 2):
        x = rng.standard_normal((n, k)).astype(np.float32)
        ref = O.mul_mat(ot, w, x, m, k)
        wd = torch.from_numpy(w).cuda(); xd = torch.from_numpy(x).cuda()
        for layout in ((ops.LAYOUT_NATIVE, ops.LAYOUT_PLANAR) if name == "q6_K" else (ops.LAYOUT_NATIVE,)):
            ww = ops.to_planar(bt, wd) if layout == ops.LAYOUT_PLANAR else wd
            y = ops.mul_mat(ww, bt, m, k, xd, layout=layout).cpu().numpy()
            err = np.abs(y - ref).max() / np.abs(ref).max()
            print(f"{name} m={m} k={k} n={n} layout={'planar' if layout else 'native'}: max rel err {err:.3g}", "OK" if err < 1e-5 else "<<<<<< MISMATCH")

#!/usr/bin/env python
"""GPU diagnostic: cycles per tcgen05.mma (M=128, K=16) for operand modes SS / SS-Bmn / TS and several N."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402

lib = mcn._capi.load()
rs = np.random.RandomState(0)
modes = [(0, "SS")]
for d_col in (256, 80, 56, 160, 320):
    modes.append((600 + d_col, "TS-inplace d%d" % d_col))
for mode, name in modes:
    for N, K in ((160, 160), (112, 112)):
        if mode == 1 and K > 128:
            continue
        a = rs.uniform(-1, 1, (128, K)).astype(np.float16).astype(np.float32)
        b = rs.uniform(-1, 1, (N, K)).astype(np.float16).astype(np.float32)
        d = np.zeros((128, N), np.float32)
        cyc = C.c_longlong()
        reps = int(os.environ.get("REPS", "50"))
        mcn._capi.check(lib.cn_debug_umma_bench(N, K, mode, reps, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                                d.ctypes.data_as(C.c_void_p), C.byref(cyc), 0))
        ref = a.astype(np.float64) @ b.astype(np.float64).T
        if 600 <= mode < 1000:       # second product: fp16(A B^T) B^T
            ref = ref.astype(np.float16).astype(np.float64) @ b.astype(np.float64).T
        err = np.max(np.abs(d - ref)) / max(1.0, np.max(np.abs(ref)))
        nm = K // 16
        print("%-14s N=%3d K=%3d  err %.2e  %7.1f cycles per layer (%d MMAs) -> %5.1f / MMA   ideal %.0f" % (
            name, N, K, err, cyc.value / reps, nm, cyc.value / reps / nm, N / 2))

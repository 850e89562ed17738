#!/usr/bin/env python
"""GPU diagnostic: the cta_group::2 building block (numerics per quadrant + cycles per layer)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402

lib = mcn._capi.load()
rs = np.random.RandomState(0)
for N, K in ((32, 16), (112, 112), (160, 32), (224, 112), (112, 160), (64, 112)):
    a = rs.uniform(-1, 1, (256, K)).astype(np.float16).astype(np.float32)
    b = rs.uniform(-1, 1, (N, K)).astype(np.float16).astype(np.float32)
    d = np.zeros((256, N), np.float32)
    cyc = C.c_longlong()
    reps = 50
    mcn._capi.check(lib.cn_selftest_umma_pair(N, K, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                              d.ctypes.data_as(C.c_void_p), reps, C.byref(cyc), 0))
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    err = np.abs(d - ref)
    q = [err[:128, :N // 2].max(), err[:128, N // 2:].max(), err[128:, :N // 2].max(), err[128:, N // 2:].max()]
    print("pair N=%3d K=%3d  quadrant max err %s  %7.1f cycles per layer (%d MMAs, ideal %d)" % (
        N, K, " ".join("%.1e" % x for x in q), cyc.value / reps, K // 16, (K // 16) * N // 2), flush=True)

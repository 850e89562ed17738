#!/usr/bin/env python
"""GPU diagnostic: tcgen05.ld bytes per cycle per SM for several warp counts and load shapes."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402

lib = mcn._capi.load()
iters = 2000
for mode, name, cols in ((0, "x32 pair", 64), (1, "x64", 64), (2, "x16", 16), (3, "x32 pair + cvt/STS", 64)):
    for nw in (1, 4, 8, 16):
        cyc = C.c_longlong()
        mcn._capi.check(lib.cn_debug_tmem_bench(mode, nw, iters, C.byref(cyc), 0))
        byts = nw * 32 * cols * 4 * iters
        print("%-20s warps %2d  %8.1f cycles/iter  %6.1f B/cycle/SM" % (name, nw, cyc.value / iters, byts / cyc.value), flush=True)

#!/usr/bin/env python
"""GPU diagnostic: cycles per phase of one tile of tc_rows_kernel (CTA 0, thread 0), via cn_debug_tc_timing."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402

NAMES = ["store X + sync", "wait mlp1.0", "epi H1 + sync", "wait mlp1.2 (+next loads)", "epi M1 + sync",
         "group mean + sync", "wait mlp2.0/attn.0 (+next features)", "epi H3/Ha1 + sync", "wait mlp2.2/attn.2",
         "P' + score + sync", "softmax + F' + sync", "wait D2", "J store + sync"]


def main(E=8192, H=5):
    w = np.load(os.path.join(ROOT, "tests", "golden", "sarl_weights_seed0.npy"))
    env = mcn.BatchedCrowdSim(E, H, auto_reset=1)
    pol = mcn.BatchedSARL(precision="f16_tc"); pol.load_weights(w)
    env.reset_device()
    out = (C.c_longlong * 256)()
    mcn._capi.check(pol.lib.cn_debug_tc_timing(pol.handle, out))      # arm
    for _ in range(3):
        pol.lookahead(env)
    mcn._capi.check(pol.lib.cn_debug_tc_timing(pol.handle, out))
    allv = np.array(list(out), dtype=np.int64)
    order = [0, 1, 2, 3, 4, 5, 14, 15, 16, 17, 6, 18, 7, 19, 8, 9, 10, 11, 12, 13]
    for name, base in (("tid0", 0), ("tid64", 64), ("tid255", 32)):
        tt = allv[base:base + 32]
        print(name, " ".join("%d:%d" % (i, tt[i] - allv[0]) for i in order))
    t = allv[:14]
    d = np.diff(t)
    print("tile total %d cycles" % (t[13] - t[0]))
    for n, c in zip(NAMES, d):
        print("  %-40s %6d  %5.1f%%" % (n, c, 100.0 * c / (t[13] - t[0])))


if __name__ == "__main__":
    main(H=int(sys.argv[1]) if len(sys.argv) > 1 else 5)

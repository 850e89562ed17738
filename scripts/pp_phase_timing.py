#!/usr/bin/env python
"""GPU diagnostic for the CTA-pair row kernel: timestamps (cycles, relative to CTA 0 / context 0's stage-0 request
of round 3) of both epilogue groups of both CTAs of cluster 0 and of the issuer warp."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 5
if len(sys.argv) > 2:                      # an A/B or ablation build instead of the in-tree library
    mcn._capi.LIB_PATH = os.path.abspath(sys.argv[2])
w = np.load(os.path.join(ROOT, "tests", "golden", "sarl_weights_seed0.npy"))
env = mcn.BatchedCrowdSim(8192, H, auto_reset=1)
pol = mcn.BatchedSARL(precision="f16_tc"); pol.load_weights(w)
env.reset_device()
out = (C.c_longlong * 256)()
mcn._capi.check(pol.lib.cn_debug_tc_timing(pol.handle, out))
for _ in range(3):
    pol.lookahead(env)
mcn._capi.check(pol.lib.cn_debug_tc_timing(pol.handle, out))
v = np.array(list(out), dtype=np.int64).reshape(8, 32)
t0 = v[0][0]
EPI = ["tile start", "done S0", "H1 stored", "done S1", "M1 stored", "mean stored", "next prepared", "done S3",
       "H3/Ha1 stored", "done S4", "wF stored", "J stored"]
for name, r in (("CTA0 ctx0", v[0]), ("CTA0 ctx1", v[1]), ("CTA1 ctx0", v[4]), ("CTA1 ctx1", v[5])):
    print(name, " ".join("%s@%d" % (EPI[i], r[i] - t0) for i in range(12)))
    print("   deltas", " ".join("%d" % (r[i + 1] - r[i]) for i in range(11)), " total", r[11] - r[0])
for name, r in (("CTA0 ctx0", v[0]), ("CTA0 ctx1", v[1])):
    print(name, "M1 stored@%d loads issued@%d barrier passed@%d mean done@%d signalled@%d | done S4@%d score@%d barrier@%d wF@%d barrier@%d" % (
        r[4] - t0, r[12] - t0, r[13] - t0, r[14] - t0, r[5] - t0, r[9] - t0, r[15] - t0, r[16] - t0, r[17] - t0, r[10] - t0))
for name, r in (("CTA0 ctx0", v[0]), ("CTA0 ctx1", v[1])):
    if r[18]:
        print(name, "split mean: barrier passed@%d partials@%d barrier@%d finished@%d barrier@%d replicated@%d" % (
            r[13] - t0, r[18] - t0, r[19] - t0, r[20] - t0, r[21] - t0, r[14] - t0))
print("ctx0 reqm arrival per warp (CTA0 hf0 q0-3, hf1 q0-3 | CTA1 ...):", " ".join(str(int(x - t0)) for x in v[6][:16]))
for c in (0, 1):
    r = v[2 + c]
    print("issuer ctx%d" % c, " ".join("s%d:ready@%d,issued@%d" % (s, r[2 * s] - t0, r[2 * s + 1] - t0) for s in range(5)))

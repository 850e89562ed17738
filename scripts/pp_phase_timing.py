#!/usr/bin/env python
"""GPU diagnostic for the ping-pong row kernel (CN_TC_VARIANT=pp): per-stage timestamps of epilogue group 0/1
(row 0) and of the issuer warp for one tile round of CTA 0."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["CN_TC_VARIANT"] = "pp"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402

w = np.load(os.path.join(ROOT, "tests", "golden", "sarl_weights_seed0.npy"))
env = mcn.BatchedCrowdSim(8192, 5, auto_reset=1)
pol = mcn.BatchedSARL(precision="f16_tc"); pol.load_weights(w)
env.reset_device()
out = (C.c_longlong * 96)()
mcn._capi.check(pol.lib.cn_debug_tc_timing(pol.handle, out))
for _ in range(3):
    pol.lookahead(env)
mcn._capi.check(pol.lib.cn_debug_tc_timing(pol.handle, out))
v = np.array(list(out), dtype=np.int64).reshape(3, 32)
t0 = v[0][31]
print("EG0: X stored/req0 at 0")
for name, r in (("EG0", v[0]), ("EG1", v[1])):
    print(name, " ".join("s%d:done@%d,req@%d" % (s, r[2 * s] - t0, r[2 * s + 1] - t0) for s in range(8)))
print("ISS ", " ".join("s%d:got@%d,commit@%d" % (s, v[2][2 * s] - t0, v[2][2 * s + 1] - t0) for s in range(8)))
print("stage-1 MMA issue times:", " ".join(str(int(x - t0)) for x in v[2][16:26]))

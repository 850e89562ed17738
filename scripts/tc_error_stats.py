#!/usr/bin/env python
"""GPU diagnostic: fp16 tensor-core lookahead vs the FP32 CUDA-core lookahead on the same evolving states.
Prints max |dV|, and argmax agreement as a function of the excluded top-2 gap. Needs a B200."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402


def main(E=2048, H=5, steps=40, weights=None, query_env=0):
    w = np.load(weights or os.path.join(ROOT, "tests", "golden", "sarl_weights_seed0.npy"))
    env = mcn.BatchedCrowdSim(E, H, auto_reset=1, seed=11)
    p32 = mcn.BatchedSARL(precision="f32"); p32.load_weights(w)
    p16 = mcn.BatchedSARL(precision="f16_tc"); p16.load_weights(w)
    env.reset_device()
    gaps = [0, 1e-5, 5e-5, 1e-4, 2e-4, 5e-4, 1e-3]
    tot = np.zeros(len(gaps)); agr = np.zeros(len(gaps))
    maxerr = 0.0; maxrel = 0.0; vmax = 0.0
    for s in range(steps):
        env.orca()
        p32.lookahead(env, query_env); b32, v32 = p32.read(env)
        p16.lookahead(env, query_env); b16, v16 = p16.read(env)
        err = np.abs(v32 - v16)
        maxerr = max(maxerr, err.max()); vmax = max(vmax, np.abs(v32).max())
        maxrel = max(maxrel, (err.max(1) / np.maximum(1.0, np.abs(v32).max(1))).max())
        srt = np.sort(v32, axis=1)
        gap = srt[:, -1] - srt[:, -2]
        for i, g in enumerate(gaps):
            m = gap > g
            tot[i] += m.sum(); agr[i] += (b32[m] == b16[m]).sum()
        env.step(update=True, read=False)        # continue with the fp16 path's actions
    print("E=%d H=%d steps=%d  max|dV|=%.3g  max rel=%.3g  max|V|=%.3g" % (E, H, steps, maxerr, maxrel, vmax))
    for g, t, a in zip(gaps, tot, agr):
        print("  gap > %-7g states %7d  argmax agreement %.5f" % (g, t, a / max(t, 1)))


if __name__ == "__main__":
    main(weights=sys.argv[1] if len(sys.argv) > 1 else None)

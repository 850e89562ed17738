#!/usr/bin/env python
"""Accuracy of the fp16 tensor-core lookahead against the REFERENCE's values on the decisive sets (tests/golden/decisive_*.npz,
trained weights): percentiles of |dv| and of |dv| / max(|v|, 0.1), worst state, argmax agreement on decidable states.
Usage: python scripts/tc_error_stats.py [lib.so]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
if len(sys.argv) > 1:
    from modelcrowdnav_b200 import _capi
    _capi.LIB_PATH = os.path.abspath(sys.argv[1])
import modelcrowdnav_b200 as mcn  # noqa: E402
from conftest import DECISIVE_NAMES, load_decisive  # noqa: E402

w = np.load(os.path.join(ROOT, "tests", "golden", "sarl_weights_trained.npy"))
for prec in ("f32", "f16_tc"):
    for name in DECISIVE_NAMES:
        d = load_decisive(name)
        N, H = d["agents"].shape[0], int(d["H"])
        env = mcn.BatchedCrowdSim(N, H)
        pol = mcn.BatchedSARL(precision=prec)
        pol.load_weights(w)
        env.set_state(d["agents"], d["time"])
        env.orca()
        pol.lookahead(env, query_env=int(d["query_env"]))
        best, values = pol.read(env)
        ref = d["values"].astype(np.float64)
        dv = np.abs(values - ref)
        rel = dv / np.maximum(np.abs(ref), 0.1)
        relstate = dv.max(1) / np.maximum(np.abs(ref).max(1), 0.1)
        top2 = np.sort(ref, axis=1)[:, -2:]
        gap = top2[:, 1] - top2[:, 0]
        line = "%-6s %-16s N=%d |dv| rms %.2e p99 %.2e p99.9 %.2e max %.2e | rel(floor .1) p99 %.2e p99.9 %.2e p99.99 %.2e max %.2e | per-state inf-norm rel max %.2e" % (
            prec, name, N, np.sqrt((dv ** 2).mean()), np.percentile(dv, 99), np.percentile(dv, 99.9), dv.max(),
            np.percentile(rel, 99), np.percentile(rel, 99.9), np.percentile(rel, 99.99), rel.max(), relstate.max())
        for thr in (2e-4, 5e-4, 1e-3):
            m = gap > thr
            line += " | gap>%.0e: %d/%d" % (thr, int((best[m] == d["best"][m]).sum()), int(m.sum()))
        regret = ref.max(1) - ref[np.arange(N), best]
        line += " | regret max %.2e" % regret.max()
        print(line, flush=True)
        env.close(); pol.close()

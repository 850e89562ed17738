#!/usr/bin/env python
"""A/B timing of whole rollout steps and whole lookaheads (no per-kernel events, so launch-level changes such as programmatic
dependent launch show up): CUDA events around each call, L2 flushed in between, median of N.
Usage: python scripts/ab_step.py lib_a.so lib_b.so ...   (each library in its own subprocess, alternating twice)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(lib):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from modelcrowdnav_b200 import _capi
    _capi.LIB_PATH = os.path.abspath(lib)
    import modelcrowdnav_b200 as mcn
    env = mcn.BatchedCrowdSim(8192, 5, auto_reset=1, seed=0)
    pol = mcn.BatchedSARL(precision="f16_tc")
    pol.load_weights(np.load(os.path.join(ROOT, "tests", "golden", "sarl_weights_seed0.npy")))
    env.reset_device()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(20):
        mcn.rollout_step(pol, env, 0)
    steps, looks = [], []
    for _ in range(40):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); mcn.rollout_step(pol, env, 0); b.record(); torch.cuda.synchronize()
        steps.append(a.elapsed_time(b))
    env.orca()
    for _ in range(40):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pol.lookahead(env, 0); b.record(); torch.cuda.synchronize()
        looks.append(a.elapsed_time(b)); env.step(update=True, read=False); env.orca()
    best, values = pol.read(env)
    print("step ms %.4f  lookahead ms %.4f  checksum %.6f" % (float(np.median(steps)), float(np.median(looks)), float(np.nansum(values))))


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2])
        sys.exit(0)
    for rep in range(2):
        for lib in sys.argv[1:]:
            r = subprocess.run([sys.executable, __file__, "--child", lib], capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED: " + r.stderr[-400:]
            print("%-30s %s" % (os.path.basename(lib), line), flush=True)

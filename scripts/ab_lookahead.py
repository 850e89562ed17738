#!/usr/bin/env python
"""A/B timing of lookahead builds on one box: for every shared library given on the command line, the per-kernel duration of
the tensor-core lookahead (CUDA events recorded by the library on its launching stream) at E envs x H humans, L2 flushed
between lookaheads.  Each library runs in its own subprocess (ctypes cannot unload), alternating twice.
Usage: python scripts/ab_lookahead.py [--envs 8192] [--humans 5] [--iters 30] lib_a.so lib_b.so ..."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(lib, E, H, iters, sim):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    from modelcrowdnav_b200 import _capi
    _capi.LIB_PATH = os.path.abspath(lib)
    import modelcrowdnav_b200 as mcn
    env = mcn.BatchedCrowdSim(E, H, auto_reset=1, seed=0, sim_rule=sim)
    pol = mcn.BatchedSARL(precision="f16_tc")
    pol.load_weights(np.load(os.path.join(ROOT, "tests", "golden", "sarl_weights_seed0.npy")))
    env.reset_device()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        mcn.rollout_step(pol, env, 0)
    pol.kernel_timing(True)
    acc = {}
    for _ in range(iters):
        flush.zero_()
        env.orca()
        pol.lookahead(env, 0)
        for k, v in pol.kernel_ms().items():
            acc.setdefault(k, []).append(v)
        env.step(update=True, read=False)
    out = {k: float(np.median(v)) for k, v in acc.items()}
    out["lookahead"] = sum(out.values())
    try:                                     # ablation builds produce no finite value: the timings are what they are for
        best, values = pol.read(env)
        out["checksum"] = float(np.nansum(values))
    except Exception as ex:
        out["checksum"] = "unreadable: %s" % type(ex).__name__
    print(json.dumps(out))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--humans", type=int, default=5)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--sim", type=int, default=0)
    ap.add_argument("--child", default=None)
    ap.add_argument("libs", nargs="*")
    a = ap.parse_args()
    if a.child:
        child(a.child, a.envs, a.humans, a.iters, a.sim)
        sys.exit(0)
    for rep in range(2):
        for lib in a.libs:
            r = subprocess.run([sys.executable, __file__, "--child", lib, "--envs", str(a.envs), "--humans", str(a.humans),
                                "--iters", str(a.iters), "--sim", str(a.sim)], capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED: " + r.stderr[-400:]
            print("%-40s E=%d H=%d  %s" % (os.path.basename(lib), a.envs, a.humans, line), flush=True)

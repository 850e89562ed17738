#!/usr/bin/env python
"""Profiling driver: a few rollout steps (ORCA + SARL lookahead + env step) of the bench workload, nothing else.
Usage: python scripts/prof_step.py [steps] [envs] [humans] [circle|square]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
E = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
H = int(sys.argv[3]) if len(sys.argv) > 3 else 5
SIM = sys.argv[4] if len(sys.argv) > 4 else "circle"
env = mcn.BatchedCrowdSim(E, H, auto_reset=1, seed=0, sim_rule=0 if SIM == "circle" else 1)
pol = mcn.BatchedSARL(precision="f16_tc")
pol.load_weights(np.load(os.path.join(ROOT, "tests", "golden", "sarl_weights_seed0.npy")))
env.reset_device()
for _ in range(steps):
    mcn.rollout_step(pol, env, 0)
env.read_outputs()
print("ok", pol.lib.cn_launch_count())

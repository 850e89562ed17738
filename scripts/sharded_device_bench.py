#!/usr/bin/env python
"""GPU diagnostic: device-resident rollout of E envs as S env shards on S streams (cn_rollout_step_sharded), env-steps/s
over K steps measured with CUDA events across all streams.  Usage: python scripts/sharded_device_bench.py [E] [K] [S ...]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
shard_counts = [int(x) for x in sys.argv[3:]] or [1, 2, 4, 8]
w = np.load(os.path.join(ROOT, "tests", "golden", "sarl_weights_seed0.npy"))
for S in shard_counts:
    pipe = mcn.PipelinedHostRollout(E, 5, w, shards=S, auto_reset=1, seed=0)
    pipe.reset_device()
    for _ in range(5):
        pipe.step_device()
    pipe.sync_device()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for st in pipe.streams:
        st.wait_event(t0)
    for _ in range(K):
        pipe.step_device()
    cur = torch.cuda.current_stream()
    for st in pipe.streams:
        cur.wait_stream(st)
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / K
    print("shards %d: %.4f ms per step of %d envs -> %.4g env-steps/s" % (S, ms, E, E / ms * 1e3), flush=True)
    pipe.close()

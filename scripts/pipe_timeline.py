#!/usr/bin/env python
"""GPU diagnostic: timeline (named CUDA events) of a few PipelinedHostRollout steps -- which shard's copies / kernels
overlap which.  Usage: python scripts/pipe_timeline.py [shards] [envs]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402

shards = int(sys.argv[1]) if len(sys.argv) > 1 else 2
E = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
w = np.load(os.path.join(ROOT, "tests", "golden", "sarl_weights_seed0.npy"))
pipe = mcn.PipelinedHostRollout(E, 5, w, shards=shards, auto_reset=1, seed=0, use_graphs=False)   # marks need eager calls
pipe.reset_device()
for _ in range(5):
    pipe.step()
pipe.sync()
lib = pipe.lib
lib.cn_debug_trace(1)
for _ in range(4):
    pipe.step()
pipe.sync()
buf = C.create_string_buffer(1 << 16)
mcn._capi.check(lib.cn_debug_trace_dump(buf, len(buf)))
lib.cn_debug_trace(0)
streams = {}
for line in buf.value.decode().splitlines():
    ms, st, name = line.split()
    k = streams.setdefault(st, "s%d" % len(streams))
    print("%9.1f us  %s  %s" % (1e3 * float(ms), k, name))

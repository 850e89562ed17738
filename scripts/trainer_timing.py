#!/usr/bin/env python
"""Per-step time of the fused SARL trainer: wall clock vs CUDA events over N back-to-back cn_trainer_step_indexed calls on a
synthetic replay memory (batch 100 x 5 humans).  Usage: python scripts/trainer_timing.py [steps]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402
from modelcrowdnav_b200.policy import make_value_network  # noqa: E402
from modelcrowdnav_b200.trainer import Trainer, sample_batches  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = make_value_network(13, 6, [150, 100], [100, 50], [150, 100, 100, 1], [100, 100, 1]).to(dev)
memory = mcn.ReplayMemory(100000, device=dev)
memory.push_batch(torch.rand((50000, 5, 13), device=dev), torch.rand((50000,), device=dev))
for mode in ("fused", "graph"):
    tr = Trainer(model, memory, dev, 100, mode=mode)
    tr.set_learning_rate(0.001)
    tr.optimize_batch(20)
    idx = sample_batches(len(memory), 100, n, dev)
    loss = torch.zeros((), device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        if mode == "fused":
            tr._fused.step_indexed(memory.states, memory.values, idx[i], loss)
        else:
            loss += tr._step(idx[i])
    e1.record()
    t_launch = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    print("%s: %d steps  host enqueue %.1f us/step  wall %.1f us/step  device (events) %.1f us/step" % (
        mode, n, 1e6 * t_launch / n, 1e6 * t_wall / n, 1e3 * e0.elapsed_time(e1) / n), flush=True)

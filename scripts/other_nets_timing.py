"""Lookahead time of the other value networks (CADRL, LSTM-RL, OM-SARL) on the tensor-core and FP32 paths.

    python scripts/other_nets_timing.py [E] [H]      (GPU box; prints one line per network and precision)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import modelcrowdnav_b200 as mcn  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def policy(net, precision):
    if net == "sarl":
        return mcn.BatchedSARL(precision=precision)
    if net == "om_sarl":
        return mcn.BatchedSARL(precision=precision, input_dim=61, with_om=1, cell_num=4, cell_size=1.0, om_channel_size=3)
    if net == "cadrl":
        return mcn.BatchedSARL(precision=precision, network="cadrl", mlp3_dims=[150, 100, 100, 1])
    return mcn.BatchedSARL(precision=precision, network="lstm_rl", mlp3_dims=[150, 100, 100, 1], lstm_hidden=50,
                           lstm_mlp1_dims=[0, 0, 0, 0])


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    nets = np.load(os.path.join(GOLDEN, "units_nets.npz"))
    om = np.load(os.path.join(GOLDEN, "units_om.npz"))
    weights = dict(sarl=np.load(os.path.join(GOLDEN, "sarl_weights_trained.npy")), om_sarl=om["om_sarl_weights"],
                   cadrl=nets["cadrl_weights"], lstm=nets["lstm_weights"])
    env = mcn.BatchedCrowdSim(E, H, auto_reset=1, seed=3)
    env.reset_device()
    for _ in range(8):
        env.orca()
        env.robot_orca(0.0)
        env.step(update=True, read=False)
    env.orca()
    for net in ("sarl", "om_sarl", "cadrl", "lstm"):
        ms = {}
        for precision in ("f16_tc", "f32"):
            pol = policy(net, precision)
            pol.load_weights(weights[net])
            for _ in range(3):
                pol.lookahead(env, 0)
            torch.cuda.synchronize()
            n = 20 if precision == "f16_tc" else 3
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                pol.lookahead(env, 0)
            b.record()
            torch.cuda.synchronize()
            ms[precision] = a.elapsed_time(b) / n
            pol.close()
        print("%-8s E=%d H=%d  lookahead f16_tc %.3f ms  f32 %.3f ms  (x%.1f)" % (net, E, H, ms["f16_tc"], ms["f32"],
                                                                               ms["f32"] / ms["f16_tc"]))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""crowd_nav/test_mul_env.py:96-113 on the B200 backend: the same trained value network evaluated over the test cases with
env.human_num = min .. max-1 (the SARL network pools any number of humans), each sweep point one batched
Explorer.run_k_episodes of all 500 test cases.

  python scripts/test_mul_env.py [--weights tests/golden/sarl_weights_trained.npy | --model_dir DIR] [--min_human_num 5]
                                 [--max_human_num 11] [--step_human_num 1] [--precision f16_tc] [--cases 500]
"""
import argparse
import json
import logging
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import modelcrowdnav_b200 as mcn  # noqa: E402
from modelcrowdnav_b200.train_loop import ENV_DEFAULT, POLICY_DEFAULT, make_config  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--weights", default=os.path.join(ROOT, "tests", "golden", "sarl_weights_trained.npy"))
    ap.add_argument("--model_dir", default=None, help="directory holding rl_model.pth (test_mul_env.py --model_dir)")
    ap.add_argument("--min_human_num", type=int, default=5)
    ap.add_argument("--max_human_num", type=int, default=11)
    ap.add_argument("--step_human_num", type=int, default=1)
    ap.add_argument("--precision", default="f16_tc", choices=["f32", "f16_tc"])
    ap.add_argument("--cases", type=int, default=500)
    ap.add_argument("--sim", default="circle_crossing")
    a = ap.parse_args()
    logging.basicConfig(level=logging.INFO, format="%(asctime)s, %(levelname)s: %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
    device = torch.device("cuda", 0)
    env_config, policy_config = make_config(ENV_DEFAULT), make_config(POLICY_DEFAULT)
    env_config.set("sim", "test_sim", a.sim)
    env_config.set("env", "test_size", str(a.cases))
    policy = mcn.policy_factory["sarl"]()
    policy.configure(policy_config)
    policy.precision = a.precision
    model = policy.get_model()
    if a.model_dir:
        model.load_state_dict(torch.load(os.path.join(a.model_dir, "rl_model.pth"), map_location="cpu"))
    else:
        flat, off, sd = np.load(a.weights), 0, {}
        for k, v in model.state_dict().items():
            sd[k] = torch.from_numpy(flat[off:off + v.numel()].reshape(tuple(v.shape)).copy())
            off += v.numel()
        model.load_state_dict(sd)
    policy.set_device(device)
    policy.set_phase("test")
    env = mcn.CrowdSim()
    env.configure(env_config)
    robot = mcn.Robot(env_config, "robot")
    robot.set_policy(policy)
    env.set_robot(robot)
    policy.set_env(env)
    explorer = mcn.Explorer(env, robot, device, gamma=0.9)
    rows = []
    for h in range(a.min_human_num, a.max_human_num, a.step_human_num):
        env.human_num = h                                   # test_mul_env.py:101
        env.case_counter["test"] = 0
        t0 = time.time()
        ret, sr, cr, tr = explorer.run_k_episodes(env.case_size["test"], "test")
        rows.append(dict(human_num=h, reward=ret, success=sr, collision=cr, timeout=tr, seconds=time.time() - t0,
                         env_steps=int(explorer.last_run["steps"].sum())))
    print(json.dumps(rows))


if __name__ == "__main__":
    main()

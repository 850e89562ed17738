"""Import the UNMODIFIED reference Python from /root/reference (build container only). ORACLE ONLY.

Adds stubs for modules absent from the image (gym, matplotlib, attrdict), the ``rvo2`` stand-in built on
the C oracle, and the ``np.NaN`` alias (explorer.py:51 predates numpy 2).  Used by
``scripts/gen_golden.py`` and by CPU tests that skip when /root/reference is missing; nothing under
``-m gpu``, ``smoke()`` or ``bench.py`` touches it.
"""
import configparser
import importlib
import os
import sys

REFERENCE_ROOT = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "crowd_sim"))


def install():
    import numpy as np
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    repo_root = os.path.dirname(os.path.dirname(_HERE))
    for p in (repo_root, os.path.join(_HERE, "stubs"), _HERE, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p) if p != REFERENCE_ROOT else sys.path.append(p)
    for name in ("gym", "matplotlib", "attrdict"):
        try:
            importlib.import_module(name)
        except ImportError:
            pass
    # crowd_sim.envs must be imported before crowd_nav.policy.* (circular import otherwise)
    importlib.import_module("crowd_sim.envs")
    importlib.import_module("crowd_sim")


def env_config(human_num=5, sim="circle_crossing", robot_visible=False, **over):
    cp = configparser.RawConfigParser()
    cp.read(os.path.join(REFERENCE_ROOT, "crowd_nav/configs/env.config"))
    cp.set("env", "look_ahead_in_sim", "false")           # missing key, crowd_sim.py:81
    cp.set("sim", "human_num", str(human_num))
    cp.set("sim", "train_val_sim", sim)
    cp.set("sim", "test_sim", sim)
    cp.set("robot", "visible", "true" if robot_visible else "false")
    for k, v in over.items():
        sec, key = k.split("__")
        cp.set(sec, key, str(v))
    return cp


def policy_config(query_env=False):
    cp = configparser.RawConfigParser()
    cp.read(os.path.join(REFERENCE_ROOT, "crowd_nav/configs/policy.config"))
    cp.set("action_space", "query_env", "true" if query_env else "false")
    return cp


def load_flat_weights(model, flat):
    """Load a flat fp32 vector (state-dict order) into a reference ValueNetwork."""
    import torch
    sd = model.state_dict()
    off = 0
    new = {}
    for k, v in sd.items():
        n = v.numel()
        new[k] = torch.from_numpy(flat[off:off + n].reshape(tuple(v.shape)).copy())
        off += n
    assert off == flat.size
    model.load_state_dict(new)


def make_env_and_sarl(human_num=5, sim="circle_crossing", query_env=False, seed=0, robot_visible=False, weights=None,
                      randomize=False, kinematics="holonomic", policy_name="sarl", policy_over=None):
    """Reference CrowdSim + Robot + SARL wired as crowd_nav/test.py:52-87 does (holonomic honoured)."""
    install()
    import torch
    import gym
    from crowd_sim.envs.utils.robot import Robot
    from crowd_nav.policy.policy_factory import policy_factory
    ecfg = env_config(human_num, sim, robot_visible, env__randomize_attributes="true" if randomize else "false")
    policy = policy_factory[policy_name]()
    torch.manual_seed(seed)
    pcfg = policy_config(query_env)
    for k, v in (policy_over or {}).items():
        sec, key = k.split("__")
        pcfg.set(sec, key, str(v))
    policy.configure(pcfg)
    # kinematics="holonomic": policy.config:14 honoured.  kinematics=None: the fork as shipped -- cadrl.py:66 comments the
    # config read out, so policy.kinematics stays None (ActionRot dynamics, theta feature zero).  "unicycle": explicit.
    policy.kinematics = kinematics
    if weights is not None:
        load_flat_weights(policy.get_model(), weights)
    env = gym.make("CrowdSim-v0")
    env.configure(ecfg)
    robot = Robot(ecfg, "robot")
    robot.set_policy(policy)
    env.set_robot(robot)
    policy.set_phase("test")
    policy.set_device(torch.device("cpu"))
    policy.set_env(env)
    return env, robot, policy

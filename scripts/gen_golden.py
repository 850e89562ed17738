#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference Python (/root/reference) in the build
container, with the ``rvo2`` stand-in from oracle/refshim (the real rvo2 is not installable here).

Everything first-party on the hot path (CrowdSim.reset/step/onestep_lookahead, SARL.predict, rotate,
compute_reward, build_action_space, ValueNetwork.forward) is executed by the reference's own code; the
fixtures therefore pin the oracle's restatement of that code.  ORCA velocities inside the fixtures come
from the oracle's rvo2 restatement (parity unpinned for that third-party piece, see crowdnav_oracle.h).

Usage:  python scripts/gen_golden.py [--episodes]   (the --episodes pass runs 500 full episodes, minutes)
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import oracle.refshim as refshim  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
INFO_CODE = {"Nothing": 0, "Danger": 1, "ReachGoal": 2, "Collision": 3, "Timeout": 4}


KIN_CODE = {"holonomic": 0, "unicycle": 1, None: 2}


def action_pair(a):
    """(vx, vy) of an ActionXY, (v, r) of an ActionRot."""
    return [a.vx, a.vy] if hasattr(a, "vx") else [a.v, a.r]


def agents_of(env):
    rows = []
    for a in [env.robot] + env.humans:
        rows.append([a.px, a.py, a.vx, a.vy, a.gx, a.gy, a.radius, a.v_pref])
    return np.array(rows, dtype=np.float64)


def run_trajectory(env, robot, policy, phase, case, max_steps=None):
    """Teacher-forced record of one episode driven exactly like explorer.py:53-69."""
    ob = env.reset(phase, case)
    rec = dict(agents=[], time=[], human_v=[], values=[], best=[], reward=[], done=[], info=[], dmin=[],
               action=[], theta=[])
    done = False
    steps = 0
    table = None
    while not done and (max_steps is None or steps < max_steps):
        rec["agents"].append(agents_of(env))
        rec["time"].append(env.global_time)
        rec["theta"].append(float(env.robot.theta))
        action = robot.act(ob)
        if table is None:
            table = np.array([action_pair(a) for a in policy.action_space])
        vals = np.array(policy.action_values, dtype=np.float64)
        rec["values"].append(vals)
        ap = action_pair(action)
        best = int(np.argmin(np.abs(table[:, 0] - ap[0]) + np.abs(table[:, 1] - ap[1])))
        rec["best"].append(best)
        rec["action"].append(ap)
        before = agents_of(env)
        ob, reward, done, info = env.step(action)
        after = agents_of(env)
        rec["human_v"].append(after[1:, 2:4].copy())  # ORCA velocities applied this step
        rec["reward"].append(reward)
        rec["done"].append(done)
        rec["info"].append(INFO_CODE[type(info).__name__])
        rec["dmin"].append(getattr(info, "min_dist", np.inf))
        assert np.array_equal(before[:, 4:], after[:, 4:])
        steps += 1
    out = {k: np.array(v) for k, v in rec.items()}
    out["table"] = table
    return out


def gen_units():
    refshim.install()
    import torch
    from crowd_sim.envs.utils.utils import point_to_segment_dist
    from crowd_sim.envs.utils.state import FullState, ObservableState
    env, robot, policy = refshim.make_env_and_sarl(seed=0)
    policy.time_step = 0.25                               # set by CrowdSim.reset (crowd_sim.py:307-309)
    rs = np.random.RandomState(1234)
    out = {}
    # action space (cadrl.py:82-102)
    policy.build_action_space(1.0)
    out["action_space"] = np.array([[a.vx, a.vy] for a in policy.action_space])
    out["gamma_bar"] = np.array(pow(policy.gamma, 0.25 * 1.0))
    # weights of the reference ValueNetwork under torch.manual_seed(0)
    sd = policy.get_model().state_dict()
    out["weight_keys"] = np.array(list(sd.keys()))
    flat = np.concatenate([v.numpy().ravel() for v in sd.values()]).astype(np.float32)
    np.save(os.path.join(GOLD, "sarl_weights_seed0.npy"), flat)
    assert np.array_equal(flat, oracle.default_sarl_weights(0)), "oracle weight init differs from reference"
    # point_to_segment_dist (utils.py:4-26)
    seg = rs.uniform(-3, 3, size=(256, 6))
    seg[:8, 2:4] = seg[:8, 0:2]           # degenerate segments
    seg[:, 4:] = 0
    out["seg_in"] = seg
    out["seg_out"] = np.array([point_to_segment_dist(*row) for row in seg])
    # rotate (cadrl.py:217-252), torch float32
    rows = rs.uniform(-5, 5, size=(512, 14)).astype(np.float32)
    rows[:, 4] = 0.3; rows[:, 13] = 0.3; rows[:, 7] = 1.0
    out["rotate_in"] = rows
    out["rotate_out"] = policy.rotate(torch.from_numpy(rows)).numpy()
    # compute_reward (multi_human_rl.py:65-88)
    cr_in, cr_out = [], []
    for i in range(256):
        nav = FullState(*rs.uniform(-2, 2, 2), 0, 0, 0.3, *rs.uniform(-2, 2, 2), 1.0, 0)
        if i % 16 == 0:
            nav = FullState(nav.gx + 0.1, nav.gy, 0, 0, 0.3, nav.gx, nav.gy, 1.0, 0)
        hs = [ObservableState(*(np.array([nav.px, nav.py]) + rs.uniform(-1.5, 1.5, 2)), 0, 0, 0.3) for _ in range(5)]
        cr_in.append([nav.px, nav.py, nav.radius, nav.gx, nav.gy] + sum([[h.px, h.py, h.radius] for h in hs], []))
        cr_out.append(policy.compute_reward(nav, hs))
    out["reward_in"] = np.array(cr_in)
    out["reward_out"] = np.array(cr_out, dtype=np.float64)
    # ValueNetwork.forward (sarl.py:28-65) on rotated random joint states, H = 5 and 10
    for H in (5, 10):
        x = policy.rotate(torch.from_numpy(rs.uniform(-4, 4, size=(64 * H, 14)).astype(np.float32)))
        x = x.reshape(64, H, 13)
        with torch.no_grad():
            v = policy.get_model()(x).numpy().ravel()
        out["vnet_in_h%d" % H] = x.numpy()
        out["vnet_out_h%d" % H] = v
    np.savez_compressed(os.path.join(GOLD, "units.npz"), **out)
    print("units.npz written;", len(out), "arrays")


TRAJ_SPECS = [
    # name, H, sim, query_env, robot_visible, [(phase, case, max_steps)]
    ("circle5_qfalse", 5, "circle_crossing", False, False, [("test", 0, 40), ("test", 1, 40), ("test", 7, 40), ("train", 3, 40)]),
    ("circle5_qtrue", 5, "circle_crossing", True, False, [("test", 2, 30), ("test", 499, 30)]),
    ("circle5_visible", 5, "circle_crossing", False, True, [("test", 5, 30)]),
    ("square10_qfalse", 10, "square_crossing", False, False, [("test", 0, 30), ("val", 11, 30)]),
    ("square10_qtrue", 10, "square_crossing", True, False, [("test", 3, 20)]),
]


# [env] randomize_attributes = true: heterogeneous human radii / preferred speeds (agent.py:39-45)
TRAJ_SPECS_RANDOM = [
    ("circle5_random", 5, "circle_crossing", False, False, [("test", 20, 30), ("test", 21, 30), ("train", 8, 30)], True),
    ("square10_random", 10, "square_crossing", True, False, [("test", 22, 20)], True),
]


# robot kinematics other than the honoured holonomic: None = the fork exactly as shipped (cadrl.py:66), "unicycle" explicit
TRAJ_SPECS_KIN = [
    ("circle5_kin_none", 5, "circle_crossing", False, False, [("test", 30, 40), ("test", 31, 40), ("train", 5, 40)], False, None),
    ("circle5_kin_none_qtrue", 5, "circle_crossing", True, False, [("test", 32, 30)], False, None),
    ("circle5_unicycle", 5, "circle_crossing", False, False, [("test", 33, 40), ("test", 34, 40)], False, "unicycle"),
    ("square10_unicycle_qtrue", 10, "square_crossing", True, False, [("test", 35, 20)], False, "unicycle"),
]


TRAJ_SPECS_TRAINED = [
    ("circle5_qfalse_trained", 5, "circle_crossing", False, False, [("test", 10, 60), ("test", 11, 60), ("test", 12, 60)]),
    ("circle5_qtrue_trained", 5, "circle_crossing", True, False, [("test", 13, 60)]),
]


# the other value networks behind the same lookahead (policy_factory: cadrl, lstm_rl): name, ..., randomize, kinematics,
# policy name, policy.config overrides
TRAJ_SPECS_NETS = [
    ("cadrl_circle5", 5, "circle_crossing", False, False, [("test", 40, 30), ("test", 41, 30)], False, "holonomic", "cadrl", None),
    ("cadrl_circle5_qtrue", 5, "circle_crossing", True, False, [("test", 42, 20)], False, "holonomic", "cadrl", None),
    ("cadrl_circle1", 1, "circle_crossing", False, False, [("test", 43, 30)], False, "holonomic", "cadrl", None),
    ("lstm_circle5", 5, "circle_crossing", False, False, [("test", 44, 30), ("test", 45, 30)], False, "holonomic", "lstm_rl", None),
    ("lstm_circle5_qtrue", 5, "circle_crossing", True, False, [("test", 46, 20)], False, "holonomic", "lstm_rl", None),
    ("lstm2_square10", 10, "square_crossing", False, False, [("test", 47, 20)], False, "holonomic", "lstm_rl",
     {"lstm_rl__with_interaction_module": "true"}),
    # the fork as shipped never reads [action_space] kinematics for ANY policy (cadrl.py:66): ActionRot planning
    ("cadrl_circle5_kin_none", 5, "circle_crossing", False, False, [("test", 48, 25)], False, None, "cadrl", None),
    ("lstm_circle5_kin_none_qtrue", 5, "circle_crossing", True, False, [("test", 49, 20)], False, None, "lstm_rl", None),
    ("lstm_circle5_unicycle", 5, "circle_crossing", False, False, [("test", 54, 25)], False, "unicycle", "lstm_rl", None),
]


def gen_net_units():
    """model(x) of the reference's CADRL / LSTM-RL value networks on rotated random joint states + their seed-0 weights."""
    import torch
    out = {}
    rs = np.random.RandomState(4321)
    for tag, pname, over, ncfg in (("cadrl", "cadrl", None, oracle.NetCfg.cadrl()),
                                   ("lstm", "lstm_rl", None, oracle.NetCfg.lstm_rl()),
                                   ("lstm2", "lstm_rl", {"lstm_rl__with_interaction_module": "true"}, oracle.NetCfg.lstm_rl(True))):
        env, robot, policy = refshim.make_env_and_sarl(seed=0, policy_name=pname, policy_over=over)
        sd = policy.get_model().state_dict()
        flat = np.concatenate([v.numpy().ravel() for v in sd.values()]).astype(np.float32)
        assert np.array_equal(flat, oracle.default_net_weights(ncfg, 0)), "oracle weight init differs from reference: " + tag
        out[tag + "_weight_keys"] = np.array(list(sd.keys()))
        out[tag + "_weights"] = flat
        for H in (1, 5):
            rows = rs.uniform(-4, 4, size=(32 * H, 14)).astype(np.float32)
            x = policy.rotate(torch.from_numpy(rows)).reshape(32, H, 13)
            with torch.no_grad():
                if pname == "cadrl":
                    v = policy.get_model()(x.reshape(-1, 13)).numpy().reshape(32, H)
                else:
                    v = policy.get_model()(x).numpy().ravel()
            out["%s_in_h%d" % (tag, H)] = x.numpy()
            out["%s_out_h%d" % (tag, H)] = v
    np.savez_compressed(os.path.join(GOLD, "units_nets.npz"), **out)
    print("units_nets.npz written;", len(out), "arrays")


# occupancy maps (with_om = true): input_dim 13 + 4 * 4 * 3 = 61
TRAJ_SPECS_OM = [
    ("om_sarl_circle5", 5, "circle_crossing", False, False, [("test", 50, 25), ("test", 51, 25)], False, "holonomic", "sarl",
     {"sarl__with_om": "true"}),
    ("om_sarl_square10_qtrue", 10, "square_crossing", True, False, [("test", 52, 15)], False, "holonomic", "sarl",
     {"sarl__with_om": "true"}),
    ("om_lstm_circle5", 5, "circle_crossing", False, False, [("test", 53, 25)], False, "holonomic", "lstm_rl",
     {"lstm_rl__with_om": "true"}),
]


def gen_om_units():
    """build_occupancy_maps on random human states, transform() with maps, and the seed-0 weights of the OM networks."""
    refshim.install()
    from crowd_sim.envs.utils.state import ObservableState, FullState, JointState
    out = {}
    rs = np.random.RandomState(777)
    for tag, pname, over in (("om_sarl", "sarl", {"sarl__with_om": "true"}), ("om_lstm", "lstm_rl", {"lstm_rl__with_om": "true"})):
        env, robot, policy = refshim.make_env_and_sarl(seed=0, policy_name=pname, policy_over=over)
        sd = policy.get_model().state_dict()
        out[tag + "_weights"] = np.concatenate([v.numpy().ravel() for v in sd.values()]).astype(np.float32)
        out[tag + "_weight_keys"] = np.array(list(sd.keys()))
    for H in (2, 5, 10):
        ins, maps, trs = [], [], []
        for _ in range(48):
            hs = rs.uniform(-2.5, 2.5, size=(H, 4))
            hs[:, 2:] = rs.uniform(-1, 1, size=(H, 2))
            if rs.rand() < 0.2:
                hs[0, 2:] = 0.0                                   # a standing human: arctan2(0, 0)
            humans = [ObservableState(*row, 0.3) for row in hs]
            maps.append(policy.build_occupancy_maps(humans).numpy())
            robot_state = FullState(*rs.uniform(-3, 3, 2), *rs.uniform(-1, 1, 2), 0.3, *rs.uniform(-3, 3, 2), 1.0, 0.0)
            policy.kinematics = "holonomic"
            trs.append(policy.transform(JointState(robot_state, humans)).numpy())
            ins.append(np.concatenate([[[robot_state.px, robot_state.py, robot_state.vx, robot_state.vy, robot_state.gx,
                                         robot_state.gy, 0.3, 1.0]],
                                       np.concatenate([hs, np.zeros((H, 2)), np.full((H, 1), 0.3), np.ones((H, 1))], axis=1)]))
        out["om_agents_h%d" % H] = np.array(ins)
        out["om_maps_h%d" % H] = np.array(maps)
        out["om_transform_h%d" % H] = np.array(trs)
    np.savez_compressed(os.path.join(GOLD, "units_om.npz"), **out)
    print("units_om.npz written;", len(out), "arrays")


# [sim] train_val_sim / test_sim = mixed: every scene draws its own human count (0-5 standing or 1-5 moving humans)
# (one case per fixture: the reference sizes human_times by the PREVIOUS scene's count, crowd_sim.py:272, and raises
# IndexError in step() as soon as a scene draws more humans than the one before)
TRAJ_SPECS_MIXED = [
    ("mixed_sarl_a", 5, "mixed", False, False, [("test", 0, 30)], False, "holonomic", "sarl", None),
    ("mixed_sarl_b", 5, "mixed", False, False, [("test", 3, 30)], False, "holonomic", "sarl", None),
    ("mixed_sarl_c", 5, "mixed", True, False, [("test", 12, 30)], False, "holonomic", "sarl", None),
]


def gen_mixed_scenes(n=64):
    """Reference reset() under sim = mixed for test cases 0..n-1 and val cases 0..15: agents per case (ragged)."""
    env, robot, policy = refshim.make_env_and_sarl(human_num=5, sim="mixed", seed=0)
    out = {}
    for phase, cases in (("test", range(n)), ("val", range(16))):
        for c in cases:
            env.reset(phase, c)
            a = agents_of(env)
            out["%s_%d" % (phase, c)] = a
            out["%s_%d_human_num" % (phase, c)] = np.array(env.human_num)
            assert np.array_equal(a, oracle.generate_scene(phase, c, human_num=5, rule="mixed")), (phase, c)
    np.savez_compressed(os.path.join(GOLD, "scenes_mixed.npz"), **out)
    print("scenes_mixed.npz written;", len(out) // 2, "scenes, human counts",
          sorted(set(int(out[k]) for k in out if k.endswith("human_num"))))


def gen_model_world():
    """ModelCrowdSim (model_crowd_sim.py) with a seed-1 AttentionWorld / MlpWorld driving the humans and the seed-0 SARL
    driving the robot: scenes from the seeded GLOBAL numpy stream, per step the world model's velocities (passed to step()
    through its own new_v argument so that they are on record), SARL's values, the env outcome and the next state."""
    refshim.install()
    import torch
    from crowd_sim.envs.model_crowd_sim import ModelCrowdSim
    from crowd_nav.policy.world_model import AttentionWorld, MlpWorld
    for tag, make_world, H, sim, qenv, np_seed in (("attn_circle5", AttentionWorld, 5, "circle_crossing", False, 11),
                                                    ("attn_square5_qtrue", AttentionWorld, 5, "square_crossing", True, 12),
                                                    ("mlp_circle3", lambda: MlpWorld(3), 3, "circle_crossing", False, 13)):
        _, robot, policy = refshim.make_env_and_sarl(human_num=H, sim=sim, query_env=qenv, seed=0)
        env = ModelCrowdSim()
        env.configure(refshim.env_config(H, sim))
        env.set_robot(robot)
        policy.set_env(env)
        torch.manual_seed(1)
        world = make_world()
        world.eval()                                   # MlpWorld carries Dropout layers
        env.sim_world, env.device = world, torch.device("cpu")
        np.random.seed(np_seed)
        ob = env.reset("test", 0)
        rec = dict(agents=[], time=[], new_v=[], values=[], best=[], action=[], reward=[], done=[], info=[], dmin=[])
        table, done, steps = None, False, 0
        while not done and steps < 25:
            rec["agents"].append(agents_of(env)); rec["time"].append(env.global_time)
            cur = torch.Tensor([[h.get_observable_state().getvalue() for h in env.humans]])
            with torch.no_grad():
                new_v = torch.reshape(world(cur.reshape(1, -1))[0], (H, 2)).tolist()
            action = robot.act(ob)
            if table is None:
                table = np.array([action_pair(a) for a in policy.action_space])
            rec["values"].append(np.array(policy.action_values, dtype=np.float64))
            ap = action_pair(action)
            rec["best"].append(int(np.argmin(np.abs(table[:, 0] - ap[0]) + np.abs(table[:, 1] - ap[1]))))
            rec["action"].append(ap); rec["new_v"].append(new_v)
            ob, reward, done, info = env.step(action, new_v=new_v)
            rec["reward"].append(reward); rec["done"].append(done); rec["info"].append(INFO_CODE[type(info).__name__])
            rec["dmin"].append(getattr(info, "min_dist", np.inf))
            steps += 1
        rec["agents"].append(agents_of(env)); rec["time"].append(env.global_time)     # state after the last step
        out = {k: np.array(v) for k, v in rec.items()}
        sd = world.state_dict()
        out.update(table=table, H=np.array(H), sim=np.array(sim), query_env=np.array(int(qenv)), np_seed=np.array(np_seed),
                   world=np.array("mlp" if tag.startswith("mlp") else "attention"),
                   world_keys=np.array(list(sd.keys())),
                   world_weights=np.concatenate([v.numpy().ravel() for v in sd.values()]).astype(np.float32))
        np.savez_compressed(os.path.join(GOLD, "model_world_%s.npz" % tag), **out)
        print("model_world_%s.npz: %d steps" % (tag, steps))


def gen_trajectories(specs=None, weights=None):
    for spec in (specs or TRAJ_SPECS):
        name, H, sim, qenv, vis, cases = spec[:6]
        randomize = bool(spec[6]) if len(spec) > 6 else False
        kinematics = spec[7] if len(spec) > 7 else "holonomic"
        pname = spec[8] if len(spec) > 8 else "sarl"
        pover = spec[9] if len(spec) > 9 else None
        env, robot, policy = refshim.make_env_and_sarl(human_num=H, sim=sim, query_env=qenv, seed=0,
                                                        robot_visible=vis, weights=weights, randomize=randomize,
                                                        kinematics=kinematics, policy_name=pname, policy_over=pover)
        out = {"H": np.array(H), "query_env": np.array(int(qenv)), "robot_visible": np.array(int(vis)),
               "sim": np.array(sim), "randomize": np.array(int(randomize)), "kinematics": np.array(KIN_CODE[kinematics]),
               "policy": np.array(pname),
               "interaction_module": np.array(int(bool(pover) and pover.get("lstm_rl__with_interaction_module") == "true")),
               "with_om": np.array(int(bool(pover) and "true" in (pover.get("sarl__with_om"), pover.get("lstm_rl__with_om"))))}
        t0 = time.time()
        for (phase, case, max_steps) in cases:
            rec = run_trajectory(env, robot, policy, phase, case, max_steps)
            key = "%s_%d" % (phase, case)
            for k, v in rec.items():
                out[key + "/" + k] = v
            # scene pin: oracle generator == reference reset
            scene = oracle.generate_scene(phase, case, human_num=H if pname != "cadrl" or phase == "test" else 1, rule=sim,
                                          randomize=randomize)
            assert np.array_equal(scene, rec["agents"][0]), (name, key)
        out["cases"] = np.array(["%s_%d" % (p, c) for p, c, _ in cases])
        np.savez_compressed(os.path.join(GOLD, "traj_%s.npz" % name), **out)
        print("traj_%s.npz written in %.1fs" % (name, time.time() - t0))


def run_episode_chunk(args):
    lo, hi, H, sim, qenv, wpath, kinematics = args
    import torch
    torch.set_num_threads(1)
    env, robot, policy = refshim.make_env_and_sarl(human_num=H, sim=sim, query_env=qenv, seed=0,
                                                    weights=np.load(wpath) if wpath else None, kinematics=kinematics)
    res = []
    for case in range(lo, hi):
        ob = env.reset("test", case)
        done, steps, too_close, ret = False, 0, 0, 0.0
        while not done:
            action = robot.act(ob)
            ob, reward, done, info = env.step(action)
            ret += pow(0.9, steps * 0.25 * 1.0) * reward
            steps += 1
            too_close += type(info).__name__ == "Danger"
        res.append((case, INFO_CODE[type(info).__name__], steps, too_close, env.global_time, ret))
    return res


def gen_episodes(n_proc, wpath=None, tag="seed0", kinematics="holonomic"):
    """crowd_nav/test.py equivalent: 500 test cases, SARL, circle_crossing, H=5.  kinematics=None: the fork as shipped."""
    import multiprocessing as mp
    chunks = [(lo, min(lo + 10, 500), 5, "circle_crossing", False, wpath, kinematics) for lo in range(0, 500, 10)]
    t0 = time.time()
    with mp.get_context("fork").Pool(n_proc) as pool:
        res = sum(pool.map(run_episode_chunk, chunks), [])
    res = np.array(sorted(res), dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, "episodes_circle5_%s.npz" % tag), case=res[:, 0].astype(np.int32),
                        info=res[:, 1].astype(np.int8), steps=res[:, 2].astype(np.int32),
                        too_close=res[:, 3].astype(np.int32), end_time=res[:, 4], ret=res[:, 5])
    k = len(res)
    print("episodes: success %.3f collision %.3f timeout %.3f  steps %d  (%.0fs, %d procs)" % (
        np.mean(res[:, 1] == 2), np.mean(res[:, 1] == 3), np.mean(res[:, 1] == 4), res[:, 2].sum(),
        time.time() - t0, n_proc))
    return k


# Decisive argmax sets: ALL 500 test cases teacher-forced by the reference with the trained SARL weights
# (|V| up to ~2, top-2 gaps far above the fp16 noise), a 1-in-`every` sample of the visited states recorded per case.
# name, H, sim, query_env, every
DECISIVE_SPECS = [
    ("circle5_qfalse", 5, "circle_crossing", False, 3),
    ("circle5_qtrue", 5, "circle_crossing", True, 6),
    ("square10_qfalse", 10, "square_crossing", False, 6),
]


def run_decisive_chunk(args):
    lo, hi, H, sim, qenv, wpath, every = args
    import torch
    torch.set_num_threads(1)
    env, robot, policy = refshim.make_env_and_sarl(human_num=H, sim=sim, query_env=qenv, seed=0, weights=np.load(wpath))
    rows = []
    for case in range(lo, hi):
        ob = env.reset("test", case)
        done, t = False, 0
        while not done:
            agents, gtime = agents_of(env), env.global_time
            action = robot.act(ob)
            if t % every == case % every:
                table = np.array([action_pair(a) for a in policy.action_space])
                ap = action_pair(action)
                best = int(np.argmin(np.abs(table[:, 0] - ap[0]) + np.abs(table[:, 1] - ap[1])))
                rows.append((case, t, agents, gtime, np.array(policy.action_values, dtype=np.float64), best))
            ob, reward, done, info = env.step(action)
            t += 1
    return rows


def gen_decisive(n_proc, only=None):
    import multiprocessing as mp
    wpath = os.path.join(GOLD, "sarl_weights_trained.npy")
    for name, H, sim, qenv, every in DECISIVE_SPECS:
        if only and name not in only:
            continue
        chunks = [(lo, min(lo + 5, 500), H, sim, qenv, wpath, every) for lo in range(0, 500, 5)]
        t0 = time.time()
        with mp.get_context("fork").Pool(n_proc) as pool:
            rows = sum(pool.map(run_decisive_chunk, chunks), [])
        rows.sort(key=lambda r: (r[0], r[1]))
        values = np.array([r[4] for r in rows])
        top2 = np.sort(values, axis=1)[:, -2:]
        np.savez_compressed(os.path.join(GOLD, "decisive_%s.npz" % name),
                            case=np.array([r[0] for r in rows], np.int16), step=np.array([r[1] for r in rows], np.int16),
                            agents=np.array([r[2] for r in rows]), time=np.array([r[3] for r in rows]),
                            values=values.astype(np.float32), gap=(top2[:, 1] - top2[:, 0]),
                            best=np.array([r[5] for r in rows], np.int16), H=np.array(H), sim=np.array(sim),
                            query_env=np.array(int(qenv)), weights=np.array("sarl_weights_trained.npy"))
        print("decisive_%s.npz: %d states of 500 cases, %d with top-2 gap > 2e-4 (%.0fs, %d procs)" % (
            name, len(rows), int(((top2[:, 1] - top2[:, 0]) > 2e-4).sum()), time.time() - t0, n_proc))



def gen_training():
    """Replay filling and SGD of the reference (explorer.py:153-186, trainer.py:61-82) -> tests/golden/training.npz.

    il_*: Explorer.run_k_episodes(8, 'train', update_memory=True, imitation_learning=True) with the ORCA robot
    (safety_space 0.15) and the seed-0 SARL as target_policy (train.py:157-169): the memory it leaves behind.
    rl_*: the same call with the trained SARL driving the robot (epsilon 0, so no random draws) and a copy of it as the
    target model: value_i = r_i + gamma_bar * V_target(s_{i+1}).
    sgd_*: the reference Trainer.optimize_batch(100) on the IL memory from the seed-0 network, lr 0.01, batch 100, with the
    100 index batches its DataLoader drew on record, and the weights after the 100 steps."""
    refshim.install()
    import torch
    from crowd_nav.utils.explorer import Explorer
    from crowd_nav.utils.memory import ReplayMemory
    from crowd_nav.utils.trainer import Trainer
    from crowd_sim.envs.policy.policy_factory import policy_factory as sim_policy_factory
    out = {}
    torch.set_num_threads(1)

    def dump(memory, tag):
        out[tag + "_states"] = torch.stack([m[0] for m in memory.memory]).numpy()
        out[tag + "_values"] = torch.stack([m[1] for m in memory.memory]).numpy().reshape(-1)

    # ---- imitation learning ----
    env, robot, policy = refshim.make_env_and_sarl(seed=0)
    memory = ReplayMemory(100000)
    explorer = Explorer(env, robot, torch.device("cpu"), memory, policy.gamma, target_policy=policy)
    il_policy = sim_policy_factory["orca"]()
    il_policy.multiagent_training = policy.multiagent_training
    il_policy.safety_space = 0.15
    robot.set_policy(il_policy)
    res = explorer.run_k_episodes(8, "train", update_memory=True, imitation_learning=True)
    dump(memory, "il")
    out["il_result"] = np.array(res, dtype=np.float64)
    print("IL memory: %d states, result %s" % (len(memory), res))

    # ---- SGD on the IL memory (reference Trainer, batches on record) ----
    class RecordingMemory(ReplayMemory):
        def __init__(self, src):
            super().__init__(src.capacity)
            self.memory, self.position, self.log = list(src.memory), src.position, []

        def __getitem__(self, item):
            self.log.append(int(item))
            return self.memory[item]

    rec = RecordingMemory(memory)
    torch.manual_seed(0)
    _, _, fresh = refshim.make_env_and_sarl(seed=0)
    model = fresh.get_model()
    w0 = np.concatenate([v.numpy().ravel() for v in model.state_dict().values()]).astype(np.float32)
    assert np.array_equal(w0, np.load(os.path.join(GOLD, "sarl_weights_seed0.npy")))
    trainer = Trainer(model, rec, torch.device("cpu"), 100)
    trainer.set_learning_rate(0.01)
    torch.manual_seed(123)
    loss = trainer.optimize_batch(100)
    idx = np.array(rec.log, dtype=np.int32).reshape(100, 100)
    out["sgd_idx"] = idx
    out["sgd_loss"] = np.array(loss)
    out["sgd_weights"] = np.concatenate([v.numpy().ravel() for v in model.state_dict().values()]).astype(np.float32)
    print("SGD: average loss %.3e, |w - w0| max %.3e" % (loss, np.abs(out["sgd_weights"] - w0).max()))

    # ---- reinforcement learning targets ----
    wtr = np.load(os.path.join(GOLD, "sarl_weights_trained.npy"))
    env, robot, policy = refshim.make_env_and_sarl(seed=0, weights=wtr)
    memory = ReplayMemory(100000)
    explorer = Explorer(env, robot, torch.device("cpu"), memory, policy.gamma, target_policy=policy)
    explorer.update_target_model(policy.get_model())
    policy.set_epsilon(0.0)
    res = explorer.run_k_episodes(8, "train", update_memory=True, episode=0, returnRate=False)
    dump(memory, "rl")
    out["rl_result"] = np.array(res, dtype=np.float64)
    print("RL memory: %d states, result %s" % (len(memory), res))
    np.savez_compressed(os.path.join(GOLD, "training.npz"), **out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--episodes", action="store_true")
    ap.add_argument("--procs", type=int, default=6)
    ap.add_argument("--skip-units", action="store_true")
    ap.add_argument("--kin-none", action="store_true", help="with --episodes: the fork's literal kinematics (None)")
    ap.add_argument("--random", action="store_true", help="only the randomize_attributes trajectories")
    ap.add_argument("--kinematics", action="store_true", help="only the kinematics = None / unicycle trajectories")
    ap.add_argument("--model-world", action="store_true", help="only the ModelCrowdSim (world-model humans) fixtures")
    ap.add_argument("--mixed", action="store_true", help="only the 'mixed' scene fixtures and trajectories")
    ap.add_argument("--om", action="store_true", help="only the occupancy-map (with_om) unit vectors and trajectories")
    ap.add_argument("--nets", action="store_true", help="only the CADRL / LSTM-RL unit vectors and trajectories")
    ap.add_argument("--training", action="store_true", help="only the replay-filling / SGD fixtures (explorer.update_memory, Trainer)")
    ap.add_argument("--decisive", action="store_true", help="only the 500-case trained-weight argmax sets (minutes per set)")
    ap.add_argument("--trained", action="store_true", help="use tests/golden/sarl_weights_trained.npy (GPU-trained SARL)")
    a = ap.parse_args()
    wtrained = os.path.join(GOLD, "sarl_weights_trained.npy")
    assert refshim.available(), "/root/reference is required"
    os.makedirs(GOLD, exist_ok=True)
    oracle.build()
    if a.training:
        gen_training()
    elif a.decisive:
        gen_decisive(a.procs, os.environ.get("GOLDEN_ONLY"))
    elif a.episodes and a.kin_none:
        gen_episodes(a.procs, wtrained if a.trained else None, "kin_none_" + ("trained" if a.trained else "seed0"), None)
    elif a.episodes:
        gen_episodes(a.procs, wtrained if a.trained else None, "trained" if a.trained else "seed0")
    elif a.model_world:
        gen_model_world()
    elif a.mixed:
        gen_mixed_scenes()
        gen_trajectories(TRAJ_SPECS_MIXED)
    elif a.om:
        gen_om_units()
        gen_trajectories(TRAJ_SPECS_OM)
    elif a.nets:
        only = os.environ.get("GOLDEN_ONLY")            # comma-separated fixture names: regenerate just those
        if only:
            gen_trajectories([sp for sp in TRAJ_SPECS_NETS if sp[0] in only.split(",")])
        else:
            gen_net_units()
            gen_trajectories(TRAJ_SPECS_NETS)
    elif a.random:
        gen_trajectories(TRAJ_SPECS_RANDOM)
    elif a.kinematics:
        gen_trajectories(TRAJ_SPECS_KIN)
    elif a.trained:
        gen_trajectories(TRAJ_SPECS_TRAINED, np.load(wtrained))
    else:
        if not a.skip_units:
            gen_units()
        gen_trajectories()
        gen_trajectories(TRAJ_SPECS_RANDOM)
        gen_trajectories(TRAJ_SPECS_KIN)

"""The reference-facing Python surface (CrowdSim / SARL / Explorer mirrors) on the GPU, checked against the
reference's own outputs (tests/golden, produced by scripts/gen_golden.py from the unmodified reference)."""
import configparser
import os

import numpy as np
import pytest

from conftest import GOLDEN, MODEL_WORLD_NAMES, load_model_world, load_traj, net_tag, weights_for

pytestmark = pytest.mark.gpu

ENV_INI = """
[env]
time_limit = 25
time_step = 0.25
val_size = 100
test_size = 500
randomize_attributes = false
[reward]
success_reward = 1
collision_penalty = -0.25
discomfort_dist = 0.2
discomfort_penalty_factor = 0.5
[sim]
train_val_sim = circle_crossing
test_sim = circle_crossing
square_width = 10
circle_radius = 4
human_num = 5
[humans]
visible = true
policy = orca
radius = 0.3
v_pref = 1
sensor = coordinates
[robot]
visible = false
policy = none
radius = 0.3
v_pref = 1
sensor = coordinates
"""
POLICY_INI = """
[rl]
gamma = 0.9
[om]
cell_num = 4
cell_size = 1
om_channel_size = 3
[action_space]
kinematics = holonomic
speed_samples = 5
rotation_samples = 16
sampling = exponential
query_env = false
[sarl]
mlp1_dims = 150, 100
mlp2_dims = 100, 50
attention_dims = 100, 100, 1
mlp3_dims = 150, 100, 100, 1
multiagent_training = true
with_om = false
with_global_state = true
[cadrl]
mlp_dims = 150, 100, 100, 1
multiagent_training = false
[lstm_rl]
global_state_dim = 50
mlp1_dims = 150, 100, 100, 50
mlp2_dims = 150, 100, 100, 1
multiagent_training = true
with_om = false
with_interaction_module = false
"""


def _cfg(text, **over):
    cp = configparser.RawConfigParser()
    cp.read_string(text)
    for k, v in over.items():
        sec, key = k.split("__")
        cp.set(sec, key, str(v))
    return cp


KIN_NAME = {0: "holonomic", 1: "unicycle", 2: None}


def _setup(weights0, precision="f32", query_env=False, human_num=5, sim="circle_crossing", randomize=False,
           kinematics="holonomic", policy_name="sarl", interaction_module=False, with_om=False):
    """Wire env, robot, policy, explorer exactly as crowd_nav/test.py:52-87 does."""
    import torch
    import modelcrowdnav_b200 as mcn
    ecfg = _cfg(ENV_INI, sim__human_num=human_num, sim__train_val_sim=sim, sim__test_sim=sim,
                env__randomize_attributes="true" if randomize else "false")
    pcfg = _cfg(POLICY_INI, action_space__query_env="true" if query_env else "false",
                action_space__kinematics=kinematics or "holonomic",
                lstm_rl__with_interaction_module="true" if interaction_module else "false",
                sarl__with_om="true" if with_om else "false", lstm_rl__with_om="true" if with_om else "false")
    policy = mcn.policy_factory[policy_name]()
    import modelcrowdnav_b200.policy as policy_mod
    policy_mod.LITERAL_FORK_KINEMATICS = kinematics is None      # None: the fork never reads the key (cadrl.py:66)
    try:
        policy.configure(pcfg)
    finally:
        policy_mod.LITERAL_FORK_KINEMATICS = False
    assert policy.kinematics == kinematics
    if policy_name == "sarl" or precision == "f32":         # CADRL / LSTM-RL pick their own default (tensor cores when supported)
        policy.precision = precision
    sd = policy.get_model().state_dict()
    off = 0
    new = {}
    for k, v in sd.items():
        n = v.numel()
        new[k] = torch.from_numpy(weights0[off:off + n].reshape(tuple(v.shape)).copy())
        off += n
    policy.get_model().load_state_dict(new)                       # test.py:59
    env = mcn.CrowdSim()
    env.configure(ecfg)
    robot = mcn.Robot(ecfg, "robot")
    robot.set_policy(policy)
    env.set_robot(robot)
    device = torch.device("cuda:0")
    explorer = mcn.Explorer(env, robot, device, gamma=0.9)
    policy.set_phase("test")
    policy.set_device(device)
    policy.set_env(env)
    return env, robot, policy, explorer


def test_state_dict_keys_match_reference(weights0, units):
    env, robot, policy, _ = _setup(weights0)
    assert list(policy.get_model().state_dict().keys()) == [str(k) for k in units["weight_keys"]]
    assert np.array_equal(policy.flat_weights(), weights0)


@pytest.mark.parametrize("name", ["circle5_qfalse", "circle5_qtrue", "circle5_qfalse_trained", "circle5_qtrue_trained",
                                  "circle5_random", "square10_random",
                                  "circle5_kin_none", "circle5_kin_none_qtrue", "circle5_unicycle", "square10_unicycle_qtrue",
                                  "cadrl_circle5", "cadrl_circle5_qtrue", "cadrl_circle1", "lstm_circle5",
                                  "lstm_circle5_qtrue", "lstm2_square10", "cadrl_circle5_kin_none",
                                  "lstm_circle5_kin_none_qtrue", "lstm_circle5_unicycle",
                                  "om_sarl_circle5", "om_sarl_square10_qtrue", "om_lstm_circle5",
                                  "mixed_sarl_a", "mixed_sarl_b", "mixed_sarl_c"])
def test_facade_replays_reference_episode(name):
    """gym-style loop (explorer.py:53-69) through the single-env façade: ob/reward/done/info, action values and
    chosen actions equal the reference's, step by step, while the façade follows its own actions."""
    import modelcrowdnav_b200 as mcn
    tr = load_traj(name)
    weights0 = weights_for(name)
    if tr["with_om"]:                                            # occupancy maps: [sarl] / [lstm_rl] with_om = true
        weights0 = np.load(os.path.join(GOLDEN, "units_om.npz"))[("om_sarl" if tr["policy"] == "sarl" else "om_lstm") + "_weights"]
    elif tr["policy"] != "sarl":                                 # CADRL / LSTM-RL: policy_factory['cadrl' | 'lstm_rl']
        weights0 = np.load(os.path.join(GOLDEN, "units_nets.npz"))[net_tag(tr) + "_weights"]
    env, robot, policy, _ = _setup(weights0, "f32", query_env=bool(tr["query_env"]),
                                   human_num=5 if tr["sim"] == "mixed" else tr["H"], sim=tr["sim"],
                                   randomize=bool(tr["randomize"]), kinematics=KIN_NAME[tr["kinematics"]],
                                   policy_name=tr["policy"], interaction_module=bool(tr["interaction_module"]),
                                   with_om=bool(tr["with_om"]))
    holonomic = tr["kinematics"] == 0
    case = [c for c in tr["cases"] if c.startswith("test_")][0]
    rec = tr["cases"][case]
    ob = env.reset("test", int(case.split("_")[1]))
    info_types = {0: mcn.Nothing, 1: mcn.Danger, 2: mcn.ReachGoal, 3: mcn.Collision, 4: mcn.Timeout}
    for t in range(len(rec["time"])):
        assert env.global_time == rec["time"][t]
        got = np.array([[o.px, o.py, o.vx, o.vy, o.radius] for o in ob])
        assert np.array_equal(got, rec["agents"][t][1:, [0, 1, 2, 3, 6]])
        action = robot.act(ob)
        ref_v = rec["values"][t]
        assert np.max(np.abs(np.array(policy.action_values) - ref_v)) <= 1e-5
        top2 = np.sort(ref_v)[-2:]
        cls = mcn.ActionXY if holonomic else mcn.ActionRot
        assert isinstance(action, cls)
        if top2[1] - top2[0] <= 2e-5:
            action = cls(*rec["action"][t])                   # tie in the reference: follow its choice
        else:
            assert tuple(action) == tuple(rec["action"][t])
        if not holonomic:
            assert abs(env.robot.theta - rec["theta"][t]) <= 1e-12
        ob, reward, done, info = env.step(action)
        assert reward == rec["reward"][t] and done == bool(rec["done"][t])
        assert isinstance(info, info_types[int(rec["info"][t])])
        if isinstance(info, mcn.Danger):
            assert info.min_dist == rec["dmin"][t]


def test_facade_errors_match_reference(weights0):
    import modelcrowdnav_b200 as mcn
    env = mcn.CrowdSim()
    env.configure(_cfg(ENV_INI))
    with pytest.raises(AttributeError):
        env.reset("test")                                          # crowd_sim.py:266-267
    env2, robot, policy, _ = _setup(weights0)
    policy.set_phase(None)
    ob = env2.reset("test", 0)
    with pytest.raises(AttributeError):
        robot.act(ob)                                              # multi_human_rl.py:17-18
    policy.set_phase("train")
    with pytest.raises(AttributeError):
        robot.act(ob)                                              # epsilon unset (multi_human_rl.py:19-20)
    with pytest.raises(AssertionError):
        env2.reset("bogus")                                        # crowd_sim.py:268
    with pytest.raises(AssertionError):
        env2.step(mcn.ActionRot(1.0, 0.0))                         # agent.py:104-108


@pytest.mark.parametrize("wset", ["seed0", "trained", "kin_none_trained"])
@pytest.mark.parametrize("precision", ["f32", "f16_tc"])
def test_explorer_500_test_episodes_match_reference(precision, wset):
    """crowd_nav/test.py equivalent: 500 test cases, SARL (random-init or GPU-trained weights), circle_crossing,
    5 humans.  The golden file holds the per-episode outcome of the REFERENCE's own run_k_episodes loop
    (scripts/gen_golden.py --episodes [--trained])."""
    g = np.load(os.path.join(GOLDEN, "episodes_circle5_%s.npz" % wset))
    weights0 = weights_for("x_" + wset)
    # "kin_none_*": the fork exactly as shipped (policy.kinematics stays None: ActionRot actions, non-holonomic robot)
    env, robot, policy, explorer = _setup(weights0, precision, kinematics=None if wset.startswith("kin_none") else "holonomic")
    ret, sr, cr, tr_, nav = explorer.run_k_episodes(env.case_size["test"], "test", print_failure=True, returnNav=True)
    run = explorer.last_run
    assert list(run["cases"]) == list(g["case"])
    ref_sr, ref_cr, ref_tr = np.mean(g["info"] == 2), np.mean(g["info"] == 3), np.mean(g["info"] == 4)
    # north star: episode-level success / collision / timeout rates within 0.5 pp over 500 episodes
    assert abs(sr - ref_sr) <= 0.005 and abs(cr - ref_cr) <= 0.005 and abs(tr_ - ref_tr) <= 0.005
    same = (run["info"] == g["info"]) & (run["steps"] == g["steps"])
    if precision == "f32":
        assert same.mean() >= 0.99, same.mean()
        assert abs(ret - float(np.mean(g["ret"]))) < 2e-3
    if wset == "trained":
        assert sr > 0.95                                             # a real policy: SARL reaches the goal
    assert env.case_counter["test"] == 0                            # 500 % 500


def test_explorer_imitation_learning_fills_memory(weights0):
    """IL phase (train.py:157-178): ORCA robot with safety_space, discounted return-to-go targets."""
    import torch
    import modelcrowdnav_b200 as mcn
    from modelcrowdnav_b200.trainer import Trainer
    env, robot, policy, _ = _setup(weights0)
    il_policy = mcn.policy_factory["orca"]()
    il_policy.multiagent_training = policy.multiagent_training
    il_policy.safety_space = 0.15
    robot.set_policy(il_policy)
    device = torch.device("cuda:0")
    memory = mcn.ReplayMemory(100000)
    explorer = mcn.Explorer(env, robot, device, memory, 0.9, target_policy=policy)
    out = explorer.run_k_episodes(64, "train", update_memory=True, imitation_learning=True)
    assert out[1] > 0.8                                              # ORCA robot mostly succeeds
    run = explorer.last_run
    kept = (run["info"] == 2) | (run["info"] == 3)
    assert len(memory) == int(run["steps"][kept].sum())
    s, v = memory[0]
    assert tuple(s.shape) == (5, 13) and tuple(v.shape) == (1,)
    # first stored episode: value_0 = sum_t gamma^(t*dt*v_pref) r_t = that episode's discounted return
    first = int(np.nonzero(kept)[0][0])
    assert abs(float(v) - run["returns"][first]) < 1e-5
    # a short supervised fit must reduce the loss (trainer.py:36-59) and refresh the GPU weights
    robot.set_policy(policy)
    trainer = Trainer(policy.get_model(), memory, device, 100, policy=policy)
    trainer.set_learning_rate(0.01)
    l0 = trainer.optimize_epoch(1)
    l1 = trainer.optimize_epoch(5)
    assert l1 < l0
    before = policy.handle(1.0).n_params
    assert before == 96502


def test_explorer_rl_td_targets(weights0):
    """RL phase (explorer.py:168-174): value_i = r_i + gamma_bar * V_target(s_{i+1}), terminal = r."""
    import torch
    import modelcrowdnav_b200 as mcn
    env, robot, policy, _ = _setup(weights0)
    device = torch.device("cuda:0")
    memory = mcn.ReplayMemory(100000)
    explorer = mcn.Explorer(env, robot, device, memory, 0.9, target_policy=policy)
    explorer.update_target_model(policy.get_model())
    policy.set_epsilon(0.0)
    # start close to the goal so that episodes end in ReachGoal within a few steps
    out = explorer.run_k_episodes(32, "train", update_memory=True)
    run = explorer.last_run
    kept = (run["info"] == 2) | (run["info"] == 3)
    assert len(memory) == int(run["steps"][kept].sum())
    if len(memory):
        states, values = memory.states[:len(memory)], memory.values[:len(memory)]
        with torch.no_grad():
            vt = explorer.target_model.to(device)(states[1:]).reshape(-1)
        # for non-terminal entries the stored value equals r + gamma_bar * V(next entry); check consistency where r = 0
        gb = 0.9 ** 0.25
        resid = (values[:-1].reshape(-1) - gb * vt).abs()
        assert (resid < 1e-4).float().mean() > 0.5


@pytest.mark.parametrize("tag,pname,im", [("cadrl", "cadrl", False), ("lstm", "lstm_rl", False), ("lstm2", "lstm_rl", True)])
def test_other_policies_state_dict_and_training_surface(tag, pname, im):
    """policy_factory['cadrl' | 'lstm_rl']: state-dict keys and flat weights equal the reference's (a reference checkpoint
    loads unchanged); the torch module used for training computes what the CUDA network computes; last_state of LSTM-RL
    comes out in predict()'s sorted human order; CADRL.transform insists on one human (cadrl.py:209)."""
    import torch
    import modelcrowdnav_b200 as mcn
    z = np.load(os.path.join(GOLDEN, "units_nets.npz"))
    w = z[tag + "_weights"]
    env, robot, policy, _ = _setup(w, policy_name=pname, interaction_module=im)
    assert list(policy.get_model().state_dict().keys()) == [str(k) for k in z[tag + "_weight_keys"]]
    assert np.array_equal(policy.flat_weights(), w)
    x = torch.from_numpy(z[tag + "_in_h5"]).cuda()
    with torch.no_grad():
        ref = policy.get_model()(x.reshape(-1, 13) if tag == "cadrl" else x).reshape(x.shape[0], -1)
    got = policy.handle(1.0).forward(x)
    want = ref.min(dim=1).values if tag == "cadrl" else ref[:, 0]
    assert float((got - want).abs().max()) <= 1e-5
    ob = env.reset("test", 3)
    policy.set_phase("train"); policy.set_epsilon(0.0)
    if tag == "cadrl":
        with pytest.raises(AssertionError):
            robot.act(ob)                                          # 5 humans: transform asserts a single human
    else:
        robot.act(ob)
        d = [np.hypot(h.px - env.robot.px, h.py - env.robot.py) for h in env.humans]
        order = sorted(range(len(d)), key=lambda i: d[i], reverse=True)
        plain = policy.transform(mcn.JointState(env.robot.get_full_state(), [h.get_observable_state() for h in env.humans]))
        assert torch.equal(policy.last_state, plain[order])


def test_explorer_groups_mixed_scenes_by_human_count(weights0):
    """[sim] mixed: run_k_episodes rolls the cases out as one batch per human count; every episode must end exactly as
    the same case driven alone through the gym-style façade loop (explorer.py:53-69)."""
    import modelcrowdnav_b200 as mcn
    env, robot, policy, explorer = _setup(weights0, "f32", sim="mixed")
    k = 12
    explorer.run_k_episodes(k, "test")
    run = explorer.last_run
    assert list(run["cases"]) == list(range(k))
    sizes = set()
    for c in range(k):
        ob = env.reset("test", c)
        sizes.add(len(ob))
        done, steps = False, 0
        while not done:
            ob, reward, done, info = env.step(robot.act(ob))
            steps += 1
        code = {mcn.ReachGoal: 2, mcn.Collision: 3, mcn.Timeout: 4}[type(info)]
        assert (code, steps) == (int(run["info"][c]), int(run["steps"][c])), c
    assert len(sizes) >= 3

    # multiagent_training = false (CADRL): train / val scenes hold ONE circle-crossing human (crowd_sim.py:272-292)
    z = np.load(os.path.join(GOLDEN, "units_nets.npz"))
    env2, robot2, policy2, _ = _setup(z["cadrl_weights"], policy_name="cadrl", sim="square_crossing")
    assert policy2.multiagent_training is False
    assert len(env2.reset("val", 0)) == 1 and env2.train_val_sim == "circle_crossing"
    assert len(env2.reset("test", 0)) == 5


@pytest.mark.parametrize("name", MODEL_WORLD_NAMES)
def test_model_crowd_sim_facade_replays_reference(weights0, name):
    """gym.make('ModelCrowdSim-v0') mirror: humans move with a world model's velocities (cn_env_set_human_actions), the robot
    with SARL.  The reference's checkpoint layout loads unchanged; the scene comes from the seeded global numpy stream; the
    façade's own world-model call agrees with the reference's to 1e-5, and under the recorded velocities values, outcomes
    and states replay exactly."""
    import torch
    import modelcrowdnav_b200 as mcn
    from modelcrowdnav_b200.world_model import AttentionWorld, MlpWorld
    g = load_model_world(name)
    H = int(g["H"])
    _, robot, policy, _ = _setup(weights0, "f32", query_env=bool(g["query_env"]), human_num=H, sim=str(g["sim"]))
    import modelcrowdnav_b200.compat as compat
    env = compat.make("ModelCrowdSim-v0")
    env.configure(_cfg(ENV_INI, sim__human_num=H, sim__train_val_sim=str(g["sim"]), sim__test_sim=str(g["sim"])))
    env.set_robot(robot)
    policy.set_env(env)
    world = MlpWorld(H) if str(g["world"]) == "mlp" else AttentionWorld()
    sd = world.state_dict()
    assert list(sd.keys()) == [str(k) for k in g["world_keys"]]
    off, new = 0, {}
    for k, v in sd.items():
        new[k] = torch.from_numpy(g["world_weights"][off:off + v.numel()].reshape(tuple(v.shape)).copy())
        off += v.numel()
    world.load_state_dict(new)
    world.eval()
    env.sim_world, env.device = world, torch.device("cuda:0")
    state = np.random.get_state()
    try:
        np.random.seed(int(g["np_seed"]))
        ob = env.reset("test", 0)
    finally:
        np.random.set_state(state)
    info_types = {0: mcn.Nothing, 1: mcn.Danger, 2: mcn.ReachGoal, 3: mcn.Collision, 4: mcn.Timeout}
    for t in range(len(g["reward"])):
        got = np.array([[o.px, o.py, o.vx, o.vy] for o in ob])
        assert np.array_equal(got, g["agents"][t][1:, :4]), t
        assert np.max(np.abs(np.array(env.world_velocities()) - g["new_v"][t])) <= 1e-5
        if g["query_env"]:
            env._human_v = None
            env._ensure_batch().set_human_actions(g["new_v"][t][None])        # the lookahead sees the recorded velocities
            env._human_v = True
        action = robot.act(ob)
        ref_v = g["values"][t]
        assert np.max(np.abs(np.array(policy.action_values) - ref_v)) <= 1e-5
        top2 = np.sort(ref_v)[-2:]
        if top2[1] - top2[0] > 2e-5:
            assert tuple(action) == tuple(g["action"][t])
        ob, reward, done, info = env.step(mcn.ActionXY(*g["action"][t]), new_v=g["new_v"][t])
        assert reward == g["reward"][t] and done == bool(g["done"][t]) and isinstance(info, info_types[int(g["info"][t])])
        assert env.global_time == g["time"][t + 1]


def test_linear_robot_policy(weights0):
    """policy_factory['linear'] (crowd_sim/envs/policy/linear.py): a robot heading straight for its goal at v_pref, both
    through the gym-style loop and through the batched Explorer; the two must end every episode identically."""
    import modelcrowdnav_b200 as mcn
    env, robot, policy, _ = _setup(weights0)
    lin = mcn.policy_factory["linear"]()
    lin.configure(None)
    robot.set_policy(lin)
    import torch
    explorer = mcn.Explorer(env, robot, torch.device("cuda:0"), gamma=0.9)
    explorer.run_k_episodes(8, "test")
    run = explorer.last_run
    for c in range(8):
        ob = env.reset("test", c)
        done, steps = False, 0
        while not done:
            action = robot.act(ob)
            assert abs(np.hypot(action.vx, action.vy) - 1.0) < 1e-12
            ob, reward, done, info = env.step(action)
            steps += 1
        code = {mcn.ReachGoal: 2, mcn.Collision: 3, mcn.Timeout: 4}[type(info)]
        assert (code, steps) == (int(run["info"][c]), int(run["steps"][c])), c


def test_explorer_over_model_crowd_sim(weights0):
    """Explorer.run_k_episodes on a ModelCrowdSim env: the batch asks the world model for every env's human velocities
    (one forward per step) instead of ORCA.  Scenes come from the global numpy stream in the order sequential resets would
    draw them; episodes must end like the same scenes driven one by one through the façade (batched and single forwards of
    the world model may round differently, so a small fraction of episodes may drift)."""
    import torch
    import modelcrowdnav_b200 as mcn
    import modelcrowdnav_b200.compat as compat
    from modelcrowdnav_b200.world_model import AttentionWorld
    g = load_model_world("attn_circle5")
    _, robot, policy, _ = _setup(weights0, "f32", human_num=5)
    env = compat.make("ModelCrowdSim-v0")
    env.configure(_cfg(ENV_INI))
    env.set_robot(robot)
    policy.set_env(env)
    world = AttentionWorld()
    off, new = 0, {}
    for k, v in world.state_dict().items():
        new[k] = torch.from_numpy(g["world_weights"][off:off + v.numel()].reshape(tuple(v.shape)).copy())
        off += v.numel()
    world.load_state_dict(new)
    env.sim_world, env.device = world, torch.device("cuda:0")
    explorer = mcn.Explorer(env, robot, torch.device("cuda:0"), gamma=0.9)
    k = 16
    state = np.random.get_state()
    try:
        np.random.seed(21)
        explorer.run_k_episodes(k, "test")
        run = explorer.last_run
        np.random.seed(21)
        same = 0
        for c in range(k):
            ob = env.reset("test", c)
            done, steps = False, 0
            while not done:
                ob, reward, done, info = env.step(robot.act(ob))
                steps += 1
            code = {mcn.ReachGoal: 2, mcn.Collision: 3, mcn.Timeout: 4}[type(info)]
            same += int((code, steps) == (int(run["info"][c]), int(run["steps"][c])))
    finally:
        np.random.set_state(state)
    assert same >= k - 2, same


@pytest.mark.parametrize("name", MODEL_WORLD_NAMES)
def test_world_model_kernel_matches_reference(name):
    """cn_world_predict (csrc/world_model.cu) against the velocities the REFERENCE's AttentionWorld / MlpWorld predicted for
    every recorded step (tests/golden/model_world_*.npz): all steps of the episode as one env batch, on the device."""
    import torch
    import modelcrowdnav_b200 as mcn
    from modelcrowdnav_b200.world_model import AttentionWorld, MlpWorld
    g = load_model_world(name)
    H, T = int(g["H"]), len(g["reward"])
    world = MlpWorld(H) if str(g["world"]) == "mlp" else AttentionWorld()
    off, new = 0, {}
    for k, v in world.state_dict().items():
        new[k] = torch.from_numpy(g["world_weights"][off:off + v.numel()].reshape(tuple(v.shape)).copy())
        off += v.numel()
    world.load_state_dict(new)
    env = mcn.BatchedCrowdSim(T, H)
    env.set_state(np.stack([g["agents"][t] for t in range(T)]))
    world.predict_into(env)
    v = env.human_actions()
    assert np.max(np.abs(v - g["new_v"][:T])) <= 2e-6
    # the module call (what env.sim_world(x) did in the fork) is the same kernel
    x = torch.tensor(np.stack([g["agents"][t][1:, :4] for t in range(T)]), dtype=torch.float32, device="cuda").reshape(T, -1)
    out = world(x).cpu().numpy().reshape(T, H, 2)
    assert np.max(np.abs(out - g["new_v"][:T])) <= 2e-6
    # in-place parameter updates are picked up (version stamp)
    with torch.no_grad():
        list(world.parameters())[-1].add_(0.25)
    world.predict_into(env)
    v2 = env.human_actions()
    if str(g["world"]) == "mlp":
        assert np.max(np.abs(v2 - v)) > 1e-3
    else:
        assert np.max(np.abs(v2 - (v + 0.25))) <= 2e-6
    env.close()


def _training_golden():
    return dict(np.load(os.path.join(GOLDEN, "training.npz"), allow_pickle=False))


def test_update_memory_il_matches_reference(weights0):
    """Explorer.update_memory, imitation learning (explorer.py:153-166): the replay memory after 8 ORCA-robot episodes equals
    the one the REFERENCE's Explorer left behind for the same cases (tests/golden/training.npz, gen_golden.py --training):
    transformed states and discounted returns-to-go, entry by entry."""
    import torch
    import modelcrowdnav_b200 as mcn
    g = _training_golden()
    env, robot, policy, _ = _setup(weights0)
    il_policy = mcn.policy_factory["orca"]()
    il_policy.multiagent_training = policy.multiagent_training
    il_policy.safety_space = 0.15
    robot.set_policy(il_policy)
    memory = mcn.ReplayMemory(100000)
    explorer = mcn.Explorer(env, robot, torch.device("cuda:0"), memory, 0.9, target_policy=policy)
    res = explorer.run_k_episodes(8, "train", update_memory=True, imitation_learning=True)
    assert len(memory) == g["il_states"].shape[0]
    s = memory.states[:len(memory)].cpu().numpy()
    v = memory.values[:len(memory)].cpu().numpy().reshape(-1)
    assert np.max(np.abs(s - g["il_states"])) < 5e-6                  # fp32 rotate: atan2 / cos / sin of two libms
    assert np.max(np.abs(v - g["il_values"])) < 1e-6
    assert np.allclose(np.array(res, dtype=np.float64), g["il_result"], rtol=0, atol=1e-9)


def test_update_memory_rl_matches_reference():
    """Explorer.update_memory, reinforcement learning (explorer.py:167-174): value_i = r_i + gamma_bar * V_target(s_{i+1}),
    terminal = r, for the 8 'train' episodes the reference rolled out with the trained SARL (epsilon 0)."""
    import torch
    import modelcrowdnav_b200 as mcn
    g = _training_golden()
    env, robot, policy, _ = _setup(np.load(os.path.join(GOLDEN, "sarl_weights_trained.npy")))
    memory = mcn.ReplayMemory(100000)
    explorer = mcn.Explorer(env, robot, torch.device("cuda:0"), memory, 0.9, target_policy=policy)
    explorer.update_target_model(policy.get_model())
    policy.set_epsilon(0.0)
    res = explorer.run_k_episodes(8, "train", update_memory=True, episode=0, returnRate=False)
    assert len(memory) == g["rl_states"].shape[0]
    s = memory.states[:len(memory)].cpu().numpy()
    v = memory.values[:len(memory)].cpu().numpy().reshape(-1)
    assert np.max(np.abs(s - g["rl_states"])) < 5e-6
    assert np.max(np.abs(v - g["rl_values"]) / np.maximum(np.abs(g["rl_values"]), 0.1)) < 1e-5
    assert np.allclose(np.array(res, dtype=np.float64), g["rl_result"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("mode", ["eager", "graph", "fused"])
def test_trainer_matches_reference_sgd(weights0, mode):
    """Trainer (trainer.py:61-82): 100 SGD-momentum steps (lr 0.01, batch 100, MSE) on the batches the REFERENCE's DataLoader
    drew, from the same seed-0 network -> the reference's weights after the 100 steps, <= 1e-5."""
    import torch
    import modelcrowdnav_b200 as mcn
    from modelcrowdnav_b200.policy import make_value_network
    from modelcrowdnav_b200.trainer import Trainer
    g = _training_golden()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = make_value_network(13, 6, [150, 100], [100, 50], [150, 100, 100, 1], [100, 100, 1]).to(dev)
    flat0 = torch.cat([p.detach().reshape(-1) for p in model.state_dict().values()]).cpu().numpy()
    assert np.array_equal(flat0, weights0)
    memory = mcn.ReplayMemory(100000, device=dev)
    memory.push_batch(torch.from_numpy(g["il_states"]).to(dev), torch.from_numpy(g["il_values"]).to(dev))
    tr = Trainer(model, memory, dev, 100, mode=mode)
    tr.set_learning_rate(0.01)
    idx = torch.from_numpy(g["sgd_idx"].astype(np.int64)).to(dev)
    loss = 0.0
    for i in range(idx.shape[0]):
        loss += float(tr._step(idx[i]))
    tr._after()
    flat = torch.cat([p.detach().reshape(-1) for p in model.state_dict().values()]).cpu().numpy()
    assert abs(loss / idx.shape[0] - float(g["sgd_loss"])) < 1e-5
    assert np.max(np.abs(flat - g["sgd_weights"])) <= 1e-5, np.max(np.abs(flat - g["sgd_weights"]))


@pytest.mark.parametrize("B,H", [(100, 5), (37, 10), (1, 1), (64, 3)])
def test_fused_trainer_gradient_matches_autograd(weights0, B, H):
    """csrc/trainer.cu (forward + hand-written backward of sarl.py:28-65 with MSELoss) against torch autograd on the same
    batch: loss and every one of the 96,502 gradient entries."""
    import ctypes as C
    import torch
    from modelcrowdnav_b200.fused_trainer import FusedSarlTrainer
    from modelcrowdnav_b200.policy import make_value_network
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = make_value_network(13, 6, [150, 100], [100, 50], [150, 100, 100, 1], [100, 100, 1]).to(dev)
    g = torch.Generator(device="cpu").manual_seed(B * 100 + H)
    x = (torch.rand((B, H, 13), generator=g) * 4 - 2).to(dev)
    x[:, :, 2] = 0.0
    y = torch.rand((B, 1), generator=g).to(dev)
    loss = torch.nn.functional.mse_loss(model(x), y)
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in model.parameters()]).cpu().numpy()
    ft = FusedSarlTrainer(model, dev)
    ft.lr = 0.01
    ft.sync_from_model()
    lossd = torch.zeros((), device=dev)
    from modelcrowdnav_b200._capi import check
    check(ft.lib.cn_trainer_step(ft.handle, C.c_void_p(ft.flat.data_ptr()), C.c_void_p(x.data_ptr()),
                                 C.c_void_p(y.reshape(-1).contiguous().data_ptr()), B, H, 0.01, 0.9,
                                 C.c_void_p(ft.grad.data_ptr()), C.c_void_p(lossd.data_ptr()), None))
    torch.cuda.synchronize()
    got = ft.grad.cpu().numpy()
    assert abs(float(lossd) - float(loss)) <= 1e-6 * max(1.0, abs(float(loss)))
    scale = np.abs(ref).max()
    assert np.max(np.abs(got - ref)) <= 2e-5 * scale, (np.max(np.abs(got - ref)), scale)
    # the parameters are views of the flat block: a step moves model.state_dict() itself
    before = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()
    ft.step(x, y)
    after = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    assert torch.allclose(after, before - 0.01 * torch.from_numpy(ref).to(dev), atol=1e-6)
    ft.close()


def test_fused_trainer_indexed_step_matches_reference_sgd(weights0):
    """cn_trainer_step_indexed (the batch gathered from the replay tensors inside the kernel, loss accumulated on the device --
    the form Trainer.optimize_batch uses in `fused` mode): the reference Trainer's 100 steps, <= 1e-5, and its average loss."""
    import torch
    import modelcrowdnav_b200 as mcn
    from modelcrowdnav_b200.policy import make_value_network
    from modelcrowdnav_b200.trainer import Trainer
    g = _training_golden()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = make_value_network(13, 6, [150, 100], [100, 50], [150, 100, 100, 1], [100, 100, 1]).to(dev)
    memory = mcn.ReplayMemory(100000, device=dev)
    memory.push_batch(torch.from_numpy(g["il_states"]).to(dev), torch.from_numpy(g["il_values"]).to(dev))
    tr = Trainer(model, memory, dev, 100, mode="fused")
    tr.set_learning_rate(0.01)
    assert tr._fused_indexed_ok()
    idx = torch.from_numpy(g["sgd_idx"].astype(np.int64)).to(dev)
    loss = torch.zeros((), device=dev)
    for i in range(idx.shape[0]):
        tr._fused.step_indexed(memory.states, memory.values, idx[i], loss)
    tr._after()
    flat = torch.cat([p.detach().reshape(-1) for p in model.state_dict().values()]).cpu().numpy()
    assert abs(float(loss) / idx.shape[0] - float(g["sgd_loss"])) < 1e-5
    assert np.max(np.abs(flat - g["sgd_weights"])) <= 1e-5, np.max(np.abs(flat - g["sgd_weights"]))
    # and the public entry point runs on it (random batches: only sanity -- finite, decreasing on a fixed memory)
    l0 = tr.optimize_batch(20)
    l1 = tr.optimize_batch(200)
    assert np.isfinite(l0) and np.isfinite(l1) and l1 < l0

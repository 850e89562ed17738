"""Pin the C oracle against fixtures produced by the reference's own Python (scripts/gen_golden.py)."""
import os
import numpy as np
import pytest

from conftest import value_errors, GOLDEN, MODEL_WORLD_NAMES, load_model_world, TRAJ_NAMES_MIXED, TRAJ_NAMES_NETS, TRAJ_NAMES_OM, net_tag, TRAJ_NAMES, TRAJ_NAMES_KIN, load_traj, weights_for


def test_action_space_matches_reference(oracle_mod, units):
    tab = oracle_mod.action_space(1.0, 5, 16)
    assert tab.shape == (81, 2)
    assert np.array_equal(tab, units["action_space"])          # bit-exact (f64)
    # speeds and gamma_bar known answers
    assert np.allclose(np.hypot(*tab[1:6].T), [0.12885, 0.28623, 0.47845, 0.71324, 1.0], atol=1e-5)
    assert float(units["gamma_bar"]) == 0.9740037464252967


def test_weight_init_matches_reference(oracle_mod, weights0):
    assert weights0.size == 96502 == oracle_mod.sarl_param_count(oracle_mod.SarlCfg.default())
    assert np.array_equal(oracle_mod.default_sarl_weights(0), weights0)


def test_point_to_segment(oracle_mod, units):
    got = np.array([oracle_mod.lib().orc_point_to_segment_dist(*map(float, r)) for r in units["seg_in"]])
    assert np.array_equal(got, units["seg_out"])               # bit-exact (f64)


def test_rotate(oracle_mod, units):
    got = oracle_mod.rotate(units["rotate_in"])
    # torch f32 atan2/cos/sin vs glibc: a few ulp on values up to ~15
    assert np.max(np.abs(got - units["rotate_out"])) < 5e-6


def test_compute_reward(oracle_mod, units):
    rin = units["reward_in"]
    got = np.array([oracle_mod.compute_reward(r[:5], r[5:].reshape(5, 3)) for r in rin])
    assert np.array_equal(got, units["reward_out"])            # bit-exact (f64)
    assert {-0.25, 1.0, 0.0} <= set(np.unique(got))            # every ladder rung exercised


@pytest.mark.parametrize("H", [5, 10])
def test_value_network(oracle_mod, units, weights0, H):
    scfg = oracle_mod.SarlCfg.default()
    x = units["vnet_in_h%d" % H]
    got = np.array([oracle_mod.sarl_forward(scfg, weights0, xi) for xi in x])
    ref = units["vnet_out_h%d" % H]
    assert value_errors(got, ref, "f32") <= 1.0


@pytest.mark.parametrize("name", TRAJ_NAMES + TRAJ_NAMES_KIN + TRAJ_NAMES_MIXED)
def test_trajectories(oracle_mod, weights0, name):
    """Teacher-forced, step by step: ORCA velocities, outcome ladder, state update, 81 values, argmax."""
    o = oracle_mod
    tr = load_traj(name)
    weights0 = weights_for(name)
    H = tr["H"]
    ecfg = o.EnvCfg.default(robot_visible=tr["robot_visible"])
    scfg = o.SarlCfg.default()
    n_steps = 0
    kin = tr["kinematics"]
    for case, rec in tr["cases"].items():
        table = rec["table"]
        assert np.array_equal(table, o.action_space(1.0, 5, 16, kinematics=kin))       # bit-exact, any kinematics
        for t in range(len(rec["time"])):
            agents = np.ascontiguousarray(rec["agents"][t])
            gt = float(rec["time"][t])
            theta = float(rec["theta"][t]) if "theta" in rec else 0.0
            hv = o.human_actions(ecfg, agents)
            assert np.array_equal(hv, rec["human_v"][t]), (case, t)
            best, values, reached = o.lookahead(ecfg, scfg, weights0, agents, gt, table, tr["query_env"], hv,
                                                kinematics=kin, theta=theta)
            assert not reached
            ref_v = rec["values"][t]
            assert value_errors(values, ref_v, "f32") <= 1.0, (case, t)
            top2 = np.sort(ref_v)[-2:]
            if top2[1] - top2[0] > 1e-5:
                assert best == int(rec["best"][t]), (case, t)
            a = table[int(rec["best"][t])]
            r, done, info, dmin = o.step_outcome(ecfg, agents, gt, a, kinematics=kin, theta=theta)
            assert r == rec["reward"][t] and done == bool(rec["done"][t]) and info == int(rec["info"][t])
            if info == o.DANGER:
                assert dmin == rec["dmin"][t]
            if t + 1 < len(rec["time"]):
                nt, nth = o.apply_step(ecfg, agents, gt, a, hv, kinematics=kin, theta=theta)
                assert nt == rec["time"][t + 1]
                assert np.array_equal(agents, rec["agents"][t + 1]), (case, t)
                if "theta" in rec:
                    assert nth == rec["theta"][t + 1], (case, t)
            n_steps += 1
    assert n_steps >= 20


def test_scene_known_answers(oracle_mod):
    """SURVEY §8(c) golden scenes captured from the reference's own reset()."""
    s = oracle_mod.generate_scene("test", 0)
    exp = [(-2.6625559084662678, -2.8379852903266491), (-3.6025107218593906, 0.15897818498977118),
           (3.7670532713727254, 0.74515630301793467), (1.8871992410374889, -3.1111986546762234),
           (-3.4340226851763447, 2.7512881890275414)]
    assert np.array_equal(s[1:, :2], np.array(exp))
    assert np.array_equal(s[1:, 4:6], -np.array(exp))
    assert tuple(oracle_mod.generate_scene("test", 1)[1, :2]) == (-1.6189844599022485, 3.4489804011123173)
    assert tuple(oracle_mod.generate_scene("test", 499)[1, :2]) == (1.1864128465707466, -3.4844207477053897)
    sq = oracle_mod.generate_scene("test", 0, human_num=10, rule="square_crossing")
    assert tuple(sq[1, :2]) == (-0.57503471562202868, 4.5028286434902451)
    assert tuple(sq[1, 4:6]) == (2.4109570071399911, 3.7247453518203533)


@pytest.mark.parametrize("name", ["circle5_random", "square10_random"])
def test_randomized_scenes_match_reference(oracle_mod, name):
    """[env] randomize_attributes: the reference's own reset() states (first record of every fixture case) against
    both scene generators -- the oracle's and the product's host generator (same MT19937 stream, bit-exact)."""
    from modelcrowdnav_b200 import scenes
    tr = load_traj(name)
    assert tr["randomize"] == 1
    for case, rec in tr["cases"].items():
        phase, idx = case.rsplit("_", 1)
        ref = rec["agents"][0]
        got = oracle_mod.generate_scene(phase, int(idx), human_num=tr["H"], rule=tr["sim"], randomize=True)
        assert np.array_equal(got, ref), case
        prod = scenes.generate_scene(phase, int(idx), human_num=tr["H"], rule=tr["sim"], randomize_attributes=True)
        assert np.array_equal(prod, ref), case
        assert np.all((ref[1:, 6] >= 0.3) & (ref[1:, 6] <= 0.5)) and np.all((ref[1:, 7] >= 0.5) & (ref[1:, 7] <= 1.5))
        assert len(set(ref[1:, 6])) > 1


def _net_cfg(o, tag):
    return {"cadrl": o.NetCfg.cadrl(), "lstm": o.NetCfg.lstm_rl(), "lstm2": o.NetCfg.lstm_rl(True)}[tag]


@pytest.mark.parametrize("tag", ["cadrl", "lstm", "lstm2"])
def test_other_value_networks(oracle_mod, units_nets, tag):
    """CADRL mlp (cadrl.py:22-30), LSTM-RL ValueNetwork1 / ValueNetwork2 (lstm_rl.py:9-66) against the reference's
    torch modules; the seed-0 weights themselves are pinned too (state-dict order and RNG stream)."""
    o = oracle_mod
    ncfg = _net_cfg(o, tag)
    w = units_nets[tag + "_weights"]
    assert w.size == o.net_param_count(ncfg)
    assert np.array_equal(w, o.default_net_weights(ncfg, 0))
    for H in (1, 5):
        x, ref = units_nets["%s_in_h%d" % (tag, H)], units_nets["%s_out_h%d" % (tag, H)]
        got = np.array([o.net_forward(ncfg, w, xi) for xi in x])
        assert got.shape == ref.shape
        assert value_errors(got, ref, "f32") <= 1.0


def test_lstm_human_order_is_stable_descending(oracle_mod):
    """lstm_rl.py:99-104: sorted(key=dist, reverse=True) keeps the env order of equidistant humans."""
    o = oracle_mod
    a = o.generate_scene("test", 0)
    a[1, :2] = a[0, :2] + [3, 0]; a[2, :2] = a[0, :2] + [0, 1]; a[3, :2] = a[0, :2] + [0, -3]
    a[4, :2] = a[0, :2] + [5, 0]; a[5, :2] = a[0, :2] + [-1, 0]
    assert list(o.lstm_human_order(a)) == [3, 0, 2, 1, 4]


@pytest.mark.parametrize("name", TRAJ_NAMES_NETS)
def test_trajectories_other_networks(oracle_mod, units_nets, name):
    """CADRL.predict (min over humans) and LstmRL.predict (humans sorted by decreasing distance unless query_env) replayed
    step by step against the reference's own episodes: values, argmax, and the env transition under that action."""
    o = oracle_mod
    tr = load_traj(name)
    tag = net_tag(tr)
    ncfg, w = _net_cfg(o, tag), units_nets[tag + "_weights"]
    ecfg = o.EnvCfg.default()
    n = 0
    for case, rec in tr["cases"].items():
        table = rec["table"]
        for t in range(len(rec["time"])):
            agents = np.ascontiguousarray(rec["agents"][t])
            gt = float(rec["time"][t])
            hv = o.human_actions(ecfg, agents)
            assert np.array_equal(hv, rec["human_v"][t]), (case, t)
            theta = float(rec["theta"][t]) if "theta" in rec else 0.0
            best, values, reached = o.lookahead_net(ecfg, ncfg, w, agents, gt, table, tr["query_env"], hv,
                                                    kinematics=tr["kinematics"], theta=theta)
            assert not reached
            ref_v = rec["values"][t]
            assert value_errors(values, ref_v, "f32") <= 1.0, (case, t)
            top2 = np.sort(ref_v)[-2:]
            if top2[1] - top2[0] > 1e-5:
                assert best == int(rec["best"][t]), (case, t)
            n += 1
    assert n >= 20


def _om_cfgs(o, policy):
    """(network cfg with input_dim 61, OmCfg) of the reference's OM-SARL / OM-LSTM-RL."""
    om = o.OmCfg.default()
    if policy == "sarl":
        cfg = o.SarlCfg.default()
    else:
        cfg = o.NetCfg.lstm_rl()
    cfg.input_dim = 13 + om.dim
    return cfg, om


@pytest.mark.parametrize("H", [2, 5, 10])
def test_occupancy_maps(oracle_mod, units_om, H):
    """build_occupancy_maps (multi_human_rl.py:109-163) and transform() with maps against the reference, including humans
    at rest (arctan2(0, 0)) and humans outside every cell."""
    o = oracle_mod
    om = o.OmCfg.default()
    agents, maps, trs = units_om["om_agents_h%d" % H], units_om["om_maps_h%d" % H], units_om["om_transform_h%d" % H]
    occupied = 0
    for a, m, t in zip(agents, maps, trs):
        got = o.occupancy_maps(om, a[1:, :4])
        assert got.shape == m.shape and np.max(np.abs(got - m)) <= 1e-6
        assert np.array_equal(got[:, 0::3], m[:, 0::3])                 # occupancy channel exact
        occupied += int(m[:, 0::3].sum())
        gt = o.transform_om(om, a)
        assert np.max(np.abs(gt - t)) <= 1e-5
    assert occupied > 0


@pytest.mark.parametrize("name", TRAJ_NAMES_OM)
def test_trajectories_with_occupancy_maps(oracle_mod, units_om, name):
    """OM-SARL / OM-LSTM-RL predict() replayed against the reference's own episodes: the maps are built once per predict
    from the first action's next human states (multi_human_rl.py:47-49), in the network's human order."""
    o = oracle_mod
    tr = load_traj(name)
    assert tr["with_om"] == 1
    cfg, om = _om_cfgs(o, tr["policy"])
    w = units_om[("om_sarl" if tr["policy"] == "sarl" else "om_lstm") + "_weights"]
    ecfg = o.EnvCfg.default()
    n = 0
    for case, rec in tr["cases"].items():
        table = rec["table"]
        for t in range(len(rec["time"])):
            agents = np.ascontiguousarray(rec["agents"][t])
            hv = o.human_actions(ecfg, agents)
            assert np.array_equal(hv, rec["human_v"][t]), (case, t)
            best, values, reached = o.lookahead_om(ecfg, cfg, om, w, agents, float(rec["time"][t]), table, tr["query_env"], hv)
            assert not reached
            ref_v = rec["values"][t]
            assert value_errors(values, ref_v, "f32") <= 1.0, (case, t)
            top2 = np.sort(ref_v)[-2:]
            if top2[1] - top2[0] > 1e-5:
                assert best == int(rec["best"][t]), (case, t)
            n += 1
    assert n >= 15


def test_mixed_scenes_match_reference(oracle_mod):
    """[sim] mixed (crowd_sim.py:111-161): 80 scenes generated by the reference's reset() -- human counts 0..5, standing and
    moving humans, the dummy human of an empty static scene -- against the oracle's and the product's host generators."""
    import os
    from modelcrowdnav_b200 import scenes
    z = np.load(os.path.join(GOLDEN, "scenes_mixed.npz"))
    counts = set()
    for key in [k for k in z.files if not k.endswith("human_num")]:
        phase, case = key.split("_")
        ref = z[key]
        assert np.array_equal(oracle_mod.generate_scene(phase, int(case), human_num=5, rule="mixed"), ref), key
        assert np.array_equal(scenes.generate_scene(phase, int(case), human_num=5, rule="mixed"), ref), key
        n = int(z[key + "_human_num"])
        assert ref.shape[0] - 1 == max(n, 1)              # 0 humans -> one dummy human parked at (0, -10)
        counts.add(n)
    assert counts == {0, 1, 2, 3, 4, 5}


@pytest.mark.parametrize("name", MODEL_WORLD_NAMES)
def test_model_crowd_sim_fixtures(oracle_mod, weights0, name):
    """ModelCrowdSim (model_crowd_sim.py:268-441) run by the reference with world-model humans: the scene from the seeded
    GLOBAL numpy stream with initial velocities, then per step SARL's values, the outcome ladder and the state update under
    the recorded world-model velocities -- oracle (and the product's host scene generator) against the fixture."""
    from modelcrowdnav_b200 import scenes
    o = oracle_mod
    g = load_model_world(name)
    H, rule = int(g["H"]), str(g["sim"])
    state = np.random.get_state()
    try:
        np.random.seed(int(g["np_seed"]))
        a0 = o.generate_scene("test", 0, human_num=H, rule=rule, rs=np.random, init_velocity=True)
        np.random.seed(int(g["np_seed"]))
        a1 = scenes.generate_scene("test", 0, human_num=H, rule=rule, rs=np.random, init_velocity=True)
    finally:
        np.random.set_state(state)
    assert np.array_equal(a0, g["agents"][0]) and np.array_equal(a1, g["agents"][0])
    assert np.all(np.max(np.abs(a0[1:, 2:4]), axis=1) == 1.0)          # gen_init_v: larger component = v_pref
    ecfg, scfg = o.EnvCfg.default(), o.SarlCfg.default()
    table = g["table"]
    for t in range(len(g["reward"])):
        agents = np.ascontiguousarray(g["agents"][t])
        gt, hv = float(g["time"][t]), np.ascontiguousarray(g["new_v"][t])
        best, values, reached = o.lookahead(ecfg, scfg, weights0, agents, gt, table, int(g["query_env"]), hv)
        ref_v = g["values"][t]
        assert value_errors(values, ref_v, "f32") <= 1.0, t
        a = g["action"][t]
        r, done, info, dmin = o.step_outcome(ecfg, agents, gt, a)
        assert (r, done, info) == (g["reward"][t], bool(g["done"][t]), int(g["info"][t])), t
        nt = o.apply_step(ecfg, agents, gt, a, hv)
        assert nt == g["time"][t + 1] and np.array_equal(agents, g["agents"][t + 1]), t


@pytest.mark.parametrize("name", __import__("conftest").DECISIVE_NAMES)
def test_decisive_sets_pin_the_oracle(oracle_mod, name):
    """tests/golden/decisive_*.npz (the reference's trained SARL teacher-forced over all 500 test cases, scripts/gen_golden.py
    --decisive): the oracle reproduces the reference's 81 values and its argmax on a strided sample of the states (the GPU
    tests take every state).  This is the pin of the checker the scale tests and bench.py's parity sample rely on."""
    from conftest import check_argmax, load_decisive
    o = oracle_mod
    d = load_decisive(name)
    w = np.load(os.path.join(GOLDEN, "sarl_weights_trained.npy"))
    ecfg, scfg = o.EnvCfg.default(), o.SarlCfg.default()
    table = o.action_space(1.0, 5, 16)
    N = d["agents"].shape[0]
    idx = np.arange(0, N, max(1, N // 160))
    bests, refs = [], []
    for i in idx:
        agents = np.ascontiguousarray(d["agents"][i])
        hv = o.human_actions(ecfg, agents)
        best, values, reached = o.lookahead(ecfg, scfg, w, agents, float(d["time"][i]), table, int(d["query_env"]), hv)
        ref = d["values"][i].astype(np.float64)
        assert value_errors(values, ref, "f32") <= 1.0, i
        bests.append(best); refs.append(ref)
    check_argmax(bests, d["best"][idx], np.stack(refs), "f32", min_decidable=int(0.8 * len(idx)))

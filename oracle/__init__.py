"""CPU ORACLE (test infrastructure, NOT product code).

ctypes front-end of ``oracle/libcrowdnav_oracle.so`` (C restatement of the reference hot path,
see ``crowdnav_oracle.h``) plus the numpy scene generators.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package; ``modelcrowdnav_b200`` never does.

All file:line citations are relative to /root/reference.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcrowdnav_oracle.so")

NOTHING, DANGER, REACHGOAL, COLLISION, TIMEOUT = 0, 1, 2, 3, 4
AGENT_STRIDE = 8  # px py vx vy gx gy radius v_pref


def build(force=False):
    """Compile the C oracle in place (gcc, seconds)."""
    src = [os.path.join(_HERE, f) for f in ("crowdnav_oracle.c", "crowdnav_oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B", "libcrowdnav_oracle.so"], check=True,
                   stdout=subprocess.DEVNULL)
    return _LIB_PATH


class EnvCfg(C.Structure):
    """crowd_nav/configs/env.config + ORCA.__init__ (orca.py:55-67)."""
    _fields_ = [("time_limit", C.c_double), ("time_step", C.c_double),
                ("success_reward", C.c_double), ("collision_penalty", C.c_double),
                ("discomfort_dist", C.c_double), ("discomfort_penalty_factor", C.c_double),
                ("neighbor_dist", C.c_double), ("max_neighbors", C.c_int),
                ("time_horizon", C.c_double), ("human_safety_space", C.c_double),
                ("robot_visible", C.c_int)]

    @classmethod
    def default(cls, **kw):
        d = dict(time_limit=25, time_step=0.25, success_reward=1, collision_penalty=-0.25,
                 discomfort_dist=0.2, discomfort_penalty_factor=0.5, neighbor_dist=10,
                 max_neighbors=10, time_horizon=5, human_safety_space=0, robot_visible=0)
        d.update(kw)
        return cls(**d)


class SarlCfg(C.Structure):
    """crowd_nav/configs/policy.config [sarl]."""
    _fields_ = [("input_dim", C.c_int), ("self_state_dim", C.c_int), ("mlp1_dims", C.c_int * 2),
                ("mlp2_dims", C.c_int * 2), ("attn_dims", C.c_int * 3), ("mlp3_dims", C.c_int * 4)]

    @classmethod
    def default(cls):
        return cls(13, 6, (C.c_int * 2)(150, 100), (C.c_int * 2)(100, 50),
                   (C.c_int * 3)(100, 100, 1), (C.c_int * 4)(150, 100, 100, 1))


NET_SARL, NET_CADRL, NET_LSTM_RL = 0, 1, 2


class NetCfg(C.Structure):
    """crowd_nav/configs/policy.config [cadrl] / [lstm_rl] (orc_net_cfg)."""
    _fields_ = [("network", C.c_int), ("input_dim", C.c_int), ("self_state_dim", C.c_int), ("mlp_dims", C.c_int * 4),
                ("lstm_hidden", C.c_int), ("mlp1_dims", C.c_int * 4)]

    @classmethod
    def cadrl(cls):
        return cls(NET_CADRL, 13, 6, (C.c_int * 4)(150, 100, 100, 1), 0, (C.c_int * 4)(0, 0, 0, 0))

    @classmethod
    def lstm_rl(cls, with_interaction_module=False):
        m1 = (150, 100, 100, 50) if with_interaction_module else (0, 0, 0, 0)
        return cls(NET_LSTM_RL, 13, 6, (C.c_int * 4)(150, 100, 100, 1), 50, (C.c_int * 4)(*m1))


class OmCfg(C.Structure):
    """crowd_nav/configs/policy.config [om]."""
    _fields_ = [("cell_num", C.c_int), ("cell_size", C.c_double), ("om_channel_size", C.c_int)]

    @classmethod
    def default(cls):
        return cls(4, 1.0, 3)

    @property
    def dim(self):
        return self.cell_num ** 2 * self.om_channel_size


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp, fp, ip = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)
        L.orc_rvo_new_velocity.argtypes = [C.c_int, fp, fp, fp, fp, fp, C.c_int, C.c_float, C.c_float,
                                           C.c_float, C.c_float, C.c_int, C.c_float, C.c_float, fp, fp, ip]
        L.orc_rvo_do_step.argtypes = [C.c_int, fp, fp, fp, fp, fp, fp, fp, fp, fp, ip, fp, C.c_float]
        L.orc_point_to_segment_dist.argtypes = [C.c_double] * 6
        L.orc_point_to_segment_dist.restype = C.c_double
        L.orc_human_actions.argtypes = [C.POINTER(EnvCfg), C.c_int, dp, dp]
        L.orc_robot_orca_action.argtypes = [C.POINTER(EnvCfg), C.c_int, dp, C.c_double, dp]
        L.orc_step_outcome.argtypes = [C.POINTER(EnvCfg), C.c_int, dp, C.c_double, C.c_double,
                                       C.c_double, dp, ip, ip, dp]
        L.orc_apply_step.argtypes = [C.POINTER(EnvCfg), C.c_int, dp, dp, C.c_double, C.c_double, dp]
        L.orc_action_space.argtypes = [C.c_double, C.c_int, C.c_int, dp]
        L.orc_action_space.restype = C.c_int
        L.orc_rotate.argtypes = [fp, fp]
        L.orc_compute_reward.argtypes = [C.c_double] * 5 + [C.c_int, dp, dp, dp, C.c_double]
        L.orc_compute_reward.restype = C.c_double
        L.orc_sarl_param_count.argtypes = [C.POINTER(SarlCfg)]
        L.orc_sarl_param_count.restype = C.c_int64
        L.orc_sarl_forward.argtypes = [C.POINTER(SarlCfg), fp, C.c_int, fp, fp]
        L.orc_sarl_forward.restype = C.c_float
        L.orc_lookahead.argtypes = [C.POINTER(EnvCfg), C.POINTER(SarlCfg), fp, C.c_int, dp, C.c_double,
                                    C.c_int, dp, C.c_int, dp, C.c_double, dp, ip]
        L.orc_lookahead.restype = C.c_int
        L.orc_transform.argtypes = [C.c_int, dp, fp]
        L.orc_action_space_k.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, dp]
        L.orc_action_space_k.restype = C.c_int
        L.orc_rotate_k.argtypes = [fp, C.c_int, fp]
        L.orc_step_outcome_k.argtypes = [C.POINTER(EnvCfg), C.c_int, dp, C.c_double, C.c_int, C.c_double, C.c_double,
                                         C.c_double, dp, ip, ip, dp]
        L.orc_apply_step_k.argtypes = [C.POINTER(EnvCfg), C.c_int, dp, dp, C.c_int, dp, C.c_double, C.c_double, dp]
        L.orc_lookahead_k.argtypes = [C.POINTER(EnvCfg), C.POINTER(SarlCfg), fp, C.c_int, dp, C.c_double, C.c_int,
                                      C.c_double, C.c_int, dp, C.c_int, dp, C.c_double, dp, ip]
        L.orc_lookahead_k.restype = C.c_int
        L.orc_transform_k.argtypes = [C.c_int, dp, C.c_int, C.c_double, fp]
        L.orc_batch_lookahead_step.argtypes = [C.POINTER(EnvCfg), C.POINTER(SarlCfg), fp, C.c_int, C.c_int,
                                               dp, dp, C.c_int, dp, C.c_int, C.c_double,
                                               C.POINTER(C.c_int32), dp, C.POINTER(C.c_uint8),
                                               C.POINTER(C.c_uint8), C.c_int]
        L.orc_net_param_count.argtypes = [C.POINTER(NetCfg)]
        L.orc_net_param_count.restype = C.c_int64
        L.orc_net_forward.argtypes = [C.POINTER(NetCfg), fp, C.c_int, fp, fp]
        L.orc_lookahead_net.argtypes = [C.POINTER(EnvCfg), C.POINTER(NetCfg), fp, C.c_int, dp, C.c_double, C.c_int,
                                        C.c_double, C.c_int, dp, C.c_int, dp, C.c_double, dp, ip]
        L.orc_lookahead_net.restype = C.c_int
        L.orc_lstm_human_order.argtypes = [C.c_int, dp, ip]
        L.orc_occupancy_maps.argtypes = [C.POINTER(OmCfg), C.c_int, dp, fp]
        L.orc_lookahead_om.argtypes = [C.POINTER(EnvCfg), C.c_int, C.POINTER(SarlCfg), C.POINTER(NetCfg), C.POINTER(OmCfg),
                                       fp, C.c_int, dp, C.c_double, C.c_int, C.c_double, C.c_int, dp, C.c_int, dp,
                                       C.c_double, dp, ip]
        L.orc_lookahead_om.restype = C.c_int
        L.orc_transform_om.argtypes = [C.POINTER(OmCfg), C.c_int, dp, ip, C.c_int, C.c_double, fp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# --------------------------------------------------------------------------------------------
# thin numpy wrappers (one env at a time, like the reference)
# --------------------------------------------------------------------------------------------

def rvo_new_velocity(pos, vel, radius, self_idx, max_speed, pref, neighbor_dist=10.0, max_neighbors=10,
                     time_horizon=5.0, time_step=0.25, return_info=False):
    """Agent::computeNeighbors + computeNewVelocity of agent ``self_idx`` (float32)."""
    pos = _f32(pos); vel = _f32(vel); radius = _f32(radius)
    n = pos.shape[0]
    px, py = np.ascontiguousarray(pos[:, 0]), np.ascontiguousarray(pos[:, 1])
    vx, vy = np.ascontiguousarray(vel[:, 0]), np.ascontiguousarray(vel[:, 1])
    ox, oy = C.c_float(), C.c_float()
    info = (C.c_int * 2)()
    lib().orc_rvo_new_velocity(n, _fp(px), _fp(py), _fp(vx), _fp(vy), _fp(radius), self_idx,
                               np.float32(max_speed), np.float32(pref[0]), np.float32(pref[1]),
                               np.float32(neighbor_dist), max_neighbors, np.float32(time_horizon),
                               np.float32(time_step), C.byref(ox), C.byref(oy), info)
    out = np.array([ox.value, oy.value], dtype=np.float32)
    return (out, (info[0], info[1])) if return_info else out


def human_actions(cfg, agents):
    agents = _f64(agents)
    H = agents.shape[0] - 1
    out = np.empty((H, 2), np.float64)
    lib().orc_human_actions(C.byref(cfg), H, _dp(agents), _dp(out))
    return out


def robot_orca_action(cfg, agents, safety_space):
    agents = _f64(agents)
    out = np.empty(2, np.float64)
    lib().orc_robot_orca_action(C.byref(cfg), agents.shape[0] - 1, _dp(agents), float(safety_space), _dp(out))
    return out


# robot kinematics (crowdnav_oracle.h): HOLONOMIC = ActionXY; UNICYCLE = ActionRot + heading feature; NONE = the fork's
# literal behaviour (cadrl.py:66 commented out): ActionRot dynamics, heading feature left at zero
KIN_HOLONOMIC, KIN_UNICYCLE, KIN_NONE = 0, 1, 2


def step_outcome(cfg, agents, global_time, action, kinematics=KIN_HOLONOMIC, theta=0.0):
    """action = (vx, vy) holonomic, (v, r) otherwise; theta = robot heading before the step."""
    agents = _f64(agents)
    r, dmin = C.c_double(), C.c_double()
    done, info = C.c_int(), C.c_int()
    lib().orc_step_outcome_k(C.byref(cfg), agents.shape[0] - 1, _dp(agents), float(global_time), int(kinematics),
                             float(theta), float(action[0]), float(action[1]), C.byref(r), C.byref(done),
                             C.byref(info), C.byref(dmin))
    return r.value, bool(done.value), info.value, dmin.value


def apply_step(cfg, agents, global_time, action, human_vxy, kinematics=KIN_HOLONOMIC, theta=None):
    """In-place update of ``agents``; returns the new global time (and the new heading when ``theta`` is given)."""
    assert agents.dtype == np.float64 and agents.flags.c_contiguous
    t = C.c_double(float(global_time))
    th = C.c_double(float(theta if theta is not None else 0.0))
    hv = _f64(human_vxy)
    lib().orc_apply_step_k(C.byref(cfg), agents.shape[0] - 1, _dp(agents), C.byref(t), int(kinematics), C.byref(th),
                           float(action[0]), float(action[1]), _dp(hv))
    return t.value if theta is None else (t.value, th.value)


def action_space(v_pref=1.0, speed_samples=5, rotation_samples=16, kinematics=KIN_HOLONOMIC):
    out = np.empty((speed_samples * rotation_samples + 1, 2), np.float64)
    n = lib().orc_action_space_k(float(v_pref), speed_samples, rotation_samples, int(kinematics), _dp(out))
    assert n == out.shape[0]
    return out


def rotate(rows14, kinematics=KIN_HOLONOMIC):
    rows14 = _f32(rows14).reshape(-1, 14)
    out = np.empty((rows14.shape[0], 13), np.float32)
    for i in range(rows14.shape[0]):
        lib().orc_rotate_k(_fp(rows14[i]), int(kinematics), _fp(out[i]))
    return out


def compute_reward(nav, humans, time_step=0.25):
    """nav = (px, py, radius, gx, gy); humans = (H,3) px py radius."""
    humans = _f64(humans)
    hx, hy, hr = (np.ascontiguousarray(humans[:, i]) for i in range(3))
    return lib().orc_compute_reward(*[float(v) for v in nav], humans.shape[0], _dp(hx), _dp(hy), _dp(hr),
                                    float(time_step))


def sarl_param_count(scfg):
    return int(lib().orc_sarl_param_count(C.byref(scfg)))


def sarl_forward(scfg, weights, x, return_attention=False):
    x = _f32(x)
    weights = _f32(weights)
    assert weights.size == sarl_param_count(scfg)
    attn = np.empty(x.shape[0], np.float32)
    v = lib().orc_sarl_forward(C.byref(scfg), _fp(weights), x.shape[0], _fp(x), _fp(attn))
    return (float(v), attn) if return_attention else float(v)


def lookahead(ecfg, scfg, weights, agents, global_time, actions, query_env, human_vxy, gamma=0.9,
              kinematics=KIN_HOLONOMIC, theta=0.0):
    agents = _f64(agents); actions = _f64(actions); weights = _f32(weights)
    hv = _f64(human_vxy if human_vxy is not None else np.zeros((agents.shape[0] - 1, 2)))
    values = np.full(actions.shape[0], np.nan)
    reached = C.c_int()
    best = lib().orc_lookahead_k(C.byref(ecfg), C.byref(scfg), _fp(weights), agents.shape[0] - 1, _dp(agents),
                                 float(global_time), int(kinematics), float(theta), actions.shape[0], _dp(actions),
                                 int(query_env), _dp(hv), float(gamma), _dp(values), C.byref(reached))
    return best, values, bool(reached.value)


def net_param_count(ncfg):
    return int(lib().orc_net_param_count(C.byref(ncfg)))


def net_forward(ncfg, weights, x):
    """model(x) for x (H, 13): CADRL -> (H,) values, LSTM-RL -> scalar."""
    x = _f32(x); weights = _f32(weights)
    assert weights.size == net_param_count(ncfg)
    out = np.zeros(max(1, x.shape[0]), np.float32)
    lib().orc_net_forward(C.byref(ncfg), _fp(weights), x.shape[0], _fp(x), _fp(out))
    return out if ncfg.network == NET_CADRL else float(out[0])


def lookahead_net(ecfg, ncfg, weights, agents, global_time, actions, query_env, human_vxy, gamma=0.9,
                  kinematics=KIN_HOLONOMIC, theta=0.0):
    """CADRL.predict / LstmRL.predict greedy branch (same return convention as lookahead)."""
    agents = _f64(agents); actions = _f64(actions); weights = _f32(weights)
    hv = _f64(human_vxy if human_vxy is not None else np.zeros((agents.shape[0] - 1, 2)))
    values = np.full(actions.shape[0], np.nan)
    reached = C.c_int()
    best = lib().orc_lookahead_net(C.byref(ecfg), C.byref(ncfg), _fp(weights), agents.shape[0] - 1, _dp(agents),
                                   float(global_time), int(kinematics), float(theta), actions.shape[0], _dp(actions),
                                   int(query_env), _dp(hv), float(gamma), _dp(values), C.byref(reached))
    return best, values, bool(reached.value)


def lstm_human_order(agents):
    agents = _f64(agents)
    order = np.zeros(agents.shape[0] - 1, np.int32)
    lib().orc_lstm_human_order(agents.shape[0] - 1, _dp(agents), order.ctypes.data_as(C.POINTER(C.c_int)))
    return order


def occupancy_maps(om, humans):
    """build_occupancy_maps (multi_human_rl.py:109-163); humans (H, 4) = px py vx vy in network order -> (H, om.dim) f32."""
    humans = _f64(humans)
    out = np.zeros((humans.shape[0], om.dim), np.float32)
    lib().orc_occupancy_maps(C.byref(om), humans.shape[0], _dp(humans), _fp(out))
    return out


def lookahead_om(ecfg, cfg, om, weights, agents, global_time, actions, query_env, human_vxy, gamma=0.9,
                 kinematics=KIN_HOLONOMIC, theta=0.0):
    """Lookahead with occupancy maps; cfg is a SarlCfg (SARL) or a NetCfg (LSTM-RL) whose input_dim = 13 + om.dim."""
    agents = _f64(agents); actions = _f64(actions); weights = _f32(weights)
    hv = _f64(human_vxy if human_vxy is not None else np.zeros((agents.shape[0] - 1, 2)))
    values = np.full(actions.shape[0], np.nan)
    reached = C.c_int()
    is_sarl = isinstance(cfg, SarlCfg)
    assert cfg.input_dim == 13 + om.dim
    best = lib().orc_lookahead_om(C.byref(ecfg), NET_SARL if is_sarl else cfg.network, C.byref(cfg) if is_sarl else None,
                                  None if is_sarl else C.byref(cfg), C.byref(om), _fp(weights), agents.shape[0] - 1,
                                  _dp(agents), float(global_time), int(kinematics), float(theta), actions.shape[0],
                                  _dp(actions), int(query_env), _dp(hv), float(gamma), _dp(values), C.byref(reached))
    return best, values, bool(reached.value)


def transform_om(om, agents, order=None, kinematics=KIN_HOLONOMIC, theta=0.0):
    agents = _f64(agents)
    H = agents.shape[0] - 1
    out = np.empty((H, 13 + om.dim), np.float32)
    op = None
    if order is not None:
        order = np.ascontiguousarray(order, dtype=np.int32)
        op = order.ctypes.data_as(C.POINTER(C.c_int))
    lib().orc_transform_om(C.byref(om), H, _dp(agents), op, int(kinematics), float(theta), _fp(out))
    return out


def default_net_weights(ncfg, seed=0):
    """Flat f32 parameters (state-dict order) of the reference's CADRL / LSTM-RL ValueNetwork under torch.manual_seed(seed):
    the construction order (and hence the RNG stream) of cadrl.py:22-26 / lstm_rl.py:9-16,37-45."""
    import torch
    import torch.nn as nn
    torch.manual_seed(seed)

    def mlp(i, dims):
        d = [i] + list(dims)
        return [nn.Linear(d[k], d[k + 1]) for k in range(len(d) - 1)]
    flat = []
    if ncfg.network == NET_CADRL:
        mods = mlp(ncfg.input_dim, ncfg.mlp_dims)
    else:
        m1 = mlp(ncfg.input_dim, ncfg.mlp1_dims) if ncfg.mlp1_dims[0] > 0 else []
        m = mlp(ncfg.self_state_dim + ncfg.lstm_hidden, ncfg.mlp_dims)
        lstm = nn.LSTM(ncfg.mlp1_dims[3] if m1 else ncfg.input_dim, ncfg.lstm_hidden, batch_first=True)
        mods = m1 + m
    for l in mods:
        flat += [l.weight.detach().numpy().ravel(), l.bias.detach().numpy().ravel()]
    if ncfg.network == NET_LSTM_RL:
        flat += [t.detach().numpy().ravel() for t in (lstm.weight_ih_l0, lstm.weight_hh_l0, lstm.bias_ih_l0, lstm.bias_hh_l0)]
    return np.concatenate(flat).astype(np.float32)


def transform(agents, kinematics=KIN_HOLONOMIC, theta=0.0):
    agents = _f64(agents)
    H = agents.shape[0] - 1
    out = np.empty((H, 13), np.float32)
    lib().orc_transform_k(H, _dp(agents), int(kinematics), float(theta), _fp(out))
    return out


def batch_lookahead_step(ecfg, scfg, weights, agents, global_time, actions, query_env, done, gamma=0.9,
                         n_threads=1):
    """In place on ``agents`` (E,H+1,8) f64, ``global_time`` (E,) f64, ``done`` (E,) u8."""
    E, A1, _ = agents.shape
    assert agents.dtype == np.float64 and agents.flags.c_contiguous
    assert global_time.dtype == np.float64 and done.dtype == np.uint8
    weights = _f32(weights); actions = _f64(actions)
    idx = np.zeros(E, np.int32); reward = np.zeros(E); info = np.zeros(E, np.uint8)
    lib().orc_batch_lookahead_step(C.byref(ecfg), C.byref(scfg), _fp(weights), E, A1 - 1, _dp(agents),
                                   _dp(global_time), actions.shape[0], _dp(actions), int(query_env),
                                   float(gamma), idx.ctypes.data_as(C.POINTER(C.c_int32)), _dp(reward),
                                   done.ctypes.data_as(C.POINTER(C.c_uint8)),
                                   info.ctypes.data_as(C.POINTER(C.c_uint8)), int(n_threads))
    return idx, reward, info


# --------------------------------------------------------------------------------------------
# scene generation (CrowdSim.reset, crowd_sim.py:165-217, 261-323) -- numpy legacy MT19937
# --------------------------------------------------------------------------------------------

COUNTER_OFFSET = {"train": 2000, "val": 0, "test": 1000}  # crowd_sim.py:68,282-283


def _norm2(a, b):
    return float(np.linalg.norm((a, b)))


def generate_scene(phase, case, human_num=5, rule="circle_crossing", circle_radius=4.0, square_width=10.0,
                   radius=0.3, v_pref=1.0, discomfort_dist=0.2, robot_radius=0.3, robot_v_pref=1.0, randomize=False,
                   rs=None, init_velocity=False):
    """Agents (H+1, 8) f64 for (phase, case); agent 0 = robot at (0,-R) -> (0,R).
    randomize = [env] randomize_attributes: every human first draws v_pref ~ U(0.5, 1.5) and radius ~ U(0.3, 0.5)
    (crowd_sim.py:167-168,190-191; agent.py:39-45), one legacy-uniform draw each."""
    # rs / init_velocity: ModelCrowdSim.reset draws from the unseeded global stream (model_crowd_sim.py:293) and starts the
    # humans towards (-px, -py) with the larger velocity component equal to v_pref (gen_init_v, :186-192,222)
    if rs is None:
        rs = np.random.RandomState(COUNTER_OFFSET[phase] + case)  # crowd_sim.py:286
    base_radius, base_v_pref = radius, v_pref

    def init_v(px, py):
        vx, vy = -px - px, -py - py
        vmax = abs(vx)
        if vmax < abs(vy):
            vmax = abs(vy)
        return v_pref * vx / vmax, v_pref * vy / vmax
    robot_row = [0, -circle_radius, 0, 0, 0, circle_radius, robot_radius, robot_v_pref]  # crowd_sim.py:284
    per_human_rule = None
    if rule == "mixed":                                            # crowd_sim.py:111-161
        static_num = {0: 0.05, 1: 0.2, 2: 0.2, 3: 0.3, 4: 0.1, 5: 0.15}
        dynamic_num = {1: 0.3, 2: 0.3, 3: 0.2, 4: 0.1, 5: 0.1}
        static = rs.random_sample() < 0.2
        prob = rs.random_sample()
        for key, value in sorted(static_num.items() if static else dynamic_num.items()):
            if prob - value <= 0:
                human_num = key
                break
            else:
                prob -= value
        if static:                                                 # standing humans in a 4 x 8 box, goal = position
            rows = [robot_row]
            if human_num == 0:
                rows.append([0, -10, 0, 0, 0, -10, base_radius, base_v_pref])
            for _ in range(human_num):
                sign = -1 if rs.random_sample() > 0.5 else 1
                while True:
                    px = rs.random_sample() * 4 * 0.5 * sign
                    py = (rs.random_sample() - 0.5) * 8
                    if not any(_norm2(px - a[0], py - a[1]) < base_radius + a[6] + discomfort_dist for a in rows):
                        break
                rows.append([px, py, 0, 0, px, py, base_radius, base_v_pref])
            return np.array(rows, dtype=np.float64)
        per_human_rule = ["circle_crossing" if i < 2 else "square_crossing" for i in range(human_num)]
    agents = np.zeros((human_num + 1, AGENT_STRIDE))
    agents[0] = robot_row
    for i in range(1, human_num + 1):
        radius, v_pref = base_radius, base_v_pref
        if per_human_rule is not None:
            rule = per_human_rule[i - 1]
        if randomize:
            v_pref = rs.uniform(0.5, 1.5)
            radius = rs.uniform(0.3, 0.5)
        if rule == "circle_crossing":                              # crowd_sim.py:165-186
            while True:
                angle = rs.random_sample() * np.pi * 2
                px_noise = (rs.random_sample() - 0.5) * v_pref
                py_noise = (rs.random_sample() - 0.5) * v_pref
                px = circle_radius * np.cos(angle) + px_noise
                py = circle_radius * np.sin(angle) + py_noise
                collide = False
                for a in agents[:i]:
                    min_dist = radius + a[6] + discomfort_dist
                    if _norm2(px - a[0], py - a[1]) < min_dist or _norm2(px - a[4], py - a[5]) < min_dist:
                        collide = True
                        break
                if not collide:
                    break
            agents[i] = [px, py, 0, 0, -px, -py, radius, v_pref]
            if init_velocity:
                agents[i, 2:4] = init_v(px, py)
        elif rule == "square_crossing":                            # crowd_sim.py:188-217
            sign = -1 if rs.random_sample() > 0.5 else 1
            while True:
                px = rs.random_sample() * square_width * 0.5 * sign
                py = (rs.random_sample() - 0.5) * square_width
                if not any(_norm2(px - a[0], py - a[1]) < radius + a[6] + discomfort_dist for a in agents[:i]):
                    break
            while True:
                gx = rs.random_sample() * square_width * 0.5 * -sign
                gy = (rs.random_sample() - 0.5) * square_width
                if not any(_norm2(gx - a[4], gy - a[5]) < radius + a[6] + discomfort_dist for a in agents[:i]):
                    break
            agents[i] = [px, py, 0, 0, gx, gy, radius, v_pref]
            if init_velocity:
                agents[i, 2:4] = init_v(px, py)
        else:
            raise ValueError("Rule doesn't exist")
    return agents


def default_sarl_weights(seed=0, scfg=None):
    """Flat f32 parameters in state-dict order, default nn.Linear init under torch.manual_seed(seed).

    Uses torch only as an RNG/initialiser so that tests, bench and the reference goldens share weights.
    """
    import torch
    import torch.nn as nn
    scfg = scfg or SarlCfg.default()
    torch.manual_seed(seed)

    def mlp(i, dims):
        layers, d = [], [i] + list(dims)
        for k in range(len(d) - 1):
            layers.append(nn.Linear(d[k], d[k + 1]))
        return layers
    m1 = mlp(scfg.input_dim, scfg.mlp1_dims)
    m2 = mlp(scfg.mlp1_dims[1], scfg.mlp2_dims)
    at = mlp(scfg.mlp1_dims[1] * 2, scfg.attn_dims)
    m3 = mlp(scfg.mlp2_dims[1] + scfg.self_state_dim, scfg.mlp3_dims)
    flat = []
    for l in m1 + m2 + at + m3:
        flat += [l.weight.detach().numpy().ravel(), l.bias.detach().numpy().ravel()]
    return np.concatenate(flat).astype(np.float32)

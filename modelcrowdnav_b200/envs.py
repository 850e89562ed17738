"""Host-side mirror of the reference environment interface (same names, argument meaning and errors):

    ActionXY / ActionRot                       crowd_sim/envs/utils/action.py:3-4
    ObservableState / FullState / JointState   crowd_sim/envs/utils/state.py:1-55
    Timeout / ReachGoal / Danger / Collision / Nothing     crowd_sim/envs/utils/info.py:1-38
    Agent / Robot / Human                      crowd_sim/envs/utils/{agent,robot,human}.py
    CrowdSim                                   crowd_sim/envs/crowd_sim.py:15-434 (reset / step / onestep_lookahead)

`CrowdSim` is the gym-style single-environment façade: a view on a one-env batch that lives on the GPU
(`BatchedCrowdSim`).  Every call goes through the C ABI; there is no Python re-implementation of the step.
Throughput code should use `BatchedCrowdSim` / `Explorer` directly -- the façade synchronises on every call.
"""
import logging
from collections import namedtuple

import numpy as np

from . import _capi, scenes
from .batch import BatchedCrowdSim

ActionXY = namedtuple("ActionXY", ["vx", "vy"])
ActionRot = namedtuple("ActionRot", ["v", "r"])
# agent.kinematics -> CN_KIN_*; None = the fork exactly as shipped (every `== 'holonomic'` test fails, agent.py:104-133)
KIN_CODE = {"holonomic": _capi.KIN_HOLONOMIC, "unicycle": _capi.KIN_UNICYCLE, None: _capi.KIN_NONE}


class FullState(object):
    def __init__(self, px, py, vx, vy, radius, gx, gy, v_pref, theta):
        self.px, self.py, self.vx, self.vy, self.radius = px, py, vx, vy, radius
        self.gx, self.gy, self.v_pref, self.theta = gx, gy, v_pref, theta
        self.position = (self.px, self.py)
        self.goal_position = (self.gx, self.gy)
        self.velocity = (self.vx, self.vy)

    def __add__(self, other):
        return other + (self.px, self.py, self.vx, self.vy, self.radius, self.gx, self.gy, self.v_pref, self.theta)

    def __str__(self):
        return " ".join(str(x) for x in [self.px, self.py, self.vx, self.vy, self.radius, self.gx, self.gy,
                                         self.v_pref, self.theta])


class ObservableState(object):
    def __init__(self, px, py, vx, vy, radius):
        self.px, self.py, self.vx, self.vy, self.radius = px, py, vx, vy, radius
        self.position = (self.px, self.py)
        self.velocity = (self.vx, self.vy)

    def __add__(self, other):
        return other + (self.px, self.py, self.vx, self.vy, self.radius)

    def __str__(self):
        return " ".join(str(x) for x in [self.px, self.py, self.vx, self.vy, self.radius])

    def getvalue(self):
        return [self.px, self.py, self.vx, self.vy]

    def getvel(self):
        return [self.vx, self.vy]


class JointState(object):
    def __init__(self, self_state, human_states):
        assert isinstance(self_state, FullState)
        for human_state in human_states:
            assert isinstance(human_state, ObservableState)
        self.self_state = self_state
        self.human_states = human_states


class Timeout(object):
    def __str__(self):
        return "Timeout"


class ReachGoal(object):
    def __str__(self):
        return "Reaching goal"


class Danger(object):
    def __init__(self, min_dist):
        self.min_dist = min_dist

    def __str__(self):
        return "Too close"


class Collision(object):
    def __str__(self):
        return "Collision"


class Nothing(object):
    def __str__(self):
        return ""


def info_from_code(code, dmin):
    """CN_* info code (include/crowdnav_b200.h) -> the reference's Info object."""
    code = int(code)
    if code == _capi.TIMEOUT:
        return Timeout()
    if code == _capi.COLLISION:
        return Collision()
    if code == _capi.REACHGOAL:
        return ReachGoal()
    if code == _capi.DANGER:
        return Danger(float(dmin))
    return Nothing()


class Agent(object):
    """Physical attributes + state accessors of a robot / human (agent.py:10-138); the kinematic update
    itself runs on the GPU."""

    def __init__(self, config, section):
        from .policy import policy_factory
        self.visible = config.getboolean(section, "visible")
        self.v_pref = config.getfloat(section, "v_pref")
        self.radius = config.getfloat(section, "radius")
        self.policy = policy_factory[config.get(section, "policy")]()
        self.sensor = config.get(section, "sensor")
        self.kinematics = self.policy.kinematics if self.policy is not None else None
        self.px = self.py = self.gx = self.gy = self.vx = self.vy = self.theta = None
        self.time_step = None

    def print_info(self):
        logging.info("Agent is {} and has {} kinematic constraint".format(
            "visible" if self.visible else "invisible", self.kinematics))

    def set_policy(self, policy):
        self.policy = policy
        self.kinematics = policy.kinematics

    def set(self, px, py, gx, gy, vx, vy, theta, radius=None, v_pref=None):
        self.px, self.py, self.gx, self.gy, self.vx, self.vy, self.theta = px, py, gx, gy, vx, vy, theta
        if radius is not None:
            self.radius = radius
        if v_pref is not None:
            self.v_pref = v_pref

    def get_observable_state(self):
        return ObservableState(self.px, self.py, self.vx, self.vy, self.radius)

    def get_full_state(self):
        return FullState(self.px, self.py, self.vx, self.vy, self.radius, self.gx, self.gy, self.v_pref, self.theta)

    def get_position(self):
        return self.px, self.py

    def get_goal_position(self):
        return self.gx, self.gy

    def get_velocity(self):
        return self.vx, self.vy

    def check_validity(self, action):
        if self.kinematics == "holonomic":
            assert isinstance(action, ActionXY)
        else:
            assert isinstance(action, ActionRot)

    def reached_destination(self):
        return np.linalg.norm(np.array(self.get_position()) - np.array(self.get_goal_position())) < self.radius

    def _row(self):
        return [self.px, self.py, self.vx, self.vy, self.gx, self.gy, self.radius, self.v_pref]

    def _load_row(self, row):
        self.px, self.py, self.vx, self.vy, self.gx, self.gy = (float(x) for x in row[:6])


class Human(Agent):
    def __init__(self, config, section):
        super().__init__(config, section)

    def act(self, ob):
        state = JointState(self.get_full_state(), ob)
        return self.policy.predict(state)


class Robot(Agent):
    def __init__(self, config, section):
        super().__init__(config, section)

    def act(self, ob):
        if self.policy is None:
            raise AttributeError("Policy attribute has to be set!")
        state = JointState(self.get_full_state(), ob)
        return self.policy.predict(state)


def batch_env_kwargs(env):
    """cn_env_cfg fields from a configured CrowdSim façade (crowd_sim.py:58-89, orca.py:55-67)."""
    cfg = env.config
    return dict(time_limit=float(env.time_limit), time_step=env.time_step, success_reward=env.success_reward,
                collision_penalty=env.collision_penalty, discomfort_dist=env.discomfort_dist,
                discomfort_penalty_factor=env.discomfort_penalty_factor,
                robot_visible=int(env.robot.visible) if env.robot is not None else 0,
                circle_radius=env.circle_radius, square_width=env.square_width,
                human_radius=cfg.getfloat("humans", "radius"), human_v_pref=cfg.getfloat("humans", "v_pref"),
                robot_radius=cfg.getfloat("robot", "radius"), robot_v_pref=cfg.getfloat("robot", "v_pref"),
                randomize_attributes=int(bool(env.randomize_attributes)),
                robot_kinematics=KIN_CODE[env.robot.kinematics] if env.robot is not None else _capi.KIN_HOLONOMIC)


class CrowdSim(object):
    """Drop-in for gym.make('CrowdSim-v0') (crowd_sim/envs/crowd_sim.py:15-434), backed by the CUDA library."""
    metadata = {"render.modes": ["human"]}

    def __init__(self):
        self.time_limit = self.time_step = self.robot = self.humans = self.global_time = self.human_times = None
        self.success_reward = self.collision_penalty = self.discomfort_dist = self.discomfort_penalty_factor = None
        self.config = self.case_capacity = self.case_size = self.case_counter = self.randomize_attributes = None
        self.train_val_sim = self.test_sim = self.square_width = self.circle_radius = self.human_num = None
        self.states = self.action_values = self.attention_weights = None
        self.device = 0
        self._batch = None
        self._human_v = None

    def configure(self, config):
        self.config = config
        self.time_limit = config.getint("env", "time_limit")
        self.time_step = config.getfloat("env", "time_step")
        self.randomize_attributes = config.getboolean("env", "randomize_attributes")
        self.success_reward = config.getfloat("reward", "success_reward")
        self.collision_penalty = config.getfloat("reward", "collision_penalty")
        self.discomfort_dist = config.getfloat("reward", "discomfort_dist")
        self.discomfort_penalty_factor = config.getfloat("reward", "discomfort_penalty_factor")
        if config.get("humans", "policy") == "orca":
            self.case_capacity = {"train": np.iinfo(np.uint32).max - 2000, "val": 1000, "test": 1000}
            self.case_size = {"train": 100, "val": config.getint("env", "val_size"),
                              "test": config.getint("env", "test_size")}          # crowd_sim.py:71 (fork: train = 100)
            self.train_val_sim = config.get("sim", "train_val_sim")
            self.test_sim = config.get("sim", "test_sim")
            self.square_width = config.getfloat("sim", "square_width")
            self.circle_radius = config.getfloat("sim", "circle_radius")
            self.human_num = config.getint("sim", "human_num")
        else:
            raise NotImplementedError
        self.case_counter = {"train": 0, "test": 0, "val": 0}
        logging.info("human number: {}".format(self.human_num))
        if self.randomize_attributes:
            logging.info("Randomize human's radius and preferred speed")
        else:
            logging.info("Not randomize human's radius and preferred speed")
        logging.info("Training simulation: {}, test simulation: {}".format(self.train_val_sim, self.test_sim))
        logging.info("Square width: {}, circle width: {}".format(self.square_width, self.circle_radius))

    def set_robot(self, robot):
        self.robot = robot

    # -- helpers -----------------------------------------------------------------------------------
    def phase_human_num(self, phase):
        """crowd_sim.py:272-275,288-292: a policy trained on single humans (CADRL, multiagent_training = false) sees ONE
        circle-crossing human in the train / val phases."""
        if phase != "test" and not self.robot.policy.multiagent_training:
            self.train_val_sim = "circle_crossing"                     # crowd_sim.py:276-277 (mutates, like the reference)
            return 1
        return self.human_num

    def scene_kwargs(self, phase):
        human_num = self.phase_human_num(phase)
        rule = self.test_sim if phase == "test" else self.train_val_sim
        return dict(human_num=human_num, rule=rule, circle_radius=self.circle_radius,
                    square_width=self.square_width, human_radius=self.config.getfloat("humans", "radius"),
                    human_v_pref=self.config.getfloat("humans", "v_pref"), discomfort_dist=self.discomfort_dist,
                    robot_radius=self.robot.radius, robot_v_pref=self.robot.v_pref,
                    randomize_attributes=bool(self.randomize_attributes))

    def next_cases(self, phase, k, test_case=None):
        """Case ids the next k reset() calls would use (crowd_sim.py:269-270,285-294); advances the counter."""
        cases = []
        for _ in range(k):
            if test_case is not None:
                self.case_counter[phase] = test_case
            cases.append(self.case_counter[phase])
            self.case_counter[phase] = (self.case_counter[phase] + 1) % self.case_size[phase]
        return cases

    def _ensure_batch(self):
        kin = KIN_CODE[self.robot.kinematics]        # changes when train.py swaps the robot's policy (ORCA -> SARL)
        H = len(self.humans) if self.humans is not None else self.human_num
        if self._batch is None or self._batch.H != H or self._batch.cfg.robot_kinematics != kin:
            if self._batch is not None:
                self._batch.close()
            self._batch = BatchedCrowdSim(1, H, device=self.device, **batch_env_kwargs(self))
        return self._batch

    def _ensure_human_actions(self, b):
        """The humans' next velocities for the current state, computed once per env step: an ORCA solve here, a world-model
        prediction in ModelCrowdSim.  Shared by step() and by a query_env lookahead (crowd_sim.py:337-342)."""
        if self._human_v is None:
            b.orca()
            self._human_v = True

    def _sync_agents(self, agents):
        for agent, row in zip([self.robot] + self.humans, agents):
            agent._load_row(row)

    # -- gym surface -------------------------------------------------------------------------------
    def reset(self, phase="test", test_case=None):
        if self.robot is None:
            raise AttributeError("robot has to be set!")
        assert phase in ["train", "val", "test"]
        if self.config.get("humans", "policy") == "trajnet":
            raise NotImplementedError
        case = self.next_cases(phase, 1, test_case)[0]
        if case < 0:
            raise NotImplementedError("debug test cases (crowd_sim.py:297-303) are not part of the hot path")
        self.global_time = 0
        kw = self.scene_kwargs(phase)
        agents = scenes.generate_scene(phase, case, **kw)
        H = agents.shape[0] - 1
        if kw["rule"] == "mixed":                                      # crowd_sim.py:124: the drawn count sticks to the env
            dummy = H == 1 and tuple(agents[1, [0, 1, 4, 5]]) == (0.0, -10.0, 0.0, -10.0)
            self.human_num = 0 if dummy else H
        self.human_times = [0] * H
        self.robot.set(*[float(agents[0, i]) for i in (0, 1, 4, 5, 2, 3)], np.pi / 2)
        self.humans = [Human(self.config, "humans") for _ in range(H)]
        for h, row in zip(self.humans, agents[1:]):
            h.set(*[float(row[i]) for i in (0, 1, 4, 5, 2, 3)], 0, radius=float(row[6]), v_pref=float(row[7]))
        for agent in [self.robot] + self.humans:
            agent.time_step = self.time_step
            agent.policy.time_step = self.time_step
        self.states = list()
        if hasattr(self.robot.policy, "action_values"):
            self.action_values = list()
        if hasattr(self.robot.policy, "get_attention_weights"):
            self.attention_weights = list()
        b = self._ensure_batch()
        b.set_state(agents[None], np.zeros(1))
        self._human_v = None
        if self.robot.sensor == "coordinates":
            return [human.get_observable_state() for human in self.humans]
        raise NotImplementedError

    def onestep_lookahead(self, action):
        return self.step(action, update=False)

    def step(self, action, update=True):
        if self.robot.kinematics not in KIN_CODE:
            raise NotImplementedError("robot kinematics must be holonomic, unicycle or None")
        self.robot.check_validity(action)
        fresh = self._batch is None or self._batch.cfg.robot_kinematics != KIN_CODE[self.robot.kinematics]
        b = self._ensure_batch()
        if fresh:                          # the robot's policy (hence kinematics) changed since reset(): re-upload
            b.set_state(np.array([[a._row() for a in [self.robot] + self.humans]]), np.array([self.global_time]))
        holonomic = self.robot.kinematics == "holonomic"
        if not holonomic:
            b.set_theta(np.array([float(self.robot.theta)]))
        self._ensure_human_actions(b)      # one ORCA solve per env step, shared by the 81 lookahead queries
        act = np.array([[action.vx, action.vy] if holonomic else [action.v, action.r]], dtype=np.float64)
        reward, done, info, dmin = b.step(act, update=update)
        info_obj = info_from_code(info[0], dmin[0])
        if update:
            self.states.append([self.robot.get_full_state(), [human.get_full_state() for human in self.humans]])
            if hasattr(self.robot.policy, "action_values"):
                self.action_values.append(self.robot.policy.action_values)
            if hasattr(self.robot.policy, "get_attention_weights"):
                self.attention_weights.append(self.robot.policy.get_attention_weights())
            agents, times = b.get_state()
            self._sync_agents(agents[0])
            if not holonomic:
                self.robot.theta = float(b.get_theta()[0])        # agent.py:131
            self.global_time = float(times[0])
            self._human_v = None
            for i, human in enumerate(self.humans):
                if self.human_times[i] == 0 and human.reached_destination():
                    self.human_times[i] = self.global_time
            ob = [human.get_observable_state() for human in self.humans]
        else:
            nob = b.next_obs()[0]
            ob = [ObservableState(*[float(x) for x in row]) for row in nob]
        return ob, float(reward[0]), bool(done[0]), info_obj

    def render(self, mode="human", output_file=None, **kw):
        raise NotImplementedError("rendering is outside the B200 hot path (SURVEY §2 #1)")


class ModelCrowdSim(CrowdSim):
    """Drop-in for gym.make('ModelCrowdSim-v0') (crowd_sim/envs/model_crowd_sim.py:15-441): CrowdSim whose humans move with
    the velocities a learned world model predicts (``sim_world``: world_model.MlpWorld / AttentionWorld) instead of ORCA.
    The prediction is one kernel on the env batch (cn_world_predict): it reads the humans' states on the device and leaves
    their next velocities where the step expects the ORCA result; collision / reward / done ladder and the state update are
    the same CUDA step kernel.  Scenes come from the GLOBAL numpy stream (the fork never seeds, :293) and humans start
    moving towards their goal (gen_init_v, :186-192)."""

    def __init__(self):
        super().__init__()
        self.sim_world = None
        self.device = None           # torch device of the world model (model_crowd_sim.py:55)
        self._cuda_device = 0

    def _ensure_batch(self):
        dev = self.device                                  # CrowdSim uses .device as the CUDA ordinal
        self.device = getattr(dev, "index", None) or self._cuda_device
        try:
            return super()._ensure_batch()
        finally:
            self.device = dev

    def _ensure_human_actions(self, b):
        if self._human_v is None:
            self.world_step_batch(b)
            self._human_v = True

    def scene_kwargs(self, phase):
        kw = super().scene_kwargs(phase)
        kw.update(rs=np.random, init_velocity=True)
        return kw

    def reset(self, phase="test", test_case=None, no_random_gen=False):
        if not no_random_gen:
            ob = super().reset(phase, test_case)
            if not hasattr(self.robot.policy, "get_attention_weights"):
                self.attention_weights = None                         # model_crowd_sim.py:319-320
            return ob
        # model_crowd_sim.py:283-285: blank humans, to be filled by set_current_state
        if self.robot is None:
            raise AttributeError("robot has to be set!")
        assert phase in ["train", "val", "test"]
        if test_case is not None:
            self.case_counter[phase] = test_case
        self.global_time = 0
        self.human_times = [0] * self.human_num
        self.humans = [Human(self.config, "humans") for _ in range(self.human_num)]
        self.robot.set(0, -self.circle_radius, 0, self.circle_radius, 0, 0, np.pi / 2)
        for agent in [self.robot] + self.humans:
            agent.time_step = self.time_step
            agent.policy.time_step = self.time_step
        self.states = list()
        self.action_values = list() if hasattr(self.robot.policy, "action_values") else self.action_values
        self.attention_weights = list() if hasattr(self.robot.policy, "get_attention_weights") else None
        self._upload()
        return [human.get_observable_state() for human in self.humans]

    def _upload(self):
        b = self._ensure_batch()
        b.set_state(np.array([[a._row() for a in [self.robot] + self.humans]]), np.array([float(self.global_time)]))
        self._human_v = None

    def set_current_state(self, obs, robot_info=None, phase="train"):
        """model_crowd_sim.py:339-345: load observed humans (position + velocity, goal at the origin) into the simulator."""
        self.human_num = len(obs)
        self.reset(phase, no_random_gen=True)
        if robot_info is not None:
            self.robot.set(robot_info.px, robot_info.py, robot_info.gx, robot_info.gy, 0, 0, np.pi / 2)
        for i, ob in enumerate(obs):
            self.humans[i].set(ob.px, ob.py, 0, 0, ob.vx, ob.vy, 0)
        self._upload()

    def world_step_batch(self, b):
        """model_crowd_sim.py:397-407 for every env of a batch, ON THE DEVICE: the world model reads the humans' states from
        the batch and leaves their next velocities where the env step expects the ORCA result (cn_world_predict)."""
        if not hasattr(self.sim_world, "predict_into"):
            raise NotImplementedError("sim_world must be a modelcrowdnav_b200.world_model MlpWorld / AttentionWorld")
        self.sim_world.predict_into(b)

    def world_velocities(self):
        """model_crowd_sim.py:397-407: (H, 2) velocities the world model predicts for the current human states."""
        b = self._ensure_batch()
        self.world_step_batch(b)
        self._human_v = True
        return b.human_actions()[0].tolist()

    def step(self, action, update=True, new_v=None):
        if new_v is not None:                      # caller-supplied velocities (model_crowd_sim.py:347,397)
            b = self._ensure_batch()
            b.set_human_actions(np.asarray(new_v, dtype=np.float64).reshape(1, len(self.humans), 2))
            self._human_v = True
        return super().step(action, update=update)

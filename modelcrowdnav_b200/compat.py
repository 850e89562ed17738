"""Install this package under the reference's module names so that the UNMODIFIED drivers
(`crowd_nav/test.py`, `crowd_nav/train.py`, `Explorer.run_k_episodes` callers) import the B200 backend:

    import modelcrowdnav_b200.compat as compat
    compat.install_as_reference()
    import gym; env = gym.make('CrowdSim-v0')                 # -> modelcrowdnav_b200.envs.CrowdSim
    from crowd_nav.policy.policy_factory import policy_factory  # -> SARL / ORCA façades on the GPU
    from crowd_nav.utils.explorer import Explorer               # -> batched Explorer

Only the hot-path surface is aliased; model-based / SGAN modules are not provided.
"""
import sys
import types


def make(env_id):
    from .envs import CrowdSim, ModelCrowdSim
    if env_id == "ModelCrowdSim-v0":
        return ModelCrowdSim()
    if env_id != "CrowdSim-v0":
        raise ValueError("only CrowdSim-v0 / ModelCrowdSim-v0 are provided by the B200 backend")
    return CrowdSim()


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_as_reference(provide_gym=True, literal_kinematics=False):
    """literal_kinematics=True reproduces the fork exactly: SARL never reads [action_space] kinematics (cadrl.py:66 is
    commented out), so it plans with ActionRot actions and non-holonomic robot dynamics whatever policy.config says."""
    from . import envs, explorer, policy, trainer
    policy.LITERAL_FORK_KINEMATICS = bool(literal_kinematics)
    _module("crowd_sim")
    from . import world_model
    _module("crowd_sim.envs", CrowdSim=envs.CrowdSim, ModelCrowdSim=envs.ModelCrowdSim)
    _module("crowd_sim.envs.model_crowd_sim", ModelCrowdSim=envs.ModelCrowdSim)
    _module("crowd_nav.policy.world_model", MlpWorld=world_model.MlpWorld, AttentionWorld=world_model.AttentionWorld,
            SGANWorld=world_model.SGANWorld, init_weight=world_model.init_weight)
    _module("crowd_sim.envs.crowd_sim", CrowdSim=envs.CrowdSim)
    _module("crowd_sim.envs.utils")
    _module("crowd_sim.envs.utils.action", ActionXY=envs.ActionXY, ActionRot=envs.ActionRot)
    _module("crowd_sim.envs.utils.state", FullState=envs.FullState, ObservableState=envs.ObservableState,
            JointState=envs.JointState)
    _module("crowd_sim.envs.utils.info", Timeout=envs.Timeout, ReachGoal=envs.ReachGoal, Danger=envs.Danger,
            Collision=envs.Collision, Nothing=envs.Nothing)
    _module("crowd_sim.envs.utils.robot", Robot=envs.Robot)
    _module("crowd_sim.envs.utils.human", Human=envs.Human)
    _module("crowd_sim.envs.utils.agent", Agent=envs.Agent)
    _module("crowd_sim.envs.policy")
    _module("crowd_sim.envs.policy.policy", Policy=policy.Policy)
    _module("crowd_sim.envs.policy.orca", ORCA=policy.ORCA)
    _module("crowd_sim.envs.policy.linear", Linear=policy.Linear)
    _module("crowd_sim.envs.policy.policy_factory", policy_factory=policy.policy_factory)
    _module("crowd_nav")
    _module("crowd_nav.policy")
    _module("crowd_nav.policy.policy_factory", policy_factory=policy.policy_factory)
    _module("crowd_nav.policy.sarl", SARL=policy.SARL)
    _module("crowd_nav.policy.cadrl", CADRL=policy.CADRL, mlp=policy.mlp)
    _module("crowd_nav.policy.lstm_rl", LstmRL=policy.LstmRL)
    _module("crowd_nav.utils")
    _module("crowd_nav.utils.explorer", Explorer=explorer.Explorer, average=explorer.average)
    _module("crowd_nav.utils.memory", ReplayMemory=explorer.ReplayMemory)
    _module("crowd_nav.utils.trainer", Trainer=trainer.Trainer)
    if provide_gym and "gym" not in sys.modules:
        try:
            import gym  # noqa: F401
        except ImportError:
            _module("gym", make=make, Env=object)
    if "gym" in sys.modules and not hasattr(sys.modules["gym"], "_b200_make"):
        g = sys.modules["gym"]
        orig = getattr(g, "make", None)

        def _make(env_id, *a, **kw):
            if env_id in ("CrowdSim-v0", "ModelCrowdSim-v0"):
                return make(env_id)
            return orig(env_id, *a, **kw)
        g.make = _make
        g._b200_make = True

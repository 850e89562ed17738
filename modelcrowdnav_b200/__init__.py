"""modelcrowdnav_b200 -- B200-native (sm_100a) hot path of minh86/ModelCrowdNav.

CrowdSim environment step (ORCA humans, kinematics, collision/discomfort, reward/done) and the SARL
one-step lookahead, batched over thousands of environments per GPU behind a C ABI
(``include/crowdnav_b200.h`` -> ``csrc/libcrowdnav_b200.so``).  No CPU fallback.
"""
from . import _capi  # noqa: F401
from ._capi import CrowdNavError  # noqa: F401
from .batch import (BatchedCrowdSim, BatchedSARL, HostStepBuffers, PackedHostStepBuffers, PipelinedHostRollout,  # noqa: F401
                    pin_to_gpu_numa, rollout_step, rollout_step_host, rollout_step_host_packed)

from .envs import (ActionRot, ActionXY, Collision, CrowdSim, Danger, FullState, Human, JointState, ModelCrowdSim,  # noqa: F401
                   Nothing, ObservableState, ReachGoal, Robot, Timeout)
from .policy import CADRL, ORCA, SARL, Linear, LstmRL, policy_factory  # noqa: F401
from .explorer import Explorer, ReplayMemory  # noqa: F401

__all__ = ["CrowdSim", "ModelCrowdSim", "Robot", "Human", "SARL", "CADRL", "LstmRL", "Linear", "ORCA", "policy_factory", "Explorer", "ReplayMemory",
           "ActionXY", "ActionRot", "FullState", "ObservableState", "JointState",
           "Timeout", "ReachGoal", "Danger", "Collision", "Nothing",
           "BatchedCrowdSim", "BatchedSARL", "HostStepBuffers", "PackedHostStepBuffers", "rollout_step", "rollout_step_host",
           "rollout_step_host_packed", "PipelinedHostRollout", "pin_to_gpu_numa",
           "CrowdNavError"]

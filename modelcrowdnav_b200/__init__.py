"""modelcrowdnav_b200 -- B200-native (sm_100a) hot path of minh86/ModelCrowdNav.

CrowdSim environment step (ORCA humans, kinematics, collision/discomfort, reward/done) and the SARL
one-step lookahead, batched over thousands of environments per GPU behind a C ABI
(``include/crowdnav_b200.h`` -> ``csrc/libcrowdnav_b200.so``).  No CPU fallback.
"""
from . import _capi  # noqa: F401
from ._capi import CrowdNavError  # noqa: F401
from .batch import BatchedCrowdSim, BatchedSARL, HostStepBuffers, rollout_step, rollout_step_host  # noqa: F401

__all__ = ["BatchedCrowdSim", "BatchedSARL", "HostStepBuffers", "rollout_step", "rollout_step_host",
           "CrowdNavError"]

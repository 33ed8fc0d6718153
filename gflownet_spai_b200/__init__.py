"""B200-native SPAI reward path of tonylizza/gflownet-spai.

    from gflownet_spai_b200 import PreconditionerEnv, GFlowNet, trajectory_balance_loss

The arithmetic lives in libspai_b200.so (CUDA, sm_100a) behind the C ABI of
include/spai_b200.h; these modules are the host-side mirror of the reference's
Python interface.
"""
from .env import Data, Env, PreconditionerEnv, SpaiContext, residual_pair  # noqa: F401
from .sampler import BackwardPolicy, GFlowNet, Log, trajectory_balance_loss  # noqa: F401

__all__ = ["Data", "Env", "PreconditionerEnv", "SpaiContext", "residual_pair", "GFlowNet", "Log",
           "trajectory_balance_loss", "BackwardPolicy"]

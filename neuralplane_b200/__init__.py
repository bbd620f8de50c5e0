"""neuralplane_b200: B200-native (sm_100a) vectorised F-16 flight-dynamics step behind NeuralPlane's env surface.

    from neuralplane_b200 import ControlEnv, GPUVecEnv
    env = ControlEnv(num_envs=1_000_000, config='heading', model='F16', random_seed=0, device='cuda:0')
    obs = env.reset()
    obs, reward, done, bad_done, exceed_time_limit, info = env.step(action)   # one kernel launch

The per-step work runs in csrc/nplane.cu through the C ABI of include/nplane.h; PyTorch only owns the device
buffers.  There is no CPU fallback.
"""
from .envs.control_env import ControlEnv  # noqa: F401
from .envs.env_base import BaseEnv  # noqa: F401
from .envs.env_wrappers import GPUVecEnv  # noqa: F401
from .envs.planning_env import PlanningEnv  # noqa: F401
from .envs.singlecombat_env import SingleCombatEnv  # noqa: F401
from .envs.multiplecombat_env import MultipleCombatEnv  # noqa: F401

__all__ = ["ControlEnv", "PlanningEnv", "SingleCombatEnv", "MultipleCombatEnv", "BaseEnv", "GPUVecEnv"]

"""Task interface (reference: envs/tasks/task_base.py:8-103).

In the reference a task owns Python lists of reward functions and termination conditions that are evaluated one
ATen op at a time.  Here the observation, the six termination predicates and the reward are evaluated inside the
fused step kernel (csrc/nplane.cu); the task object carries the yaml parameters, the target tensors (views of the
kernel's target rows) and the gym spaces the runners read.
"""
import numpy as np

try:  # gym is optional: only the Box spaces are used (task_base.py:29-43)
    from gym import spaces as _spaces
except Exception:  # pragma: no cover - gym absent in this image
    _spaces = None


class Box:
    """Stand-in for gym.spaces.Box when gym is not installed."""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def sample(self):
        return np.random.uniform(-1.0, 1.0, size=self.shape).astype(self.dtype)


def make_box(dim):
    if _spaces is not None:
        return _spaces.Box(low=-np.inf, high=np.inf, shape=(dim,))
    return Box(-np.inf, np.inf, (dim,))


class BaseTask:
    task_id = None                      # NP_TASK_* of include/nplane.h
    target_names = ()                   # names of the three target rows, in kernel order
    termination_names = ()
    reward_names = ()

    def __init__(self, config, n, device, random_seed, tgt_rows):
        self.config = config
        self.n = n
        self.device = device
        self.num_observation = getattr(self.config, 'num_observation', 12)
        self.num_actions = getattr(self.config, 'num_actions', 5)
        self.noise_scale = getattr(self.config, 'noise_scale', 0.01)
        self.observation_space = make_box(self.num_observation)
        self.action_space = make_box(self.num_actions)
        for name, row in zip(self.target_names, tgt_rows):
            setattr(self, name, row)

    # reference API; all three are produced by the fused kernel and cached on the env
    def reset(self, env):
        env.reset()

    def get_obs(self, env):
        return env.last_obs

    def get_reward(self, env):
        return env.last_reward

    def get_termination(self, env, info={}):
        return env.is_done, env.bad_done, env.exceed_time_limit, info

"""Aircraft plug-in interface (reference: envs/models/model_base.py:7-250).

A model owns the per-aircraft state/control tensors and exposes the getter set that tasks, reward functions,
termination conditions and the PID controllers (algorithms/pid/*) call.  Concrete models bind their tensors to
the device buffers the native step kernel works on.
"""
from abc import ABC, abstractmethod


class BaseModel(ABC):
    def __init__(self, config, n, device, random_seed):
        self.config = config
        self.n = n
        self.device = device
        self.random_seed = random_seed

    @abstractmethod
    def reset(self, env):
        raise NotImplementedError

    @abstractmethod
    def get_extended_state(self):
        raise NotImplementedError

    @abstractmethod
    def update(self, action):
        raise NotImplementedError

    # the 22 getters of the reference interface
    GETTERS = ("get_state", "get_control", "get_position", "get_ground_speed", "get_climb_rate", "get_posture",
               "get_euler_angular_velocity", "get_vt", "get_TAS", "get_EAS", "get_AOA", "get_AOS",
               "get_angular_velocity", "get_thrust", "get_control_surface", "get_velocity", "get_acceleration",
               "get_G", "get_EAS2TAS", "get_accels", "get_atmos", "get_extended_state")

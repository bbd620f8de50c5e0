"""Attitude-control task parameters/targets (reference: envs/tasks/control_task.py:23-68)."""
from .task_base import BaseTask


class ControlTask(BaseTask):
    task_id = 1
    target_names = ("target_pitch", "target_heading", "target_vt")
    reward_names = ("PostureReward", "EventDrivenReward")
    termination_names = ("Overload", "LowAltitude", "HighSpeed", "LowSpeed", "ExtremeState", "UnreachPosture")

    def __init__(self, config, n, device, random_seed, tgt_rows):
        super().__init__(config, n, device, random_seed, tgt_rows)
        self.max_pitch_increment = getattr(self.config, 'max_pitch_increment', 0.3)
        self.max_heading_increment = getattr(self.config, 'max_heading_increment', 0.3)
        self.max_velocities_u_increment = getattr(self.config, 'max_velocities_u_increment', 100)

"""Heading task parameters/targets (reference: envs/tasks/heading_task.py:23-69)."""
from .task_base import BaseTask


class HeadingTask(BaseTask):
    task_id = 0
    target_names = ("target_altitude", "target_heading", "target_vt")
    reward_names = ("HeadingReward", "EventDrivenReward")
    termination_names = ("Overload", "LowAltitude", "HighSpeed", "LowSpeed", "ExtremeState", "UnreachHeading")

    def __init__(self, config, n, device, random_seed, tgt_rows):
        super().__init__(config, n, device, random_seed, tgt_rows)
        self.max_heading_increment = getattr(self.config, 'max_heading_increment', 0.3)
        self.max_altitude_increment = getattr(self.config, 'max_altitude_increment', 500)
        self.max_velocities_u_increment = getattr(self.config, 'max_velocities_u_increment', 100)

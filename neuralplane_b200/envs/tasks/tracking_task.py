"""Waypoint-tracking task parameters/targets (reference: envs/tasks/tracking_task.py:22-71)."""
from .task_base import BaseTask


class TrackingTask(BaseTask):
    task_id = 2
    target_names = ("target_npos", "target_epos", "target_altitude")
    reward_names = ("PositionReward", "EventDrivenReward")
    termination_names = ("Overload", "LowAltitude", "HighSpeed", "LowSpeed", "ExtremeState", "UnreachTarget")

    def __init__(self, config, n, device, random_seed, tgt_rows):
        super().__init__(config, n, device, random_seed, tgt_rows)
        self.max_distance = getattr(self.config, 'max_distance', 2000)
        self.min_distance = getattr(self.config, 'min_distance', 2000)

"""ControlEnv (reference: envs/control_env.py:12-35): picks the aircraft model and the task by name."""
from .env_base import BaseEnv
from .models.F16_model import F16Model, F16TablesModel
from .models.UAV_model import UAVModel
from .tasks.control_task import ControlTask
from .tasks.heading_task import HeadingTask
from .tasks.tracking_task import TrackingTask


class ControlEnv(BaseEnv):
    """Single-agent fly-control env; same constructor signature as the reference plus sharding keywords."""

    def __init__(self, num_envs=1, config='heading', model='F16', random_seed=None, device="cuda:0", **kw):
        super().__init__(num_envs, config, model, random_seed, device, **kw)

    def load(self, random_seed, config, model):
        if model == 'F16':
            self.model = F16Model(self.config, self.n, self.device, random_seed, ld=self.ld)
        elif model == 'F16_tables':   # the F-16 with the table aero back-end (no reference counterpart in envs/)
            self.model = F16TablesModel(self.config, self.n, self.device, random_seed, ld=self.ld)
        elif model == 'UAV':
            self.model = UAVModel(self.config, self.n, self.device, random_seed, ld=self.ld)
        else:
            raise NotImplementedError(f"model {model!r}: the native plug-ins are 'F16', 'F16_tables' and 'UAV' (control_env.py:22-27)")
        rows = [self._tgt[j, :self.n] for j in range(3)]
        name = config if config in ('heading', 'control', 'tracking') else getattr(self.config, 'task', None)
        if name is None:  # a yaml path: recognise the task from its file name
            import os
            name = os.path.splitext(os.path.basename(str(config)))[0]
        if name == 'heading':
            self.task = HeadingTask(self.config, self.n, self.device, random_seed, rows)
        elif name == 'control':
            self.task = ControlTask(self.config, self.n, self.device, random_seed, rows)
        elif name == 'tracking':
            self.task = TrackingTask(self.config, self.n, self.device, random_seed, rows)
        else:
            raise NotImplementedError(name)

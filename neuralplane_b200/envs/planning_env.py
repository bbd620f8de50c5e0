"""PlanningEnv (reference: envs/planning_env.py:31-177): hierarchical tracking env.  The high-level 3-D action sets
pitch / heading / speed targets that a low-level controller tracks for 50 FDM sub-steps; the whole env step is ONE
kernel launch (np_env_plan_step) with the aircraft state held in registers across the sub-steps.

Low-level controller: the reference loads a GRU PPO actor from a checkpoint that is not part of its repository
(planning_env.py:16,41-44), so the env cannot be constructed there.  Here the low level is the reference's own PID
stack (algorithms/pid: roll / pitch / yaw rate loops, L1 heading hold, a TAS loop), fused into the kernel
(csrc/ctrl_device.cuh); parity is pinned against those reference classes (tests/golden/planning_pid_traj.npz).
"""
import torch

from .. import _native as nv
from .env_base import BaseEnv
from .models.F16_model import F16Model, F16TablesModel
from .tasks.tracking_task import TrackingTask
from .utils.utils import wrap_PI


class PlanningEnv(BaseEnv):
    action_width = 3

    def __init__(self, num_envs=1, config='tracking', model='F16', random_seed=None, device="cuda:0", n_substeps=50, **kw):
        self.n_substeps = int(n_substeps)
        super().__init__(num_envs, config, model, random_seed, device, **kw)

    def load(self, random_seed, config, model):
        if model not in ('F16', 'F16_tables'):
            raise NotImplementedError("the fused PID low-level controller flies the F16 plug-in (MLP or table aero back-end)")
        cls = F16Model if model == 'F16' else F16TablesModel
        self.model = cls(self.config, self.n, self.device, random_seed, ld=self.ld)
        rows = [self._tgt[j, :self.n] for j in range(3)]
        self.task = TrackingTask(self.config, self.n, self.device, random_seed, rows)

    @property
    def pid_state(self):
        """[n, 12] view of the fused controller's state: {roll, pitch, yaw, speed} x {error, integrator, last_out}."""
        off = nv.lib().np_env_pid_offset_bytes(self._cfg)
        blk = self._workspace[off: off + 12 * self.ld * 4].view(torch.float32).view(12, self.ld)
        return blk.t()[:self.n]

    def reset_controller(self):
        """Zero the controller state and re-arm the PIDs' first-call initialisation (pid.py:13,22-27)."""
        self.pid_state.zero_()
        nv.check(nv.lib().np_env_set_pid_started(self._handle, 0), "np_env_set_pid_started")

    def low_level_obs(self, target_pitch, target_heading, target_vt):
        """The 22-D low-level observation the reference feeds its actor (planning_env.py:60-142): the control-task
        layout without noise.  Not used by the fused PID controller; kept for callers that bring their own policy."""
        m = self.model
        npos, epos, altitude = m.get_position()
        roll, pitch, heading = m.get_posture()
        vt, EAS, alpha, beta = m.get_vt(), m.get_EAS(), m.get_AOA(), m.get_AOS()
        P, Q, R = m.get_angular_velocity()
        T = m.get_thrust()
        el, ail, rud, lef = m.get_control_surface()
        cols = [wrap_PI(pitch - target_pitch), wrap_PI(heading - target_heading), (vt - target_vt) * 0.3048 / 340,
                altitude * 0.3048 / 5000, torch.sin(roll), torch.cos(roll), torch.sin(pitch), torch.cos(pitch),
                EAS * 0.3048 / 340, torch.sin(alpha), torch.cos(alpha), torch.sin(beta), torch.cos(beta), P, Q, R,
                T / 0.225 / 76300 * 0.3048, el / 45, ail / 45, rud / 45, lef / 45, m.get_EAS2TAS()]
        return torch.stack(cols, dim=1)

    def step(self, action, render=False, count=0, reset_draws=None, noise=None):
        """PlanningEnv.step (planning_env.py:144-177): action [n, 3] in [-1, 1]."""
        if not torch.is_tensor(action):
            action = torch.as_tensor(action, dtype=torch.float32, device=self.device)
        if action.dim() != 2 or action.shape[0] != self.n or action.shape[1] < 3:
            raise ValueError(f"action must have shape [{self.n}, 3], got {tuple(action.shape)}")
        if action.shape[1] != 3 or action.dtype != torch.float32 or not action.is_contiguous() or action.device != self.device:
            action = action[:, :3].to(device=self.device, dtype=torch.float32).contiguous()
        self._sync_cfg()
        st = nv.lib().np_env_plan_step(self._handle, action.data_ptr(), self.n_substeps,
                                       self._ptr(reset_draws, (self.n, nv.NUM_DRAWS), "reset_draws"),
                                       self._ptr(noise, (self.n, nv.NUM_OBS), "noise"), self._stream())
        nv.check(st, "np_env_plan_step")
        return self._obs, self._reward, self.is_done, self.bad_done, self.exceed_time_limit, {}

"""SingleCombatEnv (reference: envs/singlecombat_env.py:25-274): 1-v-1 combat, two aircraft per env, stored as adjacent
pairs (ego = 2e, enemy = 2e + 1).  One kernel launch per env step (np_env_combat_step): env-level reset, 5 FDM sub-steps
under the PID attitude controller, 15-D observation with the relative geometry, AO/TA/range reward, blood model and the
Crash / Shutdown / Timeout terminations.  Both aircraft of a pair live in one thread, so the pairwise terms need no
exchange when pairs are kept on one rank (`neuralplane_b200.sharding.shard_range` never splits a pair); the
role-sharded layout (egos and opponents on different ranks) exchanges 8-float records with an all-gather
(`neuralplane_b200/combat_exchange.py`).

The reference class is stale at the surveyed commit and cannot be constructed; what it computes is pinned through its
own obs / reward / geometry / termination code (tests/golden/combat*_traj.npz), the orchestration is re-derived
(oracle/combat_oracle.py lists every decision).
"""
import torch

from .. import _native as nv
from .env_base import BaseEnv
from .models.F16_model import F16Model
from .tasks.task_base import BaseTask


class CombatTask(BaseTask):
    """Carries the yaml parameters and gym spaces; everything it would compute runs in the kernel."""
    task_id = 0
    target_names = ()
    reward_names = ("PostureReward(AO, TA, R)",)
    termination_names = ("Overload", "LowAltitude", "HighSpeed", "LowSpeed", "ExtremeState", "Crash", "Timeout", "Shutdown")

    def __init__(self, config, n, device, random_seed, tgt_rows):
        super().__init__(config, n, device, random_seed, tgt_rows)
        self.noise_scale = 0.0          # the combat observation carries no noise (singlecombat_env.py:64-138)


class SingleCombatEnv(BaseEnv):
    native_obs_dim = nv.NUM_OBS_COMBAT
    n_substeps = 5                      # singlecombat_env.py:244
    # the device counters' cause bits 5 / 6 carry Crash-or-ego-Shutdown / enemy-Shutdown here (crash.py:29-42, shutdown.py:30-40)
    COUNTER_NAMES = ("overload", "low_altitude", "high_speed", "low_speed", "extreme_state", "crash_or_shutdown",
                     "enemy_shutdown", "resets")

    def __init__(self, num_envs=1, config='selfplay', random_seed=None, device="cuda:0", **kw):
        super().__init__(num_envs, config, 'F16', random_seed, device, **kw)
        if self.num_agents != 2:
            raise NotImplementedError("Singlecombat number of agents must be 2!")
        off = nv.lib().np_env_blood_offset_bytes(self._cfg)
        self.blood = self._workspace[off: off + self.ld * 4].view(torch.float32)[:self.n]
        self.blood.fill_(100.0)

    def load(self, random_seed, config, model):
        self.model = F16Model(self.config, self.n, self.device, random_seed, ld=self.ld)
        self.task = CombatTask(self.config, self.n, self.device, random_seed, ())

    @property
    def ctrl_state(self):
        """[n, 12] view: rows 0..8 roll / pitch / yaw PID {error, integrator, last_out}, 9 roll_dem, 10 pitch_dem."""
        off = nv.lib().np_env_pid_offset_bytes(self._cfg)
        return self._workspace[off: off + 12 * self.ld * 4].view(torch.float32).view(12, self.ld).t()[:self.n]

    def _combat(self, action, n_sub, reset_draws):
        st = nv.lib().np_env_combat_step(self._handle, None if action is None else action.data_ptr(), n_sub,
                                         self._ptr(reset_draws, (self.n, nv.NUM_DRAWS), "reset_draws"), self._stream())
        nv.check(st, "np_env_combat_step")

    def reset(self, reset_draws=None, noise=None):
        """SingleCombatEnv.reset (singlecombat_env.py:183-205): every pair is re-initialised."""
        self._flags.fill_(1)
        self._combat(None, 0, reset_draws)
        return self._obs

    def step(self, action, render=False, count=0, reset_draws=None, noise=None):
        """SingleCombatEnv.step (singlecombat_env.py:240-274): action [n, 4] = [throttle, roll_dem, pitch_dem, yaw]."""
        if not torch.is_tensor(action):
            action = torch.as_tensor(action, dtype=torch.float32, device=self.device)
        if action.dim() != 2 or action.shape[0] != self.n or action.shape[1] < 4:
            raise ValueError(f"action must have shape [{self.n}, 4], got {tuple(action.shape)}")
        if action.shape[1] != 4 or action.dtype != torch.float32 or not action.is_contiguous() or action.device != self.device:
            action = action[:, :4].to(device=self.device, dtype=torch.float32).contiguous()
        self._combat(action, self.n_substeps, reset_draws)
        return self._obs, self._reward, self.is_done, self.bad_done, self.exceed_time_limit, {}

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
from .models.F16_model import F16Model, F16TablesModel
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
    """layout='pair' (default): both aircraft of an env on this rank, stored adjacently; one launch per step, no exchange.
    layout='role': this rank holds ONE aircraft of every env -- all egos (role=0) or all opponents (role=1); local aircraft i
    belongs to global env `first_env + i`.  A step is then  local half (np_env_combat_role_local: flies this rank's aircraft,
    publishes a 28-float record each) -> exchange (a cross-device barrier for NVLink peer slabs, or an NCCL all-gather) ->
    pair half (np_env_combat_role_pair: pulls the partner records and produces Crash / Shutdown / obs / reward / blood /
    final flags).  Outputs are bit-identical to the pair layout's.  Connect an exchange with `connect()` before reset()."""
    native_obs_dim = nv.NUM_OBS_COMBAT
    n_substeps = 5                      # singlecombat_env.py:244
    combat_pairs_per_env = 1            # duels per env (MultipleCombatEnv: 2)
    combat_reward_scale = 0.01          # singlecombat_env.py:176-177
    # the device counters' cause bits 5 / 6 carry Crash-or-ego-Shutdown / enemy-Shutdown here (crash.py:29-42, shutdown.py:30-40)
    COUNTER_NAMES = ("overload", "low_altitude", "high_speed", "low_speed", "extreme_state", "crash_or_shutdown",
                     "enemy_shutdown", "resets")

    def __init__(self, num_envs=1, config='selfplay', random_seed=None, device="cuda:0", layout='pair', role=None,
                 first_env=0, model='F16', **kw):
        if layout not in ('pair', 'role'):
            raise ValueError("layout must be 'pair' or 'role'")
        self.layout, self.role, self.exchange = layout, role, None
        if model not in ('F16', 'F16_tables') or (model != 'F16' and layout == 'role'):
            raise NotImplementedError("combat flies the F16 plug-in ('F16'; 'F16_tables' in the pair layout)")
        if layout == 'role':
            if role not in (0, 1):
                raise ValueError("layout='role' needs role=0 (egos) or role=1 (opponents)")
            if num_envs % 2:
                raise ValueError("layout='role' needs an even number of local envs (the step kernel moves aircraft in pairs)")
            kw.update(local_agents=1, index_base=2 * int(first_env) + role, index_stride=2)
        super().__init__(num_envs, config, model, random_seed, device, **kw)
        if layout == 'pair' and self.num_agents != 2 * self.combat_pairs_per_env:
            raise NotImplementedError("Singlecombat number of agents must be 2!")
        off = nv.lib().np_env_blood_offset_bytes(self._cfg)
        self.blood = self._workspace[off: off + self.ld * 4].view(torch.float32)[:self.n]
        self.blood.fill_(100.0)
        off = nv.lib().np_env_pair_reset_offset_bytes(self._cfg)
        self._pair_reset = self._workspace[off: off + self.ld]

    def load(self, random_seed, config, model):
        cls = F16Model if model == 'F16' else F16TablesModel
        self.model = cls(self.config, self.n, self.device, random_seed, ld=self.ld)
        self.task = CombatTask(self.config, self.n, self.device, random_seed, ())

    @property
    def ctrl_state(self):
        """[n, 12] view: rows 0..8 roll / pitch / yaw PID {error, integrator, last_out}, 9 roll_dem, 10 pitch_dem."""
        off = nv.lib().np_env_pid_offset_bytes(self._cfg)
        return self._workspace[off: off + 12 * self.ld * 4].view(torch.float32).view(12, self.ld).t()[:self.n]

    def connect(self, exchange):
        """layout='role': the object that makes the partner rank's records readable (neuralplane_b200.combat_exchange:
        PeerSlabExchange, AllGatherExchange or, for two envs in one process, LocalPairExchange)."""
        if self.layout != 'role':
            raise RuntimeError("connect() is for layout='role'")
        self.exchange = exchange
        return self

    def _combat(self, action, n_sub, reset_draws):
        draws = self._ptr(reset_draws, (self.n, nv.NUM_DRAWS), "reset_draws")
        a = None if action is None else action.data_ptr()
        if self.layout == 'pair':
            st = nv.lib().np_env_combat_step(self._handle, a, n_sub, draws, self._stream())
            nv.check(st, "np_env_combat_step")
            return
        self.step_local(action, n_sub, reset_draws)
        self.exchange.sync(self)
        self.step_pair(n_sub)

    # ---- role layout, split phases (two envs of one process interleave them: combat_exchange.step_both) ---------------
    def step_local(self, action, n_sub, reset_draws=None):
        if self.exchange is None:
            raise RuntimeError("layout='role': connect() an exchange first")
        rec = self.exchange.own_slab(self)
        st = nv.lib().np_env_combat_role_local(self._handle, None if action is None else action.data_ptr(), n_sub,
                                               self._ptr(reset_draws, (self.n, nv.NUM_DRAWS), "reset_draws"), rec.data_ptr(),
                                               self._stream())
        nv.check(st, "np_env_combat_role_local")

    def step_pair(self, n_sub):
        ex = self.exchange
        st = nv.lib().np_env_combat_role_pair(self._handle, ex.own_slab(self).data_ptr(), ex.partner_ptr(self), 1 if ex.peer else 0,
                                              self.role, n_sub, self._stream())
        nv.check(st, "np_env_combat_role_pair")
        ex.advance(self)

    def _normalise_action(self, action):
        if not torch.is_tensor(action):
            action = torch.as_tensor(action, dtype=torch.float32, device=self.device)
        if action.dim() != 2 or action.shape[0] != self.n or action.shape[1] < 4:
            raise ValueError(f"action must have shape [{self.n}, 4], got {tuple(action.shape)}")
        if action.shape[1] != 4 or action.dtype != torch.float32 or not action.is_contiguous() or action.device != self.device:
            action = action[:, :4].to(device=self.device, dtype=torch.float32).contiguous()
        return action

    def reset(self, reset_draws=None, noise=None):
        """SingleCombatEnv.reset (singlecombat_env.py:183-205): every pair is re-initialised."""
        self._flags.fill_(1)
        self._pair_reset.fill_(1)
        self._combat(None, 0, reset_draws)
        return self._obs

    def step(self, action, render=False, count=0, reset_draws=None, noise=None):
        """SingleCombatEnv.step (singlecombat_env.py:240-274): action [n, 4] = [throttle, roll_dem, pitch_dem, yaw]."""
        self._combat(self._normalise_action(action), self.n_substeps, reset_draws)
        return self._obs, self._reward, self.is_done, self.bad_done, self.exceed_time_limit, {}


def role_sharded_bench(dev, rank, world, pairs_total, K, W, barrier, max_over_ranks):
    """bench.py side line `combat_role_sharded`: 5 x 10^5 envs, egos on ranks [0, world/2), opponents on the rest, the
    exchange inside the timed region; both exchanges (NVLink peer slabs pulled by the pair kernel / NCCL all-gather) and a
    bit-identity check of every output against the pair-sharded env on the same envs."""
    import torch.distributed as dist
    from ..combat_exchange import AllGatherExchange, PeerSlabExchange, role_block
    role, first_env, n_env = role_block(pairs_total, rank, world)
    acts = [torch.rand((n_env, 4), device=dev, generator=torch.Generator(device=dev).manual_seed(100 + k)) * 2 - 1 for k in range(2)]
    out = {"layout": f"egos on ranks 0..{world // 2 - 1}, opponents on ranks {world // 2}..{world - 1}; {n_env} envs per rank",
           "record_bytes_per_aircraft": 4 * nv.COMBAT_RECORD_FLOATS}

    def timed(ex_cls, label):
        env = SingleCombatEnv(num_envs=n_env, config="selfplay", random_seed=0, device=dev, layout='role', role=role, first_env=first_env)
        ex = ex_cls(env)
        env.connect(ex)
        env.reset()
        for k in range(W):
            env.step(acts[k % 2])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(K):
            env.step(acts[k % 2])
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / K
        xs = max_over_ranks(ex.time_exchange(env, 20))
        out[label] = {"ms_per_env_step": ms, "env_steps_per_s": 2 * pairs_total / (ms * 1e-3),
                      "fdm_steps_per_s": 2 * pairs_total * env.n_substeps / (ms * 1e-3), "exchange_ms": xs,
                      "link_bytes_per_step_per_rank": ex.link_bytes(env)}
        return env

    env_p = timed(PeerSlabExchange, "peer_slabs")
    env_g = timed(AllGatherExchange, "all_gather")
    # bit identity: the same envs stepped pair-sharded on this rank (both aircraft local), compared on this rank's role
    chk_env, steps = min(n_env, 20_000), 12
    pair = SingleCombatEnv(num_envs=chk_env, config="selfplay", random_seed=0, device=dev, index_base=2 * first_env)
    rolee = SingleCombatEnv(num_envs=chk_env, config="selfplay", random_seed=0, device=dev, layout='role', role=role, first_env=first_env)
    rolee.connect(PeerSlabExchange(rolee))
    same = bool(torch.equal(pair.reset()[role::2], rolee.reset()))
    gen = torch.Generator(device=dev).manual_seed(7)          # the same seed on every rank: both halves see the same actions
    for k in range(steps):
        a_pair = torch.rand((2 * chk_env, 4), device=dev, generator=gen) * 2 - 1
        rp, rr = pair.step(a_pair), rolee.step(a_pair[role::2].contiguous())
        same = same and all(bool(torch.equal(x[role::2], y)) for x, y in zip(rp[:5], rr[:5]))
        same = same and bool(torch.equal(pair.model.s[role::2], rolee.model.s)) and bool(torch.equal(pair.blood[role::2], rolee.blood))
    flag = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["bit_identical_to_pair_sharded"] = bool(flag.item())
    out["bit_identity_check"] = f"{steps} steps x {chk_env} envs per rank, obs / reward / flags / state / blood"
    del env_p, env_g
    return out

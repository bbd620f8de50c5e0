"""Batched gym-style env (reference: envs/env_base.py:13-109) whose step() is one native kernel launch.

Keeps the reference surface: n = num_envs * num_agents aircraft, `reset() -> obs[n,D]`,
`step(action[n,A]) -> (obs, reward, done, bad_done, exceed_time_limit, info)`, attributes `model`, `task`,
`step_count`, `is_done`, `bad_done`, `exceed_time_limit`, the gym spaces and `num_observation` / `num_actions`.
Differences a caller can observe: returned tensors are the env's persistent device buffers (overwritten by the
next step) and `step_count` is int32.
"""
import ctypes as C
import os

import numpy as np
import torch

from .. import _native as nv
from . import _soa
from .utils.utils import parse_config

_CFG_KEYS = (  # yaml key, default (the reference's getattr defaults), ctypes field
    ("dt", 0.02), ("airspeed", 0), ("noise_scale", 0.01),
    ("altitude_limit", 2500.0), ("acceleration_limit", 300.0), ("max_velocity", 3), ("min_velocity", 0.01),
    ("min_alpha", -20), ("max_alpha", 45), ("min_beta", -30), ("max_beta", 30),
    ("max_heading_increment", 0.3), ("max_pitch_increment", 0.3), ("max_velocities_u_increment", 100),
    ("max_distance", 2000), ("min_distance", 2000),
    ("max_check_interval", 1500), ("min_check_interval", 300),
    ("max_altitude", 20000), ("min_altitude", 19000), ("max_vt", 1200), ("min_vt", 1000),
    # combat (selfplay.yaml; singlecombat_env.py:33-44, crash.py:16, timeout.py:16)
    ("max_steps", 500), ("distance_limit", 200), ("target_dist", 3), ("max_heading", 0.5), ("min_heading", -0.5),
    ("max_npos", 5000), ("min_npos", -5000), ("max_epos", 5000), ("min_epos", -5000),
)


class BaseEnv:
    metadata = {}
    native_obs_dim = nv.NUM_OBS

    def __init__(self, num_envs=10, config='heading', model='F16', random_seed=None, device="cuda:0",
                 index_base=0, use_coef_cache=True, local_agents=None, index_stride=1):
        """index_base / index_stride: global aircraft index of local aircraft i = index_base + index_stride * i (rank
        sharding; the in-kernel RNG streams are keyed by global index).  local_agents: agents of each env that live on THIS
        rank when fewer than the yaml's num_agents do (the role-sharded combat layout keeps one of the two)."""
        self.config = parse_config(config)
        self.num_envs = num_envs
        self.num_agents = getattr(self.config, 'num_agents', 100) if local_agents is None else int(local_agents)
        self.index_stride = int(index_stride)
        self.n = self.num_agents * self.num_envs
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError(f"neuralplane_b200 envs run on CUDA devices only (got {device}); no CPU fallback exists")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.random_seed = random_seed
        self._seed = int(random_seed) if random_seed is not None else int.from_bytes(os.urandom(8), "little")
        self.index_base = int(index_base)
        self.use_coef_cache = bool(use_coef_cache)
        self.ld = _soa.pitch(self.n)
        self._handle = C.c_void_p()

        n, ld, dev = self.n, self.ld, self.device
        self._tgt = torch.zeros((3, ld), device=dev)
        self._step_count = torch.zeros(ld, dtype=torch.int32, device=dev)
        self._flags = torch.ones((3, ld), dtype=torch.uint8, device=dev)   # everyone resets first (env_base.py:31-33)
        self.load(random_seed, config, model)                              # builds self.model / self.task
        self._obs = torch.zeros((n, self.num_observation), device=dev)
        self._reward = torch.zeros(n, device=dev)
        self.step_count = self._step_count[:n]
        self._flag_views = [self._flags[j, :n].view(torch.bool) for j in range(3)]
        self.create_records = False
        self._native_create()

    # ---- construction ------------------------------------------------------------------------------------
    def load(self, random_seed, config, model):
        raise NotImplementedError

    def _cfg_struct(self):
        c = nv.EnvCfg()
        c.n, c.ld, c.task = self.n, self.ld, self.task.task_id
        c.model = self.model.model_id
        c.use_coef_cache = 1 if self.use_coef_cache else 0
        c.seed, c.index_base = self._seed & (2 ** 64 - 1), self.index_base
        c.index_stride = self.index_stride
        c.combat_pairs_per_env = getattr(self, "combat_pairs_per_env", 1)
        c.combat_reward_scale = getattr(self, "combat_reward_scale", 0.01)
        for key, default in _CFG_KEYS:
            setattr(c, key, getattr(self.config, key, default))
        c.noise_scale = self.task.noise_scale
        c.airspeed = self.model.airspeed
        init_state = getattr(self.config, 'init_state', None)
        c.init_T = init_state['init_T'] if init_state else getattr(self.config, 'init_T', 2000)
        return c

    def _native_create(self):
        if self.num_observation != self.native_obs_dim:
            raise NotImplementedError(f"this env's native kernel produces {self.native_obs_dim}-D observations")
        L = nv.lib()
        self._cfg = self._cfg_struct()
        with torch.cuda.device(self.device):
            aero = self.model.aero.handle if self.model.aero is not None else None
            if getattr(self.model, "aero_backend", "mlp") == "tables":
                nv.check(L.np_env_create_tables(C.byref(self._cfg), aero, C.byref(self._handle)), "np_env_create_tables")
            else:
                nv.check(L.np_env_create(C.byref(self._cfg), aero, C.byref(self._handle)), "np_env_create")
            nbytes = L.np_env_workspace_bytes(C.byref(self._cfg))
            self._workspace = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
            b = nv.Buffers(self.model._s.data_ptr(), self.model._u.data_ptr(), self._tgt.data_ptr(),
                           self._step_count.data_ptr(), self._flags.data_ptr(), self._obs.data_ptr(),
                           self._reward.data_ptr(), self._workspace.data_ptr())
            nv.check(L.np_env_bind(self._handle, C.byref(b), self._stream()), "np_env_bind")

    def _sync_cfg(self):
        ns = float(self.task.noise_scale)
        if ns != self._cfg.noise_scale:
            self._cfg.noise_scale = ns
            nv.check(nv.lib().np_env_set_cfg(self._handle, C.byref(self._cfg)), "np_env_set_cfg")

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                nv.lib().np_env_destroy(self._handle)
        except Exception:
            pass

    def seed(self, random_seed):
        self._seed = int(random_seed)
        self._cfg.seed = self._seed & (2 ** 64 - 1)
        nv.check(nv.lib().np_env_set_cfg(self._handle, C.byref(self._cfg)), "np_env_set_cfg")

    # ---- reference attribute surface ----------------------------------------------------------------------
    @property
    def is_done(self):
        return self._flag_views[0]

    @property
    def bad_done(self):
        return self._flag_views[1]

    @property
    def exceed_time_limit(self):
        return self._flag_views[2]

    @property
    def observation_space(self):
        return self.task.observation_space

    @property
    def action_space(self):
        return self.task.action_space

    @property
    def num_observation(self):
        return self.task.num_observation

    @property
    def num_actions(self):
        return self.task.num_actions

    @property
    def last_obs(self):
        return self._obs

    @property
    def last_reward(self):
        return self._reward

    def obs(self):
        return self._obs

    def reward(self):
        return self._reward

    def info(self):
        return {}

    def get_number_of_agents(self):
        return self.n

    COUNTER_NAMES = nv.COUNTER_NAMES     # what the eight device counters mean for this env (combat renames two)

    def termination_counters(self):
        """Per-cause termination counts since construction (device counters; one sync)."""
        out = (C.c_uint64 * nv.NUM_COUNTERS)()
        nv.check(nv.lib().np_env_counters(self._handle, out, self._stream()), "np_env_counters")
        return dict(zip(self.COUNTER_NAMES, [int(x) for x in out]))

    def advance_rng(self, delta):
        """CUDA-graph capture freezes the host-side RNG counter that every reset() / step() call advances, so each replay of a
        captured graph would draw the same reset states and observation noise.  Capture `env.advance_rng(K)` at the end of a
        graph of K step calls: replay r then continues exactly the sequence an eager loop of r * K steps produces."""
        nv.check(nv.lib().np_env_rng_advance(self._handle, int(delta), self._stream()), "np_env_rng_advance")

    # ---- reset / step ---------------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    @staticmethod
    def _ptr(t, shape, what):
        if t is None:
            return None
        if t.dtype != torch.float32 or not t.is_cuda or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
            raise ValueError(f"{what} must be a contiguous float32 CUDA tensor of shape {tuple(shape)}")
        return t.data_ptr()

    def reset(self, reset_draws=None, noise=None):
        """BaseEnv.reset (env_base.py:83-97).  reset_draws [n,5] / noise [n,22] inject the random numbers
        (parity runs); by default they come from the in-kernel Philox stream keyed by (seed, global index)."""
        self._sync_cfg()
        st = nv.lib().np_env_reset(self._handle, self._ptr(reset_draws, (self.n, nv.NUM_DRAWS), "reset_draws"),
                                   self._ptr(noise, (self.n, nv.NUM_OBS), "noise"), self._stream())
        nv.check(st, "np_env_reset")
        return self._obs

    def step(self, action, render=False, count=0, reset_draws=None, noise=None):
        """BaseEnv.step (env_base.py:99-109): one kernel launch, no host synchronisation."""
        if not torch.is_tensor(action):
            action = torch.as_tensor(np.asarray(action), dtype=torch.float32, device=self.device)
        if action.dim() != 2 or action.shape[0] != self.n or action.shape[1] < 4:
            raise ValueError(f"action must have shape [{self.n}, >=4], got {tuple(action.shape)}")
        if action.shape[1] != 4 or action.dtype != torch.float32 or not action.is_contiguous() or action.device != self.device:
            action = action[:, :4].to(device=self.device, dtype=torch.float32).contiguous()
        self._sync_cfg()
        st = nv.lib().np_env_step(self._handle, action.data_ptr(),
                                  self._ptr(reset_draws, (self.n, nv.NUM_DRAWS), "reset_draws"),
                                  self._ptr(noise, (self.n, nv.NUM_OBS), "noise"), self._stream())
        nv.check(st, "np_env_step")
        if render:
            self.render(count=count)
        return self._obs, self._reward, self.is_done, self.bad_done, self.exceed_time_limit, {}

    def step_range(self, action, first, count, advance=True, reset_draws=None, noise=None):
        """One logical step may be issued as several aircraft ranges (on different streams): `action` is the full
        [n, 4] device tensor, the range is [first, first + count); advance=True on the first range only."""
        self._sync_cfg()
        if action.device != self.device:
            raise ValueError(f"action must live on {self.device}, got {action.device}")
        st = nv.lib().np_env_step_range(self._handle, self._ptr(action, (self.n, 4), "action"),
                                        self._ptr(reset_draws, (self.n, nv.NUM_DRAWS), "reset_draws"),
                                        self._ptr(noise, (self.n, nv.NUM_OBS), "noise"), int(first), int(count),
                                        1 if advance else 0, self._stream())
        nv.check(st, "np_env_step_range")

    def step_host(self, action_np, act_pinned, act_dev, obs_pinned, rew_pinned, flags_pinned, edges, n_chunks):
        """The numpy boundary in one native call (np_env_step_host): host actions in, pinned host obs [n, D] / reward [n] /
        flags [3, n] out, pipelined over `n_chunks` aircraft ranges with edges `edges` (ctypes int array).  Returns when
        the host buffers are ready."""
        self._sync_cfg()
        n, D = self.n, self.num_observation
        self._check_host_action(action_np)
        self._check_pinned(act_pinned, (n, 4), torch.float32, "act_pinned")
        self._check_pinned(obs_pinned, (n, D), torch.float32, "obs_pinned")
        self._check_pinned(rew_pinned, (n,), torch.float32, "rew_pinned")
        self._check_pinned(flags_pinned, (3, n), torch.uint8, "flags_pinned")
        self._ptr(act_dev, (n, 4), "act_dev")
        st = nv.lib().np_env_step_host(self._handle, action_np.ctypes.data, act_pinned.data_ptr(), act_dev.data_ptr(),
                                       obs_pinned.data_ptr(), rew_pinned.data_ptr(), flags_pinned.data_ptr(), edges,
                                       int(n_chunks), self._stream())
        nv.check(st, "np_env_step_host")

    def _check_host_action(self, a):
        if not isinstance(a, np.ndarray) or a.dtype != np.float32 or a.shape != (self.n, 4) or not a.flags.c_contiguous:
            raise ValueError(f"host actions must be a C-contiguous float32 numpy array of shape ({self.n}, 4)")

    def _check_pinned(self, t, shape, dtype, what):
        key = (t.data_ptr(), tuple(shape), dtype)
        seen = self.__dict__.setdefault("_pinned_ok", set())      # a staging buffer is validated once (is_pinned() is a driver query)
        if key in seen:
            return
        if t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous() or t.is_cuda or not t.is_pinned():
            raise ValueError(f"{what} must be a contiguous pinned host tensor of shape {tuple(shape)}, dtype {dtype}")
        seen.add(key)

    def step_mapped(self, action_np, act_pinned, obs_pinned, rew_pinned, flags_pinned, act_dev=None):
        """The numpy boundary with no copy engine in the path (np_env_step_mapped): the step kernel reads the actions from
        and writes obs [n, D] / reward [n] / flags [3, ld] straight into pinned (device-mapped) host memory.  Returns when
        the host buffers are ready.  `last_obs` / `last_reward` (the device buffers) are not updated by this call."""
        self._sync_cfg()
        n, D = self.n, self.num_observation
        self._check_host_action(action_np)
        self._check_pinned(act_pinned, (n, 4), torch.float32, "act_pinned")
        self._check_pinned(obs_pinned, (n, D), torch.float32, "obs_pinned")
        self._check_pinned(rew_pinned, (n,), torch.float32, "rew_pinned")
        self._check_pinned(flags_pinned, (3, self.ld), torch.uint8, "flags_pinned")
        st = nv.lib().np_env_step_mapped(self._handle, action_np.ctypes.data, act_pinned.data_ptr(),
                                         None if act_dev is None else self._ptr(act_dev, (n, 4), "act_dev"),
                                         obs_pinned.data_ptr(), rew_pinned.data_ptr(), flags_pinned.data_ptr(), self.ld,
                                         self._stream())
        nv.check(st, "np_env_step_mapped")

    def launch_info(self):
        g, b, s, m = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        nv.check(nv.lib().np_env_launch_info(self._handle, C.byref(g), C.byref(b), C.byref(s), C.byref(m)), "np_env_launch_info")
        return {"grid": g.value, "block": b.value, "smem_bytes": s.value, "num_sms": m.value}

    def render(self, count, filename='./tracks/F16SimRecording-', max_aircraft=16):
        """Tacview track of the first `max_aircraft` aircraft (env_base.py:111-151); host file I/O, outside the hot path."""
        from .acmi import AcmiWriter
        w = getattr(self, "_acmi", None)
        if w is None or w.prefix != filename or w.max_aircraft != max_aircraft:
            w = self._acmi = AcmiWriter(filename, max_aircraft)
        m = min(max_aircraft, self.n)
        npos, epos, alt = (x[:m].cpu().numpy() for x in self.model.get_position())
        roll, pitch, yaw = (x[:m].cpu().numpy() for x in self.model.get_posture())
        flags = self._flags[:, :self.n]
        ended = bool(flags.any())
        return w.write(count, float(self.step_count[0]) * self.model.dt, npos, epos, alt, roll, pitch, yaw, ended)

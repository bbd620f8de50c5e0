"""GPUVecEnv (reference: envs/env_wrappers.py:84-124): the numpy boundary the runners call.

Same shapes as the reference: actions (num_envs, agents, A) in; obs (num_envs, agents, D), rewards / dones /
bad_dones / exceed_time_limits (num_envs, agents, 1) out.  Host<->device traffic goes through pinned staging
buffers; large ControlEnv populations are pipelined in aircraft chunks on side streams (upload / kernel / download
of different chunks overlap; the 88 B/aircraft observation download is what bounds this boundary), with one
synchronise per step.  The returned arrays alias pinned buffers that are reused every OTHER step
(double-buffered), so a result stays valid until the step after next.
"""
import ctypes as C

import numpy as np
import torch



DEFAULT_PIPELINE = (1, 2, 3, 4, 4, 4)   # relative chunk sizes; see profiles/r01_variants.txt for the sweep


MAX_PIPELINE_CHUNKS = 16   # np_env_step_host (include/nplane.h)


def pipeline_edges(n, pattern):
    """Aircraft ranges [(i0, i1), ...] of the pipelined numpy step, or None for a single launch.  `pattern` is a chunk
    count (equal chunks) or a sequence of relative chunk sizes.  Interior edges fall on multiples of 256 aircraft (whole
    pairs; 16-byte aligned rows for the TMA-staged UAV kernel); empty chunks are dropped."""
    w = [1.0] * int(pattern) if np.isscalar(pattern) else [float(x) for x in pattern]
    if len(w) > MAX_PIPELINE_CHUNKS or any(x <= 0 for x in w):
        raise ValueError(f"pipeline_chunks: 1..{MAX_PIPELINE_CHUNKS} positive chunk sizes, got {pattern!r}")
    if len(w) <= 1:
        return None
    cum = np.cumsum([0.0] + w) / sum(w)
    edges = [min(n, 256 * int(round(n * f / 256))) for f in cum[:-1]] + [n]
    chunks = [(edges[c], edges[c + 1]) for c in range(len(w)) if edges[c + 1] > edges[c]]
    return chunks if len(chunks) > 1 else None


class GPUVecEnv:
    def __init__(self, env_fns, device_tensors=False, pipeline_chunks=None):
        """device_tensors=True (SURVEY f-2): step()/reset() take and return torch CUDA tensors in the same
        (num_envs, agents, .) shapes, with no host round trip and no synchronisation -- for policies that live on
        the same GPU.  The default reproduces the reference's numpy boundary.
        pipeline_chunks: number of aircraft chunks the numpy step is pipelined over (default 4 from 2x10^5 aircraft)."""
        self.device_tensors = bool(device_tensors)
        assert len(env_fns) == 1, "Number of create env funcitions must be 1!"
        self.gpu_vec_env = env_fns[0]()
        assert hasattr(self.gpu_vec_env, "num_envs"), "Parameter of env must contain num_envs!"
        e = self.gpu_vec_env
        self.num_envs = e.num_envs
        self.observation_space = e.observation_space
        self.action_space = e.action_space
        self.agents = e.num_agents
        self.closed = False
        n, D = e.n, e.num_observation
        self._A = A = getattr(e, "action_width", 4)     # PlanningEnv takes 3-D actions
        self._flip = 0
        self.h2d_bytes_per_step = 0 if self.device_tensors else n * A * 4
        self.d2h_bytes_per_step = 0 if self.device_tensors else n * D * 4 + n * 4 + 3 * n
        self._chunks = None
        if self.device_tensors:
            return
        # pipelined boundary for large single-step envs (ControlEnv): aircraft chunks through np_env_step_host
        # pipeline_chunks: a chunk count (equal chunks) or a sequence of relative chunk sizes, e.g. (1, 2, 3, 5, 5): a small
        # first chunk starts the observation download -- the resource this boundary is bound by -- sooner
        if pipeline_chunks is None:
            pipeline_chunks = DEFAULT_PIPELINE if n >= 200_000 else 1
        self._chunks = pipeline_edges(n, pipeline_chunks) if hasattr(e, "step_host") and type(e).__name__ == "ControlEnv" else None
        if self._chunks is not None:
            self._edges = (C.c_int * (len(self._chunks) + 1))(*([c[0] for c in self._chunks] + [n]))
        self._act_h = torch.empty((n, A), dtype=torch.float32).pin_memory()
        self._act_d = torch.empty((n, A), dtype=torch.float32, device=e.device)
        self._out = [dict(obs=torch.empty((n, D), dtype=torch.float32).pin_memory(),
                          rew=torch.empty(n, dtype=torch.float32).pin_memory(),
                          flags=torch.empty((3, n), dtype=torch.uint8).pin_memory()) for _ in range(2)]

    def _download(self, with_rest=True):
        e = self.gpu_vec_env
        o = self._out[self._flip]
        self._flip ^= 1
        o["obs"].copy_(e.last_obs, non_blocking=True)
        if with_rest:
            o["rew"].copy_(e.last_reward, non_blocking=True)
            o["flags"].copy_(e._flags[:, :e.n], non_blocking=True)
        torch.cuda.current_stream(e.device).synchronize()
        return o

    def _step_device(self, actions):
        e = self.gpu_vec_env
        import torch as _t
        a = _t.as_tensor(actions, device=e.device, dtype=_t.float32).reshape(self.num_envs * self.agents, -1)
        obs, rew, done, bad, exc, info = e.step(a)
        shp = (self.num_envs, self.agents, 1)
        return (obs.view(self.num_envs, self.agents, e.num_observation), rew.view(shp), done.view(shp), bad.view(shp),
                exc.view(shp), info)

    def step(self, actions):
        if self.device_tensors:
            return self._step_device(actions)
        e = self.gpu_vec_env
        a = np.asarray(actions, dtype=np.float32).reshape(self.num_envs * self.agents, -1)
        if self._chunks is None:
            self._act_h.copy_(torch.from_numpy(a[:, :self._A]))
            self._act_d.copy_(self._act_h, non_blocking=True)
            e.step(self._act_d)
            o = self._download()
        else:
            o = self._step_pipelined(a)
        shp = (self.num_envs, self.agents, 1)
        obs = o["obs"].numpy().reshape(self.num_envs, self.agents, e.num_observation)
        flags = o["flags"].numpy().view(np.bool_)
        return (obs, o["rew"].numpy().reshape(shp), flags[0].reshape(shp), flags[1].reshape(shp),
                flags[2].reshape(shp), {})

    def _step_pipelined(self, a):
        """Aircraft are independent, so the step is pipelined chunk by chunk over three in-order streams -- upload, kernels,
        download: while chunk c's observations travel to the host (the 88 B/aircraft D2H is what bounds this boundary),
        chunk c+1 runs and chunk c+2's actions are staged and uploaded.  The whole pipeline is ONE native call
        (np_env_step_host): issuing a chunk from Python cost ~110 us of interpreter time, which delayed the first download.
        Same results as the single launch (one RNG counter for all chunks)."""
        e = self.gpu_vec_env
        o = self._out[self._flip]
        self._flip ^= 1
        if a.shape[1] != self._A or not a.flags.c_contiguous:
            a = np.ascontiguousarray(a[:, :self._A])
        e.step_host(a, self._act_h, self._act_d, o["obs"], o["rew"], o["flags"], self._edges, len(self._chunks))
        return o

    def reset(self):
        e = self.gpu_vec_env
        if self.device_tensors:
            return e.reset().view(self.num_envs, self.agents, e.num_observation)
        e.reset()
        o = self._download(with_rest=False)
        return o["obs"].numpy().reshape(self.num_envs, self.agents, e.num_observation)

    def step_async(self, actions):
        pass

    def step_wait(self):
        pass

    def close_extras(self):
        pass

    def close(self):
        if self.closed:
            return
        self.close_extras()
        self.closed = True

"""GPUVecEnv (reference: envs/env_wrappers.py:84-124): the numpy boundary the runners call.

Same shapes as the reference: actions (num_envs, agents, A) in; obs (num_envs, agents, D), rewards / dones /
bad_dones / exceed_time_limits (num_envs, agents, 1) out.  Three implementations of the host boundary:

  'pipelined' (default from 2 x 10^5 aircraft) upload / kernel / download of aircraft chunks overlapped on three streams by
              one native call (np_env_step_host); the copy engine moves the 88 B/aircraft observation block at 56 GB/s.
  'mapped'    (default for smaller F16 populations) ONE kernel launch that reads the actions from and writes observation
              rows, rewards and flags straight into pinned, device-mapped host memory (np_env_step_mapped): no copy
              engine, nothing to issue but the launch -- the lowest latency; but SM / TMA posted writes into host memory
              reach only ~31 GB/s on this platform, so it loses to the pipeline at large n (3.1 vs 2.3 ms at 10^6).
  'copy'      single launch, then device-to-host copies (small populations of the other kinds).

Lifetime of the returned arrays.  The reference returns fresh arrays every step (`_t2n`).  Here small populations
(< COPY_BELOW_BYTES of observations) also get fresh arrays; large ones get views of a ring of `ring` pinned buffers
(default 2: a result stays valid until the step after next) because a 95 MB host memcpy per step would cost more than
the step itself.  `copy=True` / `copy=False` forces either behaviour.
"""
import ctypes as C
import os

import numpy as np
import torch


COPY_BELOW_BYTES = 8 << 20   # observation block size below which step() returns fresh arrays by default

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
    def __init__(self, env_fns, device_tensors=False, pipeline_chunks=None, boundary=None, copy=None, ring=2):
        """device_tensors=True (SURVEY f-2): step()/reset() take and return torch CUDA tensors in the same
        (num_envs, agents, .) shapes, with no host round trip and no synchronisation -- for policies that live on
        the same GPU (the flag tensors are fresh clones, like the reference's; obs / reward are the env's persistent
        buffers, overwritten by the next step).  The default reproduces the reference's numpy boundary.
        boundary: 'mapped' | 'pipelined' | 'copy' | None (choose by plug-in and population, see the module docstring).
        pipeline_chunks: chunk count or relative chunk sizes of the 'pipelined' boundary.
        copy: return fresh numpy arrays (True), views of the pinned ring (False), or decide by size (None).
        ring: number of pinned output buffer sets the views rotate through."""
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
        self.boundary = "device" if self.device_tensors else None
        if self.device_tensors:
            return
        single_step_env = type(e).__name__ == "ControlEnv"
        is_f16 = getattr(e.model, "model_id", None) == 0
        boundary = boundary or os.environ.get("NPLANE_BOUNDARY") or None
        if boundary is None:
            # measured on B200 / PCIe Gen5 (profiles/r02_e2e_boundaries.txt): SM- or TMA-issued posted writes into host
            # memory run at ~31 GB/s, the copy engine at 56 GB/s, so large populations take the chunk pipeline; for small
            # ones the single mapped launch has the lowest latency (no separate copies to issue)
            if single_step_env and hasattr(e, "step_host") and (pipeline_chunks is not None or n >= 200_000):
                boundary = "pipelined"
            elif single_step_env and is_f16 and hasattr(e, "step_mapped"):
                boundary = "mapped"
            else:
                boundary = "copy"
        if boundary not in ("mapped", "pipelined", "copy"):
            raise ValueError(f"boundary must be 'mapped', 'pipelined' or 'copy', got {boundary!r}")
        if boundary == "mapped" and not (single_step_env and is_f16):
            raise ValueError("the 'mapped' boundary serves ControlEnv with the F16 plug-in")
        if boundary == "pipelined":
            if not single_step_env:
                raise ValueError("the 'pipelined' boundary serves ControlEnv")
            # a chunk count (equal chunks) or relative chunk sizes, e.g. (1, 2, 3, 5, 5): a small first chunk starts the
            # observation download -- the resource this boundary is bound by -- sooner
            if pipeline_chunks is None and os.environ.get("NPLANE_PIPELINE"):       # e.g. "1,2,3,4,4,4" (experiments)
                pipeline_chunks = tuple(float(x) for x in os.environ["NPLANE_PIPELINE"].split(","))
            self._chunks = pipeline_edges(n, DEFAULT_PIPELINE if pipeline_chunks is None else pipeline_chunks)
            if self._chunks is None:
                boundary = "copy"
            else:
                self._edges = (C.c_int * (len(self._chunks) + 1))(*([c[0] for c in self._chunks] + [n]))
        self.boundary = boundary
        self._copy = (n * D * 4 < COPY_BELOW_BYTES) if copy is None else bool(copy)
        self._zero_copy_actions = os.environ.get("NPLANE_MAPPED_ACTIONS", "1") != "0"
        frows = e.ld if boundary == "mapped" else n      # the mapped kernel mirrors the flag rows at the SoA pitch
        self._act_h = torch.empty((n, A), dtype=torch.float32).pin_memory()
        self._act_d = torch.empty((n, A), dtype=torch.float32, device=e.device)
        self._out = [dict(obs=torch.zeros((n, D), dtype=torch.float32).pin_memory(),
                          rew=torch.zeros(n, dtype=torch.float32).pin_memory(),
                          flags=torch.zeros((3, frows), dtype=torch.uint8).pin_memory()) for _ in range(max(1, int(ring)))]
        shp = (self.num_envs, self.agents, 1)
        for o in self._out:      # numpy views of the pinned buffers in the shapes the runners expect, built once
            fl = o["flags"].numpy().view(np.bool_)[:, :n]
            o["views"] = (o["obs"].numpy().reshape(self.num_envs, self.agents, D), o["rew"].numpy().reshape(shp),
                          fl[0].reshape(shp), fl[1].reshape(shp), fl[2].reshape(shp))

    def _next_out(self):
        o = self._out[self._flip]
        self._flip = (self._flip + 1) % len(self._out)
        return o

    def _download(self, with_rest=True):
        e = self.gpu_vec_env
        o = self._next_out()
        o["obs"].copy_(e.last_obs, non_blocking=True)
        if with_rest:
            o["rew"].copy_(e.last_reward, non_blocking=True)
            o["flags"][:, :e.n].copy_(e._flags[:, :e.n], non_blocking=True)
        torch.cuda.current_stream(e.device).synchronize()
        return o

    def _step_device(self, actions):
        e = self.gpu_vec_env
        a = torch.as_tensor(actions, device=e.device, dtype=torch.float32).reshape(self.num_envs * self.agents, -1)
        obs, rew, done, bad, exc, info = e.step(a)
        shp = (self.num_envs, self.agents, 1)
        # the reference's `self.is_done = self.is_done + done` yields new tensors every step: callers keep `dones` while stepping
        return (obs.view(self.num_envs, self.agents, e.num_observation), rew.view(shp), done.clone().view(shp),
                bad.clone().view(shp), exc.clone().view(shp), info)

    def step(self, actions):
        if self.device_tensors:
            return self._step_device(actions)
        e = self.gpu_vec_env
        a = np.asarray(actions, dtype=np.float32).reshape(self.num_envs * self.agents, -1)
        if self.boundary == "copy":
            self._act_h.copy_(torch.from_numpy(a[:, :self._A]))
            self._act_d.copy_(self._act_h, non_blocking=True)
            e.step(self._act_d)
            o = self._download()
        else:
            if a.shape[1] != self._A or not a.flags.c_contiguous:
                a = np.ascontiguousarray(a[:, :self._A])
            o = self._next_out()
            if self.boundary == "mapped":
                e.step_mapped(a, self._act_h, o["obs"], o["rew"], o["flags"], None if self._zero_copy_actions else self._act_d)
            else:
                # Aircraft are independent, so the step is pipelined chunk by chunk over three in-order streams -- upload,
                # kernels, download -- by ONE native call (issuing a chunk from Python cost ~110 us of interpreter time).
                # Same results as the single launch (one RNG counter for all chunks).
                e.step_host(a, self._act_h, self._act_d, o["obs"], o["rew"], o["flags"], self._edges, len(self._chunks))
        v = o["views"]
        if self._copy:
            return (v[0].copy(), v[1].copy(), v[2].copy(), v[3].copy(), v[4].copy(), {})
        return (v[0], v[1], v[2], v[3], v[4], {})

    def reset(self):
        e = self.gpu_vec_env
        if self.device_tensors:
            return e.reset().view(self.num_envs, self.agents, e.num_observation)
        e.reset()
        o = self._download(with_rest=False)
        obs = o["views"][0]
        return obs.copy() if self._copy else obs

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

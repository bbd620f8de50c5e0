"""GPUVecEnv (reference: envs/env_wrappers.py:84-124): the numpy boundary the runners call.

Same shapes as the reference: actions (num_envs, agents, A) in; obs (num_envs, agents, D), rewards / dones /
bad_dones / exceed_time_limits (num_envs, agents, 1) out.  Host<->device traffic goes through pinned staging
buffers; large ControlEnv populations are pipelined in aircraft chunks on side streams (upload / kernel / download
of different chunks overlap; the 88 B/aircraft observation download is what bounds this boundary), with one
synchronise per step.  The returned arrays alias pinned buffers that are reused every OTHER step
(double-buffered), so a result stays valid until the step after next.
"""
import numpy as np
import torch


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
        # pipelined boundary for large single-step envs (ControlEnv): chunks of whole pairs on side streams
        k = int(pipeline_chunks) if pipeline_chunks is not None else (4 if n >= 200_000 else 1)
        if k > 1 and hasattr(e, "step_range") and type(e).__name__ == "ControlEnv":
            edges = [min(n, 2 * ((n // 2 + k - 1) // k) * c) for c in range(k)] + [n]
            self._chunks = [(edges[c], edges[c + 1]) for c in range(k) if edges[c + 1] > edges[c]]
            self._streams = [torch.cuda.Stream(device=e.device) for _ in self._chunks]
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
        """Aircraft are independent, so the step is issued chunk by chunk on side streams: while chunk c's observations
        travel to the host (the 88 B/aircraft D2H dominates the boundary), chunk c+1 runs and chunk c+2's actions are
        staged and uploaded.  Same results as the single launch (one RNG counter for all chunks)."""
        e = self.gpu_vec_env
        o = self._out[self._flip]
        self._flip ^= 1
        main = torch.cuda.current_stream(e.device)
        start = main.record_event()
        for c, (i0, i1) in enumerate(self._chunks):
            st = self._streams[c]
            st.wait_event(start)
            self._act_h[i0:i1].copy_(torch.from_numpy(a[i0:i1, :self._A]))          # host memcpy, overlaps the GPU
            with torch.cuda.stream(st):
                self._act_d[i0:i1].copy_(self._act_h[i0:i1], non_blocking=True)
                e.step_range(self._act_d, i0, i1 - i0, advance=(c == 0))
                o["obs"][i0:i1].copy_(e.last_obs[i0:i1], non_blocking=True)
                o["rew"][i0:i1].copy_(e.last_reward[i0:i1], non_blocking=True)
                for j in range(3):
                    o["flags"][j, i0:i1].copy_(e._flags[j, i0:i1], non_blocking=True)
        for st in self._streams:
            main.wait_stream(st)
        main.synchronize()
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

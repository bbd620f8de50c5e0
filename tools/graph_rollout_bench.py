#!/usr/bin/env python
"""The reference's training population (scripts/train_heading.sh: 3 000 envs, PPO actor / critic MLP 128-128) as a device-resident
rollout: a policy in plain torch, env.step writing into the DeviceRolloutBuffer in place, masks, inserts.  us per env step
  eager   : one Python call per step (launch-bound: ~10 small kernels per step),
  graph   : the whole T-step rollout captured once in a CUDA graph (env.advance_rng(T) inside) and replayed,
and the same loop through the numpy boundary the reference's runner uses (GPUVecEnv.step, policy still on the GPU)."""
import json
import os
import sys
import time
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import ControlEnv, GPUVecEnv  # noqa: E402
from neuralplane_b200.rollout import DeviceRolloutBuffer  # noqa: E402

T = 100
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
dev = "cuda:0"
torch.manual_seed(0)
actor = torch.nn.Sequential(torch.nn.Linear(22, 128), torch.nn.ReLU(), torch.nn.Linear(128, 128), torch.nn.ReLU(), torch.nn.Linear(128, 4), torch.nn.Tanh()).to(dev)
critic = torch.nn.Sequential(torch.nn.Linear(22, 128), torch.nn.ReLU(), torch.nn.Linear(128, 128), torch.nn.ReLU(), torch.nn.Linear(128, 1)).to(dev)
args = types.SimpleNamespace(buffer_size=T, n_rollout_threads=N, gamma=0.99, use_proper_time_limits=True, use_gae=True, gae_lambda=0.95,
                             recurrent_hidden_size=1, recurrent_hidden_layers=1)
env = ControlEnv(num_envs=N, config="heading", model="F16", random_seed=0, device=dev)
buf = DeviceRolloutBuffer(args, env.num_agents, env.observation_space, env.action_space, dev)
buf.attach(env)
env.reset()


@torch.no_grad()
def rollout():
    for t in range(T):
        x = buf.obs[t].view(-1, 22)
        buf.step_env(env, actor(x), None, critic(x).view(N, 1, 1))


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / (reps * T) * 1e6


out = {"n": N, "T": T, "policy": "actor 22-128-128-4 + critic 22-128-128-1, fp32 torch"}
out["eager_us_per_step"] = round(timed(rollout, 3), 2)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    rollout()
    env.advance_rng(T)
out["graph_us_per_step"] = round(timed(g.replay, 20), 2)
# env alone under the same graph treatment (no policy): what the step kernel + masks + inserts cost
a_fix = torch.rand((N, 4), device=dev) * 2 - 1
v_fix = torch.zeros((N, 1, 1), device=dev)
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2):
    for t in range(T):
        buf.step_env(env, a_fix, None, v_fix)
    env.advance_rng(T)
out["graph_env_only_us_per_step"] = round(timed(g2.replay, 20), 2)
# the reference runner's shape: numpy in / out of the vec env every step, policy on the GPU (F16sim_runner.py:120-160)
venv = GPUVecEnv([lambda: ControlEnv(num_envs=N, config="heading", model="F16", random_seed=0, device=dev)])
obs = venv.reset()


@torch.no_grad()
def numpy_loop():
    global obs
    for t in range(T):
        x = torch.from_numpy(obs.reshape(-1, 22)).to(dev)
        a = actor(x); critic(x)
        obs = venv.step(a.cpu().numpy().reshape(N, 1, 4))[0]


out["numpy_boundary_us_per_step"] = round(timed(numpy_loop, 3), 2)
out["env_steps_per_s_graph"] = round(N / (out["graph_us_per_step"] * 1e-6))
print(json.dumps(out))

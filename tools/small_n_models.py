#!/usr/bin/env python
"""us per step (CUDA-graph replay) of the device-resident step at small populations for the other plug-ins / back-ends:
UAV (K2), F16 with the table aero back-end (K1t), 1-v-1 combat (5 sub-steps): K1c<MODE_COMBAT> (default up to 18 944 aircraft) and
K5 with 128-thread CTAs (NPLANE_COOP_PAIRS=0, what ran before)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import ControlEnv, SingleCombatEnv  # noqa: E402
dev = torch.device("cuda:0")


def timed(env, a):
    env.reset()
    for _ in range(3):
        env.step(a)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            env.step(a)
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    li = env.launch_info()
    return {"us_per_step": round(e0.elapsed_time(e1) * 1e3 / 200, 2), "launch": (li["grid"], li["block"])}


for n in (1000, 3000, 10_000, 40_000):
    row = {}
    a = torch.rand((n, 4), device=dev) * 2 - 1
    row["UAV control"] = timed(ControlEnv(num_envs=n, config="control", model="UAV", random_seed=0, device=dev), a)
    row["F16_tables heading"] = timed(ControlEnv(num_envs=n, config="heading", model="F16_tables", random_seed=0, device=dev), a)
    row["combat 1v1 (n/2 envs, 5 sub-steps)"] = timed(SingleCombatEnv(num_envs=n // 2, config="selfplay", random_seed=0, device=dev), a - 0.5)
    os.environ["NPLANE_COOP_PAIRS"] = "0"
    row["combat 1v1, K5 128-thread"] = timed(SingleCombatEnv(num_envs=n // 2, config="selfplay", random_seed=0, device=dev), a - 0.5)
    del os.environ["NPLANE_COOP_PAIRS"]
    print(n, json.dumps(row), flush=True)

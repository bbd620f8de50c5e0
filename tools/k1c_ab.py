#!/usr/bin/env python
"""us per step (CUDA-graph replay) of the device-resident heading step at small populations, for the library named by NPLANE_LIB
(same-box A/B of K1c variants)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import ControlEnv  # noqa: E402

dev = torch.device("cuda:0")
out = {}
for n in (256, 3000, 10_000, 18_944):
    best = []
    for rep in range(3):
        env = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device=dev)
        env.reset()
        a = torch.rand((n, 4), device=dev) * 2 - 1
        for _ in range(3):
            env.step(a)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                env.step(a)
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        best.append(round(e0.elapsed_time(e1) * 1e3 / 400, 2))
    out[n] = best
print(os.environ.get("NPLANE_LIB", "in-tree"), json.dumps(out))

#!/usr/bin/env python
"""How much of a K1c step is cold start (launch, image staging, first pass through ~120 KB of straight-line code) and how much
a warm pass: n = 64 * m aircraft on ONE CTA (NPLANE_COOP_GRID=1) -> m iterations of the CTA's loop."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import ControlEnv  # noqa: E402
dev = torch.device("cuda:0")
out = {}
for grid in ("1", "0"):
    os.environ["NPLANE_COOP_GRID"] = grid
    for m in (1, 2, 4, 8):
        n = 64 * m
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
        out[f"grid_cap={grid} n={n} launch={env.launch_info()['grid']}"] = round(e0.elapsed_time(e1) * 1e3 / 400, 2)
print(json.dumps(out, indent=1))

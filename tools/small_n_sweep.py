#!/usr/bin/env python
"""us per step of the device-resident F16 heading step (CUDA-graph replay) over small and mid populations, by CTA shape
(NPLANE_BLOCK = 128 / 384, K1c = the cooperative kernel forced, auto = the library's choice): where should the dispatch switch?"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import ControlEnv  # noqa: E402

dev = torch.device("cuda:0")
out = {}
for n in (256, 1000, 3000, 10_000, 18_944, 25_000, 37_888, 50_000, 75_000, 100_000, 150_000, 200_000):
    row = {}
    for blk in ("128", "384", "K1c", "auto"):
        if blk in ("128", "384"):
            os.environ["NPLANE_BLOCK"] = blk
        if blk == "K1c":
            os.environ["NPLANE_COOP_PAIRS"] = "100000000"
        env = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device=dev)
        os.environ.pop("NPLANE_BLOCK", None)
        os.environ.pop("NPLANE_COOP_PAIRS", None)
        env.reset()
        a = torch.rand((n, 4), device=dev) * 2 - 1
        for _ in range(3):
            env.step(a)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                env.step(a)
        for _ in range(2):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        row[blk] = round(e0.elapsed_time(e1) * 1e3 / 100, 2)
        row[blk + "_launch"] = (env.launch_info()["grid"], env.launch_info()["block"])
    out[n] = row
    print(n, row, flush=True)
print(json.dumps(out))

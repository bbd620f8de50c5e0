#!/usr/bin/env python
"""ms per step of ControlEnv.step (F16 heading, n = 10^6, random policy) as a function of the step index since reset():
a freshly reset population flies ~40 steps before the first episodes end, so a short timed window sees no episodic
resets at all; the stationary regime (1.8 % of the aircraft re-initialised per step) is reached after ~150 steps."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import ControlEnv  # noqa: E402

n, window, windows = 1_000_000, 20, 25
dev = torch.device("cuda:0")
env = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device=dev)
env.reset()
g = torch.Generator(device=dev).manual_seed(1)
acts = [torch.rand((n, 4), device=dev, generator=g) * 2 - 1 for _ in range(8)]
out, last = [], env.termination_counters()["resets"]
for w in range(windows):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(window):
        env.step(acts[k % 8])
    e1.record()
    torch.cuda.synchronize()
    r = env.termination_counters()["resets"]
    out.append({"steps": [w * window, (w + 1) * window], "ms_per_step": e0.elapsed_time(e1) / window,
                "resets_per_step_pct": 100.0 * (r - last) / window / n})
    last = r
print(json.dumps(out))

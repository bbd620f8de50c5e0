#!/usr/bin/env python
"""PlanningEnv.step (50 FDM sub-steps under the fused PID controller) at the populations the reference trains it at
(scripts/train_tracking.sh: 10 000 envs): ms per env step (CUDA events, eager launches) by kernel:
K1c<PLAN> (default up to 18 944 aircraft), K1<MODE_PLAN> with 128-thread CTAs (NPLANE_COOP_PAIRS=0), with 384 (NPLANE_BLOCK=384 = round 2 before)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import PlanningEnv  # noqa: E402
dev = torch.device("cuda:0")
out = {}
for n in (1000, 3000, 10_000, 18_944, 40_000):
    row = {}
    for name, var, val in (("K1c", None, None), ("K1_128", "NPLANE_COOP_PAIRS", "0"), ("K1_384", "NPLANE_BLOCK", "384")):
        if var:
            os.environ[var] = val
        env = PlanningEnv(num_envs=n, config="tracking", model="F16", random_seed=0, device=dev, n_substeps=50)
        if var:
            del os.environ[var]
        env.reset()
        a = torch.rand((n, 3), device=dev) * 2 - 1
        for _ in range(5):
            env.step(a)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40):
            env.step(a)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 40
        li = env.launch_info()
        row[name] = {"ms_per_env_step": round(ms, 4), "us_per_fdm_substep": round(ms * 1e3 / 50, 2), "fdm_steps_per_s": round(n * 50 / (ms * 1e-3)),
                     "launch": (li["grid"], li["block"])}
    out[n] = row
    print(n, json.dumps(row), flush=True)

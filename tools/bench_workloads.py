#!/usr/bin/env python
"""Side measurements of the other BASELINE configs (not the bench.py headline line):
   config 3  ControlEnv(config='control', model='UAV'), n = 10^5 and 10^6   (HBM-bound kernel: GB/s vs roofline)
   config 4  PlanningEnv(config='tracking'), n = 10^6, 50 fused sub-steps per env step
Prints one JSON object per workload.  CUDA events around K launches after W warm-ups, inputs resident in HBM."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import ControlEnv, PlanningEnv  # noqa: E402

HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def timed(env, acts, K, W):
    for k in range(W):
        env.step(acts[k % len(acts)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        env.step(acts[k % len(acts)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K


def main():
    out = []
    for n in (100_000, 1_000_000, 8_000_000):
        env = ControlEnv(num_envs=n, config="control", model="UAV", random_seed=0, device="cuda:0")
        env.reset()
        acts = [torch.rand((n, 4), device="cuda") * 2 - 1 for _ in range(4)]
        ms = timed(env, acts, 200, 20)
        gbs = 268.0 * n / (ms * 1e-3) / 1e9
        out.append({"workload": f"UAV control task, n={n}", "ms_per_step": ms, "aircraft_steps_per_s": n / (ms * 1e-3),
                    "algorithmic_bytes_per_aircraft_step": 268, "achieved_GBps": gbs, "hbm_peak_GBps": HBM, "frac": gbs / HBM,
                    "note": "n = 10^5 (27 MB) and 10^6 (268 MB x 2 directions) partly live in the 126 MB L2; 8x10^6 is HBM-resident"})
        del env
    n = 1_000_000
    env = PlanningEnv(num_envs=n, config="tracking", random_seed=0, device="cuda:0")
    env.reset()
    acts = [torch.rand((n, 3), device="cuda") * 2 - 1 for _ in range(4)]
    ms = timed(env, acts, 8, 2)
    out.append({"workload": "PlanningEnv tracking, F16, fused PID low level, n=10^6, 50 sub-steps per env step", "ms_per_env_step": ms,
                "aircraft_fdm_steps_per_s": 50 * n / (ms * 1e-3), "env_steps_per_s": n / (ms * 1e-3)})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Side measurement of the device-resident rollout buffer (SURVEY f-2): the return scan against the HBM roofline and a
full zero-copy collection loop (env.step + masks) at n = 10^6.  One JSON object per line."""
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import ControlEnv  # noqa: E402
from neuralplane_b200.rollout import DeviceRolloutBuffer  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def ev():
    return torch.cuda.Event(enable_timing=True)


def main():
    T, N = 64, 1_000_000
    for use_gae, proper, bytes_per in ((True, True, 20), (True, False, 16), (False, False, 12)):
        a = types.SimpleNamespace(buffer_size=T, n_rollout_threads=N, gamma=0.99, use_proper_time_limits=proper, use_gae=use_gae,
                                  gae_lambda=0.95, recurrent_hidden_size=1, recurrent_hidden_layers=1)
        buf = DeviceRolloutBuffer(a, 1, 1, 1, "cuda:0")       # obs / action width 1: only the scalar rows matter here
        buf.rewards.normal_(); buf.value_preds.normal_()
        buf.masks.copy_((torch.rand_like(buf.masks) > 0.02).float()); buf.bad_masks.copy_((torch.rand_like(buf.masks) > 0.02).float())
        nxt = torch.zeros((N, 1, 1), device="cuda")
        for _ in range(3):
            buf.compute_returns(nxt)
        torch.cuda.synchronize()
        e0, e1 = ev(), ev(); e0.record()
        for _ in range(20):
            buf.compute_returns(nxt)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        gbs = bytes_per * T * N / ms / 1e6
        print(json.dumps({"workload": f"np_rollout_returns T={T} M={N} use_gae={use_gae} proper={proper}", "ms": ms,
                          "algorithmic_bytes_per_element": bytes_per, "achieved_GBps": gbs, "hbm_peak_GBps": HBM, "frac": gbs / HBM,
                          "elements_per_s": T * N / ms * 1e3}))
        del buf
    T, N = 16, 1_000_000
    env = ControlEnv(num_envs=N, config="heading", model="F16", random_seed=0, device="cuda:0")
    a = types.SimpleNamespace(buffer_size=T, n_rollout_threads=N, gamma=0.99, use_proper_time_limits=True, use_gae=True,
                              gae_lambda=0.95, recurrent_hidden_size=1, recurrent_hidden_layers=1)
    buf = DeviceRolloutBuffer(a, env.num_agents, env.observation_space, env.action_space, "cuda:0")
    buf.attach(env); env.reset()
    acts = [torch.rand((N, 4), device="cuda") * 2 - 1 for _ in range(4)]
    val = torch.zeros((N, 1, 1), device="cuda")
    for k in range(T):
        buf.step_env(env, acts[k % 4], val, val)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev(); e0.record()
    for it in range(10):
        for k in range(T):
            buf.step_env(env, acts[k % 4], val, val)
        buf.compute_returns(val); buf.after_update()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (10 * T)
    print(json.dumps({"workload": f"zero-copy collection loop: env.step + masks + insert of actions/logp/values, T={T}, n={N}; "
                                  "compute_returns + after_update per T steps included", "ms_per_step": ms, "aircraft_steps_per_s": N / ms * 1e3}))


if __name__ == "__main__":
    main()

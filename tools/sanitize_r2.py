#!/usr/bin/env python
"""Driver for compute-sanitizer over the kernels added in round 2 (small sizes; run under memcheck / racecheck / synccheck):
K1 with the staged observation tile + TMA bulk store (384 / 512 / 128-thread shapes, all three tasks), K1t (table back-end, one
aircraft per thread), the mapped host boundary, model.update() kernels, the role-sharded combat halves, 2-v-2 combat,
planning / combat on the table back-end."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import ControlEnv, GPUVecEnv, MultipleCombatEnv, PlanningEnv, SingleCombatEnv  # noqa: E402
from neuralplane_b200.combat_exchange import LocalPairExchange  # noqa: E402

dev = "cuda:0"
for n in (257, 3000 + 11, 60_000 + 1):          # 128-thread latency shape, 128, 384; odd tails
    for task in ("heading", "control", "tracking"):
        for model in ("F16", "F16_tables"):
            e = ControlEnv(num_envs=n, config=task, model=model, random_seed=1, device=dev)
            e.reset()
            for k in range(3):
                e.step(torch.rand((n, 4), device=dev) * 2 - 1)
            e.model.update(torch.rand((n, 4), device=dev) * 2 - 1)
            torch.cuda.synchronize()
# n = 257 / 3011 above ran on K1c (the cooperative small-population kernel); K1c over several waves (tile / slot reuse across
# iterations), and K1's 128-thread shape it replaced at these sizes
for var, val, n in (("NPLANE_COOP_PAIRS", "100000000", 40_001), ("NPLANE_COOP_PAIRS", "0", 3011)):
    os.environ[var] = val
    e = ControlEnv(num_envs=n, config="tracking", model="F16", random_seed=1, device=dev)
    del os.environ[var]
    e.reset()
    for k in range(3):
        e.step(torch.rand((n, 4), device=dev) * 2 - 1)
    torch.cuda.synchronize()
os.environ["NPLANE_BLOCK"] = "512"
e = ControlEnv(num_envs=5000, config="heading", model="F16", random_seed=1, device=dev)
del os.environ["NPLANE_BLOCK"]
e.reset()
for k in range(3):
    e.step(torch.rand((5000, 4), device=dev) * 2 - 1)
u = ControlEnv(num_envs=1000, config="control", model="UAV", random_seed=1, device=dev)
u.reset()
u.model.update(torch.rand((1000, 4), device=dev))
for model in ("F16", "F16_tables"):
    n = 3000 + 11
    for boundary in ("mapped", "pipelined", "copy"):
        v = GPUVecEnv([lambda: ControlEnv(num_envs=n, config="heading", model=model, random_seed=2, device=dev)], boundary=boundary,
                      pipeline_chunks=(1, 2, 3))
        v.reset()
        for k in range(3):
            out = v.step(np.random.rand(n, 1, 4).astype(np.float32) * 2 - 1)
        assert np.isfinite(out[0]).all()
E = 1000
mk = lambda r: SingleCombatEnv(num_envs=E, config="selfplay", random_seed=3, device=dev, layout="role", role=r, first_env=10)
e0, e1 = mk(0), mk(1)
ex = LocalPairExchange(e0, e1)
ex.reset_both()
for k in range(3):
    ex.step_both(torch.rand((E, 4), device=dev) - 0.5, torch.rand((E, 4), device=dev) - 0.5)
m = MultipleCombatEnv(num_envs=500, random_seed=1, device=dev)
m.reset()
for k in range(3):
    m.step(torch.rand((m.n, 4), device=dev) - 0.5)
for n_plan, var in ((513, None), (12_001, None), (513, "NPLANE_COOP_PAIRS")):   # K1c<PLAN> with 8 / 4 warps, K1<128, MODE_PLAN>
    if var:
        os.environ[var] = "0"
    p = PlanningEnv(num_envs=n_plan, config="tracking", model="F16", random_seed=1, device=dev, n_substeps=3)
    if var:
        del os.environ[var]
    p.reset()
    for k in range(2):
        p.step(torch.rand((n_plan, 3), device=dev) - 0.5)
p = PlanningEnv(num_envs=513, config="tracking", model="F16_tables", random_seed=1, device=dev, n_substeps=3)
p.reset()
p.step(torch.rand((513, 3), device=dev) - 0.5)
for E_c, var in ((300, None), (6001, None), (300, "NPLANE_COOP_PAIRS")):   # K1c<MODE_COMBAT> with 8 / 4 warps, K5 with 128-thread CTAs
    if var:
        os.environ[var] = "0"
    c = SingleCombatEnv(num_envs=E_c, config="selfplay", random_seed=1, device=dev)
    if var:
        del os.environ[var]
    c.reset()
    for k in range(2):
        c.step(torch.rand((2 * E_c, 4), device=dev) - 0.5)
c = SingleCombatEnv(num_envs=300, config="selfplay", model="F16_tables", random_seed=1, device=dev)
c.reset()
c.step(torch.rand((600, 4), device=dev) - 0.5)
torch.cuda.synchronize()
print("sanitize_r2 done")

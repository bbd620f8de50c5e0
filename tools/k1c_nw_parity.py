#!/usr/bin/env python
"""Bit-identity of a forced K1c CTA shape (NPLANE_COOP_WARPS) against K1, all three modes, a few steps with resets."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuralplane_b200 import ControlEnv, PlanningEnv, SingleCombatEnv  # noqa: E402
nw = sys.argv[1]
for name, mk, n, A in (("control", lambda: ControlEnv(num_envs=3001, config="control", model="F16", random_seed=6, device="cuda:0"), 3001, 4),
                       ("planning", lambda: PlanningEnv(num_envs=3001, config="tracking", model="F16", random_seed=6, device="cuda:0", n_substeps=5), 3001, 3),
                       ("combat", lambda: SingleCombatEnv(num_envs=1500, config="selfplay", random_seed=6, device="cuda:0"), 3000, 4)):
    os.environ["NPLANE_COOP_PAIRS"] = "0"
    ref = mk()
    del os.environ["NPLANE_COOP_PAIRS"]
    os.environ["NPLANE_COOP_WARPS"] = nw
    coop = mk()
    del os.environ["NPLANE_COOP_WARPS"]
    assert torch.equal(coop.reset(), ref.reset())
    g = torch.Generator(device="cuda").manual_seed(1)
    for k in range(12):
        a = torch.rand((n, A), device="cuda", generator=g) * 2 - 1
        for x, y in zip(coop.step(a)[:5], ref.step(a)[:5]):
            assert torch.equal(x, y), (name, k)
        if k % 4 == 3:
            for e in (coop, ref):
                e.is_done[::5] = True
    assert torch.equal(coop.model.s, ref.model.s) and coop.termination_counters() == ref.termination_counters()
    print(name, "NW", nw, "bit-identical; launch", coop.launch_info())

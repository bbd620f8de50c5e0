"""The combat oracle (oracle/combat_oracle.py) against fixtures produced by the reference's own obs / reward /
geometry / termination / model / controller code (tests/golden/make_golden.py RefCombat)."""
import os

import numpy as np
import pytest
import torch

from oracle import tapes
from oracle.combat_oracle import CombatOracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("fixture,close,steps", [("combat_traj.npz", False, 12), ("combat_close_traj.npz", True, 8)])
def test_combat_trajectory_matches_reference_components(fixture, close, steps):
    g = np.load(os.path.join(GOLDEN, fixture))
    num_envs, _, seed = [int(x) for x in g["meta"]]
    o = CombatOracle(num_envs)
    n = o.n
    obs0 = o.reset(torch.from_numpy(tapes.reset_draw_tape(seed, 0, n)))
    assert np.array_equal(obs0.numpy(), g["obs0"])
    o.s = torch.from_numpy(g["s_start"].copy()); o.blood = torch.from_numpy(g["blood_start"].copy())
    events = 0
    for k in range(1, steps + 1):
        obs, rew, done, bad, exc = o.step(torch.from_numpy(tapes.action_tape(seed, k, n, 0.2 if close else 1.0)),
                                          torch.from_numpy(tapes.reset_draw_tape(seed, k, n)))
        assert np.array_equal(o.s.numpy(), g[f"k{k}_s"]), k
        assert np.array_equal(o.u.numpy(), g[f"k{k}_u"]), k
        assert np.array_equal(o.ctrl_state().numpy(), g[f"k{k}_ctrl"]), k
        assert np.array_equal(obs.numpy(), g[f"k{k}_obs"]), k
        assert np.array_equal(rew.numpy(), g[f"k{k}_reward"]), k
        assert np.array_equal(o.blood.numpy(), g[f"k{k}_blood"]), k
        assert np.array_equal(bad.numpy(), g[f"k{k}_bad"]) and np.array_equal(done.numpy(), g[f"k{k}_done"]), k
        assert np.array_equal(exc.numpy(), g[f"k{k}_exc"]), k
        assert np.array_equal(o.step_count.numpy().astype(np.int32), g[f"k{k}_step_count"]), k
        events += int(bad.sum()) + int(done.sum())
    assert (events > 0) == close

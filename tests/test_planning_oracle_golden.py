"""The planning oracle (oracle/planning_oracle.py: PlanningEnv.step + PID low-level controller) against the fixture
produced by running the reference's own F16Model / TrackingTask / termination / reward / PID controller classes
(tests/golden/make_golden.py RefPidPlanner)."""
import os

import numpy as np
import torch

from oracle import tapes
from oracle.planning_oracle import PlanningOracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_planning_trajectory_matches_reference_components():
    g = np.load(os.path.join(GOLDEN, "planning_pid_traj.npz"))
    n, steps, seed = [int(x) for x in g["meta"]]
    steps = min(steps, 6)                      # 300 reference sub-steps keep the CPU suite short
    o = PlanningOracle(n)
    obs0 = o.reset(torch.from_numpy(tapes.reset_draw_tape(seed, 0, n)))
    assert np.array_equal(obs0.numpy(), g["obs0"])
    for k in range(1, steps + 1):
        obs, rew, done, bad, exc = o.plan_step(torch.from_numpy(tapes.action_tape(seed, k, n, float(g["scale"]), num_actions=3)),
                                               torch.from_numpy(tapes.reset_draw_tape(seed, k, n)))
        assert np.array_equal(o.targets.numpy(), g[f"k{k}_targets"]), k
        assert np.array_equal(o.s.numpy(), g[f"k{k}_s"]), k
        assert np.array_equal(o.u.numpy(), g[f"k{k}_u"]), k
        assert np.array_equal(o.pid_state().numpy(), g[f"k{k}_pid"]), k
        assert np.array_equal(obs.numpy(), g[f"k{k}_obs"]), k
        assert np.array_equal(rew.numpy(), g[f"k{k}_reward"]), k
        assert np.array_equal(bad.numpy(), g[f"k{k}_bad"]) and np.array_equal(done.numpy(), g[f"k{k}_done"]), k
        assert np.array_equal(o.step_count.numpy().astype(np.int32), g[f"k{k}_step_count"]), k
    assert int(sum(g[f"k{k}_bad"].sum() for k in range(1, steps + 1))) > 0

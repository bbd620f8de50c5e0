"""The UAV oracle (oracle/uav_oracle.py) against fixtures generated from the unmodified reference's
ControlEnv(model='UAV') (tests/golden/make_golden.py uav): bit-exact in float32 on the same torch build."""
import os

import numpy as np
import pytest
import torch

from oracle import tapes
from oracle.uav_oracle import UAVEnvOracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("task,fixture", [("control", "uav_control_traj.npz"), ("heading", "uav_heading_traj.npz")])
def test_uav_trajectory_is_bit_exact(task, fixture):
    g = np.load(os.path.join(GOLDEN, fixture))
    n, steps, seed = [int(x) for x in g["meta"]]
    scale = float(g["scale"])
    o = UAVEnvOracle(n, task)
    obs0 = o.reset(torch.from_numpy(tapes.reset_draw_tape(seed, 0, n)))
    assert np.array_equal(obs0.numpy(), g["obs0"])
    checked = 0
    for k in range(1, steps + 1):
        obs, rew, done, bad, exc = o.step(torch.from_numpy(tapes.action_tape(seed, k, n, scale)),
                                          torch.from_numpy(tapes.reset_draw_tape(seed, k, n)))
        if f"k{k}_s" not in g.files:
            continue
        checked += 1
        assert np.array_equal(o.s.numpy(), g[f"k{k}_s"]), k
        assert np.array_equal(o.u.numpy(), g[f"k{k}_u"]), k
        assert np.array_equal(o.tgt.numpy(), g[f"k{k}_tgt"]), k
        assert np.array_equal(obs.numpy(), g[f"k{k}_obs"]), k
        assert np.array_equal(rew.numpy(), g[f"k{k}_reward"]), k
        assert np.array_equal(bad.numpy(), g[f"k{k}_bad"]) and np.array_equal(done.numpy(), g[f"k{k}_done"]), k
        assert np.array_equal(o.step_count.numpy().astype(np.int32), g[f"k{k}_step_count"]), k
    assert checked >= 10 and int(g["n_bad"].sum()) > 0      # the fixture exercises termination + reset

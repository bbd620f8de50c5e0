"""Flight metrics (SURVEY f-4; reference renders/evaluate_result.py:29-43): the host reduction against the values the
reference's own lines produce on the reference's own recording (tests/golden/eval_metrics.npz, make_golden.py), and
the device-side recorder against per-step getters."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_flight_metrics_match_reference_lines():
    from neuralplane_b200.envs.metrics import METRIC_NAMES, flight_metrics
    g = np.load(os.path.join(GOLDEN, "eval_metrics.npz"))
    rec = np.load(os.path.join(GOLDEN, "ref_recorded_trajectory.npz"))
    head = int(g["head"][0])
    got = flight_metrics(*(rec[k][:head] for k in ("G", "vt", "pitch", "alpha", "beta", "altitude")))
    assert tuple(g["names"]) == METRIC_NAMES
    for name, want in zip(g["names"], g["values"]):
        assert abs(got[str(name)] - want) <= 1e-12 * max(1.0, abs(want)), (name, got[str(name)], want)


def test_flight_metrics_shapes_and_ranges():
    from neuralplane_b200.envs.metrics import flight_metrics
    r = np.random.default_rng(0)
    m = flight_metrics(r.uniform(0.5, 2, (50, 4)), r.uniform(900, 1100, (50, 4)), r.uniform(-.1, .1, (50, 4)),
                       r.uniform(0, .2, (50, 4)), r.uniform(-.05, .05, (50, 4)), r.uniform(15000, 20000, (50, 4)))
    assert set(m) == {"G", "TAS", "RoC", "AOA", "ASM", "SSM", "OSM", "AOASM", "AOSSM"}
    assert abs(m["G"] + m["OSM"] - 1.0) < 1e-12          # the two overload figures are complements by construction
    assert all(np.isfinite(v) for v in m.values())


@pytest.mark.gpu
def test_flight_recorder_matches_per_step_getters():
    import torch
    from neuralplane_b200 import ControlEnv
    from neuralplane_b200.envs.metrics import FlightRecorder, flight_metrics
    n, m, steps = 256, 8, 40
    env = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=3, device="cuda:0")
    env.reset()
    rec = FlightRecorder(env, max_aircraft=m)
    G, vt, alt = [], [], []
    g = torch.Generator(device="cuda").manual_seed(0)
    for _ in range(steps):
        env.step((torch.rand((n, 4), device="cuda", generator=g) * 2 - 1) * 0.3)
        rec.record()
        G.append(env.model.get_G()[:m].cpu().numpy()); vt.append(env.model.get_vt()[:m].cpu().numpy())
        alt.append(env.model.get_position()[2][:m].cpu().numpy())
    s = rec.series()
    assert s["G"].shape == (steps, m) and len(rec) == steps
    assert np.allclose(s["G"], np.stack(G), rtol=1e-6, atol=1e-6)        # one batched nlplant == 40 per-step ones
    assert np.array_equal(s["vt"], np.stack(vt)) and np.array_equal(s["altitude"], np.stack(alt))
    want = flight_metrics(np.stack(G), s["vt"], s["pitch"], s["alpha"], s["beta"], s["altitude"])
    got = rec.metrics()
    assert all(abs(got[k] - want[k]) < 1e-6 for k in want)

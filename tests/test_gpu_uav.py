"""GPU parity of the UAV plug-in (uav_env_kernel / np_uav_nlplant through the C ABI) against the reference fixtures
(tests/golden/uav_*_traj.npz) and the live oracle.  fp32 tolerance: single steps 2e-6 relative (the kernel's
sincosf/tanf/powf differ from torch's CPU libm by an ulp); along the 400-step fixtures the median state error must
stay below 1e-5 (north_star's figure) on aircraft with the same reset history."""
import os

import numpy as np
import pytest
import torch

from oracle import tapes

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NOM = np.array([300, 300, 6000, 0.5, 0.5, 1.0, 300, 5, 5, 0.5, 0.5, 0.5], dtype=np.float64) * 1e-3   # error floors (SI)


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _env(n, task):
    from neuralplane_b200 import ControlEnv
    env = ControlEnv(num_envs=n, config=task, model="UAV", random_seed=0, device="cuda:0")
    env.task.noise_scale = 0.0
    return env


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return (np.abs(a - b) / (np.abs(b) + NOM)).max(axis=1)


@pytest.mark.parametrize("task,fixture", [("control", "uav_control_traj.npz"), ("heading", "uav_heading_traj.npz")])
def test_uav_trajectory_vs_reference_fixture(task, fixture):
    g = np.load(os.path.join(GOLDEN, fixture))
    n, steps, seed = [int(x) for x in g["meta"]]
    scale = float(g["scale"])
    env = _env(n, task)
    obs0 = env.reset(reset_draws=_cuda(tapes.reset_draw_tape(seed, 0, n)))
    assert np.allclose(obs0.cpu().numpy(), g["obs0"], rtol=1e-6, atol=1e-7)
    n_bad = 0
    for k in range(1, steps + 1):
        obs, rew, done, bad, exc, _ = env.step(_cuda(tapes.action_tape(seed, k, n, scale)),
                                               reset_draws=_cuda(tapes.reset_draw_tape(seed, k, n)))
        n_bad += int(bad.sum())
        if f"k{k}_s" not in g.files:
            continue
        # The reference's UAV spins up at a constant 1 rad/s^2 about all three axes (UAV_dynamics.py:72-78), so its
        # Euler-angle kinematics cross the theta = -pi/2 singularity near step 105 and again near 180: the
        # reference's OWN 1-ulp twin is 4e-4 away at step 100 and O(1) away by step 200 (measured with the oracle).
        # Every aircraft then overloads around step 250 and is re-initialised from the shared draw tape, which
        # re-synchronises the runs.  Bars: strict up to step 50 and after the common reset, 2e-3 at step 100, no state
        # bar at step 200.
        same = env.step_count.cpu().numpy() == g[f"k{k}_step_count"]
        if k == 200:
            continue
        assert same.mean() >= (0.9 if k <= 100 else 0.5), (k, same.mean())
        err = _rel(env.model.s.cpu().numpy()[same], g[f"k{k}_s"][same])
        bar = 2e-3 if k == 100 else 1e-5
        assert np.median(err) <= bar and err.max() <= 100 * bar, (k, np.median(err), err.max())
        assert np.allclose(env.model.u.cpu().numpy()[same], g[f"k{k}_u"][same], rtol=1e-6, atol=1e-3)
        assert np.median(np.abs(obs.cpu().numpy()[same] - g[f"k{k}_obs"][same]).max(axis=1)) <= (1e-2 if k == 100 else 1e-4)
        if k <= 50:      # well before the singularity every flag / reward agrees
            assert same.all()
            assert np.array_equal(bad.cpu().numpy(), g[f"k{k}_bad"]) and np.array_equal(done.cpu().numpy(), g[f"k{k}_done"])
            assert np.allclose(rew.cpu().numpy(), g[f"k{k}_reward"], rtol=1e-5, atol=1e-5)
    ref_bad = int(g["n_bad"].sum())
    assert abs(n_bad - ref_bad) <= max(3, 0.1 * ref_bad), (n_bad, ref_bad)


@pytest.mark.parametrize("task", ["heading", "control", "tracking"])
def test_uav_single_step_vs_oracle(task):
    """One step from random states / forces / targets: state, obs, reward and flags against the live oracle."""
    from oracle.uav_oracle import UAVEnvOracle
    n = 20001      # odd on purpose
    rng = np.random.RandomState(7)
    s = np.zeros((n, 12), np.float32)
    s[:, :2] = rng.uniform(-3e4, 3e4, (n, 2)); s[:, 2] = rng.uniform(500, 9000, n)
    s[:, 3:6] = rng.uniform(-1.2, 1.2, (n, 3)); s[:, 6] = rng.uniform(100, 450, n); s[:, 7:9] = rng.uniform(-30, 30, (n, 2))
    s[:, 9:12] = rng.uniform(-2, 2, (n, 3))
    F = rng.uniform(-20000, 20000, (n, 3)).astype(np.float32)
    env, orc = _env(n, task), UAVEnvOracle(n, task)
    d0 = tapes.reset_draw_tape(3, 0, n)
    env.reset(reset_draws=_cuda(d0)); orc.reset(torch.from_numpy(d0))
    tgt = orc.tgt.numpy().copy()
    tgt += rng.uniform(-1, 1, tgt.shape).astype(np.float32) * (np.abs(tgt) * 0.01 + 0.01)
    steps = rng.choice([5, 299, 300, 2498, 2499, 2600], n).astype(np.int64)
    env.model.s[:] = _cuda(s); env.model.u[:] = _cuda(F); env._tgt[:, :n] = _cuda(tgt.T.copy()); env.step_count[:] = _cuda(steps.astype(np.int32))
    orc.s = torch.from_numpy(s.copy()); orc.u = torch.from_numpy(F.copy()); orc.tgt = torch.from_numpy(tgt.copy())
    orc.step_count = torch.from_numpy(steps.copy())
    a, d = tapes.action_tape(3, 1, n, 1.0), tapes.reset_draw_tape(3, 1, n)
    obs, rew, done, bad, exc, _ = env.step(_cuda(a), reset_draws=_cuda(d))
    o_obs, o_rew, o_done, o_bad, o_exc = orc.step(torch.from_numpy(a), torch.from_numpy(d))
    err = _rel(env.model.s.cpu().numpy(), orc.s.numpy())
    assert np.percentile(err, 99) <= 2e-6 and err.max() <= 1e-4, (np.percentile(err, 99), err.max())
    assert np.allclose(obs.cpu().numpy(), o_obs.numpy(), rtol=1e-5, atol=2e-6)
    near = np.abs(orc.last_accel.numpy() - 300.0) / 300.0 < 1e-5
    assert ((bad.cpu().numpy() != o_bad.numpy()) & ~near).sum() <= 2
    assert (done.cpu().numpy() != o_done.numpy()).sum() <= 2
    ok = bad.cpu().numpy() == o_bad.numpy()
    assert np.allclose(rew.cpu().numpy()[ok], o_rew.numpy()[ok], rtol=1e-5, atol=1e-5)
    assert 0 < int(o_bad.sum()) < n
    assert np.array_equal(env.step_count.cpu().numpy(), orc.step_count.numpy().astype(np.int32))
    assert not env.model._s[:, n:].any()                      # row padding untouched


def test_uav_getters_vs_oracle():
    from oracle.uav_oracle import UAVEnvOracle, uav_nlplant
    n = 4096
    env, orc = _env(n, "control"), UAVEnvOracle(n, "control")
    d0 = tapes.reset_draw_tape(9, 0, n)
    env.reset(reset_draws=_cuda(d0)); orc.reset(torch.from_numpy(d0))
    for k in range(1, 30):
        a, d = tapes.action_tape(9, k, n, 1.0), tapes.reset_draw_tape(9, k, n)
        env.step(_cuda(a), reset_draws=_cuda(d)); orc.step(torch.from_numpy(a), torch.from_numpy(d))
    orc.s = env.model.s.cpu().clone(); orc.u = env.model.u.cpu().clone()
    xd = env.model.get_extended_state().cpu().numpy()
    ref = uav_nlplant(orc.s, orc.u).numpy()
    assert np.allclose(xd, ref, rtol=2e-6, atol=1e-5)
    for got, want in zip(env.model.get_acceleration(), orc.acceleration()):
        assert np.allclose(got.cpu().numpy(), want.numpy(), rtol=1e-5, atol=1e-3)
    assert np.allclose(env.model.get_vt().cpu().numpy(), orc.vt().numpy(), rtol=1e-6)
    assert np.allclose(env.model.get_EAS2TAS().cpu().numpy(), orc.eas2tas().numpy(), rtol=1e-6)
    assert env.model.u.shape == (n, 3) and env.model.get_AOA().abs().max() == 0


def _snapshot(env):
    return [t.clone() for t in (env.model._s, env.model._u, env._tgt, env._step_count, env._flags, env._obs, env._reward)]


@pytest.mark.parametrize("task", ["control", "heading"])
def test_uav_slab_kernel_is_bit_identical_to_per_thread_kernel(task, monkeypatch):
    """The TMA-staged slab kernel (uav_step_slab_kernel) and the per-thread kernel share one arithmetic body: with the
    in-kernel Philox resets and observation noise on, 40 steps of a ragged population (full slabs + a 99-aircraft tail)
    must agree bit for bit in every buffer, as must a step issued as aligned and unaligned ranges."""
    from neuralplane_b200 import ControlEnv
    n = 148 * 4 * 256 * 2 + 99     # more slabs than resident CTAs: every CTA pipelines >= 2 slabs
    envs = []
    for _ in range(2):
        e = ControlEnv(num_envs=n, config=task, model="UAV", random_seed=5, device="cuda:0")
        e.reset()
        envs.append(e)
    g = torch.Generator(device="cuda").manual_seed(11)
    for k in range(40):
        a = torch.rand((n, 4), device="cuda", generator=g) * 2 - 1
        monkeypatch.delenv("NPLANE_UAV_SCALAR", raising=False)
        envs[0].step(a)
        assert envs[0].launch_info()["smem_bytes"] > 40000, "slab kernel did not run"
        monkeypatch.setenv("NPLANE_UAV_SCALAR", "1")
        envs[1].step(a)
        assert envs[1].launch_info()["smem_bytes"] == 0
        for x, y in zip(_snapshot(envs[0]), _snapshot(envs[1])):
            assert torch.equal(x, y), k
    if task == "control":
        assert envs[0].termination_counters()["resets"] > n   # resets happened inside the slab kernel
    # one logical step as three ranges: slab path (aligned), per-thread fallback (first % 16 != 0), slab path with a tail
    monkeypatch.delenv("NPLANE_UAV_SCALAR", raising=False)
    a = torch.rand((n, 4), device="cuda", generator=g) * 2 - 1
    cut1, cut2 = 4096 + 2, 4096 + 2 + 1022
    envs[0].step_range(a, 0, cut1, advance=True)
    envs[0].step_range(a, cut1, cut2 - cut1, advance=False)
    envs[0].step_range(a, cut2, n - cut2, advance=False)
    monkeypatch.setenv("NPLANE_UAV_SCALAR", "1")
    envs[1].step(a)
    torch.cuda.synchronize()
    for x, y in zip(_snapshot(envs[0]), _snapshot(envs[1])):
        assert torch.equal(x, y)


def test_uav_full_size_sampled_parity_and_invariants():
    """BASELINE config 3 and beyond (n = 10^6 through the slab kernel): aircraft are independent, so a random sample of
    the population is replayed through the oracle from the same pre-step state; plus population-wide invariants."""
    from oracle.uav_oracle import UAVEnvOracle
    n, m = 1_000_000, 4096
    env = _env(n, "control")
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(30):
        env.step(torch.rand((n, 4), device="cuda", generator=g) * 2 - 1)
    assert env.launch_info()["smem_bytes"] > 40000                # the TMA-staged slab kernel ran
    idx = torch.from_numpy(np.sort(np.random.default_rng(0).choice(n, m, replace=False))).cuda()
    orc = UAVEnvOracle(m, "control")
    orc.s = env.model.s[idx].cpu(); orc.u = env.model.u[idx].cpu()
    orc.tgt = env._tgt[:, idx].t().contiguous().cpu()
    orc.step_count = env.step_count[idx].cpu().long()
    orc.is_done = env.is_done[idx].cpu().clone(); orc.bad_done = env.bad_done[idx].cpu().clone()
    orc.exceed_time_limit = env.exceed_time_limit[idx].cpu().clone()
    a = torch.rand((n, 4), device="cuda", generator=g) * 2 - 1
    d = torch.rand((n, 5), device="cuda", generator=g)
    obs, rew, done, bad, exc, _ = env.step(a, reset_draws=d)
    o_obs, o_rew, o_done, o_bad, o_exc = orc.step(a[idx].cpu(), d[idx].cpu())
    err = _rel(env.model.s[idx].cpu().numpy(), orc.s.numpy())
    assert np.percentile(err, 99) <= 2e-6 and err.max() <= 1e-4, (np.percentile(err, 99), err.max())
    assert np.allclose(obs[idx].cpu().numpy(), o_obs.numpy(), rtol=1e-5, atol=2e-6)
    near = np.abs(orc.last_accel.numpy() - 300.0) / 300.0 < 1e-5
    assert ((bad[idx].cpu().numpy() != o_bad.numpy()) & ~near).sum() <= 1
    assert np.array_equal(env.step_count[idx].cpu().numpy(), orc.step_count.numpy().astype(np.int32))
    assert torch.isfinite(env.model.s).all() and torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert int(env._flags.max()) <= 1 and not bool(exc.any())
    assert env.termination_counters()["resets"] >= n

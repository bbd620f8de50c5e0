"""GPU parity tests: the CUDA path (through the C ABI / the Python env mirror) against
  * golden fixtures generated from the unmodified reference (tests/golden/*.npz), and
  * the CPU oracle (oracle/f16_oracle.py) evaluated live on the same seeded tapes.

Tolerances (fp32; metric in tests/_metrics.py): every bar is stated against the float64 evaluation of the same
formulas ("truth"): the CUDA path must be no further from truth than 2x (single evaluations) / 1.5x (trajectories)
the reference's own fp32 arithmetic is, and within max(1e-5, 1.5 x that figure) of the reference itself along
1000-step trajectories (north_star's 1e-5 is where the reference's fp32 sits from truth after 1000 steps, so it
is a noise floor, not a margin -- see _trajectory).  Flags are compared exactly on aircraft whose history matches.
"""
import os

import numpy as np
import pytest
import torch

from oracle import tapes
from _metrics import state_rel_err

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_err(a, b, floor):
    return np.abs(a - b) / (np.abs(b) + floor)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


@pytest.fixture(params=["K1c", "K1"])
def step_kernel(request, monkeypatch):
    """Populations up to 18 944 aircraft step on K1c (coop_step_kernel.cuh) by default; the tests that hold the step against the
    oracle / the reference fixtures at such sizes run once on it and once on K1 (the kernel of the headline number)."""
    if request.param == "K1":
        monkeypatch.setenv("NPLANE_COOP_PAIRS", "0")
    return request.param


def _env(n, task="heading", **kw):
    from neuralplane_b200 import ControlEnv
    env = ControlEnv(num_envs=n, config=task, model="F16", random_seed=0, device="cuda:0", **kw)
    env.task.noise_scale = 0.0
    return env


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


# --------------------------------------------------------------------------------------------------------
# single evaluations against the reference's known answers
# --------------------------------------------------------------------------------------------------------
def test_coefficient_nets_kat(dev):
    """np_f16_coeffs: all 43 MLPs vs the reference's hifi_F16 methods on 2048 random (alpha, beta, el)."""
    import ctypes as C
    from neuralplane_b200 import _native as nv
    from neuralplane_b200.aero import get_aero
    g = np.load(os.path.join(GOLDEN, "f16_nlplant_kat.npz"))
    n, seed = [int(x) for x in g["meta"]]
    s, u = tapes.random_envelope_states(seed, n)
    r2d = np.float32(180.0 / np.pi)
    a, b, e = _cuda(s[:, 7] * r2d), _cuda(s[:, 8] * r2d), _cuda(u[:, 1])
    out = torch.zeros((43, n), device=dev)
    aero = get_aero(dev)
    nv.check(nv.lib().np_f16_coeffs(aero.handle, a.data_ptr(), b.data_ptr(), e.data_ptr(), out.data_ptr(), n, n,
                                    torch.cuda.current_stream().cuda_stream), "np_f16_coeffs")
    got = out.t().cpu().numpy()
    scale = np.abs(g["coefs"]).max(axis=0)
    err = np.abs(got - g["coefs"]) / scale
    assert err.max() < 2e-6, (err.max(), aero.names[int(err.max(axis=0).argmax())])


def test_nlplant_and_getters_kat(dev):
    """np_f16_nlplant + the F16Model getters built on it vs the reference F16Dynamics / F16Model."""
    g = np.load(os.path.join(GOLDEN, "f16_nlplant_kat.npz"))
    n, seed = [int(x) for x in g["meta"]]
    s, u = tapes.random_envelope_states(seed, n)
    env = _env(n)
    env.model.s[:] = _cuda(s)
    env.model.u[:] = _cuda(u)
    xdot = env.model.get_extended_state().cpu().numpy()
    assert xdot.shape == (n, 17) and not xdot[:, 12:].any()
    # truth = the same formulas in float64 (weights cast up); the bar is "no further from truth than the reference's
    # own fp32 evaluation is" (x2), since sums like Cl_tot cancel and amplify ulp-level MLP differences in both.
    import torch as _t
    from oracle.f16_oracle import AeroNets, body_accel as o_accel, nlplant as o_nlplant
    a64 = AeroNets(dtype=_t.float64)
    s64, u64 = _t.from_numpy(s).double(), _t.from_numpy(u).double()
    truth = o_nlplant(a64, s64, u64).numpy()
    floor = 1e-3 * np.median(np.abs(truth), axis=0) + 1e-12
    e_ours, e_ref = rel_err(xdot[:, :12], truth, floor), rel_err(g["xdot"].astype(np.float64), truth, floor)
    print("\nnlplant vs fp64: ours p50 %.2e p99 %.2e max %.2e | reference p50 %.2e p99 %.2e max %.2e" % (
        np.median(e_ours), np.percentile(e_ours, 99), e_ours.max(), np.median(e_ref), np.percentile(e_ref, 99), e_ref.max()))
    assert np.percentile(e_ours, 99) <= 2 * np.percentile(e_ref, 99) and np.median(e_ours) <= 2 * np.median(e_ref) + 1e-9
    assert e_ours.max() <= 3 * e_ref.max()
    err = rel_err(xdot[:, :12], g["xdot"], floor)
    assert np.percentile(err, 99) < 1e-5, np.percentile(err, 99)
    ax, ay, az = env.model.get_acceleration()
    acc = torch.stack((ax, ay, az), 1).cpu().numpy()
    t_acc = np.stack([x.numpy() for x in o_accel(a64, s64, u64)], 1)
    afloor = 1e-3 * np.median(np.abs(t_acc))
    ea_ours, ea_ref = rel_err(acc, t_acc, afloor), rel_err(g["accel"].astype(np.float64), t_acc, afloor)
    assert np.percentile(ea_ours, 99) <= 2 * np.percentile(ea_ref, 99) and ea_ours.max() <= 3 * ea_ref.max()
    assert np.allclose(env.model.get_EAS2TAS().cpu().numpy(), g["eas2tas"], rtol=1e-6)
    assert np.allclose(env.model.get_G().cpu().numpy(), g["G"], rtol=2e-5, atol=1e-5)


# --------------------------------------------------------------------------------------------------------
# trajectories against the reference fixtures
# --------------------------------------------------------------------------------------------------------
def _trajectory(task, fixture):
    """Replay a fixture's tapes through the CUDA env.  At every checkpoint, over aircraft whose reset history equals
    the reference's:
      * distance to the reference-fp32 state: median <= max(1e-5, 1.5 x the reference's own distance to the float64
        truth at that checkpoint) -- 1e-5 is north_star's figure; it is also the level at which the reference's
        fp32 arithmetic itself sits from exact arithmetic after 1000 steps (fixture key k*_ref_vs_truth), so a
        tighter bar would measure rounding luck, not correctness;
      * distance to the float64 truth: median <= 1.5 x the reference's (we must not be further from truth);
      * before trajectories can fork (k <= 20) every flag and reward agrees."""
    g = np.load(os.path.join(GOLDEN, fixture))
    n, steps, seed = [int(x) for x in g["meta"]]
    scale = float(g["scale"])
    env = _env(n, task)
    obs0 = env.reset(reset_draws=_cuda(tapes.reset_draw_tape(seed, 0, n)))
    assert np.allclose(obs0.cpu().numpy(), g["obs0"], rtol=1e-6, atol=1e-7)
    n_bad = []
    report = []
    for k in range(1, steps + 1):
        obs, rew, done, bad, exc, _ = env.step(_cuda(tapes.action_tape(seed, k, n, scale)),
                                               reset_draws=_cuda(tapes.reset_draw_tape(seed, k, n)))
        n_bad.append(int(bad.sum()))
        if f"k{k}_s" not in g.files:
            continue
        sc = env.step_count.cpu().numpy()
        same = sc == g[f"k{k}_step_count"]
        s = env.model.s.cpu().numpy()
        err = state_rel_err(s[same], g[f"k{k}_s"][same])
        same64 = sc == g[f"k{k}_step_count64"]
        err64 = state_rel_err(s[same64], g[f"k{k}_s64"][same64])
        ref64 = float(g[f"k{k}_ref_vs_truth"][0])
        report.append((k, float(same.mean()), float(np.median(err)), float(err.max()), float(np.median(err64)), ref64))
        assert same.mean() >= 0.80 and same64.mean() >= 0.80, report[-1]
        assert np.median(err) <= max(1e-5, 1.5 * ref64), report[-1]
        assert np.median(err64) <= 1.5 * ref64 + 1e-7, report[-1]
        oerr = np.abs(obs.cpu().numpy()[same] - g[f"k{k}_obs"][same])
        assert np.median(oerr.max(axis=1)) <= 1e-4, (k, np.median(oerr.max(axis=1)))
        if k <= 20:  # before histories can fork every flag and reward must agree
            assert same.all()
            assert np.array_equal(bad.cpu().numpy(), g[f"k{k}_bad"]) and np.array_equal(done.cpu().numpy(), g[f"k{k}_done"])
            assert np.allclose(rew.cpu().numpy(), g[f"k{k}_reward"], rtol=1e-5, atol=1e-5)
    print(f"\n{fixture}: step | same-history | ours-vs-ref32 median, max | ours-vs-fp64 median | ref32-vs-fp64 median")
    for r in report:
        print("   k=%4d same=%.3f  %.2e %.2e | %.2e | %.2e" % r)
    # episode statistics must agree with the reference run (reset events are chaotic individually, not in bulk)
    ref_bad, got_bad = int(g["n_bad"][:steps].sum()), int(np.sum(n_bad))
    assert abs(got_bad - ref_bad) <= max(5, 0.03 * ref_bad), (got_bad, ref_bad)


def test_heading_1000_steps_small_actions(dev, step_kernel):
    """BASELINE config 1: F16 heading, n=128, 1000 steps, actions 0.3*U(-1,1)."""
    _trajectory("heading", "heading_traj_a03.npz")


def test_heading_1000_steps_full_actions(dev, step_kernel):
    """Same with full-scale actions (the reset-exercising tape): ~2200 terminations/resets in 1000 steps."""
    _trajectory("heading", "heading_traj_a10.npz")


def test_control_task_trajectory(dev, step_kernel):
    _trajectory("control", "control_traj.npz")


def test_tracking_task_trajectory(dev, step_kernel):
    _trajectory("tracking", "tracking_traj.npz")


def test_done_branch(dev, step_kernel):
    """Target reached -> done, +200, and the reset at the top of the next step (unreach_heading.py:49-53)."""
    g = np.load(os.path.join(GOLDEN, "heading_done_branch.npz"))
    n, steps, seed = [int(x) for x in g["meta"]]
    env = _env(n, "heading")
    env.reset(reset_draws=_cuda(tapes.reset_draw_tape(seed, 0, n)))
    env.task.target_altitude[:] = env.model.s[:, 2]
    env.task.target_heading[:] = env.model.s[:, 5]
    env.task.target_vt[:] = env.model.s[:, 6]
    env.task.target_altitude[::4] += 500.0
    env.step_count[:] = 298
    env.step_count[1::8] = 2499
    env.step_count[::8] = 2499
    for k in range(1, steps + 1):
        obs, rew, done, bad, exc, _ = env.step(_cuda(tapes.action_tape(seed, k, n, 0.02)),
                                               reset_draws=_cuda(tapes.reset_draw_tape(seed, k, n)))
        assert np.array_equal(done.cpu().numpy(), g[f"k{k}_done"]), k
        assert np.array_equal(bad.cpu().numpy(), g[f"k{k}_bad"]), k
        assert np.array_equal(env.step_count.cpu().numpy(), g[f"k{k}_step_count"]), k
        assert np.allclose(rew.cpu().numpy(), g[f"k{k}_reward"], rtol=1e-5, atol=1e-5), k
        assert np.allclose(obs.cpu().numpy(), g[f"k{k}_obs"], rtol=1e-5, atol=1e-6), k
    assert int(g["k2_done"].sum()) == 20 and int(g["k1_bad"].sum()) == 4


# --------------------------------------------------------------------------------------------------------
# against the live oracle
# --------------------------------------------------------------------------------------------------------
def _load_state(env, orc, s, u, tgt, steps):
    env.model.s[:] = _cuda(s); env.model.u[:] = _cuda(u)
    env._tgt[:, :env.n] = _cuda(tgt.T.copy())
    env.step_count[:] = _cuda(steps.astype(np.int32))
    env._flags.zero_()
    orc.s = torch.from_numpy(s.copy()); orc.u = torch.from_numpy(u.copy()); orc.tgt = torch.from_numpy(tgt.copy())
    orc.step_count = torch.from_numpy(steps.astype(np.int64))
    orc.is_done[:] = False; orc.bad_done[:] = False; orc.exceed_time_limit[:] = False


@pytest.mark.parametrize("task", ["heading", "control", "tracking"])
def test_single_step_random_envelope(dev, task):
    """One step from 20000 random in-envelope (s, u, target, step_count): state p99 <= 2e-6, obs/reward close,
    flags equal except where the oracle's own margin to a threshold is below 1e-5 relative."""
    from oracle.f16_oracle import F16EnvOracle
    n = 20000
    s, u = tapes.random_envelope_states(77, n)
    r = tapes.uniform01(78, 1, (n, 4))
    if task == "heading":
        tgt = np.stack([s[:, 2] + (r[:, 0] - 0.5) * 400, s[:, 5] + (r[:, 1] - 0.5) * 0.3, s[:, 6] + (r[:, 2] - 0.5) * 60], 1)
    elif task == "control":
        tgt = np.stack([s[:, 4] + (r[:, 0] - 0.5) * 0.3, s[:, 5] + (r[:, 1] - 0.5) * 0.3, s[:, 6] + (r[:, 2] - 0.5) * 60], 1)
    else:
        tgt = np.stack([s[:, 0] + (r[:, 0] - 0.5) * 400, s[:, 1] + (r[:, 1] - 0.5) * 400, s[:, 2] + (r[:, 2] - 0.5) * 400], 1)
    tgt = tgt.astype(np.float32)
    steps = (r[:, 3] * 2600).astype(np.int64)
    env, orc = _env(n, task), F16EnvOracle(n, task)
    _load_state(env, orc, s, u, tgt, steps)
    a = tapes.action_tape(79, 1, n, 1.0)
    d = tapes.reset_draw_tape(79, 1, n)
    obs, rew, done, bad, exc, _ = env.step(_cuda(a), reset_draws=_cuda(d))
    o_obs, o_rew, o_done, o_bad, o_exc = orc.step(torch.from_numpy(a), torch.from_numpy(d))
    s_new, s_ref = env.model.s.cpu().numpy(), orc.s.numpy()
    # truth: the same step in float64.  Bar: our distance to truth <= 2x the reference-fp32 distance to truth.
    o64 = F16EnvOracle(n, task, dtype=torch.float64)
    o64.s = torch.from_numpy(s).double(); o64.u = torch.from_numpy(u).double(); o64.tgt = torch.from_numpy(tgt).double()
    o64.step_count = torch.from_numpy(steps.astype(np.int64))
    o64.is_done[:] = False; o64.bad_done[:] = False; o64.exceed_time_limit[:] = False
    o64.step(torch.from_numpy(a).double(), torch.from_numpy(d).double())
    truth = o64.s.numpy()
    e_ours, e_ref = state_rel_err(s_new, truth), state_rel_err(s_ref, truth)
    print("\n%s one step vs fp64: ours p50 %.2e p99 %.2e max %.2e | reference-fp32 p50 %.2e p99 %.2e max %.2e" % (
        task, np.median(e_ours), np.percentile(e_ours, 99), e_ours.max(), np.median(e_ref), np.percentile(e_ref, 99), e_ref.max()))
    assert np.percentile(e_ours, 99) <= 2 * np.percentile(e_ref, 99) and np.median(e_ours) <= 2 * np.median(e_ref)
    # the tail is a handful of samples where Cl_tot / Cn_tot cancel to ~0 and ulp-level coefficient differences are
    # amplified (in the reference's fp32 exactly as in ours): bar the 99.9th percentile at 2x and the single worst
    # sample of 20 000 (a noisy statistic) at 5x the reference's
    assert np.percentile(e_ours, 99.9) <= 2 * np.percentile(e_ref, 99.9)
    assert e_ours.max() <= 5 * e_ref.max()
    err = state_rel_err(s_new, s_ref)
    assert np.percentile(err, 99) <= 2e-5 and np.median(err) <= 2e-6, (np.median(err), np.percentile(err, 99), err.max())
    assert np.allclose(env.model.u.cpu().numpy(), orc.u.numpy(), rtol=1e-6, atol=1e-6)
    assert np.allclose(obs.cpu().numpy(), o_obs.numpy(), rtol=1e-5, atol=2e-6)
    # flags: exact except aircraft sitting on a threshold in the oracle itself
    acc_margin = np.abs(orc.last_accel.numpy() - 300.0) / 300.0
    near = acc_margin < 1e-5
    mism_bad = (bad.cpu().numpy() != o_bad.numpy()) & ~near
    assert mism_bad.sum() <= 2, int(mism_bad.sum())
    assert (done.cpu().numpy() != o_done.numpy()).sum() <= 2
    ok = bad.cpu().numpy() == o_bad.numpy()
    assert np.allclose(rew.cpu().numpy()[ok], o_rew.numpy()[ok], rtol=1e-5, atol=1e-5)
    assert 0 < int(o_bad.sum()) < n and int(o_done.sum()) > 0      # the case exercises both branches


def test_coef_cache_is_bit_identical(dev, step_kernel):
    """The (alpha,beta)-coefficient cache must not change a single bit of any output (200 steps with resets)."""
    n, seed = 512, 31
    e1, e2 = _env(n, "heading", use_coef_cache=True), _env(n, "heading", use_coef_cache=False)
    d0 = _cuda(tapes.reset_draw_tape(seed, 0, n))
    assert torch.equal(e1.reset(reset_draws=d0), e2.reset(reset_draws=d0))
    for k in range(1, 201):
        a, d = _cuda(tapes.action_tape(seed, k, n, 1.0)), _cuda(tapes.reset_draw_tape(seed, k, n))
        r1, r2 = e1.step(a, reset_draws=d), e2.step(a, reset_draws=d)
        for x, y in zip(r1[:5], r2[:5]):
            assert torch.equal(x, y), k
        if k == 100:  # an external write to the state must invalidate the cache (key mismatch path)
            for e in (e1, e2):
                e.model.s[::3, 7] += 0.01
                e.model.s[1::3, 8] -= 0.02
    assert torch.equal(e1.model.s, e2.model.s) and torch.equal(e1.model.u, e2.model.u)
    assert e1.termination_counters() == e2.termination_counters()
    assert e1.termination_counters()["resets"] > n


@pytest.mark.parametrize("n", [1, 31, 77, 257])
def test_ragged_population_sizes(dev, n, step_kernel):
    """n not a multiple of the warp / block size: tail lanes must neither read nor write out of range."""
    from oracle.f16_oracle import F16EnvOracle
    env, orc = _env(n, "heading"), F16EnvOracle(n, "heading")
    d0 = tapes.reset_draw_tape(5, 0, n)
    env.reset(reset_draws=_cuda(d0)); orc.reset(torch.from_numpy(d0))
    for k in range(1, 6):
        a, d = tapes.action_tape(5, k, n, 1.0), tapes.reset_draw_tape(5, k, n)
        obs, rew, done, bad, exc, _ = env.step(_cuda(a), reset_draws=_cuda(d))
        o = orc.step(torch.from_numpy(a), torch.from_numpy(d))
        assert np.allclose(obs.cpu().numpy(), o[0].numpy(), rtol=1e-5, atol=2e-6)
        assert np.allclose(rew.cpu().numpy(), o[1].numpy(), rtol=1e-5, atol=1e-5)
    assert not env._tgt[:, n:].any() and not env.model._s[:, n:].any()      # padding untouched


def test_full_size_sampled_parity_and_invariants(dev):
    """BASELINE config 2 size (n = 10^6): aircraft are independent, so a random sample of the population is
    replayed through the oracle from the same pre-step state; plus population-wide invariants."""
    from oracle.f16_oracle import F16EnvOracle
    n, m = 1_000_000, 4096
    env = _env(n, "heading")
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(30):
        env.step(torch.rand((n, 4), device=dev, generator=g) * 2 - 1)
    idx = torch.from_numpy(np.sort(np.random.default_rng(0).choice(n, m, replace=False))).cuda()
    orc = F16EnvOracle(m, "heading")
    orc.s = env.model.s[idx].cpu(); orc.u = env.model.u[idx].cpu()
    orc.tgt = env._tgt[:, idx].t().contiguous().cpu()
    orc.step_count = env.step_count[idx].cpu().long()
    orc.is_done = env.is_done[idx].cpu().clone(); orc.bad_done = env.bad_done[idx].cpu().clone()
    orc.exceed_time_limit = env.exceed_time_limit[idx].cpu().clone()
    a = torch.rand((n, 4), device=dev, generator=g) * 2 - 1
    d = torch.rand((n, 5), device=dev, generator=g)
    obs, rew, done, bad, exc, _ = env.step(a, reset_draws=d)
    o_obs, o_rew, o_done, o_bad, o_exc = orc.step(a[idx].cpu(), d[idx].cpu())
    assert int(orc.step_count.eq(1).sum()) > 0                    # the sample contains aircraft that just reset
    assert np.allclose(obs[idx].cpu().numpy(), o_obs.numpy(), rtol=2e-5, atol=5e-6)
    assert (bad[idx].cpu() != o_bad).sum() <= 1
    s_all = env.model.s
    assert torch.isfinite(s_all).all() and torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert int(env._flags.max()) <= 1 and not bool(exc.any())
    c = env.termination_counters()
    assert c["resets"] >= n and c["overload"] > 0


def test_sharding_is_world_size_invariant(dev):
    """Two half-populations with index_base 0 / n/2 reproduce the single population bit for bit (in-kernel
    Philox streams are keyed by the global aircraft index), noise included."""
    from neuralplane_b200 import ControlEnv
    n = 4096
    full = ControlEnv(num_envs=n, config="heading", random_seed=7, device="cuda:0")
    halves = [ControlEnv(num_envs=n // 2, config="heading", random_seed=7, device="cuda:0", index_base=b)
              for b in (0, n // 2)]
    of = full.reset().clone()
    oh = torch.cat([h.reset().clone() for h in halves])
    assert torch.equal(of, oh)
    for k in range(1, 40):
        a = _cuda(tapes.action_tape(9, k, n, 1.0))
        rf = [x.clone() for x in full.step(a)[:4]]
        rh = [h.step(a[i * n // 2:(i + 1) * n // 2].contiguous())[:4] for i, h in enumerate(halves)]
        for j in range(4):
            assert torch.equal(rf[j], torch.cat([r[j] for r in rh])), (k, j)


def test_observation_noise_statistics(dev):
    n = 65536
    from neuralplane_b200 import ControlEnv
    e0 = ControlEnv(num_envs=n, config="heading", random_seed=3, device="cuda:0")
    e1 = ControlEnv(num_envs=n, config="heading", random_seed=3, device="cuda:0")
    e0.task.noise_scale = 0.0
    d0 = _cuda(tapes.reset_draw_tape(1, 0, n))
    a = _cuda(tapes.action_tape(1, 1, n, 0.5))
    e0.reset(reset_draws=d0); e1.reset(reset_draws=d0)
    z = (e1.step(a, reset_draws=d0)[0] - e0.step(a, reset_draws=d0)[0]) / 0.01
    assert abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    assert float(z.abs().max()) < 6.5
    c = np.corrcoef(z[:, :6].t().cpu().numpy())
    assert np.abs(c - np.eye(6)).max() < 0.02
    # injected normals reproduce obs + noise * scale exactly (heading_task.py:152: obs + randn_like(obs) * noise_scale): the two
    # envs hold the same state (noise never feeds back into it), one steps with injected normals, the other without noise
    assert torch.equal(e0.model.s, e1.model.s)
    nz = torch.randn((n, 22), device=dev)
    e0.task.noise_scale = 0.01
    e1.task.noise_scale = 0.0
    o_inj = e0.step(a, reset_draws=d0, noise=nz)[0].clone()
    o_clean = e1.step(a, reset_draws=d0)[0].clone()
    assert torch.equal(o_inj, o_clean + nz * 0.01)


@pytest.mark.parametrize("model,task", [("F16", "heading"), ("F16_tables", "control"), ("UAV", "control")])
def test_pipelined_numpy_boundary_equals_single_launch(dev, model, task):
    """The three host boundaries of GPUVecEnv -- single launch + copy, the chunk pipeline (np_env_step_host) and, for the F16
    plug-in, the kernel writing straight into mapped pinned memory (np_env_step_mapped, zero-copy action reads or an
    explicit upload) -- must agree bit for bit, odd population and in-kernel Philox resets / noise included."""
    from neuralplane_b200 import ControlEnv, GPUVecEnv
    ne = 20_001
    mk = lambda: ControlEnv(num_envs=ne, config=task, model=model, random_seed=9, device="cuda:0")
    v1, v4 = GPUVecEnv([mk], boundary="copy"), GPUVecEnv([mk], boundary="pipelined", pipeline_chunks=4)
    assert v1._chunks is None and len(v4._chunks) == 4 and v4._chunks[0][0] == 0 and v4._chunks[-1][1] == ne
    assert all(c[0] % 256 == 0 and c[1] > c[0] for c in v4._chunks) and all(a[1] == b[0] for a, b in zip(v4._chunks, v4._chunks[1:]))
    others = [v4]
    if model != "UAV":
        vm, vu = GPUVecEnv([mk]), GPUVecEnv([mk], boundary="mapped")
        assert vm.boundary == "mapped"           # the default for F16 populations below 2 x 10^5
        vu._zero_copy_actions = False            # explicit H2D copy of the actions before the launch
        others += [vm, vu]
    else:
        with pytest.raises(ValueError):
            GPUVecEnv([mk], boundary="mapped")
    o1 = v1.reset()
    for v in others:
        assert np.array_equal(o1, v.reset())
    kept = None
    for k in range(1, 60):
        a = tapes.action_tape(9, k, ne, 1.0).reshape(ne, 1, 4)
        r1 = v1.step(a)
        for v in others:
            for x, y in zip(r1[:5], v.step(a)[:5]):
                assert x.shape == y.shape and np.array_equal(x, y), (k, v.boundary)
        if k == 10:                              # small populations return fresh arrays, like the reference's _t2n
            kept = (r1[0], r1[0].copy(), r1[2], r1[2].copy())
    assert np.array_equal(kept[0], kept[1]) and np.array_equal(kept[2], kept[3])
    for v in others:
        assert v1.gpu_vec_env.termination_counters() == v.gpu_vec_env.termination_counters()
    assert v1.gpu_vec_env.termination_counters()["resets"] > ne


def test_numpy_boundary_ring_lifetime(dev):
    """copy=False hands out views of a ring of pinned buffers: a result stays intact for ring - 1 further steps."""
    from neuralplane_b200 import ControlEnv, GPUVecEnv
    ne = 4096
    v = GPUVecEnv([lambda: ControlEnv(num_envs=ne, config="heading", model="F16", random_seed=3, device="cuda:0")], copy=False, ring=3)
    v.reset()
    acts = [tapes.action_tape(3, k, ne, 1.0).reshape(ne, 1, 4) for k in range(1, 5)]
    first = v.step(acts[0])[0]
    snap = first.copy()
    v.step(acts[1]); v.step(acts[2])
    assert np.array_equal(first, snap)           # two more steps: still the ring slot of step 1
    v.step(acts[3])
    assert not np.array_equal(first, snap)       # the fourth step reuses it

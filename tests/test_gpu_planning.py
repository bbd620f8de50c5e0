"""GPU parity of the fused planning step (np_env_plan_step: PlanningEnv.step + PID low-level controller).

The closed loop is violently sensitive: the reference's rate loops run at Kp = 10 x scaler^2 into +-45 deg clamps, and
the oracle's own 1-ulp twin (s * (1 + 2^-23) after the first reset) is a MEDIAN 1e-4 away after one planning step
(50 sub-steps), 9e-3 after three and O(0.1) from the fourth on (measured, /oracle/planning_oracle.py).  Parity is
therefore established where it is meaningful:
  * teacher-forced: every sub-step (and every 5 sub-steps) starts from the oracle's exact state / controls / PID
    state, and the kernel's next state, controls, controller state, flags and reward must match to fp32 rounding;
  * the reference-component fixture (tests/golden/planning_pid_traj.npz) at planning step 1, at 3x the twin's error;
  * population statistics over the 16 fixture steps.
"""
import os

import numpy as np
import pytest
import torch

from oracle import tapes
from _metrics import state_rel_err

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _env(n, n_sub):
    from neuralplane_b200 import PlanningEnv
    return PlanningEnv(num_envs=n, config="tracking", model="F16", random_seed=0, device="cuda:0", n_substeps=n_sub)


def _push(env, orc, first):
    """Copy the oracle's full state into the GPU env."""
    n = env.n
    env.model.s[:] = _cuda(orc.s.numpy()); env.model.u[:] = _cuda(orc.u.numpy())
    env._tgt[:, :n] = _cuda(orc.tgt.numpy().T.copy())
    env.step_count[:] = _cuda(orc.step_count.numpy().astype(np.int32))
    env._flags[0, :n] = _cuda(orc.is_done.numpy().astype(np.uint8)); env._flags[1, :n] = _cuda(orc.bad_done.numpy().astype(np.uint8))
    env._flags[2, :n] = 0
    env.pid_state[:] = _cuda(orc.pid_state().numpy())
    from neuralplane_b200 import _native as nv
    nv.check(nv.lib().np_env_set_pid_started(env._handle, 0 if first else 1), "np_env_set_pid_started")


@pytest.fixture(params=["K1c", "K1"])
def plan_kernel(request, monkeypatch):
    """Planning populations up to 18 944 aircraft (train_tracking.sh: 10 000) run on K1c's PLAN instantiation by default; the
    oracle tests at such sizes run once on it and once on K1's MODE_PLAN."""
    if request.param == "K1":
        monkeypatch.setenv("NPLANE_COOP_PAIRS", "0")
    return request.param


@pytest.mark.parametrize("n,n_sub", [(3000, 50), (10_001, 7), (63, 3), (18_944, 2)])
def test_cooperative_planning_kernel_is_bit_identical(n, n_sub):
    """K1c<PLAN> (coop_step_kernel.cuh) against K1<MODE_PLAN>: the same env steps -- PID stack, control lag, n_sub FDM sub-steps
    with frozen terminated aircraft, episodic resets -- must agree in every output, the state, the controls, the controller
    state and the counters, bit for bit."""
    from neuralplane_b200 import PlanningEnv
    kw = dict(num_envs=n, config="tracking", model="F16", random_seed=3, device="cuda:0", n_substeps=n_sub)
    coop = PlanningEnv(**kw)
    os.environ["NPLANE_COOP_PAIRS"] = "0"
    try:
        ref = PlanningEnv(**kw)
    finally:
        del os.environ["NPLANE_COOP_PAIRS"]
    coop.reset(); ref.reset()
    for k in range(1, 13):
        a = _cuda(tapes.action_tape(3, k, n, 1.0, num_actions=3))
        for x, y in zip(coop.step(a)[:5], ref.step(a)[:5]):
            assert torch.equal(x, y), k
        assert torch.equal(coop.pid_state, ref.pid_state), k
        if k % 4 == 0:
            for e in (coop, ref):
                e.is_done[::7] = True
                e.bad_done[3::64] = True
    li, lr = coop.launch_info(), ref.launch_info()
    assert li["block"] == (256 if n <= 148 * 64 else 128) and li["grid"] == (n + 63) // 64, li
    assert lr["block"] == 128 and lr["grid"] == (n + 255) // 256, lr
    assert torch.equal(coop.model.s, ref.model.s) and torch.equal(coop.model.u, ref.model.u)
    assert torch.equal(coop.step_count, ref.step_count)
    assert coop.termination_counters() == ref.termination_counters() and coop.termination_counters()["resets"] > 0


@pytest.mark.parametrize("n_sub", [1, 5])
def test_teacher_forced_substeps_vs_oracle(n_sub, plan_kernel):
    from oracle.planning_oracle import PlanningOracle
    n, seed, iters = 2048, 23, 40 if n_sub == 1 else 12
    env, orc = _env(n, n_sub), PlanningOracle(n)
    orc.N_SUB = n_sub
    d0 = tapes.reset_draw_tape(seed, 0, n)
    env.reset(reset_draws=_cuda(d0)); orc.reset(torch.from_numpy(d0))
    worst = 0.0
    for k in range(1, iters + 1):
        _push(env, orc, first=(k == 1))
        a = tapes.action_tape(seed, k, n, 1.0, num_actions=3)
        d = tapes.reset_draw_tape(seed, k, n)
        obs, rew, done, bad, exc, _ = env.step(_cuda(a), reset_draws=_cuda(d))
        o_obs, o_rew, o_done, o_bad, o_exc = orc.plan_step(torch.from_numpy(a), torch.from_numpy(d))
        # aircraft whose actuator demand sits on a clamp / integrator-limit edge can flip a branch on an ulp: compare
        # the population by percentiles, flags exactly but for at most a handful
        err = state_rel_err(env.model.s.cpu().numpy(), orc.s.numpy())
        # (the F-16 single-step test bars ours-vs-reference-fp32 at median 2e-6 / p99 2e-5 on random in-envelope states;
        #  here the surfaces are driven hard against their +-45 deg clamps, so the moment sums cancel less kindly and
        #  the tail is wider; right after the common reset every aircraft sits at alpha = beta = 0 and the rounding
        #  differences are correlated.  Controls and controller state are barred separately below.)
        bar50, bar99 = (5e-6, 1e-4) if n_sub == 1 else (2e-5, 2e-4)
        assert np.percentile(err, 99) <= bar99 and np.median(err) <= bar50, (k, np.median(err), np.percentile(err, 99), err.max())
        du = np.abs(env.model.u.cpu().numpy() - orc.u.numpy()) / (np.abs(orc.u.numpy()) + np.array([50, .05, .05, .05, 1], np.float32))
        assert np.percentile(du, 99) <= 1e-4, (k, np.percentile(du, 99))
        dp = np.abs(env.pid_state.cpu().numpy() - orc.pid_state().numpy()) / (np.abs(orc.pid_state().numpy()) + 1e-2)
        assert np.percentile(dp, 99) <= 1e-3, (k, np.percentile(dp, 99))
        assert (bad.cpu().numpy() != o_bad.numpy()).sum() <= 3 and (done.cpu().numpy() != o_done.numpy()).sum() <= 3, k
        ok = (bad.cpu().numpy() == o_bad.numpy()) & (done.cpu().numpy() == o_done.numpy())
        assert np.allclose(rew.cpu().numpy()[ok], o_rew.numpy()[ok], rtol=1e-4, atol=1e-5), k
        assert np.median(np.abs(obs.cpu().numpy() - o_obs.numpy()).max(axis=1)) <= 1e-5, k
        assert np.array_equal(env.step_count.cpu().numpy(), orc.step_count.numpy().astype(np.int32)), k
        worst = max(worst, float(np.percentile(err, 99)))
    print(f"\nteacher-forced n_sub={n_sub}: worst p99 state error {worst:.2e}; bad so far {int(orc.bad_done.sum())}")


def test_planning_fixture_first_step_and_statistics():
    g = np.load(os.path.join(GOLDEN, "planning_pid_traj.npz"))
    n, steps, seed = [int(x) for x in g["meta"]]
    env = _env(n, 50)
    obs0 = env.reset(reset_draws=_cuda(tapes.reset_draw_tape(seed, 0, n)))
    assert np.allclose(obs0.cpu().numpy(), g["obs0"], rtol=1e-6, atol=1e-7)
    bad_tot, ref_tot = 0, 0
    for k in range(1, steps + 1):
        obs, rew, done, bad, exc, _ = env.step(_cuda(tapes.action_tape(seed, k, n, float(g["scale"]), num_actions=3)),
                                               reset_draws=_cuda(tapes.reset_draw_tape(seed, k, n)))
        bad_tot += int(bad.sum()); ref_tot += int(g[f"k{k}_bad"].sum())
        if k == 1:
            err = state_rel_err(env.model.s.cpu().numpy(), g["k1_s"])
            print("\nplanning step 1 vs reference components: median %.2e p90 %.2e (oracle 1-ulp twin: 9.9e-05 / 2.0e-03)" % (
                np.median(err), np.percentile(err, 90)))
            assert np.median(err) <= 3e-4 and np.percentile(err, 90) <= 6e-3
            assert np.array_equal(env.step_count.cpu().numpy(), g["k1_step_count"])
            assert (bad.cpu().numpy() != g["k1_bad"]).sum() <= 2
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all()
        assert int(env.step_count.max()) <= 50 * k
    assert abs(bad_tot - ref_tot) <= max(10, 0.35 * ref_tot), (bad_tot, ref_tot)


def test_planning_full_size_invariants():
    """n = 10^5 x 50 sub-steps: finite outputs, flags consistent with rewards, frozen aircraft keep their state."""
    n = 100_000
    env = _env(n, 50)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    for k in range(3):
        a = torch.rand((n, 3), device="cuda", generator=g) * 2 - 1
        obs, rew, done, bad, exc, _ = env.step(a)
        assert torch.isfinite(obs).all() and torch.isfinite(rew).all() and torch.isfinite(env.model.s).all()
        # reward = position term (|.| < 1 here) + 200 done - 200 bad (event_driven_reward.py:28): the flags are readable off it
        assert torch.equal(rew < -100, bad & ~done) and torch.equal(rew > 100, done & ~bad)
        assert bool((rew[~bad & ~done].abs() < 100).all())
        assert int(env.step_count.min()) >= 1 and int(env.step_count.max()) <= 50 * (k + 1)
        assert not bool(exc.any())
    c = env.termination_counters()
    assert c["resets"] >= n and c["overload"] + c["extreme_state"] + c["low_altitude"] > 0


def test_planning_full_size_sampled_parity():
    """BASELINE configs[3] size (n = 10^6, 50 sub-steps): aircraft are independent, so a random sample of the population is
    replayed through the planning oracle from the same pre-step state, controls, targets, flags and controller state.  The
    closed loop amplifies an ulp to ~1e-4 within one planning step (see the module docstring), hence the percentile bars
    of the fixture test; flags must agree but for a handful of clamp-edge aircraft."""
    from oracle.planning_oracle import PlanningOracle
    n, m = 1_000_000, 1024
    env = _env(n, 50)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(4)
    for _ in range(2):
        env.step(torch.rand((n, 3), device="cuda", generator=g) * 2 - 1)
    idx = torch.from_numpy(np.sort(np.random.default_rng(1).choice(n, m, replace=False))).cuda()
    orc = PlanningOracle(m)
    orc.s = env.model.s[idx].cpu(); orc.u = env.model.u[idx].cpu()
    orc.tgt = env._tgt[:, idx].t().contiguous().cpu()
    orc.step_count = env.step_count[idx].cpu().long()
    orc.is_done = env.is_done[idx].cpu().clone(); orc.bad_done = env.bad_done[idx].cpu().clone()
    orc.exceed_time_limit = env.exceed_time_limit[idx].cpu().clone()
    ps = env.pid_state[idx].cpu()
    for j, k in enumerate(("roll", "pitch", "yaw", "speed")):
        orc.pid[k].error, orc.pid[k].integrator, orc.last_out[k] = ps[:, 3 * j].clone(), ps[:, 3 * j + 1].clone(), ps[:, 3 * j + 2].clone()
        orc.pid[k].first = False
    a = torch.rand((n, 3), device="cuda", generator=g) * 2 - 1
    d = torch.rand((n, 5), device="cuda", generator=g)
    obs, rew, done, bad, exc, _ = env.step(a, reset_draws=d)
    o_obs, o_rew, o_done, o_bad, o_exc = orc.plan_step(a[idx].cpu(), d[idx].cpu())
    assert int((orc.step_count <= 50).sum()) > 0                      # the sample contains aircraft that were just re-initialised
    err = state_rel_err(env.model.s[idx].cpu().numpy(), orc.s.numpy())
    assert np.median(err) <= 3e-4 and np.percentile(err, 90) <= 6e-3, (np.median(err), np.percentile(err, 90))
    assert (bad[idx].cpu() != o_bad).sum() <= 4 and (done[idx].cpu() != o_done).sum() <= 4
    assert np.array_equal(env.step_count[idx].cpu().numpy(), orc.step_count.numpy().astype(np.int32))
    assert torch.isfinite(env.model.s).all() and torch.isfinite(obs).all() and torch.isfinite(rew).all()
    assert torch.equal(rew < -100, bad & ~done) and torch.equal(rew > 100, done & ~bad)

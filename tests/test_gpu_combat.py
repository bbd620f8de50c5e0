"""GPU parity of the fused 1-v-1 combat step (np_env_combat_step) against the combat oracle, which is itself pinned
bit-exact to the reference's own obs / reward / geometry / termination / controller code (tests/golden/combat*_traj.npz).
As for the planning step, the PID loops make free-running trajectories diverge at fp32 noise level within tens of
sub-steps, so every env step (5 sub-steps) is teacher-forced from the oracle's exact state; the fixtures are replayed
for their first steps and for their discrete events (Crash, Shutdown bad / done, env-level resets)."""
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


def _env(num_envs):
    from neuralplane_b200 import SingleCombatEnv
    return SingleCombatEnv(num_envs=num_envs, config="selfplay", random_seed=0, device="cuda:0")


def _obs_close(got, want, what):
    """15-D combat obs: columns 11, 12 are acos() angles, ill-conditioned near 0 and pi (error = eps / sin(angle)): they
    get 5e-4 absolute; everything else fp32-tight."""
    d = np.abs(got - want)
    ang = [11, 12]
    rest = [j for j in range(15) if j not in ang]
    assert d[:, rest].max() <= 2e-5 + 2e-5 * np.abs(want[:, rest]).max(), (what, d[:, rest].max(axis=0))
    assert d[:, ang].max() <= 5e-4 and np.median(d[:, ang]) <= 2e-6, (what, d[:, ang].max(), np.median(d[:, ang]))


def _push(env, orc, first):
    from neuralplane_b200 import _native as nv
    n = env.n
    env.model.s[:] = _cuda(orc.s.numpy()); env.model.u[:] = _cuda(orc.u.numpy())
    env.step_count[:] = _cuda(orc.step_count.numpy().astype(np.int32))
    for j, f in enumerate((orc.is_done, orc.bad_done, orc.exceed_time_limit)):
        env._flags[j, :n] = _cuda(f.numpy().astype(np.uint8))
    env.blood[:] = _cuda(orc.blood.numpy())
    cs = orc.ctrl_state().numpy()                       # roll_dem, pitch_dem, then 3 x (err, int, last)
    st = np.zeros((n, 12), np.float32)
    st[:, 0:9] = cs[:, 2:11]; st[:, 9] = cs[:, 0]; st[:, 10] = cs[:, 1]
    env.ctrl_state[:] = _cuda(st)
    nv.check(nv.lib().np_env_set_pid_started(env._handle, 0 if first else 1), "np_env_set_pid_started")


def _compare(env, orc, out, ref, k, strict_flags=True):
    obs, rew, done, bad, exc, _ = out
    o_obs, o_rew, o_done, o_bad, o_exc = ref
    err = state_rel_err(env.model.s.cpu().numpy(), orc.s.numpy())
    assert np.median(err) <= 2e-5 and np.percentile(err, 99) <= 3e-4, (k, np.median(err), np.percentile(err, 99))
    same = (bad.cpu().numpy() == o_bad.numpy()) & (done.cpu().numpy() == o_done.numpy())
    assert (~same).sum() <= (0 if strict_flags else 2), (k, int((~same).sum()))
    assert np.array_equal(exc.cpu().numpy(), o_exc.numpy()), k
    do = np.abs(obs.cpu().numpy() - o_obs.numpy())
    assert np.median(do.max(axis=1)) <= 2e-5 and np.percentile(do.max(axis=1), 99) <= 2e-3, (k, np.median(do.max(axis=1)), do.max())
    assert np.allclose(rew.cpu().numpy(), o_rew.numpy(), rtol=2e-4, atol=2e-6), (k, np.abs(rew.cpu().numpy() - o_rew.numpy()).max())
    assert np.allclose(env.blood.cpu().numpy(), orc.blood.numpy(), rtol=1e-5, atol=2e-3), k
    assert np.array_equal(env.step_count.cpu().numpy(), orc.step_count.numpy().astype(np.int32)), k


def test_combat_teacher_forced_vs_oracle():
    from oracle.combat_oracle import CombatOracle
    num_envs, seed = 1024, 31
    env, orc = _env(num_envs), CombatOracle(num_envs)
    n = env.n
    d0 = tapes.reset_draw_tape(seed, 0, n)
    obs0 = env.reset(reset_draws=_cuda(d0))
    o0 = orc.reset(torch.from_numpy(d0))
    _obs_close(obs0.cpu().numpy(), o0.numpy(), "reset")
    assert torch.equal(env.blood, torch.full_like(env.blood, 100.0)) and int(env.step_count.max()) == 0
    # bring a third of the pairs into gun range so blood / Shutdown / Crash fire, with a few nearly-dead aircraft
    ego, enm = orc.ego, orc.enm
    gap = torch.linspace(50.0, 12000.0, num_envs)
    close = torch.arange(num_envs) % 3 == 0
    orc.s[enm[close], 0] = orc.s[ego[close], 0] + gap[close]
    orc.s[enm[close], 1] = orc.s[ego[close], 1] + 0.03 * gap[close]
    orc.s[enm[close], 2] = orc.s[ego[close], 2] + 15.0
    orc.s[ego[close], 5] = 0.0; orc.s[enm[close], 5] = 0.0
    orc.blood[enm[::9]] = 0.4; orc.blood[ego[3::27]] = 0.3
    events = 0
    for k in range(1, 13):
        _push(env, orc, first=(k == 1))
        a = tapes.action_tape(seed, k, n, 0.5)
        d = tapes.reset_draw_tape(seed, k, n)
        out = env.step(_cuda(a), reset_draws=_cuda(d))
        ref = orc.step(torch.from_numpy(a), torch.from_numpy(d))
        _compare(env, orc, out, ref, k, strict_flags=False)
        events += int(ref[2].sum()) + int(ref[3].sum())
    assert events > 20
    c = env.termination_counters()
    assert c["crash_or_shutdown"] > 0 and c["enemy_shutdown"] > 0 and "unreach" not in c   # combat names its own counters


@pytest.mark.parametrize("fixture,close", [("combat_traj.npz", False), ("combat_close_traj.npz", True)])
def test_combat_fixture_first_steps_and_events(fixture, close):
    g = np.load(os.path.join(GOLDEN, fixture))
    num_envs, steps, seed = [int(x) for x in g["meta"]]
    env = _env(num_envs)
    n = env.n
    obs0 = env.reset(reset_draws=_cuda(tapes.reset_draw_tape(seed, 0, n)))
    _obs_close(obs0.cpu().numpy(), g["obs0"], "reset")
    env.model.s[:] = _cuda(g["s_start"]); env.blood[:] = _cuda(g["blood_start"])
    for k in range(1, 4):
        obs, rew, done, bad, exc, _ = env.step(_cuda(tapes.action_tape(seed, k, n, 0.2 if close else 1.0)),
                                               reset_draws=_cuda(tapes.reset_draw_tape(seed, k, n)))
        # discrete events are robust (blood thresholds, 200 ft crash radius); states only while fp32 noise is small
        assert np.array_equal(done.cpu().numpy(), g[f"k{k}_done"]) and np.array_equal(bad.cpu().numpy(), g[f"k{k}_bad"]), k
        assert np.array_equal(env.step_count.cpu().numpy(), g[f"k{k}_step_count"]), k
        assert np.allclose(env.blood.cpu().numpy(), g[f"k{k}_blood"], rtol=1e-4, atol=5e-3), k
        if k == 1:
            err = state_rel_err(env.model.s.cpu().numpy(), g["k1_s"])
            assert np.median(err) <= 2e-5, np.median(err)
            assert np.allclose(rew.cpu().numpy(), g["k1_reward"], rtol=5e-4, atol=5e-6)


def test_combat_full_size_invariants():
    num_envs = 250_000
    env = _env(num_envs)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(11)
    for k in range(4):
        a = torch.rand((env.n, 4), device="cuda", generator=g) * 2 - 1
        obs, rew, done, bad, exc, _ = env.step(a)
        assert obs.shape == (env.n, 15) and torch.isfinite(obs).all() and torch.isfinite(rew).all()
        # pair-level conditions hit both aircraft of a pair: Shutdown-done (the enemy is shot down) sets `done` on both, so the
        # done flags agree pairwise; a reset pair starts the next step with both step counts at the same value
        assert torch.equal(done[0::2], done[1::2])
        assert torch.equal(env.step_count[0::2], env.step_count[1::2])
        assert bool((obs[0::2, 13] == obs[1::2, 13]).all()) and bool((obs[0::2, 14] == -obs[1::2, 14]).all())
        assert float(env.blood.max()) <= 100.0
        assert int(env.step_count.max()) <= 5 * (k + 1)


def test_records_and_relgeo_match_the_fused_obs():
    """The exchange path (records -> gather -> np_combat_relgeo) must reproduce the pairwise columns the fused kernel
    writes into the observation, bit for bit (same device code on the same inputs)."""
    from neuralplane_b200.combat_exchange import gather_records, local_records, relative_geometry
    num_envs = 4096
    env = _env(num_envs)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(2)
    for k in range(3):
        obs, *_ = env.step(torch.rand((env.n, 4), device="cuda", generator=g) * 2 - 1)
    rec = gather_records(local_records(env))                 # single rank: identity gather
    assert rec.shape == (env.n, 8)
    assert torch.equal(rec[:, 0], env.model.s[:, 0]) and torch.equal(rec[:, 7], env.blood)
    ego = torch.arange(num_envs) * 2
    geo = relative_geometry(rec, ego, ego + 1)
    o_ego, o_enm = obs[0::2], obs[1::2]
    assert torch.equal(geo[:, 3], o_ego[:, 11]) and torch.equal(geo[:, 4], o_ego[:, 12])           # AO2, TA2
    assert torch.equal(geo[:, 6], o_ego[:, 14])                                                    # side flag
    assert torch.allclose(geo[:, 5] * 0.3048 / 10000, o_ego[:, 13], rtol=1e-6, atol=0)
    assert torch.allclose(geo[:, 7] * 0.3048 / 340, o_ego[:, 9], rtol=1e-6, atol=1e-9)
    assert torch.allclose(3.141592653589793 - geo[:, 4], o_enm[:, 11], atol=1e-6)
    geo_r = relative_geometry(rec, ego + 1, ego)              # swapped roles: same range, mirrored side
    assert torch.equal(geo_r[:, 2], geo[:, 2]) and torch.equal(geo_r[:, 5], geo[:, 5])


def test_relgeo_peers_kernel_equals_relgeo_on_the_gathered_array():
    """np_combat_relgeo_peers addresses record g at slabs[g / n_local] + 8 (g % n_local): with the slabs of three virtual
    ranks as ordinary tensors of this GPU (the pointer array a symmetric-memory rendezvous would supply), it must return the
    bits of np_combat_relgeo on the concatenated array.  The real 2-GPU NVLink run is tools/combat_exchange_check.py
    (profiles/r01e_combat_exchange_2gpu_p2p.json)."""
    from neuralplane_b200 import _native as nv
    from neuralplane_b200.combat_exchange import local_records, relative_geometry
    num_envs = 3 * 1000
    env = _env(num_envs)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(5)
    for k in range(2):
        env.step(torch.rand((env.n, 4), device="cuda", generator=g) * 2 - 1)
    rec = local_records(env)                                   # [6000, 8]
    world, n_local = 3, env.n // 3
    slabs = [rec[r * n_local:(r + 1) * n_local].clone() for r in range(world)]     # three separate allocations
    ptrs = torch.tensor([t.data_ptr() for t in slabs], dtype=torch.int64, device="cuda")
    ego = torch.randint(0, env.n, (5000,), generator=torch.Generator().manual_seed(1))
    enm = torch.randint(0, env.n, (5000,), generator=torch.Generator().manual_seed(2))
    want = relative_geometry(rec, ego, enm)
    got = torch.empty_like(want)
    e32, m32 = ego.to("cuda", torch.int32), enm.to("cuda", torch.int32)
    st = nv.lib().np_combat_relgeo_peers(ptrs.data_ptr(), world, n_local, e32.data_ptr(), m32.data_ptr(), got.data_ptr(), ego.numel(),
                                         torch.cuda.current_stream().cuda_stream)
    nv.check(st, "np_combat_relgeo_peers")
    torch.cuda.synchronize()
    assert torch.equal(got, want)


def test_combat_full_size_sampled_parity():
    """BASELINE configs[4] size (5 x 10^5 pairs): pairs are independent, so a random sample of pairs is replayed through the
    combat oracle from the same pre-step state, controls, blood, flags and controller state (bars of the teacher-forced test)."""
    from oracle.combat_oracle import CombatOracle
    from neuralplane_b200 import _native as nv
    E, m = 500_000, 1024
    env = _env(E)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(12)
    for _ in range(3):
        env.step(torch.rand((env.n, 4), device="cuda", generator=g) * 2 - 1)
    pe = np.sort(np.random.default_rng(2).choice(E, m, replace=False))
    idx = torch.from_numpy(np.stack((2 * pe, 2 * pe + 1), 1).reshape(-1)).cuda()          # both aircraft of the sampled pairs
    orc = CombatOracle(m)
    orc.s = env.model.s[idx].cpu(); orc.u = env.model.u[idx].cpu()
    orc.step_count = env.step_count[idx].cpu().long()
    orc.is_done = env.is_done[idx].cpu().clone(); orc.bad_done = env.bad_done[idx].cpu().clone()
    orc.exceed_time_limit = env.exceed_time_limit[idx].cpu().clone()
    orc.blood = env.blood[idx].cpu().clone()
    cs = env.ctrl_state[idx].cpu()
    orc.roll_dem, orc.pitch_dem = cs[:, 9].clone(), cs[:, 10].clone()
    for j, k in enumerate(("roll", "pitch", "yaw")):
        orc.pid[k].error, orc.pid[k].integrator, orc.last_out[k] = cs[:, 3 * j].clone(), cs[:, 3 * j + 1].clone(), cs[:, 3 * j + 2].clone()
        orc.pid[k].first = False
    a = torch.rand((env.n, 4), device="cuda", generator=g) * 2 - 1
    d = torch.rand((env.n, 5), device="cuda", generator=g)
    obs, rew, done, bad, exc, _ = env.step(a, reset_draws=d)
    o_obs, o_rew, o_done, o_bad, o_exc = orc.step(a[idx].cpu(), d[idx].cpu())
    err = state_rel_err(env.model.s[idx].cpu().numpy(), orc.s.numpy())
    assert np.median(err) <= 2e-5 and np.percentile(err, 99) <= 3e-4, (np.median(err), np.percentile(err, 99))
    assert (bad[idx].cpu() != o_bad).sum() <= 2 and (done[idx].cpu() != o_done).sum() <= 2
    do = np.abs(obs[idx].cpu().numpy() - o_obs.numpy()).max(axis=1)
    assert np.median(do) <= 2e-5 and np.percentile(do, 99) <= 2e-3, (np.median(do), do.max())
    ok = ((bad[idx].cpu() == o_bad) & (done[idx].cpu() == o_done)).numpy()
    assert np.allclose(rew[idx].cpu().numpy()[ok], o_rew.numpy()[ok], rtol=2e-4, atol=2e-6)
    assert np.allclose(env.blood[idx].cpu().numpy(), orc.blood.numpy(), rtol=1e-5, atol=2e-3)
    assert np.array_equal(env.step_count[idx].cpu().numpy(), orc.step_count.numpy().astype(np.int32))
    assert torch.isfinite(env.model.s).all() and torch.isfinite(obs).all() and torch.isfinite(rew).all()


def _role_pair(num_envs, seed=0, first_env=0):
    from neuralplane_b200 import SingleCombatEnv
    from neuralplane_b200.combat_exchange import LocalPairExchange
    mk = lambda r: SingleCombatEnv(num_envs=num_envs, config="selfplay", random_seed=seed, device="cuda:0", layout="role", role=r,
                                   first_env=first_env)
    e0, e1 = mk(0), mk(1)
    return e0, e1, LocalPairExchange(e0, e1)


def test_role_sharded_step_is_bit_identical_to_pair_sharded():
    """SingleCombatEnv(layout='role') -- one aircraft of every env per population, the partner's 28-float record pulled between
    the local half and the pair half of the step -- against the pair-sharded env on the same envs: obs / reward / flags /
    state / controls / blood / step counts bit-equal for 20 steps, with Crash, Shutdown and env-level resets occurring, under
    in-kernel Philox draws keyed by global aircraft index (index_base + 2 i).  Two role populations in one process here
    (LocalPairExchange); the 2-GPU NVLink / NCCL run is tools/combat_role_check.py."""
    from neuralplane_b200 import SingleCombatEnv
    E, first = 3000, 1234
    pair = SingleCombatEnv(num_envs=E, config="selfplay", random_seed=5, device="cuda:0", index_base=2 * first)
    e0, e1, ex = _role_pair(E, seed=5, first_env=first)
    o = pair.reset()
    o0, o1 = ex.reset_both()
    assert torch.equal(o[0::2], o0) and torch.equal(o[1::2], o1)

    def force_events():
        # bring a third of the pairs within the 200 ft crash radius / gun range, with some nearly dead aircraft
        close = torch.arange(E, device="cuda") % 3 == 0
        gap = torch.linspace(30.0, 9000.0, E, device="cuda")
        s = pair.model.s
        s[1::2][close, 0] = s[0::2][close, 0] + gap[close]
        s[1::2][close, 1] = s[0::2][close, 1] + 0.03 * gap[close]
        s[1::2][close, 2] = s[0::2][close, 2] + 15.0
        s[0::2][close, 5] = 0.0
        s[1::2][close, 5] = 0.0
        pair.blood[1::18] = 0.4
        pair.blood[6::54] = 0.3
        for r, e in enumerate((e0, e1)):
            e.model.s.copy_(pair.model.s[r::2])
            e.blood.copy_(pair.blood[r::2])
    force_events()
    g = torch.Generator(device="cuda").manual_seed(3)
    events = resets = 0
    for k in range(20):
        a = torch.rand((2 * E, 4), device="cuda", generator=g) - 0.5
        before = int(pair.termination_counters()["resets"])
        rp = pair.step(a)
        r0, r1 = ex.step_both(a[0::2].contiguous(), a[1::2].contiguous())
        for r, (rr, e) in enumerate(((r0, e0), (r1, e1))):
            for x, y, what in zip(rp[:5], rr, ("obs", "reward", "done", "bad", "exc")):
                assert torch.equal(x[r::2], y), (k, r, what)
            assert torch.equal(pair.model.s[r::2], e.model.s) and torch.equal(pair.model.u[r::2], e.model.u), (k, r)
            assert torch.equal(pair.blood[r::2], e.blood) and torch.equal(pair.step_count[r::2], e.step_count), (k, r)
            assert torch.equal(pair.ctrl_state[r::2], e.ctrl_state), (k, r)
        events += int(rp[2].sum()) + int(rp[3].sum())
        resets += int(pair.termination_counters()["resets"]) - before
        if k == 8:
            force_events()
    assert events > 200 and resets > 200          # Crash / Shutdown fired and env-level resets happened inside the compared window
    cp, c0, c1 = pair.termination_counters(), e0.termination_counters(), e1.termination_counters()
    for name in ("overload", "low_altitude", "high_speed", "low_speed", "extreme_state", "resets"):
        assert cp[name] == c0[name] + c1[name], name
    assert cp["crash_or_shutdown"] == c0["crash_or_shutdown"] + c1["crash_or_shutdown"]


@pytest.mark.parametrize("cls,E", [("single", 3000), ("single", 7001), ("multi", 1500)])
def test_combat_kernel_shape_is_bit_identical(cls, E):
    """Small combat populations run the pair-sharded step on K1c<MODE_COMBAT> (up to 18 944 aircraft) or on K5 with 128-thread
    CTAs (latency bound, like the step).  Neither may change a bit: the default dispatch against K5 with 128-thread CTAs
    (NPLANE_COOP_PAIRS=0) and against the 384-thread shape (NPLANE_BLOCK=384), 1-v-1 and 2-v-2, with Crash / Shutdown /
    env-level resets occurring."""
    import os
    from neuralplane_b200 import MultipleCombatEnv, SingleCombatEnv
    mk = (lambda: SingleCombatEnv(num_envs=E, config="selfplay", random_seed=11, device="cuda:0")) if cls == "single" else \
         (lambda: MultipleCombatEnv(num_envs=E, random_seed=11, device="cuda:0"))
    envs = [mk()]
    for var, val in (("NPLANE_COOP_PAIRS", "0"), ("NPLANE_BLOCK", "384")):
        os.environ[var] = val
        try:
            envs.append(mk())
        finally:
            del os.environ[var]
    coop, k128, k384 = envs
    n = coop.n
    o = [e.reset() for e in envs]
    assert torch.equal(o[0], o[1]) and torch.equal(o[0], o[2])
    for e in envs:                             # a third of the duels inside the crash radius / gun range, some nearly dead
        close = torch.arange(n // 2, device="cuda") % 3 == 0
        s = e.model.s
        s[1::2][close, 0] = s[0::2][close, 0] + 150.0
        s[1::2][close, 1] = s[0::2][close, 1]
        s[1::2][close, 2] = s[0::2][close, 2] + 15.0
        e.blood[1::18] = 0.4
    g = torch.Generator(device="cuda").manual_seed(4)
    for k in range(12):
        a = torch.rand((n, 4), device="cuda", generator=g) - 0.5
        r = [e.step(a) for e in envs]
        for other in r[1:]:
            for x, y in zip(r[0][:5], other[:5]):
                assert torch.equal(x, y), k
    li = coop.launch_info()
    assert li["block"] == (256 if n <= 148 * 64 else 128) and li["grid"] == (n + 63) // 64, li
    assert k128.launch_info()["block"] == 128 and k384.launch_info()["block"] == 384
    for other in (k128, k384):
        assert torch.equal(coop.model.s, other.model.s) and torch.equal(coop.model.u, other.model.u)
        assert torch.equal(coop.blood, other.blood) and torch.equal(coop.ctrl_state, other.ctrl_state)
        assert torch.equal(coop.step_count, other.step_count)
        assert coop.termination_counters() == other.termination_counters()
    c = coop.termination_counters()
    assert c["crash_or_shutdown"] > 0 and c["resets"] > n


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_role_sharded_step_two_ranks():
    """The same comparison across two processes / GPUs with both real exchanges (NVLink peer slabs; NCCL all-gather)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29631", os.path.join(root, "tools", "combat_role_check.py"), "--envs", "20000", "--steps", "20"],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert d["peer_slabs"]["bit_identical"] and d["all_gather"]["bit_identical"] and d["events"] > 100, d

"""GPU tests of the plug-in surface added in ABI v7: the stand-alone model.update() (np_f16_update / np_f16_table_update /
np_uav_update) against what the fused env step does to the same (s, u, action) and against the oracle; the device guard
(an env on cuda:1 while the current device is 0); the block-size choice for strong-scaling shards.
Bit-exact unless stated."""
import numpy as np
import pytest
import torch

from oracle import tapes

pytestmark = pytest.mark.gpu


def _cuda(x, dev="cuda:0"):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dev)


@pytest.mark.parametrize("model", ["F16", "F16_tables", "UAV"])
def test_model_update_equals_fused_step(model):
    """F16Model.update / UAVModel.update (F16_model.py:51-67, UAV_model.py:51-62; caller planning_env.py:161): the state
    and controls after model.update(a) are bit-identical to those after env.step(a) from the same start (no aircraft
    resets inside the compared step), and recent_s / recent_u hold what the update started from."""
    from neuralplane_b200 import ControlEnv
    n = 1001                                        # odd: the tail pair path
    mk = lambda: ControlEnv(num_envs=n, config="control", model=model, random_seed=4, device="cuda:0")
    e_step, e_upd = mk(), mk()
    d0 = _cuda(tapes.reset_draw_tape(4, 0, n))
    e_step.reset(reset_draws=d0)
    e_upd.reset(reset_draws=d0)
    assert torch.equal(e_step.model.s, e_upd.model.s)
    for k in range(1, 6):
        a = _cuda(tapes.action_tape(4, k, n, 0.3))
        s_before, u_before = e_upd.model.s.clone(), e_upd.model.u.clone()
        live = ~(e_step.is_done | e_step.bad_done | e_step.exceed_time_limit)   # aircraft the step will not re-initialise
        e_step.step(a, reset_draws=_cuda(tapes.reset_draw_tape(4, k, n)))
        e_upd.model.update(a)
        assert torch.equal(e_upd.model.recent_s, s_before)
        if model != "UAV":
            assert torch.equal(e_upd.model.recent_u[:, :4], u_before[:, :4])
        assert live.float().mean() > 0.9
        assert torch.equal(e_upd.model.s[live], e_step.model.s[live]), k
        assert torch.equal(e_upd.model.u[live], e_step.model.u[live]), k
        # keep the two populations identical for the next round (the env may have flagged aircraft for reset)
        e_upd.model.s.copy_(e_step.model.s)
        e_upd.model.u.copy_(e_step.model.u)


def test_f16_update_vs_oracle():
    """np_f16_update against the CPU oracle's F16 update (oracle/f16_oracle.py, pinned to the reference) from 20000 random
    in-envelope (s, u): the same bar as the fused step's single-step test -- our distance to the float64 truth is at most
    2x the reference-fp32 distance (p50 and p99); controls to 1e-6."""
    from neuralplane_b200 import ControlEnv
    from oracle.f16_oracle import AeroNets, euler_step, lowpass_controls
    from _metrics import state_rel_err
    n = 20000
    s0, u0 = tapes.random_envelope_states(41, n)
    a = tapes.action_tape(42, 1, n, 1.0)
    env = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=0, device="cuda:0")
    env.model.s[:] = _cuda(s0); env.model.u[:] = _cuda(u0)
    env.model.update(_cuda(a))
    res = {}
    for dt in (torch.float32, torch.float64):
        u = lowpass_controls(torch.from_numpy(u0).to(dt), torch.from_numpy(a).to(dt))          # F16_model.py:51-63
        res[dt] = (euler_step(AeroNets(dtype=dt), torch.from_numpy(s0).to(dt), u, 0.02).numpy(), u.numpy())   # :64-67
    truth = res[torch.float64][0]
    e_ours, e_ref = state_rel_err(env.model.s.cpu().numpy(), truth), state_rel_err(res[torch.float32][0], truth)
    assert np.percentile(e_ours, 99) <= 2 * np.percentile(e_ref, 99) and np.median(e_ours) <= 2 * np.median(e_ref), (
        np.median(e_ours), np.percentile(e_ours, 99), np.median(e_ref), np.percentile(e_ref, 99))
    assert np.allclose(env.model.u.cpu().numpy()[:, :4], res[torch.float32][1][:, :4], rtol=1e-6, atol=1e-6)
    assert torch.equal(env.model.recent_s.cpu(), torch.from_numpy(s0))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_env_on_another_device_than_current():
    """an env created with device='cuda:1' must work while the process's current device is 0 (train and eval envs
    on different GPUs in one process) -- every native entry point switches to the env's device and back."""
    from neuralplane_b200 import ControlEnv, GPUVecEnv, PlanningEnv
    torch.cuda.set_device(0)
    n = 3000
    e0 = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=5, device="cuda:0")
    e1 = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=5, device="cuda:1")
    assert torch.cuda.current_device() == 0
    o0, o1 = e0.reset(), e1.reset()
    assert o1.device.index == 1 and torch.equal(o0.cpu(), o1.cpu())
    for k in range(1, 8):
        a = tapes.action_tape(5, k, n, 1.0)
        r0, r1 = e0.step(_cuda(a)), e1.step(_cuda(a, "cuda:1"))
        assert torch.cuda.current_device() == 0
        assert torch.equal(r0[0].cpu(), r1[0].cpu()) and torch.equal(r0[3].cpu(), r1[3].cpu())
    assert e0.termination_counters() == e1.termination_counters()
    assert torch.equal(e0.model.get_extended_state().cpu(), e1.model.get_extended_state().cpu())
    v1 = GPUVecEnv([lambda: ControlEnv(num_envs=n, config="heading", model="F16", random_seed=5, device="cuda:1")])
    v0 = GPUVecEnv([lambda: ControlEnv(num_envs=n, config="heading", model="F16", random_seed=5, device="cuda:0")])
    assert np.array_equal(v0.reset(), v1.reset())
    a = tapes.action_tape(5, 1, n, 1.0).reshape(n, 1, 4)
    assert np.array_equal(v0.step(a)[0], v1.step(a)[0])
    p1 = PlanningEnv(num_envs=256, config="tracking", random_seed=1, device="cuda:1", n_substeps=3)
    p1.reset()
    p1.step(torch.zeros((256, 3), device="cuda:1"))
    assert torch.isfinite(p1.last_obs).all() and torch.cuda.current_device() == 0


def test_block_choice_is_bit_identical_and_saves_a_wave():
    """A strong-scaling shard of 125 000 aircraft is 1.1 waves of 148 x 384 pairs: pick_block() takes 512-thread CTAs (one
    wave) there and 384 at 10^6.  The block size must not change a single bit of any output."""
    import os
    from neuralplane_b200 import ControlEnv
    n = 125_000
    env = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=2, device="cuda:0")
    os.environ["NPLANE_BLOCK"] = "384"
    try:
        ref = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=2, device="cuda:0")
    finally:
        del os.environ["NPLANE_BLOCK"]
    env.reset(); ref.reset()
    for k in range(1, 6):
        a = _cuda(tapes.action_tape(2, k, n, 1.0))
        r, q = env.step(a), ref.step(a)
        for x, y in zip(r[:5], q[:5]):
            assert torch.equal(x, y), k
    assert env.launch_info()["block"] == 512 and ref.launch_info()["block"] == 384
    assert torch.equal(env.model.s, ref.model.s) and env.termination_counters() == ref.termination_counters()
    # mid-size populations are latency bound: 128-thread CTAs (one warp per scheduler, 3x as many SMs), same bits
    n = 30_000
    small = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=2, device="cuda:0")
    os.environ["NPLANE_BLOCK"] = "384"
    try:
        ref = ControlEnv(num_envs=n, config="heading", model="F16", random_seed=2, device="cuda:0")
    finally:
        del os.environ["NPLANE_BLOCK"]
    small.reset(); ref.reset()
    for k in range(1, 60):
        a = _cuda(tapes.action_tape(2, k, n, 1.0))
        for x, y in zip(small.step(a)[:5], ref.step(a)[:5]):
            assert torch.equal(x, y), k
    assert small.launch_info()["block"] == 128 and small.launch_info()["grid"] == 118 and ref.launch_info()["grid"] == 40
    assert small.termination_counters() == ref.termination_counters()
    big = ControlEnv(num_envs=1_000_000, config="heading", model="F16", random_seed=2, device="cuda:0")
    big.reset()
    big.step(torch.zeros((1_000_000, 4), device="cuda:0"))
    assert big.launch_info()["block"] == 384


@pytest.mark.parametrize("config,n", [("heading", 3000), ("heading", 1), ("heading", 63), ("control", 2999), ("tracking", 9472),
                                      ("control", 9473), ("tracking", 18_944), ("heading", 18_945)])
def test_cooperative_small_population_kernel_is_bit_identical(config, n):
    """K1c (coop_step_kernel.cuh) deals a pair's 21 MLP evaluations over the four warps of a CTA to cut the latency of a step
    at the reference's training sizes (3 000 envs).  Same device functions on the same operands: every output, the state, the
    coefficient cache (through the following steps) and the counters must equal K1's bit for bit -- across episodic resets
    (terminations of the random actions plus flags raised by hand), ragged and odd populations, all three tasks."""
    import os
    from neuralplane_b200 import ControlEnv
    kw = dict(num_envs=n, config=config, model="F16", random_seed=5, device="cuda:0")
    coop = ControlEnv(**kw)
    os.environ["NPLANE_COOP_PAIRS"] = "0"
    try:
        ref = ControlEnv(**kw)
    finally:
        del os.environ["NPLANE_COOP_PAIRS"]
    coop.reset(); ref.reset()
    for k in range(1, 80):
        a = _cuda(tapes.action_tape(5, k, n, 1.5))
        for x, y in zip(coop.step(a)[:5], ref.step(a)[:5]):
            assert torch.equal(x, y), k
        if k % 20 == 0:                                     # some lanes of a warp reset, the others hit the coefficient cache
            for e in (coop, ref):
                e.is_done[::7] = True
                e.bad_done[3::64] = True
    li, lr = coop.launch_info(), ref.launch_info()
    if n <= 18_944:                                         # eight warps per CTA while one CTA per SM covers the population
        assert li["block"] == (256 if n <= 148 * 64 else 128) and li["grid"] == (n + 63) // 64, li
        assert lr["block"] == 128 and lr["grid"] == (n + 255) // 256, lr          # K1's 128-thread CTAs: 256 aircraft each
    else:
        assert li == lr                                     # above one wave of K1c CTAs both run K1
    assert torch.equal(coop.model.s, ref.model.s) and torch.equal(coop.model.u, ref.model.u)
    assert torch.equal(coop.step_count, ref.step_count)
    assert coop.termination_counters() == ref.termination_counters()
    assert coop.termination_counters()["resets"] > 0


@pytest.mark.parametrize("env_name", ["control", "planning", "combat"])
def test_cooperative_kernel_over_several_waves(env_name):
    """The dispatch gives K1c at most one wave of CTAs, but its loop must also be right when a CTA flies several groups of 32
    pairs (tile, coefficient slots and exchange rows reused across iterations): forced onto 45 001 / 40 000 aircraft
    (NPLANE_COOP_PAIRS=10^8, four warps, 296 CTAs -> three iterations) against K1, bit for bit."""
    import os
    from neuralplane_b200 import ControlEnv, PlanningEnv, SingleCombatEnv
    if env_name == "control":
        n, A = 45_001, 4
        mk = lambda: ControlEnv(num_envs=n, config="control", model="F16", random_seed=6, device="cuda:0")  # noqa: E731
    elif env_name == "planning":
        n, A = 45_001, 3
        mk = lambda: PlanningEnv(num_envs=n, config="tracking", model="F16", random_seed=6, device="cuda:0", n_substeps=3)  # noqa: E731
    else:
        n, A = 40_000, 4
        mk = lambda: SingleCombatEnv(num_envs=n // 2, config="selfplay", random_seed=6, device="cuda:0")  # noqa: E731
    ref = mk()                                   # 45 001 aircraft: K1 / K4 / K5 with 128-thread CTAs
    os.environ["NPLANE_COOP_PAIRS"] = "100000000"
    try:
        coop = mk()
    finally:
        del os.environ["NPLANE_COOP_PAIRS"]
    assert torch.equal(coop.reset(), ref.reset())
    for k in range(1, 9):
        a = _cuda(tapes.action_tape(6, k, n, 1.0, num_actions=A)) if A == 3 else _cuda(tapes.action_tape(6, k, n, 1.0))
        if env_name == "combat":
            a = a * 0.5
        for x, y in zip(coop.step(a)[:5], ref.step(a)[:5]):
            assert torch.equal(x, y), k
        if k % 3 == 0:
            for e in (coop, ref):
                e.is_done[::7] = True
    li = coop.launch_info()
    assert li["block"] == 128 and li["grid"] == 296 and ref.launch_info()["grid"] == (n + 255) // 256, (li, ref.launch_info())
    assert torch.equal(coop.model.s, ref.model.s) and torch.equal(coop.model.u, ref.model.u)
    assert coop.termination_counters() == ref.termination_counters()


@pytest.mark.parametrize("n", [3000, 40_000])
def test_cuda_graph_replay_equals_eager_loop(n):
    """A rollout loop captured in a CUDA graph: K1c (n = 3 000) is launched with programmatic stream serialisation (the next
    step's CTAs start behind this step's dependency wait), K1 (n = 40 000) with ordinary launches.  The RNG counter the step
    calls advance is a host-side launch argument, frozen by the capture; env.advance_rng(K), captured at the end of the K
    steps, moves a device-side epoch instead.  Three replays of a 10-step graph must leave exactly the state, outputs (with
    in-kernel observation noise and reset draws) and counters of 30 eager steps with ordinary launches (NPLANE_PDL=0)."""
    import os
    from neuralplane_b200 import ControlEnv
    kw = dict(num_envs=n, config="heading", model="F16", random_seed=9, device="cuda:0")
    env = ControlEnv(**kw)
    os.environ["NPLANE_PDL"] = "0"
    try:
        ref = ControlEnv(**kw)
    finally:
        del os.environ["NPLANE_PDL"]
    env.reset(); ref.reset()
    a = _cuda(tapes.action_tape(9, 1, n, 1.0))
    for e in (env, ref):                     # warm-up outside the capture (first launch sets the kernel attributes)
        e.step(a)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            out = env.step(a)
        env.advance_rng(10)
    for _ in range(3):
        g.replay()
    for _ in range(30):
        exp = ref.step(a)
    torch.cuda.synchronize()
    for x, y in zip(out[:5], exp[:5]):
        assert torch.equal(x, y)
    assert torch.equal(env.model.s, ref.model.s) and torch.equal(env.step_count, ref.step_count)
    c = env.termination_counters()
    assert c == ref.termination_counters() and c["resets"] > n      # full-scale actions: episodes end, fresh draws each replay


def test_rollout_attach_rejects_odd_population():
    """rewards[t] of an odd population is only 4-byte aligned for odd t; attach() must say so up front."""
    import types
    from neuralplane_b200 import ControlEnv
    from neuralplane_b200.rollout import DeviceRolloutBuffer
    env = ControlEnv(num_envs=33, config="heading", model="F16", random_seed=0, device="cuda:0")
    args = types.SimpleNamespace(buffer_size=8, n_rollout_threads=33, gamma=0.99, use_proper_time_limits=False, use_gae=True,
                                 gae_lambda=0.95, recurrent_hidden_size=8, recurrent_hidden_layers=1)
    buf = DeviceRolloutBuffer(args, 1, env.observation_space, env.action_space, "cuda:0")
    with pytest.raises(ValueError, match="even"):
        buf.attach(env)

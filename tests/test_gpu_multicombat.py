"""GPU parity of MultipleCombatEnv (2-v-2: np_env_combat_step with combat_pairs_per_env = 2, one FDM step per env step)
against oracle/combat_oracle.py:MultiCombatOracle.  Every per-duel formula is the 1-v-1 code already pinned to the
reference's own functions (tests/golden/combat*_traj.npz, tests/test_gpu_combat.py); what is checked here is the restated
orchestration -- two adjacent duels per env, env-level reset over all four aircraft, reward without the 0.01 factor,
+-10 000 ft initial box (parity unpinned: the reference file cannot run, see neuralplane_b200/envs/multiplecombat_env.py).
Teacher-forced per env step like the 1-v-1 test; same fp32 bars."""
import numpy as np
import pytest
import torch

from oracle import tapes
from _metrics import state_rel_err

pytestmark = pytest.mark.gpu


def _cuda(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def _push(env, orc, first):
    from neuralplane_b200 import _native as nv
    n = env.n
    env.model.s[:] = _cuda(orc.s.numpy()); env.model.u[:] = _cuda(orc.u.numpy())
    env.step_count[:] = _cuda(orc.step_count.numpy().astype(np.int32))
    for j, f in enumerate((orc.is_done, orc.bad_done, orc.exceed_time_limit)):
        env._flags[j, :n] = _cuda(f.numpy().astype(np.uint8))
    env.blood[:] = _cuda(orc.blood.numpy())
    cs = orc.ctrl_state().numpy()
    st = np.zeros((n, 12), np.float32)
    st[:, 0:9] = cs[:, 2:11]; st[:, 9] = cs[:, 0]; st[:, 10] = cs[:, 1]
    env.ctrl_state[:] = _cuda(st)
    nv.check(nv.lib().np_env_set_pid_started(env._handle, 0 if first else 1), "np_env_set_pid_started")


def test_multicombat_teacher_forced_vs_oracle():
    from neuralplane_b200 import MultipleCombatEnv
    from oracle.combat_oracle import MultiCombatOracle
    E, seed = 512, 17
    env = MultipleCombatEnv(num_envs=E, random_seed=0, device="cuda:0")
    orc = MultiCombatOracle(E)
    n = env.n
    assert n == 4 * E and env.num_agents == 4 and env.n_substeps == 1
    d0 = tapes.reset_draw_tape(seed, 0, n)
    obs0 = env.reset(reset_draws=_cuda(d0))
    o0 = orc.reset(torch.from_numpy(d0))
    assert np.abs(obs0.cpu().numpy() - o0.numpy())[:, [j for j in range(15) if j not in (11, 12)]].max() < 5e-5
    assert float(env.model.s[:, 0].abs().max()) > 5000.0          # the +-10 000 ft box of multiple_selfplay.yaml
    ego, enm = orc.ego, orc.enm
    close = torch.arange(2 * E) % 3 == 0                              # a third of the duels into gun / crash range
    gap = torch.linspace(50.0, 9000.0, 2 * E)
    orc.s[enm[close], 0] = orc.s[ego[close], 0] + gap[close]
    orc.s[enm[close], 1] = orc.s[ego[close], 1] + 0.03 * gap[close]
    orc.s[enm[close], 2] = orc.s[ego[close], 2] + 15.0
    orc.s[ego[close], 5] = 0.0; orc.s[enm[close], 5] = 0.0
    orc.blood[enm[::9]] = 0.4; orc.blood[ego[3::27]] = 0.3
    events = group_resets = 0
    for k in range(1, 25):
        _push(env, orc, first=(k == 1))
        flagged = (orc.is_done | orc.bad_done | orc.exceed_time_limit).reshape(E, 4)
        a = tapes.action_tape(seed, k, n, 0.5)
        d = tapes.reset_draw_tape(seed, k, n)
        obs, rew, done, bad, exc, _ = env.step(_cuda(a), reset_draws=_cuda(d))
        o_obs, o_rew, o_done, o_bad, o_exc = orc.step(torch.from_numpy(a), torch.from_numpy(d))
        # an env with any flagged aircraft restarts ALL FOUR (step_count 1 after the step), the others keep counting
        restarted = (env.step_count.cpu().reshape(E, 4) == 1)
        if k > 1:
            assert torch.equal(restarted.all(dim=1), flagged.any(dim=1)) and torch.equal(restarted.any(dim=1), restarted.all(dim=1)), k
        group_resets += int((flagged.any(dim=1) & ~flagged.all(dim=1)).sum())
        err = state_rel_err(env.model.s.cpu().numpy(), orc.s.numpy())
        assert np.median(err) <= 5e-6 and np.percentile(err, 99) <= 1e-4, (k, np.median(err), np.percentile(err, 99))
        same = (bad.cpu().numpy() == o_bad.numpy()) & (done.cpu().numpy() == o_done.numpy())
        assert (~same).sum() <= 2, (k, int((~same).sum()))
        do = np.abs(obs.cpu().numpy() - o_obs.numpy()).max(axis=1)
        assert np.median(do) <= 2e-5 and np.percentile(do, 99) <= 2e-3, (k, np.median(do), do.max())
        assert np.allclose(rew.cpu().numpy(), o_rew.numpy(), rtol=2e-4, atol=2e-4), k          # rewards are O(1) here (no 0.01)
        assert np.allclose(env.blood.cpu().numpy(), orc.blood.numpy(), rtol=1e-5, atol=2e-3), k
        assert np.array_equal(env.step_count.cpu().numpy(), orc.step_count.numpy().astype(np.int32)), k
        events += int(o_done.sum()) + int(o_bad.sum())
    assert events > 20 and group_resets > 5          # resets triggered by ONE duel of an env re-initialised the other duel too
    assert float(rew.abs().max()) > 0.2              # the reward is not scaled by 0.01


def test_multicombat_shapes_and_invariants():
    from neuralplane_b200 import GPUVecEnv, MultipleCombatEnv
    E = 50_000
    env = MultipleCombatEnv(num_envs=E, random_seed=1, device="cuda:0")
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(3)
    for k in range(5):
        obs, rew, done, bad, exc, _ = env.step(torch.rand((env.n, 4), device="cuda", generator=g) * 2 - 1)
        assert obs.shape == (4 * E, 15) and torch.isfinite(obs).all() and torch.isfinite(rew).all()
        assert bool((obs[0::2, 13] == obs[1::2, 13]).all()) and bool((obs[0::2, 14] == -obs[1::2, 14]).all())   # per-duel mirroring
        assert torch.equal(done[0::2], done[1::2])
        sc = env.step_count.reshape(E, 4)
        assert bool((sc == sc[:, :1]).all())          # the four aircraft of an env always share their episode clock
    v = GPUVecEnv([lambda: MultipleCombatEnv(num_envs=64, random_seed=1, device="cuda:0")])
    o = v.reset()
    r = v.step(np.zeros((64, 4, 4), np.float32))
    assert o.shape == (64, 4, 15) and r[0].shape == (64, 4, 15) and r[1].shape == (64, 4, 1) and r[2].dtype == np.bool_
    with pytest.raises(NotImplementedError):       # a 2-agent yaml is not a 2-v-2 config
        MultipleCombatEnv(num_envs=2, config="selfplay", random_seed=0, device="cuda:0")

"""GPU parity of the device-resident rollout buffer (SURVEY f-2) against the reference's ReplayBuffer fixture
(tests/golden/rollout_returns.npz) and the numpy oracle: returns bit-exact (fp32, same operation order), masks exact,
slot conventions, the mini-batch generator, and the zero-copy coupling with a native env."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import rollout_oracle as ro

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rollout_returns.npz"))


def _args(T, N, use_gae=True, proper=True):
    return types.SimpleNamespace(buffer_size=T, n_rollout_threads=N, gamma=float(G["gamma"][0]), use_proper_time_limits=proper,
                                 use_gae=use_gae, gae_lambda=float(G["gae_lambda"][0]), recurrent_hidden_size=8, recurrent_hidden_layers=1)


def _filled(use_gae, proper):
    from neuralplane_b200.rollout import DeviceRolloutBuffer
    T, N, A, D, act = [int(x) for x in G["meta"][:5]]
    buf = DeviceRolloutBuffer(_args(T, N, use_gae, proper), A, D, act, "cuda:0")
    buf.obs[0].copy_(torch.from_numpy(G["tape_obs0"]))
    for t in range(T):
        masks, bad, reset_env = ro.runner_masks(G["tape_dones"][t], G["tape_bad_dones"][t], G["tape_exceed"][t])
        ra, rc = G["tape_rnn_a"][t].copy(), G["tape_rnn_c"][t].copy()
        ra[reset_env] = 0; rc[reset_env] = 0
        buf.insert(G["tape_obs"][t], G["tape_actions"][t], G["tape_rewards"][t], masks, G["tape_logp"][t], G["tape_values"][t], ra, rc, bad)
    buf.compute_returns(G["tape_next_value"])
    return buf


@pytest.mark.parametrize("use_gae", [True, False])
@pytest.mark.parametrize("proper", [True, False])
def test_compute_returns_bit_exact_vs_reference(use_gae, proper):
    buf = _filled(use_gae, proper)
    key = f"gae{int(use_gae)}_proper{int(proper)}_"
    assert np.array_equal(buf.returns.cpu().numpy(), G[key + "returns"])
    assert np.allclose(buf.advantages.cpu().numpy(), G[key + "advantages"], rtol=1e-5, atol=1e-6)


def test_slots_after_update_and_generator():
    buf = _filled(True, True)
    for k in ("obs", "actions", "rewards", "masks", "bad_masks", "action_log_probs", "value_preds", "rnn_states_actor", "rnn_states_critic"):
        assert np.array_equal(getattr(buf, k).cpu().numpy(), G["buf_" + k]), k
    assert buf.step == int(G["buf_step_after"][0])
    torch.manual_seed(5)
    from neuralplane_b200.rollout import DeviceRolloutBuffer
    batches = list(DeviceRolloutBuffer.recurrent_generator(buf, 2, 4))
    assert len(batches) == 2
    for bi, batch in enumerate(batches):
        for name, x in zip(("obs", "actions", "masks", "logp", "adv", "returns", "values", "rnn_a", "rnn_c"), batch):
            want = G[f"gen{bi}_{name}"]
            assert tuple(x.shape) == want.shape, (name, x.shape, want.shape)
            if name == "adv":
                assert np.allclose(x.cpu().numpy(), want, rtol=1e-5, atol=1e-6)
            else:
                assert np.array_equal(x.cpu().numpy(), want), name
    buf.after_update()
    assert np.array_equal(buf.obs[0].cpu().numpy(), G["after_obs0"]) and np.array_equal(buf.masks[0].cpu().numpy(), G["after_masks0"])
    assert np.array_equal(buf.bad_masks[0].cpu().numpy(), G["after_bad_masks0"])
    assert np.array_equal(buf.rnn_states_actor[0].cpu().numpy(), G["after_rnn_a0"])


def test_returns_large_random_vs_oracle():
    """T = 64, M = 50 000 columns (ragged vs the block size), every variant, against the numpy oracle: bit-exact."""
    from neuralplane_b200.rollout import DeviceRolloutBuffer
    T, N, A = 64, 25_000, 2
    rng = np.random.default_rng(4)
    r = rng.standard_normal((T, N, A, 1)).astype(np.float32)
    v = rng.standard_normal((T + 1, N, A, 1)).astype(np.float32)
    m = (rng.random((T + 1, N, A, 1)) > 0.05).astype(np.float32)
    b = (rng.random((T + 1, N, A, 1)) > 0.05).astype(np.float32)
    nxt = rng.standard_normal((N, A, 1)).astype(np.float32)
    for use_gae in (True, False):
        for proper in (True, False):
            buf = DeviceRolloutBuffer(_args(T, N, use_gae, proper), A, 22, 4, "cuda:0")
            buf.rewards.copy_(torch.from_numpy(r)); buf.value_preds.copy_(torch.from_numpy(v))
            buf.masks.copy_(torch.from_numpy(m)); buf.bad_masks.copy_(torch.from_numpy(b))
            buf.compute_returns(nxt)
            want, _ = ro.compute_returns(r, v, m, b, nxt, buf.gamma, buf.gae_lambda, use_gae, proper)
            assert np.array_equal(buf.returns.cpu().numpy(), want), (use_gae, proper)


def test_zero_copy_rollout_with_native_env():
    """attach(env): the step kernel writes obs / reward straight into the buffer slots; masks from the flag rows equal the
    runner's derivation; the result equals a plain env driven with the same seed and actions."""
    from neuralplane_b200 import ControlEnv
    from neuralplane_b200.rollout import DeviceRolloutBuffer
    T, N = 24, 512
    mk = lambda: ControlEnv(num_envs=N, config="heading", model="F16", random_seed=7, device="cuda:0")  # noqa: E731
    env, plain = mk(), mk()
    A = env.num_agents
    buf = DeviceRolloutBuffer(_args(T, N), A, env.observation_space, env.action_space, "cuda:0")
    buf.attach(env)
    o0 = env.reset()
    p0 = plain.reset().clone()
    assert o0.data_ptr() == buf.obs[0].data_ptr() and torch.equal(buf.obs[0].view(-1, 22), p0)
    g = torch.Generator(device="cuda").manual_seed(1)
    for t in range(T):
        a = torch.rand((env.n, 4), device="cuda", generator=g) * 2 - 1
        logp, val = torch.randn((N, A, 1), device="cuda", generator=g), torch.randn((N, A, 1), device="cuda", generator=g)
        ha = torch.ones((N, A, 1, 8), device="cuda")
        obs, rew, done, bad, exc, _ = buf.step_env(env, a, logp, val, ha, ha)
        po, pr, pd, pb, pe, _ = plain.step(a)
        assert torch.equal(buf.obs[t + 1].view(-1, 22), po) and torch.equal(buf.rewards[t].view(-1), pr)
        assert torch.equal(done, pd) and torch.equal(bad, pb)
        wm, wb, wr = ro.runner_masks(pd.view(N, A, 1).cpu().numpy(), pb.view(N, A, 1).cpu().numpy(), pe.view(N, A, 1).cpu().numpy())
        assert np.array_equal(buf.masks[t + 1].cpu().numpy(), wm) and np.array_equal(buf.bad_masks[t + 1].cpu().numpy(), wb)
        assert np.array_equal(buf.reset_env.cpu().numpy().astype(bool), wr)
        assert np.array_equal(buf.rnn_states_actor[t + 1, :, 0, 0, 0].cpu().numpy(), (~wr).astype(np.float32))
        assert torch.equal(buf.actions[t].view(-1, 4), a) and torch.equal(buf.value_preds[t], val)
    assert buf.step == 0 and int((buf.bad_masks == 0).sum()) > 0
    buf.compute_returns(torch.zeros((N, A, 1), device="cuda"))
    want, _ = ro.compute_returns(buf.rewards.cpu().numpy(), buf.value_preds.cpu().numpy(), buf.masks.cpu().numpy(),
                                 buf.bad_masks.cpu().numpy(), np.zeros((N, A, 1), np.float32), buf.gamma, buf.gae_lambda, True, True)
    assert np.array_equal(buf.returns.cpu().numpy(), want)
    buf.after_update()
    assert torch.equal(buf.obs[0], buf.obs[-1])


def test_whole_rollout_in_one_cuda_graph():
    """The B200-idiomatic way to run the reference's training population (3 000 envs): the whole rollout -- a device policy,
    env.step writing obs / reward in place, mask derivation, value / action inserts, for all T slots -- captured ONCE in a CUDA
    graph and replayed per PPO iteration (env.advance_rng(T) inside the graph keeps the reset / noise streams moving).
    Two replays, with compute_returns + after_update between them, must fill the buffer exactly like 2 x T eager steps."""
    from neuralplane_b200 import ControlEnv
    from neuralplane_b200.rollout import DeviceRolloutBuffer
    T, N = 32, 3000
    mk = lambda: ControlEnv(num_envs=N, config="heading", model="F16", random_seed=7, device="cuda:0")  # noqa: E731
    W = torch.randn((22, 4), device="cuda", generator=torch.Generator(device="cuda").manual_seed(2)) * 8.0   # saturating: episodes end

    def policy(obs):                       # a deterministic stand-in for the actor: any capturable torch code
        x = obs.view(-1, 22)
        return torch.tanh(x @ W), (x[:, :1] * 0.1).view(N, 1, 1)

    def rollout(buf, env):
        for t in range(T):
            a, v = policy(buf.obs[t])
            buf.step_env(env, a, None, v)

    envs, bufs = [mk(), mk()], []
    for env in envs:
        buf = DeviceRolloutBuffer(_args(T, N), env.num_agents, env.observation_space, env.action_space, "cuda:0")
        buf.attach(env)
        env.reset()
        bufs.append(buf)
    (env_g, env_e), (buf_g, buf_e) = envs, bufs
    # one warm-up rollout on both (first launches configure the kernels), then capture the graph on env_g
    for env, buf in zip(envs, bufs):
        rollout(buf, env)
        buf.compute_returns(torch.zeros((N, 1, 1), device="cuda")); buf.after_update()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        rollout(buf_g, env_g)
        env_g.advance_rng(T)
    assert buf_g.step == 0
    for it in range(2):
        g.replay()
        rollout(buf_e, env_e)
        for buf in bufs:
            buf.compute_returns(torch.zeros((N, 1, 1), device="cuda"))
        torch.cuda.synchronize()
        for name in ("obs", "rewards", "masks", "bad_masks", "actions", "value_preds", "returns"):
            assert torch.equal(getattr(buf_g, name), getattr(buf_e, name)), (it, name)
        for buf in bufs:
            buf.after_update()
    assert torch.equal(env_g.model.s, env_e.model.s) and env_g.termination_counters() == env_e.termination_counters()
    assert env_g.termination_counters()["resets"] > N + 100          # episodes ended and were reset inside the graph

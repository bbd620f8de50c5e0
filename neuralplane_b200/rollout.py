"""Device-resident rollout buffer (SURVEY f-2): the reference's ReplayBuffer (algorithms/utils/buffer.py:27-282) with the
same array names, shapes and slot conventions, kept in HBM so a GPU policy never round-trips through numpy.

    buf = DeviceRolloutBuffer(args, num_agents, obs_space, act_space, device)     # same constructor + device
    buf.attach(env); obs0 = env.reset()                                           # env writes obs / reward IN PLACE
    for t in range(buf.buffer_size):
        ... policy on buf.obs[t] ...
        buf.step_env(env, actions, action_log_probs, values, rnn_a, rnn_c)        # env.step + insert, zero copies of obs
    buf.compute_returns(next_value); ...; buf.after_update()

`attach` points the env's obs / reward outputs at slot step+1 / step of this buffer (np_env_rebind_outputs), so the
step kernel's 88 B/aircraft observation store IS the rollout insert; the masks come from the env's flag rows in one
small kernel (np_rollout_masks = F16SimRunner.insert, runner/F16sim_runner.py:141-157) and compute_returns is a
backward-scan kernel (np_rollout_returns = buffer.py:139-172, bit-exact with numpy's fp32 evaluation).
"""
import numpy as np
import torch

from . import _native as nv


def _shape_of(space):
    """get_shape_from_space (algorithms/utils/utils.py:16-28) for the spaces the envs expose; also accepts an int / tuple."""
    if isinstance(space, int):
        return (space,)
    if isinstance(space, (tuple, list)):
        return tuple(space)
    if type(space).__name__ == "Discrete":
        return (1,)
    return tuple(space.shape)


class DeviceRolloutBuffer:
    @staticmethod
    def _flatten(T, N, x):
        return x.reshape(T * N, *x.shape[2:])

    @staticmethod
    def _cast(x):
        """[T, N, A, ...] -> [N * A * T, ...] (buffer.py:34-36)."""
        return x.permute(1, 2, 0, *range(3, x.dim())).reshape(-1, *x.shape[3:])

    def __init__(self, args, num_agents, obs_space, act_space, device="cuda:0"):
        self.buffer_size = args.buffer_size
        self.n_rollout_threads = args.n_rollout_threads
        self.num_agents = num_agents
        self.gamma = args.gamma
        self.use_proper_time_limits = args.use_proper_time_limits
        self.use_gae = args.use_gae
        self.gae_lambda = args.gae_lambda
        self.recurrent_hidden_size = args.recurrent_hidden_size
        self.recurrent_hidden_layers = args.recurrent_hidden_layers
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceRolloutBuffer lives on a CUDA device (the numpy ReplayBuffer is the host counterpart)")
        T, N, A = self.buffer_size, self.n_rollout_threads, num_agents
        obs_shape, act_shape = _shape_of(obs_space), _shape_of(act_space)
        z = dict(dtype=torch.float32, device=self.device)
        self.obs = torch.zeros((T + 1, N, A, *obs_shape), **z)
        self.actions = torch.zeros((T, N, A, *act_shape), **z)
        self.rewards = torch.zeros((T, N, A, 1), **z)
        self.masks = torch.ones((T + 1, N, A, 1), **z)
        self.bad_masks = torch.ones((T + 1, N, A, 1), **z)
        self.action_log_probs = torch.zeros((T, N, A, 1), **z)
        self.value_preds = torch.zeros((T + 1, N, A, 1), **z)
        self.returns = torch.zeros((T + 1, N, A, 1), **z)
        self.rnn_states_actor = torch.zeros((T + 1, N, A, self.recurrent_hidden_layers, self.recurrent_hidden_size), **z)
        self.rnn_states_critic = torch.zeros_like(self.rnn_states_actor)
        self.reset_env = torch.zeros(N, dtype=torch.uint8, device=self.device)
        self.step = 0
        self._env = None

    # ---- reference surface (buffer.py:73-138) ---------------------------------------------------------------
    @property
    def advantages(self):
        adv = self.returns[:-1] - self.value_preds[:-1]
        return (adv - adv.mean()) / (adv.std(unbiased=False) + 1e-5)

    def insert(self, obs, actions, rewards, masks, action_log_probs, value_preds, rnn_states_actor, rnn_states_critic,
               bad_masks=None, **kwargs):
        """ReplayBuffer.insert (buffer.py:82-117) for device tensors (numpy arrays are uploaded)."""
        t = self.step
        put = self._put
        put(self.obs[t + 1], obs); put(self.actions[t], actions); put(self.rewards[t], rewards)
        put(self.masks[t + 1], masks); put(self.action_log_probs[t], action_log_probs); put(self.value_preds[t], value_preds)
        put(self.rnn_states_actor[t + 1], rnn_states_actor); put(self.rnn_states_critic[t + 1], rnn_states_critic)
        if bad_masks is not None:
            put(self.bad_masks[t + 1], bad_masks)
        self.step = (self.step + 1) % self.buffer_size
        self._retarget()

    def _put(self, dst, src):
        if src is None:
            return
        if not torch.is_tensor(src):
            src = torch.from_numpy(np.ascontiguousarray(src))
        if src.data_ptr() != dst.data_ptr():      # the env may already have written this slot in place
            dst.copy_(src.reshape(dst.shape), non_blocking=True)

    def after_update(self):
        """buffer.py:119-125."""
        self.obs[0].copy_(self.obs[-1]); self.masks[0].copy_(self.masks[-1]); self.bad_masks[0].copy_(self.bad_masks[-1])
        self.rnn_states_actor[0].copy_(self.rnn_states_actor[-1]); self.rnn_states_critic[0].copy_(self.rnn_states_critic[-1])

    def clear(self):
        """buffer.py:127-138."""
        self.step = 0
        for x in (self.obs, self.actions, self.rewards, self.action_log_probs, self.value_preds, self.returns,
                  self.rnn_states_actor, self.rnn_states_critic):
            x.zero_()
        self.masks.fill_(1.0); self.bad_masks.fill_(1.0)
        self._retarget()

    def compute_returns(self, next_value):
        """buffer.py:139-172 on the device: one backward-scan kernel over [T][N*A] columns."""
        nv_t = next_value if torch.is_tensor(next_value) else torch.from_numpy(np.ascontiguousarray(next_value))
        (self.value_preds if self.use_gae else self.returns)[-1].copy_(nv_t.reshape(self.returns[-1].shape))
        T, M = self.buffer_size, self.n_rollout_threads * self.num_agents
        st = nv.lib().np_rollout_returns(self.rewards.data_ptr(), self.value_preds.data_ptr(), self.masks.data_ptr(),
                                         self.bad_masks.data_ptr(), self.returns.data_ptr(), T, M, float(self.gamma),
                                         float(self.gae_lambda), int(bool(self.use_gae)), int(bool(self.use_proper_time_limits)),
                                         torch.cuda.current_stream(self.device).cuda_stream)
        nv.check(st, "np_rollout_returns")

    # ---- zero-copy coupling with a native env ---------------------------------------------------------------------
    def attach(self, env):
        """The env's next reset() writes obs[step] and every step() writes obs[step + 1] / rewards[step] of this buffer."""
        if env.n != self.n_rollout_threads * self.num_agents or env.num_observation != int(np.prod(self.obs.shape[3:])):
            raise ValueError("the env's population / observation width does not match this buffer")
        if env.n % 2:
            # rewards[t] starts t * n * 4 bytes into the buffer: with an odd population every odd t is only 4-byte aligned,
            # and the step kernel stores rewards as 8-byte aircraft pairs (np_env_rebind_outputs would refuse mid-rollout)
            raise ValueError("attach() needs an even aircraft population (the step kernel writes reward pairs in place)")
        self._env = env
        self._scratch_reward = torch.zeros(env.n, dtype=torch.float32, device=self.device)
        self._retarget(for_reset=True)

    def _bind(self, obs_slot, reward_slot):
        env = self._env
        env._obs = obs_slot.view(env.n, env.num_observation)
        env._reward = reward_slot.view(env.n)
        nv.check(nv.lib().np_env_rebind_outputs(env._handle, env._obs.data_ptr(), env._reward.data_ptr()), "np_env_rebind_outputs")

    def _retarget(self, for_reset=False):
        if self._env is None:
            return
        if for_reset:      # reset() produces o_0 (buffer.obs[0], F16sim_runner.py:114-118); its reward output is unused
            self._bind(self.obs[self.step], self._scratch_reward)
        else:
            self._bind(self.obs[self.step + 1], self.rewards[self.step])

    def step_env(self, env, actions, action_log_probs=None, value_preds=None, rnn_states_actor=None, rnn_states_critic=None):
        """env.step(actions) + F16SimRunner.insert (F16sim_runner.py:141-157) without leaving the device: obs / reward
        land in place, masks / bad_masks come from the env's flag rows, recurrent states of reset envs are zeroed."""
        if env is not self._env:
            raise RuntimeError("attach(env) first")
        t = self.step
        self._retarget()
        a = actions if torch.is_tensor(actions) else torch.from_numpy(np.ascontiguousarray(actions)).to(self.device)
        out = env.step(a.reshape(env.n, -1))
        st = nv.lib().np_rollout_masks(env._flags.data_ptr(), env.ld, self.n_rollout_threads, self.num_agents,
                                       self.masks[t + 1].data_ptr(), self.bad_masks[t + 1].data_ptr(), self.reset_env.data_ptr(),
                                       torch.cuda.current_stream(self.device).cuda_stream)
        nv.check(st, "np_rollout_masks")
        if rnn_states_actor is not None or rnn_states_critic is not None:
            keep = (self.reset_env == 0).to(torch.float32).view(-1, 1, 1, 1)
            if rnn_states_actor is not None:
                self.rnn_states_actor[t + 1].copy_(rnn_states_actor.reshape(self.rnn_states_actor[t + 1].shape) * keep)
            if rnn_states_critic is not None:
                self.rnn_states_critic[t + 1].copy_(rnn_states_critic.reshape(self.rnn_states_critic[t + 1].shape) * keep)
        self._put(self.actions[t], a)
        self._put(self.action_log_probs[t], action_log_probs)
        self._put(self.value_preds[t], value_preds)
        self.step = (self.step + 1) % self.buffer_size
        self._retarget()
        return out

    # ---- mini-batch generator (buffer.py:174-282) ------------------------------------------------------------------
    @staticmethod
    def recurrent_generator(buffer, num_mini_batch, data_chunk_length):
        """Same chunking / shuffling as the reference (torch.randperm on the CPU generator, so a seeded run draws the
        same permutation); yields device tensors."""
        buffers = [buffer] if isinstance(buffer, DeviceRolloutBuffer) else list(buffer)
        b0 = buffers[0]
        N, T = b0.n_rollout_threads, b0.buffer_size * len(buffers)
        assert N * T >= data_chunk_length
        cat = lambda f: torch.cat([DeviceRolloutBuffer._cast(f(b)) for b in buffers], 0)  # noqa: E731
        obs, actions = cat(lambda b: b.obs[:-1]), cat(lambda b: b.actions)
        masks, logp = cat(lambda b: b.masks[:-1]), cat(lambda b: b.action_log_probs)
        adv, returns = cat(lambda b: b.advantages), cat(lambda b: b.returns[:-1])
        values = cat(lambda b: b.value_preds[:-1])
        rnn_a, rnn_c = cat(lambda b: b.rnn_states_actor[:-1]), cat(lambda b: b.rnn_states_critic[:-1])
        data_chunks = N * T // data_chunk_length
        mini_batch_size = data_chunks // num_mini_batch
        rand = torch.randperm(data_chunks)
        L = data_chunk_length
        for i in range(num_mini_batch):
            idx = rand[i * mini_batch_size:(i + 1) * mini_batch_size].to(b0.device) * L
            rows = (idx.view(1, -1) + torch.arange(L, device=b0.device).view(-1, 1)).reshape(-1)      # (L, N) order
            take = lambda x: x[rows].reshape(L * mini_batch_size, *x.shape[1:])  # noqa: E731
            yield (take(obs), take(actions), take(masks), take(logp), take(adv), take(returns), take(values),
                   rnn_a[idx].reshape(mini_batch_size, *rnn_a.shape[1:]), rnn_c[idx].reshape(mini_batch_size, *rnn_c.shape[1:]))

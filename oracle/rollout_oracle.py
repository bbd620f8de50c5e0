"""CPU oracle of the rollout buffer arithmetic (SURVEY f-2).  TEST INFRASTRUCTURE ONLY -- never imported by the product.

Restates, in numpy float32 with the reference's operation order,
  * ReplayBuffer.compute_returns (algorithms/utils/buffer.py:139-172), all four (use_gae, use_proper_time_limits) variants;
  * ReplayBuffer.insert / after_update slot conventions (buffer.py:82-124): obs / masks / bad_masks / rnn states go to
    slot step + 1, actions / rewards / log-probs / values to slot step;
  * the mask derivation of F16SimRunner.insert (runner/F16sim_runner.py:141-157).
Pinned bit-exactly to the reference's own ReplayBuffer (imported from /root/reference by tests/golden/make_golden.py ->
tests/golden/rollout_returns.npz; tests/test_rollout_oracle_golden.py).
"""
import numpy as np


def compute_returns(rewards, value_preds, masks, bad_masks, next_value, gamma, gae_lambda, use_gae, use_proper_time_limits):
    """rewards [T, ...], value_preds / masks / bad_masks [T+1, ...] float32 -> returns [T+1, ...] (buffer.py:139-172).
    value_preds is NOT modified (the reference writes next_value into value_preds[-1]; the copy here carries it)."""
    rewards, masks, bad_masks = (np.asarray(x, dtype=np.float32) for x in (rewards, masks, bad_masks))
    value_preds = np.array(value_preds, dtype=np.float32, copy=True)
    returns = np.zeros_like(value_preds)
    T = rewards.shape[0]
    if use_gae:
        value_preds[-1] = next_value                                               # :147 / :161
        gae = 0
        for step in reversed(range(T)):
            td_delta = rewards[step] + gamma * value_preds[step + 1] * masks[step + 1] - value_preds[step]
            gae = td_delta + gamma * gae_lambda * masks[step + 1] * gae
            if use_proper_time_limits:
                gae = gae * bad_masks[step + 1]                                    # :153
            returns[step] = gae + value_preds[step]
    else:
        returns[-1] = next_value                                                   # :156 / :169
        for step in reversed(range(T)):
            if use_proper_time_limits:                                             # :158-159
                returns[step] = (returns[step + 1] * gamma * masks[step + 1] + rewards[step]) * bad_masks[step + 1] \
                    + (1 - bad_masks[step + 1]) * value_preds[step]
            else:
                returns[step] = returns[step + 1] * gamma * masks[step + 1] + rewards[step]   # :171
    return returns, value_preds


def advantages(returns, value_preds):
    """ReplayBuffer.advantages (buffer.py:73-79): normalised over the whole buffer."""
    adv = returns[:-1] - value_preds[:-1]
    return (adv - adv.mean()) / (adv.std() + 1e-5)


def runner_masks(dones, bad_dones, exceed_time_limits):
    """F16SimRunner.insert (F16sim_runner.py:141-157): [N, A, 1] bool flags -> masks, bad_masks [N, A, 1] f32, reset_env [N]."""
    n, a = dones.shape[0], dones.shape[1]
    dones_env = np.any(dones.squeeze(axis=-1), axis=-1)
    bad_env = np.any(bad_dones.squeeze(axis=-1), axis=-1)
    reset_env = np.any((dones + bad_dones + exceed_time_limits).squeeze(axis=-1), axis=-1)
    masks = np.ones((n, a, 1), dtype=np.float32)
    masks[dones_env] = 0.0
    bad_masks = np.ones((n, a, 1), dtype=np.float32)
    bad_masks[bad_env] = 0.0
    return masks, bad_masks, reset_env

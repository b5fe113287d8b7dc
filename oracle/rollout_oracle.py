"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

Restatement of the rsl_rl PPO rollout-storage path of the reference:
standalone/rsl_rl/ext/storage/rollout_storage.py:71-88 (add_transitions), :113-127
(GAE ``compute_returns`` + advantage normalisation), :152-191 (mini-batch gather) and
the time-out bootstrap of standalone/rsl_rl/ext/algorithms/ppo.py:85-97.
Pinned against the unmodified reference class in tests/test_oracle_vs_reference.py
and by tests/golden/gae_*.pt.
"""
from __future__ import annotations

import torch


def bootstrap_rewards(rewards, values, time_outs, gamma):
    """ppo.py:86-92: r += gamma * squeeze(V * time_outs.unsqueeze(1), 1)."""
    r = rewards.clone()
    r += gamma * torch.squeeze(values * time_outs.unsqueeze(1).to(values.dtype), 1)
    return r


def compute_returns(rewards, values, dones, last_values, gamma, lam):
    """rollout_storage.py:113-127.  rewards/values [T,N,1] float, dones [T,N,1] uint8, last_values [N,1].
    Returns (returns, normalised advantages), both [T,N,1]."""
    T = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    advantage = 0
    for step in reversed(range(T)):
        next_values = last_values if step == T - 1 else values[step + 1]
        next_is_not_terminal = 1.0 - dones[step].float()
        delta = rewards[step] + next_is_not_terminal * gamma * next_values - values[step]
        advantage = delta + next_is_not_terminal * gamma * lam * advantage
        returns[step] = advantage + values[step]
    advantages = returns - values
    advantages = (advantages - advantages.mean()) / (advantages.std() + 1e-8)
    return returns, advantages


def mini_batches(fields, indices, num_mini_batches, num_epochs):
    """rollout_storage.py:152-191 with the ``torch.randperm`` draw passed in as ``indices``.
    ``fields`` is a list of [T,N,...] tensors; yields lists of gathered batches."""
    flat = [f.flatten(0, 1) for f in fields]
    mb = indices.numel() // num_mini_batches
    for _ in range(num_epochs):
        for i in range(num_mini_batches):
            idx = indices[i * mb:(i + 1) * mb]
            yield [f[idx] for f in flat]

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


# ---- recurrent mini-batches -------------------------------------------------------------------------------------------
# rsl_rl.utils.split_and_pad_trajectories / unpad_trajectories: third party (rsl-rl-lib 2.x, unpinned, not under
# /root/reference; call sites rollout_storage.py:9,197-199).  Restated from the published algorithm; known-answer: the a/b
# example of its docstring (tests/test_recurrent_batches.py).  The reference's own reccurent_mini_batch_generator
# (rollout_storage.py:194-254) is then executed UNMODIFIED with this function injected in place of the absent import.

def split_and_pad_trajectories(tensor, dones):
    dones = dones.clone()
    dones[-1] = 1
    flat_dones = dones.transpose(1, 0).reshape(-1, 1)
    done_indices = torch.cat((flat_dones.new_tensor([-1], dtype=torch.int64), flat_dones.nonzero()[:, 0]))
    trajectory_lengths = done_indices[1:] - done_indices[:-1]
    trajectories = torch.split(tensor.transpose(1, 0).flatten(0, 1), trajectory_lengths.tolist())
    trajectories = trajectories + (torch.zeros(tensor.shape[0], *tensor.shape[2:], device=tensor.device),)       # at least one full-length row
    padded = torch.nn.utils.rnn.pad_sequence(trajectories)[:, :-1]
    masks = trajectory_lengths > torch.arange(0, tensor.shape[0], device=tensor.device).unsqueeze(1)
    return padded, masks


def unpad_trajectories(trajectories, masks):
    return trajectories.transpose(1, 0)[masks.transpose(1, 0)].view(-1, trajectories.shape[0], trajectories.shape[-1]).transpose(1, 0)


def recurrent_mini_batches(obs, critic_obs, row_fields, dones, saved_hidden_a, saved_hidden_c, num_mini_batches, num_epochs):
    """rollout_storage.py:194-254.  row_fields: list of [T,N,...] tensors sliced per env range; saved_hidden_*: lists of [T,L,N,H]."""
    padded_obs, masks = split_and_pad_trajectories(obs, dones)
    padded_critic = split_and_pad_trajectories(critic_obs, dones)[0] if critic_obs is not None else padded_obs
    N = obs.shape[1]
    mb = N // num_mini_batches
    for _ in range(num_epochs):
        first = 0
        for i in range(num_mini_batches):
            start, stop = i * mb, (i + 1) * mb
            d = dones.squeeze(-1)
            last_was_done = torch.zeros_like(d, dtype=torch.bool)
            last_was_done[1:] = d[:-1]
            last_was_done[0] = True
            last = first + int(torch.sum(last_was_done[:, start:stop]))
            lwd = last_was_done.permute(1, 0)
            hid_a = [h.permute(2, 0, 1, 3)[lwd][first:last].transpose(1, 0).contiguous() for h in saved_hidden_a]
            hid_c = [h.permute(2, 0, 1, 3)[lwd][first:last].transpose(1, 0).contiguous() for h in saved_hidden_c]
            yield (padded_obs[:, first:last], padded_critic[:, first:last], [f[:, start:stop] for f in row_fields], hid_a, hid_c, masks[:, first:last])
            first = last

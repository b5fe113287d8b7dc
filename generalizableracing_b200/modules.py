"""State-only policy modules of the racing task.

The reference uses ``rsl_rl.modules.ActorCritic`` (third party, not vendored) with
``actor_hidden_dims=[128,128], critic_hidden_dims=[128,128], activation="lrelu", init_noise_std=1.0``
(QD/agents/rsl_rl_ppo_cfg.py:22-27) for PPO and ``BaseModel`` = ActorCritic whose ``act`` uses ``rsample``
(standalone/diff_rl/algorithms/model.py:63-99) for BPTT.  This file restates that public surface in plain torch;
the MLP stays in torch / cuBLAS (north star: tensor cores only if profiling shows the MLP is the bottleneck).
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.distributions import Normal


def get_activation(name: str) -> nn.Module:
    return {"elu": nn.ELU, "selu": nn.SELU, "relu": nn.ReLU, "lrelu": nn.LeakyReLU, "tanh": nn.Tanh, "sigmoid": nn.Sigmoid}[name]()


def _mlp(n_in, hidden, n_out, act):
    layers, d = [], n_in
    for h in hidden:
        layers += [nn.Linear(d, h), get_activation(act)]
        d = h
    layers.append(nn.Linear(d, n_out))
    return nn.Sequential(*layers)


class ActorCritic(nn.Module):
    is_recurrent = False

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=(128, 128), critic_hidden_dims=(128, 128),
                 activation="lrelu", init_noise_std=1.0, noise_std_type: str = "scalar", **kwargs):
        super().__init__()
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str(list(kwargs)))
        self.actor = _mlp(num_actor_obs, list(actor_hidden_dims), num_actions, activation)
        self.critic = _mlp(num_critic_obs, list(critic_hidden_dims), 1, activation)
        self.noise_std_type = noise_std_type
        if noise_std_type == "scalar":
            self.std = nn.Parameter(init_noise_std * torch.ones(num_actions))
        elif noise_std_type == "log":
            self.log_std = nn.Parameter(torch.log(init_noise_std * torch.ones(num_actions)))
        else:
            raise ValueError(f"Invalid noise_std_type: {noise_std_type}")
        self.distribution = None
        Normal.set_default_validate_args(False)

    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    @property
    def action_mean(self):
        return self.distribution.mean

    @property
    def action_std(self):
        return self.distribution.stddev

    @property
    def entropy(self):
        return self.distribution.entropy().sum(dim=-1)

    def update_distribution(self, observations):
        mean = self.actor(observations)
        std = self.std.expand_as(mean) if self.noise_std_type == "scalar" else torch.exp(self.log_std).expand_as(mean)
        self.distribution = Normal(mean, std)

    def act(self, observations, **kwargs):
        self.update_distribution(observations)
        return self.distribution.sample()

    def get_actions_log_prob(self, actions):
        return self.distribution.log_prob(actions).sum(dim=-1)

    def act_inference(self, observations):
        return self.actor(observations)

    def evaluate(self, critic_observations, **kwargs):
        return self.critic(critic_observations)


class BaseModel(ActorCritic):
    """standalone/diff_rl/algorithms/model.py:63-99: reparameterised sampling keeps the action differentiable."""

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=(256, 256, 256), critic_hidden_dims=(256, 256, 256),
                 activation="elu", init_noise_std=1.0, noise_std_type: str = "scalar", **kwargs):
        super().__init__(num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims, critic_hidden_dims, activation, init_noise_std,
                         noise_std_type, **kwargs)

    def act(self, observations, **kwargs):
        self.update_distribution(observations)
        return self.distribution.rsample()

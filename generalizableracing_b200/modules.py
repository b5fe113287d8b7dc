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


class Memory(nn.Module):
    """rsl_rl.modules.actor_critic_recurrent.Memory (third party rsl-rl-lib 2.x, restated): the recurrent core in front of
    the MLPs.  Inference mode carries its own hidden state; batch mode (PPO update) runs whole padded trajectories from the
    saved start states and un-pads the result (gr_traj_unpad)."""

    def __init__(self, input_size, type="lstm", num_layers=1, hidden_size=256):
        super().__init__()
        rnn_cls = nn.GRU if type.lower() == "gru" else nn.LSTM
        self.rnn = rnn_cls(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers)
        self.hidden_states = None

    def forward(self, input, masks=None, hidden_states=None):
        if masks is not None:                                 # batch mode
            if hidden_states is None:
                raise ValueError("Hidden states not passed to memory module during policy update")
            from .trajectories import unpad_trajectories
            out, _ = self.rnn(input, hidden_states)
            return unpad_trajectories(out, masks)
        out, self.hidden_states = self.rnn(input.unsqueeze(0), self.hidden_states)
        return out

    def reset(self, dones=None, hidden_states=None):
        if dones is None:
            self.hidden_states = hidden_states
        elif self.hidden_states is not None:
            hs = self.hidden_states if isinstance(self.hidden_states, tuple) else (self.hidden_states,)
            keep = (dones == 0).to(hs[0].dtype).view(1, -1, 1)
            new = tuple(h * keep for h in hs)                  # zero the state of the envs that finished (no host sync, no in-place on saved states)
            self.hidden_states = new if isinstance(self.hidden_states, tuple) else new[0]

    def detach_hidden_states(self, dones=None):
        if self.hidden_states is not None:
            hs = self.hidden_states
            self.hidden_states = tuple(h.detach() for h in hs) if isinstance(hs, tuple) else hs.detach()


class ActorCriticRecurrent(ActorCritic):
    """rsl_rl.modules.ActorCriticRecurrent (restated): Memory -> MLP for actor and critic."""
    is_recurrent = True

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=(256, 256, 256), critic_hidden_dims=(256, 256, 256),
                 activation="elu", rnn_type="lstm", rnn_hidden_dim=256, rnn_num_layers=1, init_noise_std=1.0, **kwargs):
        if "rnn_hidden_size" in kwargs:
            rnn_hidden_dim = kwargs.pop("rnn_hidden_size")
        super().__init__(rnn_hidden_dim, rnn_hidden_dim, num_actions, actor_hidden_dims, critic_hidden_dims, activation, init_noise_std, **kwargs)
        self.memory_a = Memory(num_actor_obs, type=rnn_type, num_layers=rnn_num_layers, hidden_size=rnn_hidden_dim)
        self.memory_c = Memory(num_critic_obs, type=rnn_type, num_layers=rnn_num_layers, hidden_size=rnn_hidden_dim)

    def reset(self, dones=None):
        self.memory_a.reset(dones)
        self.memory_c.reset(dones)

    def act(self, observations, masks=None, hidden_states=None):
        return super().act(self.memory_a(observations, masks, hidden_states).squeeze(0))

    def act_inference(self, observations):
        return super().act_inference(self.memory_a(observations).squeeze(0))

    def evaluate(self, critic_observations, masks=None, hidden_states=None):
        return super().evaluate(self.memory_c(critic_observations, masks, hidden_states).squeeze(0))

    def get_hidden_states(self):
        return self.memory_a.hidden_states, self.memory_c.hidden_states


class EmpiricalNormalization(nn.Module):
    """rsl_rl.modules.EmpiricalNormalization (third-party, rsl-rl-lib 2.x; restated from its published algorithm): running
    mean / variance of the observations seen in training mode, batched Welford merge per call, ``(x - mean) / (std + eps)``.
    Used by OnPolicyRunner when ``empirical_normalization`` is set (on_policy_runner.py:67-73,151-153).  The ``until`` bound is
    checked against a host-side mirror of ``count`` so that the rollout loop stays free of device->host reads."""

    def __init__(self, shape, eps: float = 1e-2, until=None):
        super().__init__()
        self.eps = eps
        self.until = until
        self.register_buffer("_mean", torch.zeros(shape).unsqueeze(0))
        self.register_buffer("_var", torch.ones(shape).unsqueeze(0))
        self.register_buffer("_std", torch.ones(shape).unsqueeze(0))
        self.register_buffer("count", torch.tensor(0, dtype=torch.long))
        self._count_host = 0

    @property
    def mean(self):
        return self._mean.squeeze(0).clone()

    @property
    def std(self):
        return self._std.squeeze(0).clone()

    def forward(self, x):
        if self.training:
            self.update(x)
        return (x - self._mean) / (self._std + self.eps)

    @torch.no_grad()
    def update(self, x):
        if self.until is not None and self._count_host >= self.until:
            return
        count_x = x.shape[0]
        self._count_host += count_x
        self.count += count_x
        rate = count_x / self._count_host
        var_x = torch.var(x, dim=0, unbiased=False, keepdim=True)
        mean_x = torch.mean(x, dim=0, keepdim=True)
        delta_mean = mean_x - self._mean
        self._mean += rate * delta_mean
        self._var += rate * (var_x - self._var + delta_mean * (mean_x - self._mean))
        self._std = torch.sqrt(self._var)

    def inverse(self, y):
        return y * (self._std + self.eps) + self._mean

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self._count_host = int(self.count)

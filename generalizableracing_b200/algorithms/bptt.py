"""BPTT of the reference (standalone/diff_rl/algorithms/bptt.py:18-63, algo.py:20-60) on the analytic reverse sweep.

``update()`` of the reference stacks the per-step losses and calls ``total_loss_mean.backward()`` through the autograd
tape of the dynamics.  Here the env keeps a tape in HBM and one launch of gr_step_bwd turns the uniform loss weights
1/(T*N) into dL/da_t for the whole window; torch only back-propagates that cotangent through the policy MLP.
(The generic path -- calling ``extras["losses"]....backward()`` yourself -- also works: see generalizableracing_b200/bptt.py.)
"""
from __future__ import annotations

import torch
from torch.optim import SGD, Adam, Adamax, AdamW, RMSprop
from torch.optim.lr_scheduler import CosineAnnealingLR, ExponentialLR, StepLR

from .. import dist_utils as D

# algo.py:44-46 selects both by name with eval(); the same names, looked up instead of evaluated
_OPTIMIZERS = {"SGD": SGD, "Adam": Adam, "Adamax": Adamax, "AdamW": AdamW, "RMSprop": RMSprop}
_SCHEDULES = {"CosineAnnealingLR": CosineAnnealingLR, "ExponentialLR": ExponentialLR, "StepLR": StepLR}


class BPTT:
    def __init__(self, actor_critic, max_iterations=1000, learning_rate=1e-3, schedule="CosineAnnealingLR", device="cuda:0",
                 optimizer="Adam", env=None, **kwargs):
        if kwargs:
            print(f"{self.__class__}.__init__ got unexpected arguments, which will be ignored: " + str(list(kwargs)))
        self.device = device
        self.learning_rate = learning_rate
        self.actor_critic = actor_critic
        self.actor_critic.to(self.device)
        D.broadcast_module(self.actor_critic)
        # algo.py:44 builds `<optimizer>(params, lr=lr)`; on CUDA the Adam family / SGD run as torch's single-kernel ("fused")
        # implementation of the same update: one launch instead of ~8 element-wise ones in an iteration that lasts < 1 ms
        opt_kw = {"fused": True} if (optimizer in ("Adam", "AdamW", "SGD") and torch.device(self.device).type == "cuda") else {}
        self.optimizer = _OPTIMIZERS[optimizer](self.actor_critic.parameters(), lr=self.learning_rate, **opt_kw)
        self.schedule = _SCHEDULES[schedule](self.optimizer, max_iterations, self.learning_rate * 0.01)
        self.env = env                      # RacingVecEnv: enables the one-launch window sweep
        self.losses, self.losses_detached, self.dones, self.rewards, self.actions = [], [], [], [], []

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    def act(self, obs, critic_obs=None):
        a = self.actor_critic.act(obs)
        self.actions.append(a)
        return a

    def process_env_step(self, losses, losses_detached, dones, rewards=None, infos=None):
        self.losses.append(losses)
        self.losses_detached.append(losses_detached)
        self.dones.append(dones)
        self.rewards.append(rewards)
        self.actor_critic.reset(dones)

    def update_fused(self, collector):
        """update() after a fused window (collect.FusedBpttCollector): one reverse sweep + ONE batched actor backward."""
        win = self.env._bptt
        T = win.t
        total_loss_mean = win.loss[:T].mean()
        self.optimizer.zero_grad(set_to_none=True)
        grad_actions = win.backward_window(grad_scale=1.0 / (T * self.env.num_envs))
        if collector.backward_kernel:
            collector.policy_backward_kernel(grad_actions)
        else:
            collector.policy_backward(grad_actions, tf32=collector.backward_tf32)
        D.allreduce_mean_grads(self.actor_critic.parameters())
        self.optimizer.step()
        self.schedule.step()
        return 0.0, total_loss_mean

    def update(self):
        losses = torch.stack([l.detach() for l in self.losses])
        losses_detached = torch.stack(self.losses_detached)
        total_loss_mean = (losses + losses_detached).mean()
        self.optimizer.zero_grad()
        win = getattr(self.env, "_bptt", None) if self.env is not None else None
        if win is not None and win.t == len(self.actions):
            _, w = D.world()
            grad_actions = win.backward_window(grad_scale=1.0 / (losses.numel()))          # dL/da_t, L = mean over [T, N_local]
            acts = [a for a in self.actions if a.requires_grad]
            torch.autograd.backward(acts, [grad_actions[t] for t, a in enumerate(self.actions) if a.requires_grad])
        else:
            (torch.stack(self.losses) + losses_detached).mean().backward()
        D.allreduce_mean_grads(self.actor_critic.parameters())
        self.optimizer.step()
        self.schedule.step()
        self.losses, self.losses_detached, self.dones, self.rewards, self.actions = [], [], [], [], []
        return 0.0, total_loss_mean

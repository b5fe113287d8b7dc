"""PPO of the reference (standalone/rsl_rl/ext/algorithms/ppo.py:14-190) on the CUDA rollout storage.

Same public methods and hyper-parameters.  Differences, all on the rollout side: ``process_env_step`` hands the
time-out mask to the storage so the bootstrap ``r += gamma*V*time_out`` (ppo.py:89-92) is fused into the single
add_transitions launch; ``compute_returns`` is two launches; mini-batches are gathered by one launch each.
Multi-GPU (env-sharded): advantages are normalised with the moments of the GLOBAL batch and the policy gradient is
all-reduced (mean) before the optimiser step, so R ranks x N envs train like one process with R*N envs.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.optim as optim

from .. import dist_utils as D
from ..storage import RolloutStorage


class PPO:
    def __init__(self, policy, env=None, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95,
                 value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True,
                 schedule="fixed", desired_kl=0.01, device="cuda:0", graphed_update=False, **kwargs):
        self.env = env
        self.device = device
        self.desired_kl = desired_kl
        self.schedule = schedule
        self.learning_rate = learning_rate
        self.policy = policy
        self.policy.to(self.device)
        D.broadcast_module(self.policy)
        self.storage = None
        self.optimizer = optim.Adam(self.policy.parameters(), lr=learning_rate)
        self.transition = RolloutStorage.Transition()
        self.clip_param = clip_param
        self.num_learning_epochs = num_learning_epochs
        self.num_mini_batches = num_mini_batches
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.gamma = gamma
        self.lam = lam
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        # opt-in: one mini-batch step (gather -> forward -> losses -> backward -> clip -> Adam, adaptive LR on the device) captured
        # once in a CUDA graph and replayed num_learning_epochs * num_mini_batches times per iteration (single-GPU runs)
        self.graphed_update = bool(graphed_update)
        self._graph = None

    def init_storage(self, training_type, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.storage = RolloutStorage(training_type, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape, self.device)

    def test_mode(self):
        self.policy.eval()

    def train_mode(self):
        self.policy.train()

    def act(self, obs, critic_obs):
        # ppo.py:71-83
        self.transition.actions = self.policy.act(obs).detach()
        self.transition.values = self.policy.evaluate(critic_obs).detach()
        self.transition.actions_log_prob = self.policy.get_actions_log_prob(self.transition.actions).detach()
        self.transition.action_mean = self.policy.action_mean.detach()
        self.transition.action_sigma = self.policy.action_std.detach()
        self.transition.observations = obs
        self.transition.privileged_observations = critic_obs
        return self.transition.actions

    def process_env_step(self, rewards, dones, infos):
        # ppo.py:85-97; the reward bootstrap on time-outs happens inside storage.add_transitions
        self.transition.rewards = rewards
        self.transition.dones = dones
        if "time_outs" in infos:
            self.transition.time_outs = infos["time_outs"]
            self.transition.gamma = self.gamma
        self.storage.add_transitions(self.transition)
        self.transition.clear()
        self.policy.reset(dones)

    def compute_returns(self, last_critic_obs, last_values=None):
        # ppo.py:99-101 (last_values: already evaluated by the fused collection kernel)
        if last_values is None:
            last_values = self.policy.evaluate(last_critic_obs).detach()
        _, w = D.world()
        if w == 1:
            self.storage.compute_returns(last_values, self.gamma, self.lam)
        else:
            self.storage.compute_returns(last_values, self.gamma, self.lam, normalize=False)
            self.storage.normalize_advantages(D.merge_moments(self.storage.moments))

    # ------------------------------------------------------------------ CUDA-graph update
    def _mini_batch_losses(self, obs_batch, critic_obs_batch, actions_batch, target_values_batch, advantages_batch, returns_batch,
                           old_actions_log_prob_batch, sample: bool):
        """ppo.py:118-171 for one mini-batch -> (loss, surrogate, value_loss, mu, sigma)"""
        if sample:
            self.policy.act(obs_batch)
        else:
            self.policy.update_distribution(obs_batch)        # same distribution; the reference's (discarded) sample draws nothing we need
        actions_log_prob_batch = self.policy.get_actions_log_prob(actions_batch)
        value_batch = self.policy.evaluate(critic_obs_batch)
        mu_batch, sigma_batch, entropy_batch = self.policy.action_mean, self.policy.action_std, self.policy.entropy
        ratio = torch.exp(actions_log_prob_batch - torch.squeeze(old_actions_log_prob_batch))
        surrogate = -torch.squeeze(advantages_batch) * ratio
        surrogate_clipped = -torch.squeeze(advantages_batch) * torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param)
        surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
        if self.use_clipped_value_loss:
            value_clipped = target_values_batch + (value_batch - target_values_batch).clamp(-self.clip_param, self.clip_param)
            value_loss = torch.max((value_batch - returns_batch).pow(2), (value_clipped - returns_batch).pow(2)).mean()
        else:
            value_loss = (returns_batch - value_batch).pow(2).mean()
        loss = surrogate_loss + self.value_loss_coef * value_loss - self.entropy_coef * entropy_batch.mean()
        return loss, surrogate_loss, value_loss, mu_batch, sigma_batch

    def _build_graph(self):
        """Static mini-batch buffers + one captured optimisation step.  Called after an eager update() (optimizer state exists)."""
        import ctypes as C
        from .. import _lib as B
        sto, dev = self.storage, self.device
        mb = sto.num_envs * sto.num_transitions_per_env // self.num_mini_batches
        od, ad, cd = sto.obs_shape[0], sto.actions_shape[0], sto.privileged_obs_shape[0]
        g = {"idx": torch.zeros(mb, dtype=torch.int64, device=dev)}
        for name, w in (("obs", od), ("critic_obs", cd), ("actions", ad), ("values", 1), ("advantages", 1), ("returns", 1), ("log_prob", 1), ("mu", ad), ("sigma", ad)):
            g[name] = torch.empty(mb, w, device=dev)
        mbs = B.GrMiniBatch()
        for name in ("obs", "critic_obs", "actions", "values", "advantages", "returns", "log_prob", "mu", "sigma"):
            setattr(mbs, name, g[name].data_ptr())
        g["sums"] = torch.zeros(2, device=dev)                               # running (value loss, surrogate loss)
        # Adam with a device-side learning rate: same update rule, graph-capturable
        state = self.optimizer.state_dict()
        lr = torch.tensor(float(self.learning_rate), device=dev)
        self.optimizer = optim.Adam(self.policy.parameters(), lr=lr, capturable=True)
        for st in state["state"].values():
            st["step"] = st["step"].to(dev) if torch.is_tensor(st["step"]) else torch.tensor(float(st["step"]), device=dev)
        state["param_groups"][0]["lr"] = lr
        state["param_groups"][0]["capturable"] = True
        self.optimizer.load_state_dict(state)
        self.optimizer.param_groups[0]["lr"] = lr
        g["lr"] = lr
        desc = sto._desc()
        lib = sto._lib
        adaptive = self.desired_kl is not None and self.schedule == "adaptive"

        def step():
            B.check(lib.gr_storage_gather(C.byref(desc), g["idx"].data_ptr(), mb, C.byref(mbs), torch.cuda.current_stream(dev).cuda_stream), "gr_storage_gather")
            loss, surrogate_loss, value_loss, mu_batch, sigma_batch = self._mini_batch_losses(
                g["obs"], g["critic_obs"], g["actions"], g["values"], g["advantages"], g["returns"], g["log_prob"], sample=False)
            if adaptive:          # ppo.py:124-141 without the host round trip
                with torch.no_grad():
                    kl = torch.sum(torch.log(sigma_batch / g["sigma"] + 1.0e-5) + (torch.square(g["sigma"]) + torch.square(g["mu"] - mu_batch))
                                   / (2.0 * torch.square(sigma_batch)) - 0.5, axis=-1)
                    kl_mean = torch.mean(kl)
                    down = (lr / 1.5).clamp(min=1e-5)
                    up = (lr * 1.5).clamp(max=1e-2)
                    lr.copy_(torch.where(kl_mean > self.desired_kl * 2.0, down, torch.where((kl_mean < self.desired_kl / 2.0) & (kl_mean > 0.0), up, lr)))
            self.optimizer.zero_grad(set_to_none=True)
            loss.backward()
            nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
            self.optimizer.step()
            g["sums"] += torch.stack([value_loss.detach(), surrogate_loss.detach()])

        g["keep"] = (mbs, desc)
        # No autograd graph of an earlier (default-stream) step may survive into the capture: its AccumulateGrad nodes are bound
        # to the stream they were created on and would make the engine synchronise across streams while capturing.
        self.policy.distribution = None
        self.optimizer.zero_grad(set_to_none=True)
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):          # warm-up on the capture stream (allocations, lazy init) before the capture
            for _ in range(3):
                step()
                self.policy.distribution = None
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph, stream=side):
            step()
        g["graph"] = graph
        return g

    def _update_graphed(self):
        sto = self.storage
        batch = sto.num_envs * sto.num_transitions_per_env
        mb = batch // self.num_mini_batches
        if self._graph is None:
            # the warm-up steps before the capture advance the parameters and the Adam moments: snapshot, build, restore in place
            import copy
            snap = [p.detach().clone() for p in self.policy.parameters()]
            opt_snap = copy.deepcopy(self.optimizer.state_dict()["state"])
            self._graph = self._build_graph()
            with torch.no_grad():
                for p, q in zip(self.policy.parameters(), snap):
                    p.copy_(q)
                live = self.optimizer.state_dict()["state"]
                for k, st in opt_snap.items():
                    for name, v in st.items():
                        if torch.is_tensor(live[k][name]):
                            live[k][name].copy_(v if torch.is_tensor(v) else torch.tensor(float(v)))
                self._graph["lr"].fill_(float(self.learning_rate))
        g = self._graph
        g["sums"].zero_()
        indices = torch.randperm(self.num_mini_batches * mb, requires_grad=False, device=self.device)
        for _ in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                g["idx"].copy_(indices[i * mb:(i + 1) * mb])
                g["graph"].replay()
        num_updates = self.num_learning_epochs * self.num_mini_batches
        out = torch.cat([g["sums"] / num_updates, g["lr"].reshape(1)]).tolist()          # the iteration's only host read
        self.learning_rate = out[2]
        self.storage.clear()
        return {"value_function": out[0], "surrogate": out[1]}

    def update(self):
        # ppo.py:103-190
        if self.graphed_update and D.world()[1] == 1 and getattr(self, "_eager_updates", 0) >= 1:
            return self._update_graphed()
        self._eager_updates = getattr(self, "_eager_updates", 0) + 1
        mean_value_loss = torch.zeros((), device=self.device)
        mean_surrogate_loss = torch.zeros((), device=self.device)
        generator = self.storage.mini_batch_generator(self.num_mini_batches, self.num_learning_epochs)
        for (obs_batch, critic_obs_batch, actions_batch, target_values_batch, advantages_batch, returns_batch,
             old_actions_log_prob_batch, old_mu_batch, old_sigma_batch, _hid, _masks) in generator:
            self.policy.act(obs_batch)
            actions_log_prob_batch = self.policy.get_actions_log_prob(actions_batch)
            value_batch = self.policy.evaluate(critic_obs_batch)
            mu_batch = self.policy.action_mean
            sigma_batch = self.policy.action_std
            entropy_batch = self.policy.entropy

            if self.desired_kl is not None and self.schedule == "adaptive":
                with torch.inference_mode():
                    kl = torch.sum(torch.log(sigma_batch / old_sigma_batch + 1.0e-5)
                                   + (torch.square(old_sigma_batch) + torch.square(old_mu_batch - mu_batch)) / (2.0 * torch.square(sigma_batch)) - 0.5, axis=-1)
                    kl_mean = torch.mean(kl)
                    _, w = D.world()
                    if w > 1:
                        torch.distributed.all_reduce(kl_mean)
                        kl_mean /= w
                    if kl_mean > self.desired_kl * 2.0:
                        self.learning_rate = max(1e-5, self.learning_rate / 1.5)
                    elif kl_mean < self.desired_kl / 2.0 and kl_mean > 0.0:
                        self.learning_rate = min(1e-2, self.learning_rate * 1.5)
                    for param_group in self.optimizer.param_groups:
                        param_group["lr"] = self.learning_rate

            ratio = torch.exp(actions_log_prob_batch - torch.squeeze(old_actions_log_prob_batch))
            surrogate = -torch.squeeze(advantages_batch) * ratio
            surrogate_clipped = -torch.squeeze(advantages_batch) * torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param)
            surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
            if self.use_clipped_value_loss:
                value_clipped = target_values_batch + (value_batch - target_values_batch).clamp(-self.clip_param, self.clip_param)
                value_losses = (value_batch - returns_batch).pow(2)
                value_losses_clipped = (value_clipped - returns_batch).pow(2)
                value_loss = torch.max(value_losses, value_losses_clipped).mean()
            else:
                value_loss = (returns_batch - value_batch).pow(2).mean()
            loss = surrogate_loss + self.value_loss_coef * value_loss - self.entropy_coef * entropy_batch.mean()

            self.optimizer.zero_grad()
            loss.backward()
            D.allreduce_mean_grads(self.policy.parameters())          # env-sharded data parallelism (NCCL over NVLink)
            nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
            self.optimizer.step()
            mean_value_loss += value_loss.detach()
            mean_surrogate_loss += surrogate_loss.detach()

        num_updates = self.num_learning_epochs * self.num_mini_batches
        self.storage.clear()
        return {"value_function": (mean_value_loss / num_updates).item(), "surrogate": (mean_surrogate_loss / num_updates).item()}

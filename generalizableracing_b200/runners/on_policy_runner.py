"""OnPolicyRunner with the reference's surface (standalone/rsl_rl/ext/runners/on_policy_runner.py:20-374):
``OnPolicyRunner(env, train_cfg, log_dir, device)``, ``learn(num_learning_iterations, init_at_random_ep_len)``,
``save/load/get_inference_policy``, ``train_mode/eval_mode``.  The rollout loop is the reference's (:135-183); the
bookkeeping that forced a host sync every step (``nonzero()`` / ``.cpu()`` at :169-173) stays on the device.
Logging is a JSON-lines file instead of TensorBoard (not installed here); the scalars are the reference's.
"""
from __future__ import annotations

import json
import os
from collections import deque

import torch

from .. import dist_utils as D
from ..algorithms import PPO
from ..modules import ActorCritic


class OnPolicyRunner:
    def __init__(self, env, train_cfg: dict, log_dir=None, device="cuda:0"):
        self.cfg = dict(train_cfg)
        self.alg_cfg = dict(train_cfg["algorithm"])
        self.policy_cfg = dict(train_cfg["policy"])
        self.device = device
        self.env = env
        if self.alg_cfg.get("class_name", "PPO") != "PPO":
            raise ValueError("only PPO is built (PPOLCP / PPOL2C2 / Distillation are vision / BC variants, out of scope)")
        self.training_type = "rl"
        obs, extras = self.env.get_observations()
        num_obs = obs.shape[1]
        self.privileged_obs_type = "critic" if "critic" in extras["observations"] else None
        num_privileged_obs = extras["observations"]["critic"].shape[1] if self.privileged_obs_type else num_obs
        policy_class = self.policy_cfg.pop("class_name", "ActorCritic")
        if policy_class == "ActorCriticRecurrent":
            from ..modules import ActorCriticRecurrent
            policy = ActorCriticRecurrent(num_obs, num_privileged_obs, self.env.num_actions, **self.policy_cfg).to(self.device)
        elif policy_class == "ActorCritic":
            policy = ActorCritic(num_obs, num_privileged_obs, self.env.num_actions, **self.policy_cfg).to(self.device)
        else:
            raise ValueError(f"policy class {policy_class!r} is not built (vision / distillation policies are out of scope)")
        self.alg_cfg.pop("class_name", None)
        self.alg = PPO(policy, env=self.env, device=self.device, **self.alg_cfg)
        self.num_steps_per_env = self.cfg["num_steps_per_env"]
        self.save_interval = self.cfg.get("save_interval", 50)
        # on_policy_runner.py:67-73 (False in the racing cfg, QD/agents/rsl_rl_ppo_cfg.py:21)
        self.empirical_normalization = bool(self.cfg.get("empirical_normalization", False))
        if self.empirical_normalization:
            if self.cfg.get("fused_collection", False):
                raise ValueError("fused_collection evaluates the policy on the raw observations inside the kernel: not with empirical_normalization")
            from ..modules import EmpiricalNormalization
            self.obs_normalizer = EmpiricalNormalization(shape=[num_obs], until=1.0e8).to(self.device)
            self.privileged_obs_normalizer = EmpiricalNormalization(shape=[num_privileged_obs], until=1.0e8).to(self.device)
        else:
            self.obs_normalizer = torch.nn.Identity()
            self.privileged_obs_normalizer = torch.nn.Identity()
        self.alg.init_storage(self.training_type, self.env.num_envs, self.num_steps_per_env, [num_obs], [num_privileged_obs], [self.env.num_actions])
        # opt-in: the whole rollout loop as one launch (collect.py / csrc/ppo_collect.cu); policy inference on the tensor cores
        self.collector = None
        if self.cfg.get("fused_collection", False):
            from ..collect import FusedCollector
            self.collector = FusedCollector(self.env, policy, self.alg.storage, gamma=self.alg.gamma, groups_per_cta=int(self.cfg.get("fused_groups_per_cta", 0)))
        self.log_dir = log_dir
        self.tot_timesteps = 0
        self.tot_time = 0
        self.current_learning_iteration = 0
        self.history = []

    def train_mode(self):
        self.alg.policy.train()
        self.obs_normalizer.train()
        self.privileged_obs_normalizer.train()

    def eval_mode(self):
        self.alg.policy.eval()
        self.obs_normalizer.eval()
        self.privileged_obs_normalizer.eval()

    def learn(self, num_learning_iterations: int, init_at_random_ep_len: bool = False):
        rank, world = D.world()
        if init_at_random_ep_len:
            self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf, high=int(self.env.max_episode_length))
        obs, extras = self.env.get_observations()
        privileged_obs = extras["observations"].get(self.privileged_obs_type, obs)
        self.train_mode()
        N = self.env.num_envs
        cur_reward_sum = torch.zeros(N, dtype=torch.float, device=self.device)
        cur_episode_length = torch.zeros(N, dtype=torch.float, device=self.device)
        ep_rew_sum = torch.zeros((), device=self.device)
        ep_len_sum = torch.zeros((), device=self.device)
        ep_count = torch.zeros((), device=self.device)
        start_iter = self.current_learning_iteration
        tot_iter = start_iter + num_learning_iterations
        for it in range(start_iter, tot_iter):
            ev_start, ev_mid, ev_end = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            ev_start.record()
            ep_infos = []
            with torch.inference_mode():
                if self.collector is not None:
                    self.collector.pack()
                    obs, privileged_obs, last_values = self.collector.collect()
                    st = self.collector.episode_stats()
                    ep_rew_sum += st[0]; ep_len_sum += st[1]; ep_count += st[2]
                    infos = self.env.extras
                for _ in range(self.num_steps_per_env if self.collector is None else 0):
                    actions = self.alg.act(obs, privileged_obs)
                    obs, rewards, dones, infos = self.env.step(actions)
                    obs = self.obs_normalizer(obs)                       # on_policy_runner.py:151-155
                    privileged_obs = self.privileged_obs_normalizer(infos["observations"][self.privileged_obs_type]) if self.privileged_obs_type else obs
                    self.alg.process_env_step(rewards, dones, infos)
                    # episode book keeping (on_policy_runner.py:167-173) without host synchronisation
                    cur_reward_sum += rewards
                    cur_episode_length += 1
                    d = dones > 0
                    ep_rew_sum += (cur_reward_sum * d).sum()
                    ep_len_sum += (cur_episode_length * d).sum()
                    ep_count += d.sum()
                    cur_reward_sum.masked_fill_(d, 0.0)
                    cur_episode_length.masked_fill_(d, 0.0)
                if self.log_dir is not None and "log" in infos:
                    ep_infos.append(infos["log"])
                ev_mid.record()
                self.alg.compute_returns(privileged_obs, last_values=None if self.collector is None else last_values)
            loss_dict = self.alg.update()                      # ends with the iteration's host read (loss sums / learning rate)
            ev_end.record()
            ev_end.synchronize()
            # device-side timing: no host wait between the rollout and the update, so the host enqueues the update's launches
            # while the (one-launch, asynchronous) fused rollout still runs
            collection_time, learn_time = ev_start.elapsed_time(ev_mid) * 1e-3, ev_mid.elapsed_time(ev_end) * 1e-3
            self.current_learning_iteration = it
            self.tot_timesteps += self.num_steps_per_env * N * world
            self.tot_time += collection_time + learn_time
            stats = torch.stack([ep_rew_sum, ep_len_sum, ep_count])
            if world > 1:
                torch.distributed.all_reduce(stats)
            n_ep = max(float(stats[2]), 1.0)
            rec = {"iteration": it, "Loss/value_function": loss_dict["value_function"], "Loss/surrogate": loss_dict["surrogate"],
                   "Loss/learning_rate": self.alg.learning_rate, "Policy/mean_noise_std": float(self.alg.policy.action_std.detach().mean()) if self.alg.policy.distribution is not None else None,
                   "Perf/total_fps": int(self.num_steps_per_env * N * world / (collection_time + learn_time)),
                   "Perf/collection time": collection_time, "Perf/learning_time": learn_time,
                   "Train/mean_reward": float(stats[0]) / n_ep, "Train/mean_episode_length": float(stats[1]) / n_ep, "Train/episodes": int(stats[2]),
                   "tot_timesteps": self.tot_timesteps}
            for info in ep_infos:
                for k, v in info.items():
                    rec[k] = float(v)
            ep_rew_sum.zero_(); ep_len_sum.zero_(); ep_count.zero_()
            self.history.append(rec)
            if rank == 0 and self.log_dir is not None:
                os.makedirs(self.log_dir, exist_ok=True)
                with open(os.path.join(self.log_dir, "progress.jsonl"), "a") as f:
                    f.write(json.dumps(rec) + "\n")
                if it % self.save_interval == 0:
                    self.save(os.path.join(self.log_dir, f"model_{it}.pt"))
        if rank == 0 and self.log_dir is not None:
            self.save(os.path.join(self.log_dir, f"model_{self.current_learning_iteration}.pt"))
        return self.history

    def save(self, path, infos=None):
        # on_policy_runner.py:288-302 (env state is never checkpointed by the reference either)
        # With `kernel_update` the optimizer's moments are views of one flat buffer and every parameter's "step" is the SAME
        # device scalar; torch.save would keep that aliasing and a later eager Adam (_foreach_add_ over 12 aliases of one
        # address) would race on it.  Checkpoints therefore hold independent copies.
        opt = self.alg.optimizer.state_dict()
        opt["state"] = {k: {n: (v.detach().clone() if torch.is_tensor(v) else v) for n, v in st.items()} for k, st in opt["state"].items()}
        saved = {"model_state_dict": self.alg.policy.state_dict(), "optimizer_state_dict": opt,
                 "iter": self.current_learning_iteration, "infos": infos}
        if self.empirical_normalization:
            saved["obs_norm_state_dict"] = self.obs_normalizer.state_dict()
            saved["privileged_obs_norm_state_dict"] = self.privileged_obs_normalizer.state_dict()
        torch.save(saved, path)

    def load(self, path, load_optimizer=True):
        loaded = torch.load(path, map_location=self.device, weights_only=False)
        self.alg.close()               # a captured update graph aliases the optimizer state: rebuild it from what is loaded
        self.alg.policy.load_state_dict(loaded["model_state_dict"])
        if self.empirical_normalization:
            self.obs_normalizer.load_state_dict(loaded["obs_norm_state_dict"])
            self.privileged_obs_normalizer.load_state_dict(loaded["privileged_obs_norm_state_dict"])
        if load_optimizer:
            self.alg.optimizer.load_state_dict(loaded["optimizer_state_dict"])
        self.current_learning_iteration = loaded["iter"]
        return loaded["infos"]

    def get_inference_policy(self, device=None):
        self.eval_mode()
        if device is not None:
            self.alg.policy.to(device)
        if self.empirical_normalization:                 # on_policy_runner.py:329-332
            if device is not None:
                self.obs_normalizer.to(device)
            return lambda x: self.alg.policy.act_inference(self.obs_normalizer(x))
        return self.alg.policy.act_inference

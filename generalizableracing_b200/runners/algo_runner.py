"""AlgoRunner (BPTT) with the reference's surface (standalone/diff_rl/algorithms/runner.py:27-302): the horizon loop
``env.unwrapped.detach()`` -> T x (act, step, process_env_step) -> ``alg.update()`` (:107-155)."""
from __future__ import annotations

import json
import os

import torch

from .. import dist_utils as D
from ..algorithms import BPTT
from ..modules import BaseModel


class AlgoRunner:
    def __init__(self, env, agent_cfg: dict, log_dir: str = None, device="cuda:0"):
        self.cfg = dict(agent_cfg)
        self.alg_cfg = dict(agent_cfg["algorithm"])
        self.policy_cfg = dict(agent_cfg["policy"])
        self.device = device
        self.env = env
        obs, extras = self.env.get_observations()
        num_obs = obs.shape[1]
        num_critic_obs = extras["observations"]["critic"].shape[1] if "critic" in extras["observations"] else num_obs
        if self.policy_cfg.pop("class_name", "BaseModel") != "BaseModel":
            raise ValueError("only the state-only BaseModel is built (vision / recurrent models need the depth image)")
        actor_critic = BaseModel(num_obs, num_critic_obs, self.env.num_actions, **self.policy_cfg).to(self.device)
        if self.alg_cfg.pop("class_name", "BPTT") != "BPTT":
            raise ValueError("only BPTT is built")
        self.alg = BPTT(actor_critic=actor_critic, max_iterations=agent_cfg["max_iterations"], device=self.device, env=self.env.unwrapped, **self.alg_cfg)
        if getattr(self.env.unwrapped, "_bptt", None) is not None:
            self.env.unwrapped._bptt.autograd = False          # the algorithm drives the one-launch window sweep itself
        self.num_steps_per_env = self.cfg["num_steps_per_env"]
        # opt-in: the forward half of the window as one launch (collect.py / csrc/bptt_collect.cu) + one batched actor backward
        self.collector = None
        if self.cfg.get("fused_collection", False):
            from ..collect import FusedBpttCollector
            self.collector = FusedBpttCollector(self.env.unwrapped, actor_critic, self.num_steps_per_env, groups_per_cta=int(self.cfg.get("fused_groups_per_cta", 0)),
                                                backward_tf32=bool(self.cfg.get("fused_backward_tf32", False)),
                                                backward_kernel=bool(self.cfg.get("fused_backward_kernel", False)))
        self.save_interval = self.cfg.get("save_interval", 200)
        if self.cfg.get("empirical_normalization", False):
            raise ValueError("AlgoRunner: empirical_normalization is not built for the BPTT runner (the racing / reach-target agent cfgs keep it off, "
                             "QD/agents/diff_rl_naive_cfg.py); pass empirical_normalization=False")
        self.log_dir = log_dir
        self.tot_timesteps = 0
        self.tot_time = 0
        self.current_learning_iteration = 0
        self.history = []

    def train_mode(self):
        self.alg.train_mode()

    def learn(self, num_learning_iterations: int, init_at_random_ep_len: bool = False):
        rank, world = D.world()
        if init_at_random_ep_len:
            self.env.episode_length_buf = torch.randint_like(self.env.episode_length_buf, high=int(self.env.max_episode_length))
        obs, extras = self.env.get_observations()
        critic_obs = extras["observations"].get("critic", obs)
        self.train_mode()
        N = self.env.num_envs
        start_iter = self.current_learning_iteration
        # Timing comes from CUDA events and the log record of an iteration is resolved when the log is flushed (every
        # ``log_interval`` iterations, default 1 = the reference's per-iteration logging, runner.py:157-191): with an interval > 1 the
        # loop has no host synchronisation, so the host enqueues iteration i+1 while the GPU still runs iteration i.
        log_interval = max(1, int(self.cfg.get("log_interval", 1)))
        last_iter = start_iter + num_learning_iterations - 1
        pending = []
        for it in range(start_iter, start_iter + num_learning_iterations):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            self.env.unwrapped.detach()                                            # runner.py:110
            rew_sum = torch.zeros((), device=self.device)
            if self.collector is not None:
                self.collector.pack()
                obs, critic_obs = self.collector.collect()
                rew_sum = self.collector.rewards.mean() * self.num_steps_per_env
            for _ in range(self.num_steps_per_env if self.collector is None else 0):
                actions = self.alg.act(obs, critic_obs)
                obs, rewards, dones, extras = self.env.step(actions)
                critic_obs = extras["observations"].get("critic", obs)
                self.alg.process_env_step(extras["losses"], extras["losses_detached"], dones, rewards, extras)
                rew_sum += rewards.mean()
            ev[1].record()
            _, total_loss_mean = self.alg.update() if self.collector is None else self.alg.update_fused(self.collector)
            ev[2].record()
            self.current_learning_iteration = it
            self.tot_timesteps += self.num_steps_per_env * N * world
            stats = torch.stack([total_loss_mean.detach(), rew_sum / self.num_steps_per_env])
            if world > 1:
                torch.distributed.all_reduce(stats)
                stats /= world
            pending.append((it, ev, stats, self.alg.optimizer.param_groups[0]["lr"]))
            # interval checkpoints are written HERE, with the weights of iteration `it` (runner.py:193-194), not from the deferred log flush
            if rank == 0 and self.log_dir is not None and it % self.save_interval == 0:
                os.makedirs(self.log_dir, exist_ok=True)
                self.save(os.path.join(self.log_dir, f"model_{it}.pt"))
            if len(pending) >= log_interval or it == last_iter:
                self._flush_log(pending, N * world, rank)
                pending = []
        if rank == 0 and self.log_dir is not None and num_learning_iterations > 0:        # the final checkpoint (runner.py:197-199)
            os.makedirs(self.log_dir, exist_ok=True)
            self.save(os.path.join(self.log_dir, f"model_{self.current_learning_iteration}.pt"))
        return self.history

    def _flush_log(self, pending, n_global: int, rank: int):
        pending[-1][1][2].synchronize()                        # the one host wait per flush
        vals = torch.stack([p[2] for p in pending]).tolist()
        for (it, ev, _, lr), v in zip(pending, vals):
            collection_time, learn_time = ev[0].elapsed_time(ev[1]) * 1e-3, ev[1].elapsed_time(ev[2]) * 1e-3
            self.tot_time += collection_time + learn_time
            rec = {"iteration": it, "Loss/mean_total_loss": v[0], "Train/mean_step_reward": v[1],
                   "Perf/total_fps": int(self.num_steps_per_env * n_global / max(collection_time + learn_time, 1e-9)),
                   "Perf/collection time": collection_time, "Perf/learning_time": learn_time, "Loss/learning_rate": lr}
            self.history.append(rec)
            if rank == 0 and self.log_dir is not None:
                os.makedirs(self.log_dir, exist_ok=True)
                with open(os.path.join(self.log_dir, "progress.jsonl"), "a") as f:
                    f.write(json.dumps(rec) + "\n")

    def save(self, path, infos=None):
        torch.save({"model_state_dict": self.alg.actor_critic.state_dict(), "optimizer_state_dict": self.alg.optimizer.state_dict(),
                    "iter": self.current_learning_iteration, "infos": infos}, path)

    def load(self, path, load_optimizer=True):
        loaded = torch.load(path, map_location=self.device, weights_only=False)
        self.alg.actor_critic.load_state_dict(loaded["model_state_dict"])
        if load_optimizer:
            self.alg.optimizer.load_state_dict(loaded["optimizer_state_dict"])
        self.current_learning_iteration = loaded["iter"]
        return loaded["infos"]

    def get_inference_policy(self, device=None):
        self.alg.test_mode()
        if device is not None:
            self.alg.actor_critic.to(device)
        return self.alg.actor_critic.act_inference

#!/usr/bin/env python
"""Generates tests/golden/ref_ppo_update.pt.  Run in the build container (needs /root/reference):
    PYTHONPATH=. python tests/golden/make_ppo_golden.py
The reference's OWN PPO (standalone/rsl_rl/ext/algorithms/ppo.py: act bookkeeping, process_env_step with the time-out
bootstrap, compute_returns, update with the adaptive-KL learning rate, clipped value loss, gradient clipping, Adam) and its
RolloutStorage, unmodified, on CPU, with the hyper-parameters of QD/agents/rsl_rl_ppo_cfg.py:35-48.  The policy class is the
repo's ActorCritic (rsl_rl's is third-party and absent); everything that consumes it is the reference's code.
Stored: the policy's initial weights, the synthetic transitions (observations, the actions the reference sampled, rewards,
dones, time-outs), the mini-batch permutation it drew, and its results (returns, advantages, per-iteration losses, final learning
rate, weights after update()) for tests/test_ppo_reference_golden.py to replay through the CUDA storage / PPO on the B200.
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from generalizableracing_b200.modules import ActorCritic  # noqa: E402
from oracle import ref_modules as RM  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
ALG = dict(value_loss_coef=1.0, use_clipped_value_loss=True, clip_param=0.2, entropy_coef=0.0, num_learning_epochs=5, num_mini_batches=4,
           learning_rate=5.0e-4, schedule="adaptive", gamma=0.99, lam=0.95, desired_kl=0.01, max_grad_norm=1.0)


def load_reference_ppo():
    ns = RM.load()
    sys.modules["rsl_rl"].modules = types.ModuleType("rsl_rl.modules")
    sys.modules["rsl_rl"].modules.ActorCritic = ActorCritic
    sys.modules["rsl_rl.modules"] = sys.modules["rsl_rl"].modules
    for name in ("standalone", "standalone.rsl_rl", "standalone.rsl_rl.ext"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sto = types.ModuleType("standalone.rsl_rl.ext.storage")
    sto.RolloutStorage = ns.RolloutStorage
    sys.modules["standalone.rsl_rl.ext.storage"] = sto
    return RM._load("_gr_ref_ppo", os.path.join(RM.REF_ROOT, "standalone/rsl_rl/ext/algorithms/ppo.py")).PPO


def main(N=128, T=24, iters=2):
    PPO = load_reference_ppo()
    torch.manual_seed(0)
    policy = ActorCritic(16, 16, 4, actor_hidden_dims=[128, 128], critic_hidden_dims=[128, 128], activation="lrelu", init_noise_std=1.0)
    init = {k: v.clone() for k, v in policy.state_dict().items()}
    alg = PPO(policy, None, device="cpu", **ALG)
    alg.init_storage("rl", N, T, [16], [16], [4])
    g = torch.Generator().manual_seed(1)
    out = dict(alg=ALG, N=N, T=T, init=init, iterations=[])
    for it in range(iters):
        rec = {k: [] for k in ("obs", "critic_obs", "actions", "rewards", "dones", "time_outs")}
        for t in range(T):
            obs = torch.randn(N, 16, generator=g)
            critic = obs + 0.1 * torch.randn(N, 16, generator=g)
            torch.manual_seed(1000 * it + t)
            actions = alg.act(obs, critic)                                   # ppo.py:71-83
            rewards = torch.randn(N, generator=g) * 0.1 + 0.02 * actions.abs().sum(-1)
            dones = (torch.rand(N, generator=g) < 0.03).long()
            time_outs = (torch.rand(N, generator=g) < 0.5) & dones.bool()
            alg.process_env_step(rewards, dones, {"time_outs": time_outs})   # ppo.py:85-97
            for k, v in (("obs", obs), ("critic_obs", critic), ("actions", actions), ("rewards", rewards), ("dones", dones), ("time_outs", time_outs)):
                rec[k].append(v.clone())
        last_critic = torch.randn(N, 16, generator=g)
        alg.compute_returns(last_critic)                                     # ppo.py:99-101
        rec = {k: torch.stack(v) for k, v in rec.items()}
        rec.update(last_critic_obs=last_critic, returns=alg.storage.returns.clone(), advantages=alg.storage.advantages.clone(),
                   stored_rewards=alg.storage.rewards.clone(), values=alg.storage.values.clone(), log_prob=alg.storage.actions_log_prob.clone())
        torch.manual_seed(77 + it)
        rec["indices"] = torch.randperm(N * T)                               # what rollout_storage.py:158 is about to draw
        torch.manual_seed(77 + it)
        loss = alg.update()                                                  # ppo.py:103-190
        rec.update(value_function=loss["value_function"], surrogate=loss["surrogate"], learning_rate=alg.learning_rate,
                   params={k: v.clone() for k, v in policy.state_dict().items()})
        out["iterations"].append(rec)
        print(f"iteration {it}: value {loss['value_function']:.5f} surrogate {loss['surrogate']:.5f} lr {alg.learning_rate:.3e}")
    path = os.path.join(OUT, "ref_ppo_update.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()

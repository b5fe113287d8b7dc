#!/usr/bin/env python
"""Generates tests/golden/reach_env_closure.pt.  Run in the build container (needs /root/reference):
    PYTHONPATH=. python tests/golden/make_reach_env_golden.py
Inputs and outputs of the reference's OWN env step for the CTBR reach-target task (ManagerBasedDiffRLEnv.step / _reset_idx,
DiffActions, UniformWorldPoseCommand, reward / loss / observation terms, unmodified, over the closure simulator of
oracle/ref_closure.py; the oracle is not involved), with normalised and with sim2real (raw a_zb / body-rate) inputs; episodes
of 1.5 s and command timers of 0.5 s so that resets and re-targets fall inside the 110 recorded steps.
tests/test_vs_reference_reach_env_golden.py replays them through the oracle and the kernels where the reference tree is absent."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from generalizableracing_b200 import layout as L_  # noqa: E402
from generalizableracing_b200.config import ReachTargetCfg  # noqa: E402
from oracle import ref_closure as RC  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
CFG_KW = dict(episode_length_s=1.5, resampling_time=0.5)


def scenario(sim2real, N, T, seed):
    cfg = ReachTargetCfg.ctbr(sim2real_test=sim2real, **CFG_KW)
    g = torch.Generator().manual_seed(seed)
    ref = RC.make_reference_reach_env(cfg, N, seed=seed)
    term, cmd, data = ref.action_manager.get_term("force_torque"), ref.command_manager.get_term("desired_pos_b"), ref.scene["robot"].data
    ids = torch.arange(N)
    rnd0 = torch.zeros(N, L_.REACH_RND_STRIDE)
    torch.manual_seed(seed)
    ref._reset_idx(ids)
    cmd._update_command()                              # reach_oracle R.5
    obs0 = ref.observation_manager.compute()["policy"].clone()
    torch.manual_seed(seed)
    RC.replay_reach_reset_draws(rnd0, ids, cfg.random_drag)
    keys = ("actions", "rnd", "policy", "reward", "reward_terms", "terminated", "time_out", "losses", "loss_terms", "root_state", "pose_command_w",
            "time_left", "episode_length")
    rec = {k: [] for k in keys}
    n_timer = 0
    for t in range(T):
        if sim2real:
            a = torch.randn(N, 4, generator=g) * torch.tensor([3.0, 1.0, 1.0, 0.5]) + torch.tensor([cfg.gravity, 0.0, 0.0, 0.0])
        else:
            a = torch.randn(N, 4, generator=g) * (1.5 if t % 9 == 0 else 0.4)
        rnd = torch.zeros(N, L_.REACH_RND_STRIDE)
        torch.manual_seed(100 * seed + t)
        with torch.no_grad():
            obs, rew, terminated, time_outs, ex = ref.step(a)
        reset_ids = ref.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        timer_ids = ref.command_manager.last_timer_ids
        n_timer += len(timer_ids)
        torch.manual_seed(100 * seed + t)
        RC.replay_reach_reset_draws(rnd, reset_ids, cfg.random_drag)
        RC.replay_reach_command_draws(rnd, timer_ids, L_.REACH_RND_CMD_TIMER)
        for k, v in zip(keys, (a, rnd, obs["policy"], rew, ref.reward_manager._step_reward, terminated, time_outs, ex["losses"], ref.loss_manager._step_loss,
                               data.root_state_w, cmd.pose_command_w, cmd.time_left, ref.episode_length_buf)):
            rec[k].append(v.detach().clone())
    out = {k: torch.stack(v) for k, v in rec.items()}
    out.update(sim2real=sim2real, cfg_kw=CFG_KW, N=N, T=T, rnd0=rnd0, policy0=obs0)
    print(f"sim2real={sim2real}: resets {int((out['terminated'] | out['time_out']).sum())} (terminated {int(out['terminated'].sum())}), timer re-targets {n_timer}")
    return out


if __name__ == "__main__":
    d = {"ctbr": scenario(False, 16, 110, seed=8), "ctbr_sim2real": scenario(True, 16, 110, seed=9)}
    path = os.path.join(OUT, "reach_env_closure.pt")
    torch.save(d, path)
    print(path, os.path.getsize(path))

#!/usr/bin/env python
"""Generates tests/golden/ref_env_closure.pt.  Run in the build container (needs /root/reference):
    PYTHONPATH=. python tests/golden/make_ref_env_golden.py
The vectors come from the reference's OWN env step (ManagerBasedDiffRLEnv.step / _reset_idx, DiffActions, RacingCommand and
the MDP term functions, unmodified, over the PhysX-free closure simulator of oracle/ref_closure.py) -- the oracle is not
involved.  Stored per scenario: the inputs (startup draws, actions, the reference's random draws laid out in the rnd[N, 52]
slots of generalizableracing_b200/layout.py, positions written before a step to put drones on gates) and the reference's outputs
and state after every step, so that tests/test_ref_env_golden.py can replay them through the oracle (CPU) and through the
kernels (B200) where the reference tree is absent.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from generalizableracing_b200 import layout as L_  # noqa: E402
from generalizableracing_b200.config import RacingCfg  # noqa: E402
from generalizableracing_b200.tracks import figure_eight_track, synthetic_track_table  # noqa: E402
from oracle import ref_closure as RC  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TERM, CMD = "force_torque", "next_gate_pose"


def scenario(stage, N, T, seed):
    cfg = RacingCfg.for_stage(stage, is_differentiable_physics=True)
    table = figure_eight_track() if stage == 0 else synthetic_track_table()
    g = torch.Generator().manual_seed(seed)
    srnd0 = torch.rand(N, L_.SRND_STRIDE, generator=g)
    srnd0[:, 12:] = torch.randn(N, 4, generator=g)
    ref, srnd = RC.make_reference_env(cfg, table, N, srnd0, seed=2000 + seed)
    term, cmd, data, ter = ref.action_manager.get_term(TERM), ref.command_manager.get_term(CMD), ref.scene["robot"].data, ref.scene.terrain

    def new_rnd():
        rnd = torch.zeros(N, L_.RND_STRIDE)
        rnd[:, L_.RND_LEVEL] = torch.rand(N, generator=g)
        ter.pending_level_u = rnd[:, L_.RND_LEVEL].clone()
        return rnd

    ids = torch.arange(N)
    rnd0 = new_rnd()
    torch.manual_seed(seed)
    ref._reset_idx(ids)
    obs0 = ref.observation_manager.compute()
    torch.manual_seed(seed)
    RC.replay_reset_draws(rnd0, ids, cfg.add_cmd_noise)
    RC.replay_obs_draws(rnd0)
    ep0 = torch.randint(cfg.max_episode_length - T // 2, cfg.max_episode_length - 1, (N,), generator=g)
    ref.episode_length_buf[:] = ep0
    rec = {k: [] for k in ("actions", "rnd", "policy", "critic", "aux", "reward", "terminated", "time_out", "achieved", "losses",
                           "root_state", "gate_id", "accumulate_gates", "terrain_levels", "episode_length", "gross_thrust", "torque")}
    teleports = {}
    for t in range(T):
        if t % 4 == 3:
            sel = torch.rand(N, generator=g) < 0.5
            off = (torch.rand(N, 3, generator=g) * 2 - 1) * 0.5 / (3 ** 0.5)
            pos = torch.where(sel[:, None], cmd.gate_pose_gt_w[:, :3] + off, data.root_pos_w)
            data.root_pos_w = pos.clone()
            term.get_state_from_sim()
            term.drone_dynamics.reset_state(term.states_all, ids)
            teleports[t] = pos.clone()
        a = torch.randn(N, 4, generator=g) * (2.0 if t % 10 == 0 else 0.5)
        rnd = new_rnd()
        torch.manual_seed(100 * seed + t)
        with torch.no_grad():
            obs, rew, terminated, time_outs, ex = ref.step(a)
        reset_ids = ref.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        achieved = ref.command_manager.last_achieved
        torch.manual_seed(100 * seed + t)
        RC.replay_reset_draws(rnd, reset_ids, cfg.add_cmd_noise)
        RC.replay_pass_draws(rnd, achieved, cfg.add_cmd_noise)
        RC.replay_obs_draws(rnd)
        for k, v in (("actions", a), ("rnd", rnd), ("policy", obs["policy"]), ("critic", obs["critic"]), ("aux", obs["auxiliary"]), ("reward", rew),
                     ("terminated", terminated), ("time_out", time_outs), ("achieved", achieved), ("losses", ex["losses"]),
                     ("root_state", data.root_state_w), ("gate_id", cmd.gate_id), ("accumulate_gates", cmd.metrics["accumulate_gates"]),
                     ("terrain_levels", ter.terrain_levels), ("episode_length", ref.episode_length_buf),
                     ("gross_thrust", term.controller.gross_thrust[:, 0]), ("torque", term.controller.torque)):
            rec[k].append(v.detach().clone())
    out = {k: torch.stack(v) for k, v in rec.items()}
    out.update(stage=stage, N=N, T=T, startup_rnd=srnd, rnd0=rnd0, episode_length0=ep0, teleports=teleports,
               policy0=obs0["policy"].clone(), critic0=obs0["critic"].clone())
    print(f"stage {stage}: resets {int((out['terminated'] | out['time_out']).sum())}, gate passes {int(out['achieved'].sum())}")
    return out


if __name__ == "__main__":
    d = {"stage0_figure8": scenario(0, 24, 64, seed=3), "stage1_table": scenario(1, 24, 64, seed=4)}
    path = os.path.join(OUT, "ref_env_closure.pt")
    torch.save(d, path)
    print(path, os.path.getsize(path))

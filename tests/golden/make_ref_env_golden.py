#!/usr/bin/env python
"""Generates tests/golden/ref_env_closure.pt.  Run in the build container (needs /root/reference):
    PYTHONPATH=. python tests/golden/make_ref_env_golden.py
The vectors come from the reference's OWN env step (ManagerBasedDiffRLEnv.step / _reset_idx, DiffActions, RacingCommand and
the MDP term functions, unmodified, over the PhysX-free closure simulator of oracle/ref_closure.py) -- the oracle is not
involved.  Stored per scenario: the inputs (startup draws, actions, the reference's random draws laid out in the rnd[N, 52]
slots of generalizableracing_b200/layout.py, positions written before a step to put drones on gates) and the reference's outputs
and state after every step (plus one differentiable 16-step window with the reference's autograd gradient of the mean loss with
respect to every action), so that tests/test_ref_env_golden.py / test_window_gradient_reference_golden.py can replay them through the oracle (CPU) and through the
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


def bptt_window(N, H, seed):
    """One differentiable window of the reference env (STAGE 0, is_differentiable_physics) with time-out resets inside it:
    d mean_{t,n}(loss) / d action from the reference's own autograd graph (S/diff_rl/algorithms/bptt.py:38-44)."""
    cfg = RacingCfg.for_stage(0, is_differentiable_physics=True)
    g = torch.Generator().manual_seed(seed)
    srnd0 = torch.rand(N, L_.SRND_STRIDE, generator=g)
    srnd0[:, 12:] = torch.randn(N, 4, generator=g)
    ref, srnd = RC.make_reference_env(cfg, figure_eight_track(), N, srnd0, seed=3000 + seed)
    term, cmd, data, ter = ref.action_manager.get_term(TERM), ref.command_manager.get_term(CMD), ref.scene["robot"].data, ref.scene.terrain
    dyn = term.drone_dynamics
    ids = torch.arange(N)
    rnd0 = torch.zeros(N, L_.RND_STRIDE)
    torch.manual_seed(seed)
    ref._reset_idx(ids)
    torch.manual_seed(seed)
    RC.replay_reset_draws(rnd0, ids, cfg.add_cmd_noise)
    RC.replay_obs_draws(rnd0)
    off = (torch.rand(N, 3, generator=g) * 2 - 1) * 1.0 / (3 ** 0.5)
    pos0 = cmd.gate_pose_gt_w[:, :3] + off
    data.root_pos_w = pos0.clone()
    term.get_state_from_sim()
    dyn.reset_state(term.states_all, ids)
    ep0 = torch.randint(cfg.max_episode_length - H - 4, cfg.max_episode_length + H, (N,), generator=g) - H
    ref.episode_length_buf[:] = ep0
    ref.detach()

    def fresh_copies(env_ids):                     # see tests/test_oracle_vs_reference_env.py: the reference's in-place reset writes
        term.thr_est_error = term.thr_est_error.clone()
        dyn.drag_coeffs, dyn.h_force_drag_coeffs = dyn.drag_coeffs.clone(), dyn.h_force_drag_coeffs.clone()
    ref.recorder_manager.pre_reset_hook = fresh_copies
    acts = [(torch.randn(N, 4, generator=g) * 0.5).requires_grad_(True) for _ in range(H)]
    rnds, losses, n_reset = [], [], 0
    for t in range(H):
        rnd = torch.zeros(N, L_.RND_STRIDE)
        rnd[:, L_.RND_LEVEL] = torch.rand(N, generator=g)
        ter.pending_level_u = rnd[:, L_.RND_LEVEL].clone()
        torch.manual_seed(500 + t)
        ex = ref.step(acts[t])[4]
        reset_ids = ref.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        n_reset += len(reset_ids)
        torch.manual_seed(500 + t)
        RC.replay_reset_draws(rnd, reset_ids, cfg.add_cmd_noise)
        RC.replay_pass_draws(rnd, ref.command_manager.last_achieved, cfg.add_cmd_noise)
        RC.replay_obs_draws(rnd)
        rnds.append(rnd)
        losses.append(ex["losses"])
    torch.stack(losses).mean().backward()
    grad = torch.stack([a.grad if a.grad is not None else torch.zeros(N, 4) for a in acts])
    print(f"bptt window: resets {n_reset}, |grad| max {float(grad.abs().max()):.3e}")
    return dict(N=N, H=H, startup_rnd=srnd, rnd0=rnd0, pos0=pos0, episode_length0=ep0, actions=torch.stack([a.detach() for a in acts]),
                rnd=torch.stack(rnds), losses=torch.stack(losses).detach(), grad_actions=grad, resets=n_reset)


if __name__ == "__main__":
    d = {"stage0_figure8": scenario(0, 24, 64, seed=3), "stage1_table": scenario(1, 24, 64, seed=4), "bptt_window": bptt_window(32, 16, seed=6)}
    path = os.path.join(OUT, "ref_env_closure.pt")
    torch.save(d, path)
    print(path, os.path.getsize(path))

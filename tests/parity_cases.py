"""Shared parity scenarios: the same seeded inputs are pushed through the oracle (CPU torch restatement of the
reference) and through the kernels (CUDA library on the GPU box, or the g++ emulation of the same sources here)."""
from __future__ import annotations

import torch

from generalizableracing_b200 import layout as L_
from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.env import RacingVecEnv
from generalizableracing_b200.tracks import figure_eight_track, synthetic_track_table
from oracle import racing_oracle as RO

# fp32 tolerance of the north star: 1e-5 relative per step (relative to the magnitude of the compared vector)
REL_TOL_STEP = 1e-5


def draw_startup(N, g):
    s = torch.rand(N, L_.SRND_STRIDE, generator=g)
    s[:, 12:] = torch.randn(N, 4, generator=g)
    return s


def draw_rnd(N, g):
    r = torch.rand(N, L_.RND_STRIDE, generator=g)
    r[:, :8] = torch.randn(N, 8, generator=g)
    return r


def make_pair(backend, stage=0, N=64, table=None, seed=0, diff=False, horizon=0, **cfg_over):
    device, lib = backend
    cfg = RacingCfg.for_stage(stage, is_differentiable_physics=diff, **cfg_over)
    table = table or (figure_eight_track() if stage == 0 else synthetic_track_table())
    g = torch.Generator().manual_seed(seed)
    srnd = draw_startup(N, g)
    orc = RO.OracleRacingEnv(cfg, table, N, srnd)
    env = RacingVecEnv(cfg, table, N, device=device, rng_mode="dense", startup_rnd=srnd, bptt_horizon=horizon, _lib=lib)
    env.export_reward_terms = True
    env.export_gate_passed = True
    return cfg, table, orc, env, g


def rel_err(ref: torch.Tensor, got: torch.Tensor) -> float:
    """max |ref-got| / max(1, max|ref|) -- per-step relative error on the scale of the compared quantity."""
    ref = ref.detach().double().cpu()
    got = got.detach().double().cpu()
    return float((ref - got).abs().max() / max(1.0, float(ref.abs().max())))


def oracle_state(orc):
    return dict(pos=orc.root_pos_w, quat=orc.root_quat_w, lin=orc.root_lin_vel_w, ang=orc.root_ang_vel_w,
                f=orc.ctrl.gross_thrust[:, 0], tau=orc.ctrl.torque, gate=orc.gate_id, acc=orc.accumulate_gates,
                level=orc.terrain_levels, eplen=orc.episode_length_buf, thr=orc.thr_est_error,
                k2=orc.dyn.drag_coeffs, k1=orc.dyn.h_force_drag_coeffs)


def kernel_state(env):
    sv = env.state_dict_view()
    from oracle import isaac_math as M
    ang_w = M.quat_rotate(sv["root_quat_w"].cpu(), sv["root_ang_vel_b"].cpu())     # the kernels keep the body-frame rate
    return dict(pos=sv["root_pos_w"], quat=sv["root_quat_w"], lin=sv["root_lin_vel_w"], ang=ang_w,
                f=sv["gross_thrust"], tau=sv["torque"], gate=sv["gate_id"], acc=sv["accumulate_gates"],
                level=sv["terrain_levels"], eplen=sv["episode_length"], thr=sv["thr_est_error"],
                k2=sv["drag_coeffs"], k1=sv["h_force_drag_coeffs"])


INT_KEYS = ("gate", "acc", "level", "eplen")


def compare_states(orc, env):
    """returns (max rel err over float columns, number of integer mismatches)"""
    o, k = oracle_state(orc), kernel_state(env)
    worst, bad = 0.0, 0
    for name in o:
        if name in INT_KEYS:
            bad += int((o[name].long().cpu() != k[name].long().cpu()).sum())
        else:
            worst = max(worst, rel_err(o[name], k[name]))
    return worst, bad


def teleport_near_gate(orc, env, g, frac=0.5, radius=0.5):
    """Move a random subset of envs to within `radius` of their current gate (same values on both sides) so that gate
    passing / success_cross / curriculum paths are exercised; the reference never reaches gates under random actions."""
    N = orc.num_envs
    sel = torch.rand(N, generator=g) < frac
    off = (torch.rand(N, 3, generator=g) * 2 - 1) * radius / (3 ** 0.5)
    new_pos = orc.gate_pose_gt_w[:, :3] + off
    pos = torch.where(sel[:, None], new_pos, orc.root_pos_w)
    orc.root_pos_w = pos.clone()
    orc._get_state_from_sim()
    orc.dyn.reset_state(orc.states_all, torch.arange(N))
    env.write_plane(L_.PL_POS, slice(0, 3), pos)
    return sel


def run_rollout(orc, env, g, steps, action_std=0.5, teleport_every=0, sync_every=0):
    """Free-running rollout with identical actions and random numbers on both sides.
    Returns per-step worst relative errors and mismatch counts."""
    N = orc.num_envs
    stats = dict(obs=0.0, critic=0.0, reward=0.0, terms=0.0, state=0.0, loss=0.0, mask_mismatch=0, int_mismatch=0,
                 resets=0, gates=0, aux_mismatch=0)
    for t in range(steps):
        if teleport_every and t % teleport_every == teleport_every - 1:
            teleport_near_gate(orc, env, g)
        a = torch.randn(N, 4, generator=g) * action_std
        r = draw_rnd(N, g)
        with torch.no_grad():
            oo, orew, oterm, oto, oex = orc.step(a, r)
        ko, krew, kdones, kex = env.step(a.to(env.device), r.to(env.device))
        stats["obs"] = max(stats["obs"], rel_err(oo["policy"], ko))
        stats["critic"] = max(stats["critic"], rel_err(oo["critic"], kex["observations"]["critic"]))
        stats["reward"] = max(stats["reward"], rel_err(orew, krew))
        stats["terms"] = max(stats["terms"], rel_err(orc.step_reward, env._last["reward_terms"]))
        stats["aux_mismatch"] += int((oo["auxiliary"].cpu() != kex["observations"]["auxiliary"].cpu()).sum())
        mm = int((oterm.cpu() != kex["terminated"].cpu()).sum()) + int((oto.cpu() != kex["time_outs"].cpu()).sum())
        mm += int(((oterm | oto).long().cpu() != kdones.cpu()).sum())
        mm += int((orc.last_achieved.cpu() != env._last["gate_passed"].bool().cpu()).sum())
        stats["mask_mismatch"] += mm
        w, bad = compare_states(orc, env)
        stats["state"] = max(stats["state"], w)
        stats["int_mismatch"] += bad
        stats["resets"] += int((oterm | oto).sum())
        stats["gates"] += int(orc.last_achieved.sum())
        if "losses" in oex and "losses" in kex:
            stats["loss"] = max(stats["loss"], rel_err(oex["losses"], kex["losses"]))
    return stats

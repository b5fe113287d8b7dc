"""Smaller pieces: the bad_pose predicate used by the kernels vs the literal reference chain, track tables,
episode-log accumulators (extras["log"]) and the config mirror."""
import math

import numpy as np
import pytest
import torch

from generalizableracing_b200 import layout as L_
from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.tracks import figure_eight_track, synthetic_track_table
from oracle import isaac_math as M
from tests.conftest import backend_params
from tests import parity_cases as PC


def test_bad_pose_predicate_equals_literal_chain():
    """kernels: bad = (1 - 2(x^2+y^2)) < 0.  Reference (termination.py:24-33): euler_xyz_from_quat -> %2pi -> wrap_to_pi
    -> |roll| > pi/2 or |pitch| > pi/2.  They may differ only in the few-ulp band around cos_roll = 0."""
    g = torch.Generator().manual_seed(0)
    q = torch.nn.functional.normalize(torch.randn(2_000_000, 4, generator=g), dim=-1)
    # add a dense cloud near the decision boundary and exact special cases
    near = q[:200_000].clone()
    near[:, 2] = 0
    near[:, 1] = torch.sqrt(torch.tensor(0.5)) + (torch.rand(200_000, generator=g) - 0.5) * 1e-4
    near[:, 0] = torch.sqrt((1 - near[:, 1] ** 2 - near[:, 3] ** 2).clamp(min=0))
    special = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0.70710678, 0.70710678, 0, 0], [0.70710678, 0, 0.70710678, 0], [0.5, 0.5, 0.5, 0.5], [0, 0, 1.0, 0]])
    q = torch.cat([q, near, special])
    roll, pitch, _ = M.euler_xyz_from_quat(q)
    lit = (M.wrap_to_pi(roll).abs() > math.pi / 2) | (M.wrap_to_pi(pitch).abs() > math.pi / 2)
    cos_roll = 1 - 2 * (q[:, 1] * q[:, 1] + q[:, 2] * q[:, 2])
    mine = cos_roll < 0
    diff = lit != mine
    assert int(diff[:2_000_000].sum()) == 0                            # random attitudes: identical
    # inside the deliberately dense boundary cloud the literal chain is a rounding lottery: every disagreement sits
    # within one ulp of cos_roll = 0 (and atan2f on the GPU differs from the CPU's by an ulp there anyway)
    assert int(diff.sum()) < 200
    assert (not diff.any()) or float(cos_roll[diff].abs().max()) < 1e-6


def test_track_tables():
    f8 = figure_eight_track()
    assert f8.gate_pose.shape == (1, 1, 6, 7) and f8.next_gate_id[0, 0] == 0
    assert np.allclose(f8.gate_pose[0, 0, :, :3], np.array([[3, 3, 0], [5, 0, 0], [3, -3, 0], [-3, 3, 0], [-5, 0, 0], [-3, -3, 0]], np.float32))
    assert np.allclose(np.abs(f8.gate_pose[0, 0, 0, 3:]), [0, 0, 0, 1], atol=1e-6) and np.allclose(f8.gate_pose[0, 0, 1, 3:], [0.70710678, 0, 0, 0.70710678], atol=1e-6)
    t = synthetic_track_table()
    assert t.gate_pose.shape == (20, 10, 8, 7) and t.terrain_origins.shape == (10, 20, 3)
    assert np.allclose(np.linalg.norm(t.gate_pose[..., 3:], axis=-1), 1.0, atol=1e-5)
    assert (t.gate_pose[..., 2] + t.terrain_origins.transpose(1, 0, 2)[:, :, None, 2] >= 0.79).all()        # gate heights clipped to [0.8, 2]
    t2 = synthetic_track_table()
    assert np.array_equal(t.gate_pose, t2.gate_pose)                   # seeded
    # tiles are laid out on a 40 m grid centred on the world origin
    rows = (np.arange(10) + 0.5) * 40.0 - 200.0
    cols = (np.arange(20) + 0.5) * 40.0 - 400.0
    assert np.all(np.abs(t.terrain_origins[:, :, 0] - rows[:, None]) <= 20.0) and np.all(np.abs(t.terrain_origins[:, :, 1] - cols[None, :]) <= 20.0)
    with pytest.raises(ValueError):
        type(t)(t.gate_pose[..., :6], t.next_gate_id, t.terrain_origins)


def test_config_stages():
    c0, c1, c2 = (RacingCfg.for_stage(s) for s in (0, 1, 2))
    assert (c0.add_cmd_noise, c0.term_out_of_bound, c0.term_bad_pose, c0.w_success, c0.w_bad_pose) == (False, True, False, 10.0, 0.0)
    assert (c1.add_cmd_noise, c1.term_bad_pose, c1.w_success, c1.w_bad_pose, c1.noise_curriculum) == (True, True, 20.0, -30.0, True)
    assert (c2.cmd_noise_pos, c2.episode_length_s, c2.max_episode_length, c2.noise_curriculum) == (0.5, 8.0, 267, False)
    assert abs(c1.gross_thrust_bound[0] + 4.6546618) < 1e-6 and abs(c1.gross_thrust_bound[1] - 86.8306905) < 1e-6
    with pytest.raises(ValueError):
        RacingCfg.for_stage(3)


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
@pytest.mark.parametrize("stage,diff", [(0, False), (1, False), (2, False), (1, True)])
def test_episode_log_accumulators(backend, stage, diff):
    """extras["log"]: means over the envs reset since the last read == the oracle's per-reset logs, aggregated -- every key the reference
    writes in _reset_idx (manager_based_diff_rl_env.py:380-407): Episode_Reward/*, Episode_Loss/* (LossManager episode sums),
    Curriculum/*, Metrics/next_gate_pose/{accumulate_gates, action_rate, avg_lin_spd, avg_ang_spd} (QD/mdp/commands.py:257-260).
    Half of the envs start one step before their time-out, so the log of the very first step after reset() is exercised too (their
    command metrics were zeroed by CommandTerm.reset and no command update has run since)."""
    N = 128
    cfg, table, orc, env, g = PC.make_pair(backend, stage=stage, N=N, seed=8 + stage, diff=diff, horizon=130)
    r0 = PC.draw_rnd(N, g)
    orc.reset(r0)
    env.reset(r0.to(env.device))
    _ = env.extras["log"]                                              # drain the reset's own log
    ep = torch.randint(0, cfg.max_episode_length, (N,), generator=g)
    ep[::2] = cfg.max_episode_length - 1
    orc.episode_length_buf[:] = ep
    env.episode_length_buf = ep
    if diff:
        orc.detach()
        env.detach()
    metric_keys = ["Metrics/next_gate_pose/" + k for k in ("accumulate_gates", "action_rate", "avg_lin_spd", "avg_ang_spd")]
    keys = ["Episode_Reward/" + k for k, w in zip(L_.REWARD_TERM_NAMES, orc.reward_weights) if w != 0.0] + ["Episode_Loss/" + k for k in L_.LOSS_TERM_NAMES] + metric_keys
    n_reset, sums = 0, {k: 0.0 for k in keys}
    for t in range(120):
        if t % 5 == 4:
            PC.teleport_near_gate(orc, env, g)
        a, r = torch.randn(N, 4, generator=g) * 0.5, PC.draw_rnd(N, g)
        with torch.no_grad():
            _, _, term, to, oex = orc.step(a, r)
        _, _, _, kex = env.step(a.to(env.device), r.to(env.device))
        k = int((term | to).sum())
        if k:
            n_reset += k
            for name in keys:
                sums[name] += float(oex["log"][name]) * k
        if t == 0:                                # the first step after reset(): N/2 time-outs whose command metrics are still zero
            assert k >= N // 2
            first = kex["log"]
            for name in keys:
                ref = sums[name] / k
                tol = (2e-3 if name.endswith("action_rate") else 1e-4) * max(1.0, abs(ref))
                assert abs(float(first[name]) - ref) < tol, (name, float(first[name]), ref)
            for name in keys:
                sums[name] = 0.0
            n_reset = 0
        if diff:
            assert [n for n, _ in kex["log_losses"]] == [n for n, _ in oex["log_losses"]]
            for (_, kv), (_, ov) in zip(kex["log_losses"], oex["log_losses"]):
                assert abs(float(kv) - ov) < 1e-5 * max(1.0, abs(ov))
    log = env.extras["log"]
    assert n_reset > 20
    assert set(keys) <= set(log.keys())
    for name in keys:
        ref = sums[name] / n_reset
        tol = (2e-3 if name.endswith("action_rate") else 1e-4) * max(1.0, abs(ref))       # action_rate rides as a 16-bit float per env (2^-12)
        assert abs(float(log[name]) - ref) < tol, (name, float(log[name]), ref)
    if diff:
        assert abs(sums["Episode_Loss/move_towards_goal"]) > 0
    else:
        assert all(float(log["Episode_Loss/" + k]) == 0.0 for k in L_.LOSS_TERM_NAMES)   # the sums only move with differentiable physics
    assert abs(float(log["Curriculum/terrain_levels"]) - float(orc.terrain_levels.float().mean())) < 1e-6
    if cfg.noise_curriculum and cfg.add_cmd_noise:
        assert abs(float(log["Curriculum/command_noise_level"]) - float(orc.noise_level.mean())) < 1e-5
    assert float(log["Episode_Termination/time_out"]) + float(log["Episode_Termination/terminated"]) >= n_reset


@pytest.mark.gpu
def test_gate_predicate_square_root_is_correctly_rounded(cuda_lib):
    """The library is built with -prec-sqrt=false -prec-div=false (fast paths for the smooth quantities); the square root behind the gate
    predicate `|gate - pos| < 0.35` must not follow: a 1-ulp error there flips gate passes against the reference.  Bit-for-bit against numpy's
    IEEE square root on a million values, dense around the threshold 0.35^2 and across the exponent range."""
    import numpy as np
    lib = cuda_lib
    rng = np.random.default_rng(0)
    thr2 = np.float32(0.35) * np.float32(0.35)
    base = np.float32(thr2).view(np.uint32)
    around = (np.arange(-200000, 200001, dtype=np.int64) + int(base)).astype(np.uint32).view(np.float32)          # 400,001 consecutive floats around 0.1225
    wide = np.exp(rng.uniform(np.log(1e-12), np.log(1e12), 600000)).astype(np.float32)
    x = np.concatenate([around, wide, np.array([0.0, 1.0, 4.0, 1e-38, 3e38], dtype=np.float32)])
    xt = torch.from_numpy(x).cuda()
    yt = torch.empty_like(xt)
    assert lib.gr_selftest_sqrt_rn(xt.data_ptr(), yt.data_ptr(), xt.numel(), torch.cuda.current_stream().cuda_stream) == 0
    got = yt.cpu().numpy()
    want = np.sqrt(x.astype(np.float32))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), int((got.view(np.uint32) != want.view(np.uint32)).sum())
    # and the decision itself on the consecutive floats around the threshold
    assert np.array_equal(got[:around.size] < np.float32(0.35), want[:around.size] < np.float32(0.35))

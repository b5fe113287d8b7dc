"""Pins the oracle's env step (SURVEY.md §8a rows a1, a5-a10 and the step order of §3.3) against the reference's OWN
``ManagerBasedDiffRLEnv.step`` / ``_reset_idx``, ``DiffActionManager``, ``LossManager``, ``DiffActions``, ``RacingCommand`` and
MDP term functions, executed unmodified where they lie under /root/reference over the PhysX-free closure simulator
(oracle/ref_closure.py).  Same actions, same random draws (the reference's global-generator calls are replayed into the
oracle's explicit ``rnd`` rows), teleports onto gates so that gate passing / curricula / resets all fire.
Skipped on the GPU box (the reference tree does not travel)."""
import pytest
import torch

from generalizableracing_b200 import layout as L_
from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.tracks import figure_eight_track, synthetic_track_table
from oracle import racing_oracle as RO
from oracle import ref_modules
from tests import parity_cases as PC

pytestmark = pytest.mark.skipif(not ref_modules.available(), reason="reference tree not present")

TERM = "force_torque"
CMD = "next_gate_pose"


def _make(stage, N, seed, diff=False):
    from oracle import ref_closure as RC
    cfg = RacingCfg.for_stage(stage, is_differentiable_physics=diff)
    table = figure_eight_track() if stage == 0 else synthetic_track_table()
    g = torch.Generator().manual_seed(seed)
    srnd0 = PC.draw_startup(N, g)
    ref, srnd = RC.make_reference_env(cfg, table, N, srnd0, seed=1000 + seed)
    orc = RO.OracleRacingEnv(cfg, table, N, srnd)
    return RC, cfg, ref, orc, g


def _reset(RC, ref, orc, g, seed):
    N = orc.num_envs
    rnd = torch.zeros(N, L_.RND_STRIDE)
    rnd[:, L_.RND_LEVEL] = torch.rand(N, generator=g)
    ref.scene.terrain.pending_level_u = rnd[:, L_.RND_LEVEL].clone()
    ids = torch.arange(N)
    torch.manual_seed(seed)
    ref._reset_idx(ids)                                       # ManagerBasedRLEnv.reset(): _reset_idx(all) then observations
    ref_obs = ref.observation_manager.compute()
    torch.manual_seed(seed)
    RC.replay_reset_draws(rnd, ids, orc.cfg.add_cmd_noise)
    RC.replay_obs_draws(rnd)
    obs, _ = orc.reset(rnd)
    return ref_obs, obs


def _step(RC, cfg, ref, orc, g, action, seed):
    N = orc.num_envs
    rnd = torch.zeros(N, L_.RND_STRIDE)
    rnd[:, L_.RND_LEVEL] = torch.rand(N, generator=g)
    ref.scene.terrain.pending_level_u = rnd[:, L_.RND_LEVEL].clone()
    torch.manual_seed(seed)
    r_obs, r_rew, r_term, r_to, r_ex = ref.step(action)       # the reference's own step, unmodified
    reset_ids = ref.reset_buf.nonzero(as_tuple=False).squeeze(-1)
    achieved = ref.command_manager.last_achieved
    torch.manual_seed(seed)
    RC.replay_reset_draws(rnd, reset_ids, cfg.add_cmd_noise)
    RC.replay_pass_draws(rnd, achieved, cfg.add_cmd_noise)
    RC.replay_obs_draws(rnd)
    o_obs, o_rew, o_term, o_to, o_ex = orc.step(action, rnd)
    return (r_obs, r_rew, r_term, r_to, r_ex), (o_obs, o_rew, o_term, o_to, o_ex), achieved


def _teleport(ref, orc, g, frac=0.5, radius=0.5):
    """Same move on both sides, through each side's own state plumbing (reference: DiffActions.get_state_from_sim +
    DroneDynamics.reset_state)."""
    N = orc.num_envs
    sel = torch.rand(N, generator=g) < frac
    off = (torch.rand(N, 3, generator=g) * 2 - 1) * radius / (3 ** 0.5)
    pos = torch.where(sel[:, None], orc.gate_pose_gt_w[:, :3] + off, orc.root_pos_w)
    orc.root_pos_w = pos.clone()
    orc._get_state_from_sim()
    orc.dyn.reset_state(orc.states_all, torch.arange(N))
    ref.scene["robot"].data.root_pos_w = pos.clone()
    term = ref.action_manager.get_term(TERM)
    term.get_state_from_sim()
    term.drone_dynamics.reset_state(term.states_all, torch.arange(N))


def _assert_state_equal(ref, orc, where):
    term = ref.action_manager.get_term(TERM)
    cmd = ref.command_manager.get_term(CMD)
    data = ref.scene["robot"].data
    ter = ref.scene.terrain
    pairs = {
        "root_pos_w": (data.root_pos_w, orc.root_pos_w), "root_quat_w": (data.root_quat_w, orc.root_quat_w),
        "root_lin_vel_w": (data.root_lin_vel_w, orc.root_lin_vel_w), "root_ang_vel_w": (data.root_ang_vel_w, orc.root_ang_vel_w),
        "body_ang_acc_w": (data.body_ang_acc_w[:, 0], orc.body_ang_acc_w),
        "dyn.pos": (term.drone_dynamics.pos, orc.dyn.pos), "dyn.quat": (term.drone_dynamics.quat, orc.dyn.quat),
        "dyn.lin_vel_b": (term.drone_dynamics.lin_vel_b, orc.dyn.lin_vel_b), "dyn.ang_vel_b": (term.drone_dynamics.ang_vel_b, orc.dyn.ang_vel_b),
        "drag2": (term.drone_dynamics.drag_coeffs, orc.dyn.drag_coeffs), "drag1": (term.drone_dynamics.h_force_drag_coeffs, orc.dyn.h_force_drag_coeffs),
        "gross_thrust": (term.controller.gross_thrust, orc.ctrl.gross_thrust), "torque": (term.controller.torque, orc.ctrl.torque),
        "rate_gain_p": (term.controller.rate_gain_p, orc.ctrl.rate_gain_p), "rate_gain_d": (term.controller.rate_gain_d, orc.ctrl.rate_gain_d),
        "thrust_ctrl_delay": (term.controller.thrust_ctrl_delay, orc.ctrl.thrust_ctrl_delay),
        "torque_ctrl_delay": (term.controller.torque_ctrl_delay, orc.ctrl.torque_ctrl_delay),
        "thr_est_error": (term.thr_est_error, orc.thr_est_error), "raw_actions": (term.raw_actions, orc.raw_actions),
        "action": (ref.action_manager.action, orc.action), "prev_action": (ref.action_manager.prev_action, orc.prev_action),
        "action_scale": (term.action_scale, orc.action_scale), "action_offset": (term.action_offset, orc.action_offset),
        "gate_id": (cmd.gate_id, orc.gate_id), "next_gate_id": (cmd.next_gate_id, orc.next_gate_id),
        "gate_pose_w": (cmd.gate_pose_w, orc.gate_pose_w), "gate_pose_gt_w": (cmd.gate_pose_gt_w, orc.gate_pose_gt_w),
        "next_gate_pose_w": (cmd.next_gate_pose_w, orc.next_gate_pose_w), "next_gate_pose_gt_w": (cmd.next_gate_pose_gt_w, orc.next_gate_pose_gt_w),
        "accumulate_gates": (cmd.metrics["accumulate_gates"], orc.accumulate_gates), 
        "metric_action_rate": (cmd.metrics["action_rate"], orc.metric_action_rate), "metric_avg_lin_spd": (cmd.metrics["avg_lin_spd"], orc.metric_avg_lin_spd),
        "metric_avg_ang_spd": (cmd.metrics["avg_ang_spd"], orc.metric_avg_ang_spd), "noise_level": (cmd.noise_level, orc.noise_level),
        "noise_range_pos_x": (cmd.noise_range_pos_x, orc.noise_range_pos_x), "noise_range_yaw": (cmd.noise_range_yaw, orc.noise_range_yaw),
        "terrain_levels": (ter.terrain_levels, orc.terrain_levels), "env_origins": (ter.env_origins, orc.env_origins),
        "episode_length_buf": (ref.episode_length_buf, orc.episode_length_buf),
    }
    for i, name in enumerate(ref.reward_manager._term_names):
        pairs["episode_sum/" + name] = (ref.reward_manager._episode_sums[name], orc.episode_sums[:, orc.reward_term_names.index(name)])
    for name, (a, b) in pairs.items():
        assert torch.equal(a.detach(), b.detach().to(a.dtype)), f"{where}: {name} differs by {(a.double() - b.double()).abs().max():.3e}"


@pytest.mark.parametrize("stage", [0, 1, 2])
def test_step_and_reset_bit_exact_with_reference_env(stage):
    """120 free-running steps with the reference's noise, curricula, resets (terminations + forced time-outs) and gate passes:
    every output and every state column of the oracle equals the reference's, bit for bit."""
    N = 96
    RC, cfg, ref, orc, g = _make(stage, N, seed=stage)
    assert ref.max_episode_length == cfg.max_episode_length                       # manager_based_diff_rl_env.py:100-102
    r_obs, o_obs = _reset(RC, ref, orc, g, seed=7)
    for k in ("policy", "critic", "auxiliary"):
        assert torch.equal(r_obs[k], o_obs[k]), k
    _assert_state_equal(ref, orc, "after reset")
    # stagger the episode clocks so that time-outs fall inside the test
    ep = torch.randint(cfg.max_episode_length - 60, cfg.max_episode_length - 1, (N,), generator=g)
    ref.episode_length_buf[:] = ep
    orc.episode_length_buf[:] = ep
    n_reset = n_pass = n_term = 0
    for t in range(120):
        if t % 4 == 3:
            _teleport(ref, orc, g)
        a = torch.randn(N, 4, generator=g) * (2.0 if t % 10 == 0 else 0.5)
        (r_obs, r_rew, r_term, r_to, r_ex), (o_obs, o_rew, o_term, o_to, o_ex), achieved = _step(RC, cfg, ref, orc, g, a, seed=100 + t)
        where = f"stage {stage} step {t}"
        assert torch.equal(r_term, o_term) and torch.equal(r_to, o_to), where
        assert torch.equal(achieved, orc.last_achieved), where
        assert torch.equal(r_rew, o_rew), f"{where}: reward differs by {(r_rew - o_rew).abs().max():.3e}"
        for i, name in enumerate(ref.reward_manager._term_names):
            assert torch.equal(ref.reward_manager._step_reward[:, i], orc.step_reward[:, orc.reward_term_names.index(name)]), (where, name)
        for k in ("policy", "critic", "auxiliary"):
            assert torch.equal(r_obs[k], o_obs[k]), f"{where}: obs[{k}] differs by {(r_obs[k] - o_obs[k]).abs().max():.3e}"
        _assert_state_equal(ref, orc, where)
        n_reset += int((r_term | r_to).sum())
        n_term += int(r_term.sum())
        n_pass += int(achieved.sum())
        if (r_term | r_to).any():                              # reset log (extras["log"]) of the step that reset
            rl, ol = r_ex["log"], o_ex["log"]
            assert float(rl["Curriculum/terrain_levels"]) == float(ol["Curriculum/terrain_levels"]), where
            assert float(rl["Metrics/next_gate_pose/accumulate_gates"]) == float(ol["Metrics/next_gate_pose/accumulate_gates"]), where
            for name in ("action_rate", "avg_lin_spd", "avg_ang_spd"):     # commands.py:258-260, logged by CommandTerm.reset
                assert float(rl["Metrics/next_gate_pose/" + name]) == float(ol["Metrics/next_gate_pose/" + name]), (where, name)
            for name in ref.reward_manager._term_names:
                assert float(rl["Episode_Reward/" + name]) == float(ol["Episode_Reward/" + name]), (where, name)
            for name in ref.loss_manager._term_names:            # LossManager.reset (loss_manager.py:71-78): the keys exist in every mode; the
                assert "Episode_Loss/" + name in rl and "Episode_Loss/" + name in ol      # harness runs the reference with the differentiable flag
                if cfg.is_differentiable_physics:                                           # on, so values compare in that mode only
                    assert float(rl["Episode_Loss/" + name]) == float(ol["Episode_Loss/" + name]), (where, name)
                else:
                    assert float(ol["Episode_Loss/" + name]) == 0.0
            if cfg.noise_curriculum:
                assert float(rl["Curriculum/command_noise_level"]) == float(ol["Curriculum/command_noise_level"]), where
    assert n_reset >= N and n_pass > 50, (n_reset, n_pass)    # the scenario exercised what it claims to
    if stage == 0:
        assert n_term > 0, "no out-of-bound termination fired"


def test_bptt_losses_and_gradient_match_reference_env():
    """Differentiable closure: the reference's LossManager terms and torch.autograd through its own step over an 8-step window,
    with time-out resets inside the window (per-env graph cuts) -- losses bit-exact, d mean(loss) / d action to fp32 round-off.
    The reference's reset writes thr_est_error (QD/mdp/diff_action.py:233) and the drag vectors (droneDynamics.py:54-57) in place
    after autograd saved them (:175, droneDynamics.py:124), so its own backward raises when an env resets mid-window; the test hands
    it fresh copies just before _reset_idx (same values), through the recorder's pre-reset call of the reference step (:236)."""
    N, H = 48, 8
    RC, cfg, ref, orc, g = _make(0, N, seed=5, diff=True)
    _reset(RC, ref, orc, g, seed=11)
    _teleport(ref, orc, g, frac=1.0, radius=1.0)
    ep = torch.randint(cfg.max_episode_length - 12, cfg.max_episode_length + 20, (N,), generator=g) - 20
    ref.episode_length_buf[:] = ep
    orc.episode_length_buf[:] = ep
    ref.detach()                                                # manager_based_diff_rl_env.py:412-416
    orc.detach()
    term = ref.action_manager.get_term(TERM)
    dyn = term.drone_dynamics

    def fresh_copies(env_ids):
        term.thr_est_error = term.thr_est_error.clone()
        dyn.drag_coeffs, dyn.h_force_drag_coeffs = dyn.drag_coeffs.clone(), dyn.h_force_drag_coeffs.clone()
    ref.recorder_manager.pre_reset_hook = fresh_copies
    acts_r = [(torch.randn(N, 4, generator=g) * 0.5).requires_grad_(True) for _ in range(H)]
    acts_o = [a.detach().clone().requires_grad_(True) for a in acts_r]
    lr, lo, n_reset = [], [], 0
    for t in range(H):
        rnd = torch.zeros(N, L_.RND_STRIDE)
        rnd[:, L_.RND_LEVEL] = torch.rand(N, generator=g)
        ref.scene.terrain.pending_level_u = rnd[:, L_.RND_LEVEL].clone()
        torch.manual_seed(300 + t)
        r_ex = ref.step(acts_r[t])[4]
        reset_ids = ref.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        n_reset += len(reset_ids)
        torch.manual_seed(300 + t)
        RC.replay_reset_draws(rnd, reset_ids, cfg.add_cmd_noise)
        RC.replay_pass_draws(rnd, ref.command_manager.last_achieved, cfg.add_cmd_noise)
        RC.replay_obs_draws(rnd)
        o_ex = orc.step(acts_o[t], rnd)[4]
        assert torch.equal(r_ex["losses"].detach(), o_ex["losses"].detach()), t
        assert torch.equal(r_ex["aligned_states"].detach(), o_ex["aligned_states"].detach()), t
        assert torch.equal(r_ex["nominal_states"].detach(), o_ex["nominal_states"].detach()), t
        assert torch.equal(ref.loss_manager._step_loss, o_ex["loss_terms"]), t
        assert r_ex["log_losses"] == o_ex["log_losses"], t                      # manager_based_diff_rl_env.py:257 (naive_train.py:173)
        for i, name in enumerate(ref.loss_manager._term_names):                 # LossManager episode sums, logged as Episode_Loss/* on reset
            assert torch.equal(ref.loss_manager._episode_sums[name], orc.loss_episode_sums[:, i]), (t, name)
            if len(reset_ids):
                assert float(r_ex["log"]["Episode_Loss/" + name]) == float(o_ex["log"]["Episode_Loss/" + name]), (t, name)
        lr.append(r_ex["losses"])
        lo.append(o_ex["losses"])
    assert 0 < n_reset < N * H // 2, n_reset
    torch.stack(lr).mean().backward()                           # S/diff_rl/algorithms/bptt.py:38-44
    torch.stack(lo).mean().backward()
    zero = torch.zeros(N, 4)
    gr = torch.stack([a.grad if a.grad is not None else zero for a in acts_r])
    go = torch.stack([a.grad if a.grad is not None else zero for a in acts_o])
    assert gr.abs().max() > 0
    assert float((gr - go).abs().max() / gr.abs().max()) < 1e-6


def test_closure_term_parameters_are_the_reference_cfg():
    """The values oracle/ref_closure.py wires into the reference's terms appear verbatim in QD/racing_ctbr_env.py."""
    src = open(ref_modules.REF_ROOT + "/" + ref_modules._QD + "/racing_ctbr_env.py").read()
    for needle in ('resampling_time_range=(20.0, 20.0)', 'consecutive_commands=True', 'pos_x=(-0.1, 0.1)', 'yaw=(-0.1, 0.1)', 'pos_x=(-0.5, 0.5)',
                   'add_noise= (STAGE != 0)', 'update_threshold=0.35', 'random_drag=True', 'action_lag=1',
                   'func=mdp.modified_base_lin_vel, params={"add_noise": True}', 'func=mdp.base_orientation_r, params={"add_noise": True}',
                   'func=mdp.modified_generated_commands, params={"command_name": "next_gate_pose"}',
                   'func=mdp.modified_generated_commands_gt, params={"command_name": "next_gate_pose"}',
                   'func=mdp.modified_last_action, params={"action_name": "force_torque"}', 'func=mdp.cross_obs, params={"reward_name": "success_cross"}',
                   '"z":(-0.5, 0.5)', '"roll": (-0.2, 0.2)', '"yaw": (-0.7, 0.7)', '"yaw": (-0.1, 0.1)',
                   'func=mdp.out_of_bound,params={"bounds":(0.00, 10.0)}', 'func=mdp.bad_pose', 'func=mdp.time_out, time_out=True',
                   '"move_on_threshold": 3', '"move_down_threshold": 2', '"enhance_threshold": 4', '"decay_threshold": 3',
                   '"enhance_percent": 0.02', '"decay_percent": 0.03', 'func=mdp.progress_reward_mine', 'weight=-0.02 if STAGE == 0 else -0.1',
                   'weight=-0.01 if STAGE == 0 else -0.05', 'func=mdp.perception_reward', 'weight=10.0 if STAGE == 0 else 20.0',
                   'func=mdp.penalize_bad_pose', 'weight=-30.0', 'func=mdp.racing_target_diff', 'func=mdp.racing_vel_diff', 'weight=0.05',
                   'func=mdp.racing_falling_diff', 'weight=0.5', 'self.episode_length_s = 6.0 if STAGE != 2 else 8.0'):
        assert needle in src, needle


@pytest.mark.parametrize("N", [1, 2])
def test_single_and_two_env_runs_match_reference_env(N):
    """With one env every reset is a reset of ALL envs -- the branch the reference's dynamics / controller take for
    ``len(idx) == num_envs`` (droneDynamics.py:59-66, controller_diff.py:147-152) -- and two envs mix both branches."""
    RC, cfg, ref, orc, g = _make(1, N, seed=40 + N)
    _reset(RC, ref, orc, g, seed=3)
    ep = torch.full((N,), cfg.max_episode_length - 9)
    ep[-1] -= 5
    ref.episode_length_buf[:] = ep
    orc.episode_length_buf[:] = ep
    n_reset = 0
    for t in range(60):
        if t % 5 == 4:
            _teleport(ref, orc, g, frac=1.0)
        a = torch.randn(N, 4, generator=g) * (3.0 if t % 7 == 0 else 0.5)
        (r_obs, r_rew, r_term, r_to, r_ex), (o_obs, o_rew, o_term, o_to, o_ex), achieved = _step(RC, cfg, ref, orc, g, a, seed=900 + t)
        where = f"N={N} step {t}"
        assert torch.equal(r_term, o_term) and torch.equal(r_to, o_to) and torch.equal(achieved, orc.last_achieved), where
        assert torch.equal(r_rew, o_rew), where
        for k in ("policy", "critic", "auxiliary"):
            assert torch.equal(r_obs[k], o_obs[k]), (where, k)
        _assert_state_equal(ref, orc, where)
        n_reset += int((r_term | r_to).sum())
    assert n_reset >= N

"""Pins oracle/reach_oracle.py's env step against the reference's OWN ``ManagerBasedDiffRLEnv.step`` / ``_reset_idx`` with
``DiffActions`` (LV / PS / CTBR command modes), ``UniformWorldPoseCommand`` and the reach-target reward / loss / observation terms,
executed unmodified over the closure simulator (oracle/ref_closure.py; substitutions R.1-R.5 of oracle/reach_oracle.py).
The reference's DiffActions cannot be constructed in the LV / PS modes (``_get_scale_factor`` repeats a 3-d tensor with two repeat
counts, QD/mdp/diff_action.py:272-280; a test below records the defect): the CTBR task runs the reference unmodified, the LV / PS
tasks run it with that one branch repaired (oracle/ref_closure.py::_repaired_diff_actions) and 3 envs (R.4).  Skipped on the GPU box."""
import pytest
import torch

from generalizableracing_b200 import layout as L_
from generalizableracing_b200.config import ReachTargetCfg
from oracle import ref_modules
from oracle.reach_oracle import OracleReachEnv

pytestmark = pytest.mark.skipif(not ref_modules.available(), reason="reference tree not present")

TERM, CMD = "force_torque", "desired_pos_b"
CASES = {"ctbr": (ReachTargetCfg.ctbr, 40), "ctbr_sim2real": (lambda: ReachTargetCfg.ctbr(sim2real_test=True), 40),
         # LV / PS: the reference's action term with its scale-factor branch repaired (oracle/ref_closure.py::_repaired_diff_actions); 3 envs because
         # the reference's outer-loop controllers only broadcast for 1 or 3 (R.4); dt = 5 ms as in tests/reach_cases.py (the shipped 20 ms diverges)
         "lv_repaired": (lambda: ReachTargetCfg.lv(decimation=1, episode_length_s=0.4, resampling_time=0.15), 3),
         "ps_repaired": (lambda: ReachTargetCfg.ps(decimation=1, episode_length_s=0.3, resampling_time=0.1), 3)}


def _assert_state_equal(ref, orc, where):
    term, cmd, data = ref.action_manager.get_term(TERM), ref.command_manager.get_term(CMD), ref.scene["robot"].data
    d, c = term.drone_dynamics, term.controller
    pairs = {"root_pos_w": (data.root_pos_w, orc.root_pos_w), "root_quat_w": (data.root_quat_w, orc.root_quat_w),
             "root_lin_vel_w": (data.root_lin_vel_w, orc.root_lin_vel_w), "root_ang_vel_w": (data.root_ang_vel_w, orc.root_ang_vel_w),
             "body_ang_acc_w": (data.body_ang_acc_w[:, 0], orc.body_ang_acc_w), "body_lin_acc_w": (data.body_lin_acc_w[:, 0], orc.body_lin_acc_w),
             "dyn.pos": (d.pos, orc.dyn.pos), "dyn.quat": (d.quat, orc.dyn.quat), "dyn.lin_vel_b": (d.lin_vel_b, orc.dyn.lin_vel_b),
             "dyn.ang_vel_b": (d.ang_vel_b, orc.dyn.ang_vel_b), "drag2": (d.drag_coeffs, orc.dyn.drag_coeffs), "drag1": (d.h_force_drag_coeffs, orc.dyn.h_force_drag_coeffs),
             "gross_thrust": (c.gross_thrust, orc.ctrl.gross_thrust),
             "thr_est_error": (term.thr_est_error, orc.thr_est_error), "raw_actions": (term.raw_actions, orc.raw_actions),
             "action": (ref.action_manager.action, orc.action), "prev_action": (ref.action_manager.prev_action, orc.prev_action),
             "action_scale": (term.action_scale, orc.action_scale), "action_offset": (term.action_offset, orc.action_offset),
             "pose_command_w": (cmd.pose_command_w, orc.pose_command_w), "pose_command_b": (cmd.pose_command_b, orc.pose_command_b),
             "time_left": (cmd.time_left, orc.time_left), "position_error": (cmd.metrics["position_error"], orc.metric_position_error),
             "episode_length_buf": (ref.episode_length_buf, orc.episode_length_buf)}
    if hasattr(c, "torque"):                                   # CTBR rate loop only; the LV / PS outer loops filter the thrust alone
        pairs["torque"] = (c.torque, orc.ctrl.torque)
    for name in ref.reward_manager._term_names:
        pairs["episode_sum/" + name] = (ref.reward_manager._episode_sums[name], orc.episode_sums[:, orc.reward_term_names.index(name)])
    for name, (a, b) in pairs.items():
        assert torch.equal(a.detach(), b.detach().to(a.dtype)), f"{where}: {name} differs by {(a.double() - b.double()).abs().max():.3e}"


@pytest.mark.parametrize("case", list(CASES))
def test_reach_step_and_reset_bit_exact_with_reference_env(case):
    from oracle import ref_closure as RC
    make, N = CASES[case]
    cfg = make()
    g = torch.Generator().manual_seed(len(case))
    ref = RC.make_reference_reach_env(cfg, N, seed=50, repair_lv_ps=case.endswith("_repaired"))
    orc = OracleReachEnv(cfg, N)
    assert list(ref.reward_manager._term_names) == list(orc.reward_term_names)
    assert ref.max_episode_length == cfg.max_episode_length
    term, cmd = ref.action_manager.get_term(TERM), ref.command_manager.get_term(CMD)
    ids = torch.arange(N)
    rnd = torch.zeros(N, L_.REACH_RND_STRIDE)
    torch.manual_seed(1)
    ref._reset_idx(ids)
    cmd._update_command()                                       # R.5 (the oracle refreshes the body-frame command on reset)
    r_obs = ref.observation_manager.compute()
    torch.manual_seed(1)
    RC.replay_reach_reset_draws(rnd, ids, cfg.random_drag)
    o_obs, _ = orc.reset(rnd)
    assert torch.equal(r_obs["policy"], o_obs["policy"])
    _assert_state_equal(ref, orc, "after reset")
    T = 160
    ep = torch.randint(max(cfg.max_episode_length - T, 0), cfg.max_episode_length - 1, (N,), generator=g)
    ref.episode_length_buf[:] = ep
    orc.episode_length_buf[:] = ep
    tl = torch.rand(N, generator=g) * T * cfg.step_dt           # stagger the command timers so that they fire inside the test
    cmd.time_left[:] = tl
    orc.time_left[:] = tl
    n_reset = n_timer = n_term = 0
    zero = torch.zeros(N, 4)
    for t in range(T):
        if cfg.sim2real_test:                                   # raw (a_zb, body rates): thrust around hover, a few crashes
            a = torch.randn(N, 4, generator=g) * torch.tensor([3.0, 1.0, 1.0, 0.5]) + torch.tensor([cfg.gravity, 0.0, 0.0, 0.0])
        else:
            a = torch.randn(N, 4, generator=g) * (1.5 if t % 9 == 0 else 0.4)
        a.requires_grad_(cfg.is_differentiable_physics)
        rnd = torch.zeros(N, L_.REACH_RND_STRIDE)
        torch.manual_seed(100 + t)
        r_obs, r_rew, r_term, r_to, r_ex = ref.step(a)
        reset_ids = ref.reset_buf.nonzero(as_tuple=False).squeeze(-1)
        timer_ids = ref.command_manager.last_timer_ids
        torch.manual_seed(100 + t)
        RC.replay_reach_reset_draws(rnd, reset_ids, cfg.random_drag)
        RC.replay_reach_command_draws(rnd, timer_ids, L_.REACH_RND_CMD_TIMER)
        o_obs, o_rew, o_term, o_to, o_ex = orc.step(a.detach().clone().requires_grad_(cfg.is_differentiable_physics), rnd)
        where = f"{case} step {t}"
        assert torch.equal(r_term, o_term) and torch.equal(r_to, o_to), where
        assert torch.equal(timer_ids, orc.last_resampled), where
        assert torch.equal(r_rew, o_rew), f"{where}: reward differs by {(r_rew - o_rew).abs().max():.3e}"
        for i, name in enumerate(ref.reward_manager._term_names):
            assert torch.equal(ref.reward_manager._step_reward[:, i], orc.step_reward[:, i]), (where, name)
        assert torch.equal(r_obs["policy"], o_obs["policy"]), f"{where}: obs differs by {(r_obs['policy'] - o_obs['policy']).abs().max():.3e}"
        assert torch.equal(r_ex["losses"].detach(), o_ex["losses"].detach()), where
        assert torch.equal(ref.loss_manager._step_loss, o_ex["loss_terms"]), where
        _assert_state_equal(ref, orc, where)
        ref.detach()
        orc.detach()
        n_reset += len(reset_ids)
        n_term += int(r_term.sum())
        n_timer += len(timer_ids)
    assert n_reset >= N and n_timer >= max(N // 4, 1), (n_reset, n_timer)


@pytest.mark.parametrize("make", [ReachTargetCfg.lv, ReachTargetCfg.ps])
def test_reference_lv_ps_action_term_cannot_be_constructed(make):
    """QD/mdp/diff_action.py:272-280: ``torch.tensor([[0, 0, 0, 0]])[None].repeat(num_envs, 1)`` raises for every num_envs."""
    from oracle import ref_closure as RC
    with pytest.raises(RuntimeError, match="repeat dims"):
        RC.make_reference_reach_env(make(), 3, seed=0)


def test_reach_closure_term_parameters_are_the_reference_cfg():
    src = open(ref_modules.REF_ROOT + "/" + ref_modules._QD + "/reach_target_lv_env.py").read()
    for needle in ('resampling_time_range=(10.0, 10.0)', 'pos_x=(-2.0,2.0),pos_y=(-2.0,2.0),pos_z=(0.5,2.5),roll=(0.0, 0.0), pitch=(0.0, 0.0), yaw=(0.0, 0.0)',
                   'command_type="LVController", controller_cfg=LVControllerCfg()', 'func=mdp.base_lin_vel', 'func=mdp.base_ang_vel',
                   'func=mdp.last_action, params={"action_name": "force_torque"}', 'func=mdp.base_orientation_q', 'func = mdp.desired_position_b',
                   'func=mdp.reset_root_state_uniform', '"z":(1.0, 2.0),"roll": (-0.5, 0.5), "pitch": (-0.5, 0.5), "yaw": (-3.14, 3.14)',
                   'func=mdp.target_reward', 'func=mdp.orientation_reward', 'func=mdp.move_in_dir', 'params={"threshold": 0.4}', 'func=mdp.action_rate_l2',
                   'func=mdp.reach_target', 'params={"threshold": 0.1}', 'func=mdp.ang_vel_reward', 'func=mdp.body_lin_acc_l2', 'func=mdp.body_ang_acc_l2',
                   'func=mdp.is_terminated', 'weight=-200', 'func=mdp.hover_state', 'params={"threshold":0.2, "ratio": 0.2}', 'func=mdp.target_diff',
                   'func=mdp.smooth_vel_diff', 'weight=0.3', 'params={"ratio": 0.5}', 'self.decimation = 4', 'self.sim.dt = 0.005'):
        assert needle in src, needle

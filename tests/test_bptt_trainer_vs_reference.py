"""Trainer-level pin of the BPTT path: the reference's OWN ``BPTT`` (standalone/diff_rl/algorithms/bptt.py,
algo.py: Adam + CosineAnnealingLR, ``update()`` = mean of the stacked losses -> autograd through its own env step) driving the
reference env over the closure simulator, against this repo's ``BPTT`` driving the kernels (g++ emulation of the CUDA sources) with the
analytic reverse sweep -- same policy class and initial weights, same exploration noise and env draws, the window loop of
standalone/diff_rl/algorithms/runner.py:107-126.  After every iteration the mean loss and the policy weights must agree
(SGD: the weight difference is lr x the gradient difference, measured 7e-9 against moves of 2e-2; Adam: 2e-7).  This test found
that the env handed the policy its ping-pong observation buffer, which the kernels rewrite two steps later behind autograd's back
(first-layer weight gradients of the step-by-step BPTT path were taken against later observations); see RacingVecEnv._grad_safe_obs.
Harness-level workarounds for the reference's own autograd hazards (values unchanged): fresh copies of the tensors
its reset writes in place, and the lag FIFO detached at the window start (the reference keeps the previous window's last action attached
to a graph whose weights the optimiser has since updated in place, so its second ``update()`` raises).  Skipped on the GPU box."""
import copy
import os
import sys
import types

import pytest
import torch

from generalizableracing_b200 import layout as L_
from generalizableracing_b200.algorithms.bptt import BPTT
from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.env import RacingVecEnv
from generalizableracing_b200.modules import BaseModel
from generalizableracing_b200.tracks import figure_eight_track, synthetic_track_table
from oracle import ref_modules as RM
from tests import parity_cases as PC

pytestmark = pytest.mark.skipif(not RM.available(), reason="reference tree not present")


def _load_reference_bptt():
    for name in ("standalone", "standalone.diff_rl"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    model = types.ModuleType("standalone.diff_rl.algorithms.model")          # model.py needs rsl_rl (third party, absent): the policy class is the repo's
    model.BaseModel, model.BaseModelRecurrent = BaseModel, type("BaseModelRecurrent", (), {})
    pkg = types.ModuleType("standalone.diff_rl.algorithms")
    pkg.__path__ = []
    sys.modules["standalone.diff_rl.algorithms"], sys.modules["standalone.diff_rl.algorithms.model"] = pkg, model
    base = os.path.join(RM.REF_ROOT, "standalone/diff_rl/algorithms")
    pkg.AlgoBase = RM._load("standalone.diff_rl.algorithms.algo", os.path.join(base, "algo.py")).AlgoBase
    return RM._load("standalone.diff_rl.algorithms.bptt", os.path.join(base, "bptt.py")).BPTT


@pytest.mark.parametrize("optimizer,trainer,stage", [("SGD", "repo", 0), ("Adam", "repo", 0), ("SGD", "reference", 0), ("Adam", "reference", 0),
                                                     ("SGD", "repo", 1), ("SGD", "repo", 2)])
def test_bptt_training_iterations_match_reference_trainer(emul_lib, optimizer, trainer, stage):
    """trainer = "repo": this repo's BPTT (one-launch window sweep) on the kernels; "reference": the reference's own, unmodified BPTT class
    on the kernels (the drop-in of INTEGRATION.md §3: its ``torch.stack(self.losses).mean().backward()`` runs the chained per-step
    reverse kernels through the autograd-connected ``extras["losses"]``).  Both against the reference's BPTT on the reference's env."""
    from oracle import ref_closure as RC
    N, H, K = 32, 8, 4
    cfg = RacingCfg.for_stage(stage, is_differentiable_physics=True)       # stages 1 / 2: command noise, curricula, bad-pose terminations
    table = figure_eight_track() if stage == 0 else synthetic_track_table()
    g = torch.Generator().manual_seed(21)
    ref, srnd = RC.make_reference_env(cfg, table, N, PC.draw_startup(N, g), seed=4000)
    env = RacingVecEnv(cfg, table, N, device="cpu", rng_mode="dense", startup_rnd=srnd, bptt_horizon=H, _lib=emul_lib)
    term, ter = ref.action_manager.get_term("force_torque"), ref.scene.terrain
    dyn = term.drone_dynamics

    def fresh_copies(env_ids):
        term.thr_est_error = term.thr_est_error.clone()
        dyn.drag_coeffs, dyn.h_force_drag_coeffs = dyn.drag_coeffs.clone(), dyn.h_force_drag_coeffs.clone()
    ref.recorder_manager.pre_reset_hook = fresh_copies
    torch.manual_seed(0)
    pol_r = BaseModel(16, 16, 4, actor_hidden_dims=[128, 128], critic_hidden_dims=[128, 128], activation="elu", init_noise_std=0.3)
    pol_k = copy.deepcopy(pol_r)
    init = [p.detach().clone() for p in pol_r.parameters()]
    hp = dict(max_iterations=K, learning_rate=1e-3 if optimizer == "Adam" else 0.05, schedule="CosineAnnealingLR", device="cpu", optimizer=optimizer)
    alg_r = _load_reference_bptt()(pol_r, **hp)
    alg_k = BPTT(pol_k, env=env, **hp) if trainer == "repo" else _load_reference_bptt()(pol_k, **hp)
    # reset (ManagerBasedRLEnv.reset): _reset_idx(all) + observations
    ids = torch.arange(N)
    rnd = torch.zeros(N, L_.RND_STRIDE)
    torch.manual_seed(1)
    ref._reset_idx(ids)
    obs_r = ref.observation_manager.compute()["policy"]
    torch.manual_seed(1)
    RC.replay_reset_draws(rnd, ids, cfg.add_cmd_noise)
    RC.replay_obs_draws(rnd)
    obs_k = env.reset(rnd)[0]
    ep = torch.randint(cfg.max_episode_length - K * H, cfg.max_episode_length - 1, (N,), generator=g)     # time-outs inside the windows
    ref.episode_length_buf[:] = ep
    env.episode_length_buf = ep
    n_reset = 0
    for it in range(K):
        ref.detach()                                           # runner.py:110
        term.action_buffer = [a.detach() for a in term.action_buffer]
        env.detach()
        for t in range(H):
            torch.manual_seed(10_000 + it * H + t)
            a_r = alg_r.act(obs_r)
            torch.manual_seed(10_000 + it * H + t)
            a_k = alg_k.act(obs_k)
            rnd = torch.zeros(N, L_.RND_STRIDE)
            rnd[:, L_.RND_LEVEL] = torch.rand(N, generator=g)
            ter.pending_level_u = rnd[:, L_.RND_LEVEL].clone()
            torch.manual_seed(20_000 + it * H + t)
            o, rew, terminated, time_outs, ex = ref.step(a_r)
            reset_ids = ref.reset_buf.nonzero(as_tuple=False).squeeze(-1)
            n_reset += len(reset_ids)
            torch.manual_seed(20_000 + it * H + t)
            RC.replay_reset_draws(rnd, reset_ids, cfg.add_cmd_noise)
            RC.replay_pass_draws(rnd, ref.command_manager.last_achieved, cfg.add_cmd_noise)
            RC.replay_obs_draws(rnd)
            obs_r = o["policy"]
            alg_r.process_env_step(ex["losses"], ex["losses_detached"], (terminated | time_outs).long(), rew, ex)   # runner.py:113-126
            obs_k, rew_k, dones_k, ex_k = env.step(a_k, rnd)
            alg_k.process_env_step(ex_k["losses"], ex_k["losses_detached"], dones_k, rew_k, ex_k)
            assert torch.equal(dones_k, (terminated | time_outs).long()), (it, t)
            assert PC.rel_err(obs_r, obs_k) < 10 * PC.REL_TOL_STEP, (it, t)
        _, loss_r = alg_r.update()                             # bptt.py:38-58
        _, loss_k = alg_k.update()
        assert abs(float(loss_r.detach()) - float(loss_k.detach())) < 1e-5 * max(1.0, abs(float(loss_r.detach()))), it
        diffs = torch.cat([(p - q).abs().flatten() for p, q in zip(pol_r.parameters(), pol_k.parameters())])
        moved = torch.cat([(p - q).abs().flatten() for p, q in zip(pol_r.parameters(), init)])
        print(f"stage {stage} {optimizer} / {trainer} trainer on the kernels, iteration {it}: loss {float(loss_r.detach()):.6f} / {float(loss_k.detach()):.6f}, weights moved by <= {float(moved.max()):.2e}, "
              f"differ by <= {float(diffs.max()):.2e}, > 2e-4: {int((diffs > 2e-4).sum())} of {diffs.numel()}")
        if optimizer == "SGD":       # w -= lr g: the weight difference IS the gradient difference (x lr, accumulated over the iterations)
            assert float(diffs.max()) < 1e-4 * float(moved.max()), it
        else:                        # measured 2e-7; Adam's first steps are sign-like, so entries at the fp32 noise floor could move by up to 2 lr
            assert float(diffs.max()) < 2e-5, it
        assert alg_r.optimizer.param_groups[0]["lr"] == pytest.approx(alg_k.optimizer.param_groups[0]["lr"], rel=1e-12)
    assert n_reset > 0


@pytest.mark.parametrize("trainer,task", [("repo", "ctbr"), ("reference", "ctbr"), ("repo", "lv"), ("repo", "ps")])
def test_reach_bptt_training_iterations_match_reference_trainer(emul_lib, trainer, task):
    """The same for the CTBR reach-target task (QD/reach_target_ctbr_env.py; 17-wide observation): the reference's BPTT on the reference's
    env against this repo's / the reference's BPTT on the reach kernels (emulation), SGD so that weight differences are gradient differences."""
    from generalizableracing_b200.config import ReachTargetCfg
    from generalizableracing_b200.reach_env import ReachTargetVecEnv
    from oracle import ref_closure as RC
    H, K = 8, 4
    if task == "ctbr":
        N, cfg = 32, ReachTargetCfg.ctbr(episode_length_s=0.6, resampling_time=0.3)      # 20-step episodes, 10-step command timers
    else:   # the task the reference's BPTT config ships with (LV) and its PS variant: repaired action term, 3 envs, dt = 5 ms (see
            # tests/test_reach_oracle_vs_reference_env.py)
        N, cfg = 3, (ReachTargetCfg.lv if task == "lv" else ReachTargetCfg.ps)(decimation=1, episode_length_s=0.1, resampling_time=0.05)
    g = torch.Generator().manual_seed(33)
    ref = RC.make_reference_reach_env(cfg, N, seed=5000, repair_lv_ps=task != "ctbr")
    env = ReachTargetVecEnv(cfg, N, device="cpu", rng_mode="dense", bptt_horizon=H, _lib=emul_lib)
    term, cmd = ref.action_manager.get_term("force_torque"), ref.command_manager.get_term("desired_pos_b")
    dyn = term.drone_dynamics

    def fresh_copies(env_ids):
        term.thr_est_error = term.thr_est_error.clone()
        dyn.drag_coeffs, dyn.h_force_drag_coeffs = dyn.drag_coeffs.clone(), dyn.h_force_drag_coeffs.clone()
    ref.recorder_manager.pre_reset_hook = fresh_copies
    torch.manual_seed(0)
    pol_r = BaseModel(17, 17, 4, actor_hidden_dims=[128, 128], critic_hidden_dims=[128, 128], activation="elu", init_noise_std=0.3)
    pol_k = copy.deepcopy(pol_r)
    init = [p.detach().clone() for p in pol_r.parameters()]
    hp = dict(max_iterations=K, learning_rate=0.02, schedule="CosineAnnealingLR", device="cpu", optimizer="SGD")
    alg_r = _load_reference_bptt()(pol_r, **hp)
    alg_k = BPTT(pol_k, env=env, **hp) if trainer == "repo" else _load_reference_bptt()(pol_k, **hp)
    ids = torch.arange(N)
    rnd = torch.zeros(N, L_.REACH_RND_STRIDE)
    torch.manual_seed(1)
    ref._reset_idx(ids)
    cmd._update_command()                                      # reach_oracle R.5
    obs_r = ref.observation_manager.compute()["policy"]
    torch.manual_seed(1)
    RC.replay_reach_reset_draws(rnd, ids, cfg.random_drag)
    obs_k = env.reset(rnd)[0]
    n_reset = 0
    for it in range(K):
        ref.detach()
        term.action_buffer = [a.detach() for a in term.action_buffer]
        env.detach()
        for t in range(H):
            torch.manual_seed(10_000 + it * H + t)
            a_r = alg_r.act(obs_r)
            torch.manual_seed(10_000 + it * H + t)
            a_k = alg_k.act(obs_k)
            rnd = torch.zeros(N, L_.REACH_RND_STRIDE)
            torch.manual_seed(20_000 + it * H + t)
            o, rew, terminated, time_outs, ex = ref.step(a_r)
            reset_ids = ref.reset_buf.nonzero(as_tuple=False).squeeze(-1)
            n_reset += len(reset_ids)
            torch.manual_seed(20_000 + it * H + t)
            RC.replay_reach_reset_draws(rnd, reset_ids, cfg.random_drag)
            RC.replay_reach_command_draws(rnd, ref.command_manager.last_timer_ids, L_.REACH_RND_CMD_TIMER)
            obs_r = o["policy"]
            alg_r.process_env_step(ex["losses"], ex["losses_detached"], (terminated | time_outs).long(), rew, ex)
            obs_k, rew_k, dones_k, ex_k = env.step(a_k, rnd)
            alg_k.process_env_step(ex_k["losses"], ex_k["losses_detached"], dones_k, rew_k, ex_k)
            assert torch.equal(dones_k != 0, terminated | time_outs), (it, t)
        _, loss_r = alg_r.update()
        _, loss_k = alg_k.update()
        assert abs(float(loss_r.detach()) - float(loss_k.detach())) < 1e-5 * max(1.0, abs(float(loss_r.detach()))), it
        diffs = torch.cat([(p - q).abs().flatten() for p, q in zip(pol_r.parameters(), pol_k.parameters())])
        moved = torch.cat([(p - q).abs().flatten() for p, q in zip(pol_r.parameters(), init)])
        print(f"reach {task} / {trainer} trainer on the kernels, iteration {it}: loss {float(loss_r.detach()):.6f} / {float(loss_k.detach()):.6f}, "
              f"weights moved by <= {float(moved.max()):.2e}, differ by <= {float(diffs.max()):.2e}")
        # LV / PS: the stiff outer loops amplify fp32 round-off (tests/reach_cases.py::check_bptt allows 2e-3 for the same reason; measured 1.5e-4)
        assert float(diffs.max()) < (1e-4 if task == "ctbr" else 2e-3) * float(moved.max()), it
    assert n_reset > 0

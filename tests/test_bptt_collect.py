"""gr_bptt_collect (fused BPTT window: tensor-core actor + rsample + differentiable env.step with tape) against the step-by-step
window: given the same actions the losses, tape-derived gradients, rewards, dones and final state agree (ids exact, floats at
the oracle tolerance: the inlined step body is contracted differently by nvcc in the two kernels), the actor outputs agree
with the fp32 module within fp16-operand accuracy, and the batched policy backward equals T per-step backwards."""
import pytest
import torch

from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.tracks import figure_eight_track, synthetic_track_table

pytestmark = pytest.mark.gpu


def _setup(N, T, stage, hidden, seed=21, groups=0):
    from generalizableracing_b200.collect import FusedBpttCollector
    from generalizableracing_b200.env import RacingVecEnv
    from generalizableracing_b200.modules import BaseModel
    cfg = RacingCfg.for_stage(stage, is_differentiable_physics=True)
    table = figure_eight_track() if stage == 0 else synthetic_track_table()
    torch.manual_seed(seed)
    pol = BaseModel(16, 16, 4, actor_hidden_dims=hidden, critic_hidden_dims=hidden, activation="lrelu").cuda()
    with torch.no_grad():
        pol.std.copy_(torch.tensor([0.3, 0.2, 0.25, 0.35]))
    envs = []
    for _ in range(2):
        e = RacingVecEnv(cfg, table, N, seed=seed, bptt_horizon=T)
        e._bptt.autograd = False
        e.reset()
        e.episode_length_buf = torch.randint(0, cfg.max_episode_length, (N,), generator=torch.Generator().manual_seed(seed), dtype=torch.int32)
        e.detach()
        envs.append(e)
    col = FusedBpttCollector(envs[0], pol, T, groups_per_cta=groups)
    col.pack()
    return cfg, pol, envs, col


@pytest.mark.parametrize("N,T,stage,hidden,groups", [(512, 16, 0, (256, 128), 0), (1000, 32, 1, (256, 128), 2), (300, 8, 1, (128, 128), 4), (4096, 32, 0, (256, 128), 1)])
def test_fused_window_matches_step_by_step(cuda_lib, N, T, stage, hidden, groups):
    cfg, pol, (env_f, env_u), col = _setup(N, T, stage, hidden, groups=groups)
    obs_u, _ = env_u.get_observations()
    obs_f, critic_f = col.collect()
    torch.cuda.synchronize()
    sigma = pol.std.detach()
    with torch.no_grad():
        for t in range(T):
            assert torch.allclose(col.obs_seq[t], obs_u, rtol=1e-5, atol=1e-5), t
            mu_ref = pol.actor(col.obs_seq[t])
            mu = col.actions[t] - sigma * col.eps_seq[t]
            assert ((mu - mu_ref).abs() <= 1e-2 + 1e-2 * mu_ref.abs()).all(), (t, float((mu - mu_ref).abs().max()))
            assert float((mu - mu_ref).abs().mean()) < 1e-3
            obs_u, rew, dones, ex = env_u.step(col.actions[t])
            assert torch.allclose(col.rewards[t], rew, rtol=1e-5, atol=1e-5), t
            assert torch.equal(col.dones[t].long(), dones), t
            assert torch.allclose(env_f._bptt.loss[t], ex["losses"], rtol=1e-5, atol=1e-5), t
    assert torch.allclose(obs_f, obs_u, rtol=1e-5, atol=1e-5)
    sf, su = env_f.state_dict_view(), env_u.state_dict_view()
    for k in ("gate_id", "accumulate_gates", "terrain_levels", "episode_length", "fresh"):
        assert torch.equal(sf[k], su[k]), k
    for k in ("root_pos_w", "root_quat_w", "root_lin_vel_w", "root_ang_vel_b", "torque", "gross_thrust"):
        assert torch.allclose(sf[k], su[k], rtol=1e-5, atol=1e-5), k
    # the reverse sweep over the fused window's tape == the sweep over the step-by-step tape
    assert env_f._bptt.t == env_u._bptt.t == T
    g_f = env_f._bptt.backward_window().clone()
    g_u = env_u._bptt.backward_window().clone()
    scale = float(g_u.abs().max())
    assert scale > 0 and float((g_f - g_u).abs().max()) < 1e-4 * scale
    # the noise is a fresh standard normal per env, step and component
    eps = col.eps_seq.reshape(-1, 4)
    tol = 5.0 / (N * T) ** 0.5
    assert float(eps.mean().abs()) < tol and float((eps.std() - 1).abs()) < tol
    # batched policy backward == the sum of T per-step backwards on the same (obs, eps, cotangent)
    pol.zero_grad()
    col.policy_backward(g_f)
    batched = [p.grad.clone() for p in pol.actor.parameters()] + [pol.std.grad.clone()]
    pol.zero_grad()
    for t in range(T):
        a = pol.actor(col.obs_seq[t]) + pol.std * col.eps_seq[t]
        torch.autograd.backward([a], [g_f[t]])
    stepwise = [p.grad.clone() for p in pol.actor.parameters()] + [pol.std.grad.clone()]
    for gb, gs in zip(batched, stepwise):
        assert float((gb - gs).abs().max()) <= 2e-4 * float(gs.abs().max()) + 1e-9


def test_fused_bptt_training_runs_and_learns(cuda_lib):
    """AlgoRunner with fused_collection: the loss falls like with the step-by-step runner."""
    from generalizableracing_b200 import make_env
    from generalizableracing_b200.runners import AlgoRunner
    cfgd = {"num_steps_per_env": 32, "max_iterations": 40, "save_interval": 10 ** 9, "empirical_normalization": False,
            "algorithm": {"class_name": "BPTT", "schedule": "CosineAnnealingLR", "optimizer": "AdamW", "learning_rate": 5e-4},
            "policy": {"class_name": "BaseModel", "actor_hidden_dims": [256, 128], "critic_hidden_dims": [256, 128], "activation": "lrelu", "init_noise_std": 0.3}}
    curves = {}
    for mode in ("step_by_step", "fused", "fused+kernel_backward"):
        torch.manual_seed(1)
        env = make_env(num_envs=2048, stage=0, track="figure8", seed=1, differentiable=True, bptt_horizon=32)
        r = AlgoRunner(env, {**cfgd, "fused_collection": mode != "step_by_step", "fused_backward_kernel": mode.endswith("backward")}, log_dir=None)
        h = r.learn(40, init_at_random_ep_len=True)
        curves[mode] = [x["Loss/mean_total_loss"] for x in h]
    for c in curves.values():
        assert sum(c[-5:]) / 5 < 0.9 * sum(c[:5]) / 5, c
    a = sum(curves["step_by_step"][-5:]) / 5
    for mode in ("fused", "fused+kernel_backward"):
        b = sum(curves[mode][-5:]) / 5
        assert abs(a - b) < 0.15 * abs(a), (mode, a, b)

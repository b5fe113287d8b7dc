"""Shared scenarios of the reach-target parity tests: the same checks run against the g++ emulation of the kernels (CPU, here)
and against libgracing.so on the GPU."""
import torch

from generalizableracing_b200 import layout as L
from generalizableracing_b200.config import ReachTargetCfg
from generalizableracing_b200.reach_env import ReachTargetVecEnv
from oracle.reach_oracle import OracleReachEnv


def make_cfg(case: str) -> ReachTargetCfg:
    # decimation=1 (dt = 5 ms): the LV / PS rate loops are unstable at the shipped 20 ms (rate_gain * dt = 4), see DESIGN.md
    if case == "lv":
        return ReachTargetCfg.lv(decimation=1, episode_length_s=0.4)
    if case == "lv_literal":         # the shipped constants: diverges within an episode -> exercises inf / NaN handling and terminations
        return ReachTargetCfg.lv(episode_length_s=0.6)
    if case == "ps":
        return ReachTargetCfg.ps(decimation=1, episode_length_s=0.3, resampling_time=0.1)
    if case == "ctbr":
        return ReachTargetCfg.ctbr(episode_length_s=1.5)
    if case == "ctbr_sim2real":
        return ReachTargetCfg.ctbr(sim2real_test=True, episode_length_s=2.0, resampling_time=0.5)
    raise ValueError(case)


def draws(N, g):
    r = torch.rand(N, L.REACH_RND_STRIDE, generator=g)
    r[:, L.REACH_RND_THR_ERR] = torch.randn(N, generator=g)
    return r


def actions_for(cfg, N, g):
    if cfg.sim2real_test:
        a = torch.randn(N, 4, generator=g) * torch.tensor([3.0, 1.0, 1.0, 1.0]) + torch.tensor([9.0, 0.0, 0.0, 0.0])
    else:
        a = torch.randn(N, 4, generator=g) * 0.6
    return a


def _close(a, b, tol, what):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    fin = torch.isfinite(b)
    assert torch.equal(torch.isfinite(a), fin), f"{what}: finiteness differs"
    err = ((a - b).abs() / (1.0 + b.abs()))[fin]
    assert err.numel() == 0 or float(err.max()) < tol, f"{what}: max rel err {float(err.max()):.3e} >= {tol}"


def check_forward(case, num_envs, steps, device, lib=None, tol=2e-4):
    if lib is None:
        tol = 5e-4          # the sm_100a build uses the fast division / square root (-prec-div=false -prec-sqrt=false)
    cfg = make_cfg(case)
    N = num_envs
    g = torch.Generator().manual_seed(3)
    ora = OracleReachEnv(cfg, N)
    env = ReachTargetVecEnv(cfg, N, device=device, rng_mode="dense", bptt_horizon=steps + 1, _lib=lib)
    env.export_reward_terms = True
    r0 = draws(N, g)
    o_ref, _ = ora.reset(r0)
    o, _ = env.reset(r0)
    _close(o, o_ref["policy"], tol, "reset obs")
    n_reset = n_term = 0
    log_ref = {}
    diverged = torch.zeros(N, dtype=torch.bool)
    for t in range(steps):
        a = actions_for(cfg, N, g)
        r = draws(N, g)
        with torch.no_grad():
            o_ref, rew_ref, term_ref, to_ref, ex_ref = ora.step(a, r)
        o, rew, dones, ex = env.step(a.to(device), r)
        # an env whose state blew up (lv_literal) follows a chaotic trajectory: compared until it diverges, masks afterwards only
        diverged |= ~torch.isfinite(rew_ref) | (ora.root_ang_vel_w.norm(dim=-1) > 1e3) | ~torch.isfinite(ora.root_pos_w).all(-1)
        ok = ~diverged
        assert torch.equal(ex["terminated"].cpu()[ok], term_ref[ok]) and torch.equal(ex["time_outs"].cpu()[ok], to_ref[ok]), f"masks differ at step {t}"
        assert torch.equal(dones.cpu()[ok] != 0, (term_ref | to_ref)[ok])
        _close(o.cpu()[ok], o_ref["policy"][ok], tol, f"obs step {t}")
        _close(rew.cpu()[ok], rew_ref[ok], tol, f"reward step {t}")
        _close(env._outs[env._flip ^ 1]["reward_terms"].cpu()[ok], ora.step_reward[ok], tol, f"reward terms step {t}")
        if cfg.is_differentiable_physics:
            _close(ex["loss_terms"].cpu()[ok], ex_ref["loss_terms"][ok], tol, f"loss terms step {t}")
            _close(ex["losses"].cpu()[ok], ex_ref["losses"][ok], tol, f"loss step {t}")
        done = term_ref | to_ref
        n_reset += int(done.sum())
        n_term += int(term_ref.sum())
        diverged &= ~done                       # a reset starts a clean trajectory
        if t % 16 == 15:
            env.detach()
    sv = env.state_dict_view()
    ok = ~diverged
    _close(sv["root_pos_w"].cpu()[ok], ora.root_pos_w[ok], tol, "pos")
    _close(sv["root_quat_w"].cpu()[ok], ora.root_quat_w[ok], tol, "quat")
    _close(sv["pose_command_w"].cpu()[ok], ora.pose_command_w[ok, :3], tol, "target")
    _close(sv["time_left"].cpu()[ok], ora.time_left[ok], 1e-5, "time_left")
    _close(sv["episode_sums"].cpu()[ok], ora.episode_sums[ok], 5 * tol, "episode sums")
    _close(sv["thr_est_error"].cpu(), ora.thr_est_error, 1e-6, "thr_est_error")
    _close(sv["drag_coeffs"].cpu(), ora.dyn.drag_coeffs, 1e-6, "drag")
    assert torch.equal(sv["episode_length"].cpu().long(), ora.episode_length_buf)
    assert n_reset >= N, "every env should have been reset at least once in this scenario"
    log = ex["log"]
    assert float(log["Episode_Termination/time_out"]) + float(log["Episode_Termination/base_contact"]) >= 1
    return n_reset, n_term


def check_bptt(case, num_envs, horizon, device, lib=None, tol=2e-3):
    """d (mean loss over the window) / d actions: analytic sweep vs autograd through the oracle (fp64 oracle as the reference)."""
    cfg = make_cfg(case)
    if case == "lv":
        cfg.episode_length_s = 0.06          # resets inside the window (adjoint cuts)
    N, T = num_envs, horizon
    g = torch.Generator().manual_seed(11)
    ora = OracleReachEnv(cfg, N, dtype=torch.float64)
    env = ReachTargetVecEnv(cfg, N, device=device, rng_mode="dense", bptt_horizon=T, _lib=lib)
    env._bptt.autograd = False
    r0 = draws(N, g)
    ora.reset(r0.double())
    env.reset(r0)
    # a few warm-up steps so that the window does not start from rest
    for _ in range(5):
        a, r = actions_for(cfg, N, g), draws(N, g)
        with torch.no_grad():
            ora.step(a.double(), r.double())
        env.step(a.to(device), r)
    ora.detach()
    env.detach()
    acts, total = [], 0.0
    for t in range(T):
        a, r = actions_for(cfg, N, g), draws(N, g)
        ad = a.double().requires_grad_(True)
        acts.append(ad)
        _, _, term, to, ex_ref = ora.step(ad, r.double())
        total = total + ex_ref["losses"].sum()
        env.step(a.to(device), r)
    (total / (T * N)).backward()
    ref = torch.stack([x.grad if x.grad is not None else torch.zeros_like(x) for x in acts])
    got = env._bptt.backward_window().cpu().double()
    scale = ref.abs().max()
    assert float(scale) > 0
    err = (got[:-1] - ref[:-1]).abs().max() / scale
    assert float(err) < tol, f"{case}: BPTT gradient max err / max |grad| = {float(err):.3e}"
    assert float(got[-1].abs().max()) == 0.0          # the last action acts in the next window (1-step lag)
    return float(err)


def check_philox(num_envs, steps, device, lib=None):
    """rng_mode='philox' == rng_mode='dense' fed with gr_reach_fill_rand of the same (seed, env, step)."""
    cfg = make_cfg("lv")
    cfg.episode_length_s = 0.05
    N = num_envs
    g = torch.Generator().manual_seed(5)
    e1 = ReachTargetVecEnv(cfg, N, device=device, rng_mode="philox", seed=7, env_id_offset=100, _lib=lib)
    e2 = ReachTargetVecEnv(cfg, N, device=device, rng_mode="dense", seed=7, env_id_offset=100, _lib=lib)
    lib_ = e1._lib

    def fill(step):
        rnd = torch.zeros(N, L.REACH_RND_STRIDE, device=device)
        rc = lib_.gr_reach_fill_rand(rnd.data_ptr(), N, 100, 7, step, e1._stream())
        assert rc == 0
        return rnd
    o1, _ = e1.reset()
    o2, _ = e2.reset(fill(0))
    assert torch.equal(o1, o2)
    for t in range(steps):
        a = (torch.randn(N, 4, generator=g) * 0.5).to(device)
        x1 = e1.step(a)
        x2 = e2.step(a, fill(t + 1))
        assert torch.equal(x1[0], x2[0]) and torch.equal(x1[1], x2[1]) and torch.equal(x1[2], x2[2])
    assert torch.equal(e1.planes, e2.planes)
    rnd = fill(3).cpu()
    u = torch.cat([rnd[:, :13], rnd[:, 14:]], dim=1)
    assert float(u.min()) >= 0.0 and float(u.max()) < 1.0


def check_masked_reset(num_envs, device, lib=None):
    cfg = make_cfg("ctbr")
    N = num_envs
    g = torch.Generator().manual_seed(9)
    ora = OracleReachEnv(cfg, N)
    env = ReachTargetVecEnv(cfg, N, device=device, rng_mode="dense", _lib=lib)
    r0 = draws(N, g)
    ora.reset(r0)
    env.reset(r0)
    for _ in range(6):
        a, r = actions_for(cfg, N, g), draws(N, g)
        with torch.no_grad():
            o_ref = ora.step(a, r)[0]
        o = env.step(a.to(device), r)[0]
    mask = torch.zeros(N, dtype=torch.bool)
    mask[::3] = True
    r = draws(N, g)
    ora._reset_idx(mask.nonzero().squeeze(-1), r)
    ora._update_command()
    o_ref = ora.compute_observations()["policy"]
    o, _ = env.reset(r, mask=mask)
    _close(o, o_ref, 2e-4, "obs after masked reset")
    obs2 = torch.zeros_like(o)
    assert env._lib.gr_reach_observe(env._gcfg, env._state, obs2.data_ptr(), env._stream()) == 0
    assert torch.equal(obs2, o)
    assert torch.equal(env.state_dict_view()["episode_length"].cpu().long(), ora.episode_length_buf)

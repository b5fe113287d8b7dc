"""gr_reach_rollout_fwd / ReachTargetVecEnv.rollout: T steps in one launch must be bit-identical to the T single steps
(rewards, masks, observations, state, BPTT tape / losses / window gradients), in Philox and dense mode, for the three command
modes, across resets and command resampling -- and single steps must continue identically afterwards (the window kernel marks
envs whose read-mostly planes it rewrote: RPL_ANGACC.w = 2)."""
import pytest
import torch

from generalizableracing_b200 import layout as L
from generalizableracing_b200.reach_env import ReachTargetVecEnv
from tests import reach_cases as RC
from tests.conftest import backend_params

pytestmark = pytest.mark.timeout(600)


def _twins(backend, case, N, rng_mode, diff=False, horizon=0):
    device, lib = backend
    cfg = RC.make_cfg(case)
    cfg.is_differentiable_physics = diff
    g = torch.Generator().manual_seed(len(case) * 10 + N)
    envs = [ReachTargetVecEnv(cfg, N, device=device, seed=5, rng_mode=rng_mode, env_id_offset=64, bptt_horizon=horizon, _lib=lib) for _ in range(2)]
    r0 = RC.draws(N, g).to(device) if rng_mode == "dense" else None
    for e in envs:
        e.reset(r0)
        e.episode_length_buf = (torch.arange(N, dtype=torch.int32) * 3) % cfg.max_episode_length
    return cfg, envs, g


def _planes_equal_but_stale_mark(a, b):
    """All state words equal; RPL_ANGACC.w may be 2 (window kernel's 'read-mostly planes rewritten' mark) where the steps leave 0."""
    pa, pb = a.planes.clone(), b.planes.clone()
    wa, wb = pa[:, L.RPL_ANGACC, :, 3], pb[:, L.RPL_ANGACC, :, 3]
    assert bool(((wa == wb) | ((wa == 0) & (wb == 2))).all())
    marked = int((wb == 2).sum())
    pa[:, L.RPL_ANGACC, :, 3] = 0
    pb[:, L.RPL_ANGACC, :, 3] = 0
    assert torch.equal(pa, pb)
    return marked


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
@pytest.mark.parametrize("case,N,rng_mode", [("lv", 130, "philox"), ("ctbr", 96, "dense"), ("ps", 63, "philox"), ("ctbr_sim2real", 64, "dense")])
def test_window_equals_single_steps(backend, case, N, rng_mode):
    cfg, (a, b), g = _twins(backend, case, N, rng_mode)
    dev = a.device
    T = 70
    acts = torch.stack([RC.actions_for(cfg, N, g) for _ in range(T)]).to(dev)
    rnd = torch.stack([RC.draws(N, g) for _ in range(T)]).to(dev) if rng_mode == "dense" else None
    ref = {"reward": [], "dones": [], "terminated": [], "time_outs": [], "obs_seq": []}
    for t in range(T):
        obs, rew, dones, ex = a.step(acts[t], None if rnd is None else rnd[t])
        ref["reward"].append(rew.clone()); ref["dones"].append(dones.bool().clone()); ref["obs_seq"].append(obs.clone())
        ref["terminated"].append(ex["terminated"].clone()); ref["time_outs"].append(ex["time_outs"].clone())
    out = b.rollout(acts, rnd, record_obs=True)
    for k, v in ref.items():
        assert torch.equal(torch.stack(v), out[k]), k
    assert int(out["dones"].sum()) > N // 2
    assert torch.equal(out["obs"], a.get_observations()[0]) and torch.equal(b.get_observations()[0], a.get_observations()[0])
    marked = _planes_equal_but_stale_mark(a, b)
    assert marked > 0                                   # some env was reset inside the window but not by its last step
    la, lb = a._log_accum.sum(0), b._log_accum.sum(0)
    assert torch.allclose(la, lb, rtol=1e-5, atol=1e-4)
    # single steps continue identically; the first one clears the marks
    for t in range(5):
        act = RC.actions_for(cfg, N, g).to(dev)
        r = RC.draws(N, g).to(dev) if rng_mode == "dense" else None
        xa, xb = a.step(act, r), b.step(act, r)
        for k in range(3):
            assert torch.equal(xa[k], xb[k]), (t, k)
        assert torch.equal(a.planes, b.planes)
    # without recorded observations only the last row is written
    act2 = torch.stack([RC.actions_for(cfg, N, g) for _ in range(3)]).to(dev)
    r2 = torch.stack([RC.draws(N, g) for _ in range(3)]).to(dev) if rng_mode == "dense" else None
    for t in range(3):
        last = a.step(act2[t], None if r2 is None else r2[t])
    o2 = b.rollout(act2, r2)
    assert "obs_seq" not in o2 and torch.equal(o2["obs"], last[0]) and torch.equal(o2["reward"][-1], last[1])


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
@pytest.mark.parametrize("case,N", [("lv", 64), ("ctbr", 100)])
def test_differentiable_window_tape_and_gradients(backend, case, N):
    H = 32
    cfg, (a, b), g = _twins(backend, case, N, "philox", diff=True, horizon=H)
    dev = a.device
    for e in (a, b):
        e._bptt.autograd = False
        e.detach()
    acts = torch.stack([RC.actions_for(cfg, N, g) for _ in range(H)]).to(dev)
    la = torch.stack([a.step(acts[t])[3]["losses"].clone() for t in range(H)])
    for t in range(4):
        b.step(acts[t])
    out = b.rollout(acts[4:])
    assert b._bptt.t == H
    assert torch.equal(la[4:], out["losses"]) and torch.equal(la, b._bptt.loss[:H])
    assert torch.equal(a._bptt.tape[:H], b._bptt.tape[:H]) and torch.equal(a._bptt.loss_terms[:H], b._bptt.loss_terms[:H])
    w = torch.rand(H, N, generator=g).to(dev)
    ga, gb = a._bptt.backward_window(grad_losses=w).clone(), b._bptt.backward_window(grad_losses=w).clone()
    assert torch.equal(ga, gb) and float(ga.abs().max()) > 0
    _planes_equal_but_stale_mark(a, b)
    with pytest.raises(RuntimeError, match="tape capacity"):
        b.rollout(acts[:1])


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_argument_errors(backend):
    cfg, (a, b), g = _twins(backend, "lv", 64, "philox")
    with pytest.raises(ValueError, match="Invalid actions shape"):
        a.rollout(torch.zeros(64, 4, device=a.device))
    with pytest.raises(ValueError, match="rnd must be"):
        a.rollout(torch.zeros(3, 64, 4, device=a.device), torch.zeros(3, 64, 5, device=a.device))
    cfg, (c, d), g = _twins(backend, "lv", 64, "dense")
    with pytest.raises(ValueError, match="needs an explicit rnd"):
        c.rollout(torch.zeros(3, 64, 4, device=c.device))


@pytest.mark.gpu
def test_full_size_window_equals_single_steps(cuda_lib):
    """65,536 envs, LV mode, in-kernel Philox: a 12-step window == 12 steps, bit for bit (up to the stale-prefetch marks)."""
    from generalizableracing_b200.config import ReachTargetCfg
    N, T = 65536, 12
    cfg = ReachTargetCfg.lv(decimation=1, is_differentiable_physics=False)
    a, b = (ReachTargetVecEnv(cfg, N, seed=9) for _ in range(2))
    for e in (a, b):
        e.reset()
        e.episode_length_buf = (torch.arange(N, dtype=torch.int32) * 13) % cfg.max_episode_length
    g = torch.Generator(device="cuda").manual_seed(0)
    acts = torch.randn(T, N, 4, device="cuda", generator=g) * 0.5
    rew, dones = [], []
    for t in range(T):
        o, r, d, _ = a.step(acts[t])
        rew.append(r.clone()); dones.append(d.bool().clone())
    out = b.rollout(acts)
    assert torch.equal(torch.stack(rew), out["reward"]) and torch.equal(torch.stack(dones), out["dones"]) and torch.equal(o, out["obs"])
    _planes_equal_but_stale_mark(a, b)
    assert int(out["dones"].sum()) > 100

"""gr_rollout_fwd / RacingVecEnv.rollout: T env.step() calls in one launch for actions known in advance must be
bit-identical to the T single steps -- rewards, masks, observations of every step, the final state planes, the episode log,
the BPTT tape / losses / window gradients -- in Philox and dense mode, across resets, and leave the env in a state from which
single steps continue identically."""
import pytest
import torch

from generalizableracing_b200 import layout as L
from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.env import RacingVecEnv
from generalizableracing_b200.tracks import figure_eight_track, synthetic_track_table
from tests import parity_cases as PC
from tests.conftest import backend_params

pytestmark = pytest.mark.timeout(600)


def _twins(backend, stage, N, rng_mode, diff=False, horizon=0, stats=True):
    device, lib = backend
    cfg = RacingCfg.for_stage(stage, is_differentiable_physics=diff)
    table = synthetic_track_table() if stage else figure_eight_track()
    g = torch.Generator().manual_seed(stage * 100 + N)
    srnd = PC.draw_startup(N, g)
    envs = [RacingVecEnv(cfg, table, N, device=device, seed=21, rng_mode=rng_mode, startup_rnd=srnd, episode_stats=stats, env_id_offset=96,
                         bptt_horizon=horizon, _lib=lib) for _ in range(2)]
    r0 = PC.draw_rnd(N, g).to(device) if rng_mode == "dense" else None
    for e in envs:
        e.reset(r0)
        e.episode_length_buf = (torch.arange(N, dtype=torch.int32) * 7) % cfg.max_episode_length        # time-outs inside the window
    return cfg, envs, g


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
@pytest.mark.parametrize("stage,N,rng_mode,stats", [(1, 130, "philox", True), (1, 96, "dense", True), (0, 63, "philox", False), (2, 64, "dense", False)])
def test_window_equals_single_steps(backend, stage, N, rng_mode, stats):
    cfg, (a, b), g = _twins(backend, stage, N, rng_mode, stats=stats)
    dev = a.device
    T = 45
    acts = (torch.randn(T, N, 4, generator=g) * 0.6).to(dev)
    rnd = torch.stack([PC.draw_rnd(N, g) for _ in range(T)]).to(dev) if rng_mode == "dense" else None
    ref = {"reward": [], "dones": [], "terminated": [], "time_outs": [], "obs_seq": []}
    for t in range(T):
        obs, rew, dones, ex = a.step(acts[t], None if rnd is None else rnd[t])
        ref["reward"].append(rew.clone()); ref["dones"].append(dones.bool().clone()); ref["obs_seq"].append(obs.clone())
        ref["terminated"].append(ex["terminated"].clone()); ref["time_outs"].append(ex["time_outs"].clone())
    out = b.rollout(acts, rnd, record_obs=True)
    for k, v in ref.items():
        assert torch.equal(torch.stack(v), out[k]), k
    assert int(out["dones"].sum()) > 0
    assert torch.equal(a.planes, b.planes)
    # episode log: the same addends, accumulated with float atomics in a different order (step-major vs env-major)
    la, lb = a._log_accum.sum(0), b._log_accum.sum(0)
    assert torch.equal(la[[0, 1, 8, 9]], lb[[0, 1, 8, 9]]) and torch.allclose(la, lb, rtol=1e-5, atol=1e-4)
    oa, ea = a.get_observations()
    ob, eb = b.get_observations()
    assert torch.equal(oa, ob) and torch.equal(out["obs"], oa)
    assert torch.equal(ea["observations"]["critic"], eb["observations"]["critic"]) and torch.equal(ea["observations"]["auxiliary"], eb["observations"]["auxiliary"])
    # single steps continue identically after a window (the read-mostly planes were rewritten without per-step prefetch flags)
    for t in range(6):
        act = (torch.randn(N, 4, generator=g) * 0.6).to(dev)
        r = PC.draw_rnd(N, g).to(dev) if rng_mode == "dense" else None
        xa, xb = a.step(act, r), b.step(act, r)
        for k in range(3):
            assert torch.equal(xa[k], xb[k]), (t, k)
    assert torch.equal(a.planes, b.planes)
    # a window of one step is a step
    act = (torch.randn(1, N, 4, generator=g) * 0.6).to(dev)
    r = PC.draw_rnd(N, g).to(dev)[None] if rng_mode == "dense" else None
    xa = a.step(act[0], None if r is None else r[0])
    ob1 = b.rollout(act, r)
    assert torch.equal(xa[0], ob1["obs"]) and torch.equal(xa[1], ob1["reward"][0]) and torch.equal(a.planes, b.planes)


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
@pytest.mark.parametrize("stage,N", [(0, 64), (1, 100)])
def test_differentiable_window_tape_and_gradients(backend, stage, N):
    H = 24
    cfg, (a, b), g = _twins(backend, stage, N, "philox", diff=True, horizon=H)
    dev = a.device
    for e in (a, b):
        e._bptt.autograd = False
        e.detach()
    acts = (torch.randn(H, N, 4, generator=g) * 0.5).to(dev)
    la = torch.stack([a.step(acts[t])[3]["losses"].clone() for t in range(H)])
    for t in range(5):                                   # a window may start with single steps and go on as one launch
        b.step(acts[t])
    out = b.rollout(acts[5:])
    assert b._bptt.t == H
    assert torch.equal(la[5:], out["losses"]) and torch.equal(la, b._bptt.loss[:H])
    assert torch.equal(a._bptt.tape[:H], b._bptt.tape[:H]) and torch.equal(a._bptt.loss_terms[:H], b._bptt.loss_terms[:H])
    w = torch.rand(H, N, generator=g).to(dev)
    ga, gb = a._bptt.backward_window(grad_losses=w).clone(), b._bptt.backward_window(grad_losses=w).clone()
    assert torch.equal(ga, gb) and float(ga.abs().max()) > 0
    assert torch.equal(a.planes, b.planes)
    with pytest.raises(RuntimeError, match="tape capacity"):
        b.rollout(acts[:1])


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_argument_errors(backend):
    cfg, (a, b), g = _twins(backend, 1, 64, "philox")
    with pytest.raises(ValueError, match="Invalid actions shape"):
        a.rollout(torch.zeros(64, 4, device=a.device))
    with pytest.raises(ValueError, match="Invalid actions shape"):
        a.rollout(torch.zeros(3, 63, 4, device=a.device))
    with pytest.raises(ValueError, match="rnd must be"):
        a.rollout(torch.zeros(3, 64, 4, device=a.device), torch.zeros(2, 64, L.RND_STRIDE, device=a.device))
    cfg, (c, d), g = _twins(backend, 1, 64, "dense")
    with pytest.raises(ValueError, match="needs an explicit rnd"):
        c.rollout(torch.zeros(3, 64, 4, device=c.device))


@pytest.mark.gpu
def test_full_size_window_equals_single_steps(cuda_lib):
    """BASELINE C4 size (65,536 envs, complex track table, stats on, in-kernel Philox): a 12-step window == 12 steps, bit for bit."""
    from generalizableracing_b200.track_gen import generate_track_table, racing_complex_cfg
    N, T = 65536, 12
    cfg, table = RacingCfg.for_stage(1), generate_track_table(racing_complex_cfg())
    a, b = (RacingVecEnv(cfg, table, N, seed=42) for _ in range(2))
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
    assert torch.equal(a.planes, b.planes)
    assert int(out["dones"].sum()) > 1000
    q = b.read_plane(L.PL_QUAT)
    assert float((q.norm(dim=-1) - 1).abs().max()) < 1e-4

"""Size-independent properties of the path (checked on the emulation here and on the B200): the reverse sweep is linear in the loss
cotangents, windows compose (rollout(T1) ; rollout(T2) == rollout(T1 + T2)), and env shards reproduce the unsharded run bit for
bit because the in-kernel Philox stream is keyed by the GLOBAL env id."""
import pytest
import torch

from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.env import RacingVecEnv, default_terrain_types
from generalizableracing_b200.tracks import synthetic_track_table
from tests import parity_cases as PC
from tests.conftest import backend_params

pytestmark = pytest.mark.timeout(600)


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_reverse_sweep_is_linear_in_the_loss_cotangents(backend):
    N, H = 96, 20
    cfg, table, orc, env, g = PC.make_pair(backend, stage=1, N=N, seed=4, diff=True, horizon=H)
    dev = env.device
    env.reset(PC.draw_rnd(N, g).to(dev))
    env.episode_length_buf = torch.randint(0, cfg.max_episode_length, (N,), generator=g)
    env._bptt.autograd = False
    env.detach()
    acts = (torch.randn(H, N, 4, generator=g) * 0.5).to(dev)
    rnd = torch.stack([PC.draw_rnd(N, g) for _ in range(H)]).to(dev)
    env.rollout(acts, rnd)
    w1, w2 = torch.rand(H, N, generator=g).to(dev), torch.randn(H, N, generator=g).to(dev)
    g1 = env._bptt.backward_window(grad_losses=w1).clone()
    g2 = env._bptt.backward_window(grad_losses=w2).clone()
    g12 = env._bptt.backward_window(grad_losses=w1 + 2.0 * w2).clone()
    ref = g1 + 2.0 * g2
    assert float(ref.abs().max()) > 0
    assert float((g12 - ref).abs().max() / ref.abs().max()) < 1e-5
    assert torch.equal(env._bptt.backward_window(grad_losses=w1), g1)                 # and deterministic
    uniform = env._bptt.backward_window().clone()                                     # BPTT.update's 1/(T*N) == explicit weights
    explicit = env._bptt.backward_window(grad_losses=torch.full((H, N), 1.0 / (H * N), device=dev))
    assert float((uniform - explicit).abs().max()) <= 1e-7 * max(1.0, float(uniform.abs().max()))


def _philox_env(backend, N, off, n_glob, types, **kw):
    device, lib = backend
    return RacingVecEnv(RacingCfg.for_stage(1), synthetic_track_table(), N, device=device, seed=77, env_id_offset=off, global_num_envs=n_glob,
                        terrain_types=types, _lib=lib, **kw)


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_windows_compose_and_shards_reproduce_the_whole(backend):
    N, T1, T2 = 192, 9, 14
    table = synthetic_track_table()
    types = default_terrain_types(N, table.num_types)
    whole = _philox_env(backend, N, 0, N, types)
    split = _philox_env(backend, N, 0, N, types)
    halves = [_philox_env(backend, N // 2, k * (N // 2), N, types[k * (N // 2):(k + 1) * (N // 2)]) for k in range(2)]
    dev = whole.device
    ep = (torch.arange(N, dtype=torch.int32) * 11) % whole.cfg.max_episode_length
    for e in (whole, split):
        e.reset()
        e.episode_length_buf = ep
    for k, h in enumerate(halves):
        h.reset()
        h.episode_length_buf = ep[k * (N // 2):(k + 1) * (N // 2)]
    g = torch.Generator().manual_seed(2)
    acts = (torch.randn(T1 + T2, N, 4, generator=g) * 0.6).to(dev)
    a = whole.rollout(acts, record_obs=True)
    b1, b2 = split.rollout(acts[:T1], record_obs=True), split.rollout(acts[T1:], record_obs=True)
    for k in ("reward", "dones", "obs_seq"):
        assert torch.equal(a[k], torch.cat([b1[k], b2[k]])), k
    assert torch.equal(whole.planes, split.planes) and int(a["dones"].sum()) > 0
    outs = [h.rollout(acts[:, k * (N // 2):(k + 1) * (N // 2)].contiguous(), record_obs=True) for k, h in enumerate(halves)]
    for k in ("reward", "dones", "obs_seq"):
        assert torch.equal(a[k], torch.cat([outs[0][k], outs[1][k]], dim=1)), k
    assert torch.equal(a["obs"], torch.cat([outs[0]["obs"], outs[1]["obs"]]))

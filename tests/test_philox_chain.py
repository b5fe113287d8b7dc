"""Random-number parity chain (SURVEY.md §7 'RNG parity'): numpy Philox4x32-10 == gr_fill_rand bits; the in-kernel
Philox path == the dense path fed by gr_fill_rand (bit-exact, same device functions); dense path == oracle
(tests/test_env_parity.py)."""
import ctypes as C

import numpy as np
import pytest
import torch

from generalizableracing_b200 import layout as L_
from oracle import philox as PH
from tests.conftest import backend_params
from tests import parity_cases as PC


def _fill(backend, N, off, seed, step):
    device, lib = backend
    out = torch.zeros(N, L_.RND_STRIDE, device=device)
    if lib is None:
        from generalizableracing_b200 import _lib as B
        B.check(B.load().gr_fill_rand(out.data_ptr(), N, off, seed, step, torch.cuda.current_stream().cuda_stream), "fill")
        torch.cuda.synchronize()
    else:
        lib.gr_fill_rand(out.data_ptr(), N, off, seed, step, None)
    return out.cpu()


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_fill_rand_matches_numpy_philox(backend):
    N, off, seed, step = 257, 1000, 0x1234_5678_9ABC_DEF0, 77
    got = _fill(backend, N, off, seed, step).numpy()
    exp = PH.rnd_rows(np.arange(off, off + N), seed, step)
    assert np.array_equal(got[:, 8:], exp[:, 8:])                     # uniforms: bit-exact (24-bit integers / 2^24)
    assert np.abs(got[:, :8] - exp[:, :8]).max() < 2e-5               # Box-Muller through MUFU vs libm
    assert got[:, 8:].min() >= 0.0 and got[:, 8:].max() < 1.0
    # known-answer vector of Philox4x32-10 (Random123 kat_vectors: counter = key = 0)
    assert PH.philox4x32_10(np.zeros((1, 4), np.uint32), np.zeros((1, 2), np.uint32))[0].tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_statistics_of_the_draws(backend):
    r = _fill(backend, 20000, 0, 42, 3)
    n, u = r[:, :8].double(), r[:, 8:].double()
    assert abs(float(n.mean())) < 0.01 and abs(float(n.std()) - 1.0) < 0.01
    assert abs(float(u.mean()) - 0.5) < 0.005 and abs(float(u.var()) - 1 / 12) < 0.002
    c = np.corrcoef(r.numpy().T)
    assert np.abs(c - np.eye(L_.RND_STRIDE)).max() < 0.04


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_philox_mode_equals_dense_mode(backend):
    """Two envs with identical startup state: one draws in-kernel, the other is fed gr_fill_rand tensors."""
    from generalizableracing_b200.config import RacingCfg
    from generalizableracing_b200.env import RacingVecEnv
    from generalizableracing_b200.tracks import synthetic_track_table
    device, lib = backend
    N, seed, off = 200, 99, 4096
    cfg, table = RacingCfg.for_stage(1), synthetic_track_table()
    g = torch.Generator().manual_seed(0)
    srnd = PC.draw_startup(N, g)
    kw = dict(device=device, seed=seed, env_id_offset=off, global_num_envs=8192, startup_rnd=srnd, _lib=lib)
    ea = RacingVecEnv(cfg, table, N, rng_mode="philox", **kw)
    eb = RacingVecEnv(cfg, table, N, rng_mode="dense", **kw)
    oa, _ = ea.reset()
    ob, _ = eb.reset(_fill(backend, N, off, seed, 0))
    assert torch.equal(oa.cpu(), ob.cpu())
    ep = torch.randint(0, cfg.max_episode_length, (N,), generator=g)
    ea.episode_length_buf = ep
    eb.episode_length_buf = ep
    for t in range(1, 60):
        a = torch.randn(N, 4, generator=g) * 0.7
        if t % 6 == 0:                                    # exercise the gate-pass draws too
            sv = ea.state_dict_view()
            tbl = torch.tensor(table.gate_pose[..., :3])
            org = torch.tensor(table.terrain_origins)
            ty, lv, gi = sv["terrain_types"].cpu().long(), sv["terrain_levels"].cpu().long(), sv["gate_id"].cpu().long()
            pos = tbl[ty, lv, gi] + org[lv, ty] + 0.1
            ea.write_plane(L_.PL_POS, slice(0, 3), pos)
            eb.write_plane(L_.PL_POS, slice(0, 3), pos)
        ra = ea.step(a.to(device))
        rb = eb.step(a.to(device), _fill(backend, N, off, seed, t))
        for x, y in zip(ra[:3], rb[:3]):
            assert torch.equal(x.cpu(), y.cpu())
        assert torch.equal(ra[3]["observations"]["critic"].cpu(), rb[3]["observations"]["critic"].cpu())
    assert torch.equal(ea.planes.cpu(), eb.planes.cpu())

"""Kernel-vs-oracle parity of the env step / reset path, on the same seeded inputs.

Each test runs twice: ``emul`` = the kernel sources compiled by g++ and executed on the CPU (logic check, runs
without a GPU), ``cuda`` = libgracing.so through the C ABI on the B200 (`-m gpu`)."""
import pytest
import torch

from tests.conftest import backend_params
from tests import parity_cases as PC

pytestmark = pytest.mark.timeout(600)


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_c1_fixed_track_rollout(backend):
    """BASELINE config C1: 64 envs, fixed figure-8 track, STAGE 0, free-running rollout with N(0, 0.5^2) actions."""
    cfg, table, orc, env, g = PC.make_pair(backend, stage=0, N=64, seed=0)
    r0 = PC.draw_rnd(64, g)
    o_obs, _ = orc.reset(r0)
    k_obs, kex = env.reset(r0.to(env.device))
    assert PC.rel_err(o_obs["policy"], k_obs) < PC.REL_TOL_STEP
    assert PC.rel_err(o_obs["critic"], kex["observations"]["critic"]) < PC.REL_TOL_STEP
    steps = 1000 if backend[0] != "cpu" else 400
    st = PC.run_rollout(orc, env, g, steps)
    print("C1", st)
    assert st["resets"] > 0
    assert st["mask_mismatch"] == 0 and st["int_mismatch"] == 0 and st["aux_mismatch"] == 0
    # free-running: error accumulates over an episode (<= 200 steps); reported, bounded at 10x the per-step tolerance
    for k in ("obs", "critic", "reward", "terms", "state"):
        assert st[k] < 10 * PC.REL_TOL_STEP, (k, st[k])


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
@pytest.mark.parametrize("stage", [0, 1, 2])
def test_gate_passing_and_curriculum(backend, stage):
    """Gate-pass detection, success_cross, accumulate_gates, terrain / noise curricula, resample noise: envs are
    teleported next to their gate every 7 steps so these rare paths fire thousands of times."""
    N = 96
    cfg, table, orc, env, g = PC.make_pair(backend, stage=stage, N=N, seed=3 + stage)
    r0 = PC.draw_rnd(N, g)
    orc.reset(r0)
    env.reset(r0.to(env.device))
    ep = torch.randint(0, cfg.max_episode_length, (N,), generator=g)
    orc.episode_length_buf[:] = ep
    env.episode_length_buf = ep
    st = PC.run_rollout(orc, env, g, 260, teleport_every=7)
    print("gates", stage, st)
    assert st["gates"] > 200 and st["resets"] > 50
    assert st["mask_mismatch"] == 0 and st["int_mismatch"] == 0 and st["aux_mismatch"] == 0
    for k in ("obs", "critic", "reward", "terms", "state"):
        assert st[k] < 10 * PC.REL_TOL_STEP, (k, st[k])


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_ragged_sizes(backend):
    """env counts that are not multiples of the block / chunk size, and a single env."""
    for N in (1, 63, 130):
        cfg, table, orc, env, g = PC.make_pair(backend, stage=1, N=N, seed=N)
        r0 = PC.draw_rnd(N, g)
        orc.reset(r0)
        env.reset(r0.to(env.device))
        st = PC.run_rollout(orc, env, g, 40, teleport_every=5)
        assert st["mask_mismatch"] == 0 and st["int_mismatch"] == 0
        assert st["obs"] < 10 * PC.REL_TOL_STEP and st["state"] < 10 * PC.REL_TOL_STEP


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_action_fifo_survives_reset(backend):
    """Appendix C.1/C.2: the action FIFO is not cleared on reset while action/prev_action are: the first step of a new
    episode applies the old action and command_rate_penalty sees a zero previous action."""
    N = 32
    cfg, table, orc, env, g = PC.make_pair(backend, stage=0, N=N, seed=11)
    r0 = PC.draw_rnd(N, g)
    orc.reset(r0)
    env.reset(r0.to(env.device))
    ep = torch.full((N,), cfg.max_episode_length - 2)
    orc.episode_length_buf[:] = ep
    env.episode_length_buf = ep
    st = PC.run_rollout(orc, env, g, 6, action_std=1.5)
    assert st["resets"] == N
    assert st["mask_mismatch"] == 0 and st["int_mismatch"] == 0
    assert st["terms"] < PC.REL_TOL_STEP * 10 and st["obs"] < PC.REL_TOL_STEP * 10

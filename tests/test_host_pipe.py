"""gr_host_pipe_* (env.step with HOST buffers): the pipelined H2D -> kernel -> D2H path returns bit-identical results to
RacingVecEnv.step on device tensors, for every pipeline depth, including ragged env counts."""
import pytest
import torch

from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.tracks import synthetic_track_table

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,depth", [(4096, 1), (4096, 2), (1000, 3), (65536, 4)])
def test_step_host_matches_step(cuda_lib, N, depth):
    from generalizableracing_b200.env import RacingVecEnv
    cfg, table = RacingCfg.for_stage(1), synthetic_track_table()
    ref = RacingVecEnv(cfg, table, N, seed=5)
    env = RacingVecEnv(cfg, table, N, seed=5)
    ref.reset()
    env.reset()
    g = torch.Generator().manual_seed(3)
    T = 12
    acts = [(torch.randn(N, 4, generator=g) * 0.5).pin_memory() for _ in range(T)]
    outs = [(torch.empty(N, 16).pin_memory(), torch.empty(N).pin_memory(), torch.empty(N, dtype=torch.int64).pin_memory(),
             torch.empty(N, 16).pin_memory(), torch.empty(N, dtype=torch.bool).pin_memory()) for _ in range(T)]
    tickets = [env.step_host(acts[t], outs[t][0], outs[t][1], outs[t][2], critic_obs=outs[t][3], time_outs=outs[t][4], depth=depth) for t in range(T)]
    for t in range(T):
        obs, rew, dones, ex = ref.step(acts[t].cuda())
        env.wait_host(tickets[t])
        assert torch.equal(outs[t][0], obs.cpu()), t
        assert torch.equal(outs[t][1], rew.cpu()), t
        assert torch.equal(outs[t][2], dones.cpu()), t
        assert torch.equal(outs[t][3], ex["observations"]["critic"].cpu()), t
        assert torch.equal(outs[t][4], ex["time_outs"].cpu()), t
    assert torch.equal(env.planes, ref.planes)
    env.close()


def test_step_host_rejects_device_or_misshaped_tensors(cuda_lib):
    from generalizableracing_b200.env import RacingVecEnv
    env = RacingVecEnv(RacingCfg.for_stage(1), synthetic_track_table(), 64)
    with pytest.raises(ValueError):
        env.step_host(torch.zeros(64, 4, device="cuda"), torch.empty(64, 16), torch.empty(64))
    with pytest.raises(ValueError):
        env.step_host(torch.zeros(64, 4), torch.empty(64, 8), torch.empty(64))


@pytest.mark.parametrize("N", [4096, 1002, 65536])
def test_contiguous_host_buffers_take_the_single_copy_path(cuda_lib, N):
    """env.host_buffers(): obs | reward | int64 dones of a set back to back in one pinned block -> gr_host_pipe_step ships them as ONE
    device->host copy; same bits as the device-tensor step, and nothing outside the three views is written (guard bytes survive)."""
    from generalizableracing_b200.env import RacingVecEnv
    cfg, table = RacingCfg.for_stage(1), synthetic_track_table()
    ref, env = RacingVecEnv(cfg, table, N, seed=8), RacingVecEnv(cfg, table, N, seed=8)
    ref.reset(); env.reset()
    sets = env.host_buffers(3)
    for b in sets:
        assert b["reward"].data_ptr() == b["obs"].data_ptr() + N * 64 and b["dones"].data_ptr() == b["obs"].data_ptr() + N * 68
        b["_block"][N * 76:] = 0x5A
    g = torch.Generator().manual_seed(4)
    tickets, acts = [], []
    for t in range(9):
        b = sets[t % 3]
        if t >= 3:
            env.wait_host(tickets[t - 3])
            obs, rew, dones, _ = ref.step(acts[t - 3].cuda())
            assert torch.equal(b["obs"], obs.cpu()) and torch.equal(b["reward"], rew.cpu()) and torch.equal(b["dones"], dones.cpu()), t
        a = torch.randn(N, 4, generator=g) * 0.5
        acts.append(a)
        b["actions"].copy_(a)
        tickets.append(env.step_host(b["actions"], b["obs"], b["reward"], b["dones"], depth=3))
    for t in range(6, 9):
        env.wait_host(tickets[t])
        obs, rew, dones, _ = ref.step(acts[t].cuda())
        b = sets[t % 3]
        assert torch.equal(b["obs"], obs.cpu()) and torch.equal(b["reward"], rew.cpu()) and torch.equal(b["dones"], dones.cpu()), t
    assert all(bool((b["_block"][N * 76:] == 0x5A).all()) for b in sets)
    assert torch.equal(env.planes, ref.planes)
    env.close()

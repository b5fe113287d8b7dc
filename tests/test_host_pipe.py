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

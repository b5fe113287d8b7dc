"""Multi-GPU plumbing: one process per GPU, envs sharded by contiguous global id ranges, no data-path collective.
The only exchanges are (1) the policy-gradient all-reduce (mean) per optimiser step and (2) three doubles per PPO
iteration to normalise advantages over the GLOBAL batch.  NCCL on GPUs, gloo in the CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def shard_range(global_num_envs: int, rank: int, world_size: int):
    """Contiguous env range of a rank: [rank*N/R, (rank+1)*N/R)."""
    if global_num_envs % world_size:
        raise ValueError("global_num_envs must be divisible by the world size")
    n = global_num_envs // world_size
    return rank * n, n


def allreduce_mean_grads(params, group=None):
    """One flat all-reduce (sum -> mean) over the gradients: equals .mean() over the global batch when every rank's
    loss is a mean over an equally sized shard."""
    _, w = world()
    if w == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(w)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def merge_moments(moments: torch.Tensor, group=None) -> torch.Tensor:
    """(count, mean, M2) of each rank's raw advantages -> moments of the global batch (Chan et al. parallel merge
    written with sums so that one all-reduce suffices): n = sum n_r, mean = sum n_r mean_r / n,
    M2 = sum (M2_r + n_r mean_r^2) - n mean^2."""
    _, w = world()
    if w == 1:
        return moments
    n, mean, m2 = moments[0], moments[1], moments[2]
    pack = torch.stack([n, n * mean, m2 + n * mean * mean]).to(torch.float64)
    dist.all_reduce(pack, op=dist.ReduceOp.SUM, group=group)
    gn = pack[0]
    gmean = pack[1] / gn
    return torch.stack([gn, gmean, pack[2] - gn * gmean * gmean])


def broadcast_module(module: torch.nn.Module, src: int = 0):
    _, w = world()
    if w == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)

"""The policy-gradient all-reduce of env-sharded training as one kernel over NVLink peer memory (csrc/peer_reduce.cu, gr_peer_allreduce).

BASELINE config C5 shards the envs over the GPUs of a node and sums the flat policy-gradient buffer once per optimiser step.  The
buffer is 152 KB and the step around it is a captured graph of short kernels, so the collective is pure latency (NCCL: 16 us at 2
GPUs, 34 us at 8).  Here every rank reads the other ranks' buffers directly between two flag barriers; the buffers live in symmetric
memory (``torch.distributed._symmetric_memory``: allocation + exchange of the peer mappings only -- the kernel is ours).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as B


class PeerAllReduce:
    """``buf`` [n] (symmetric memory: hand it to whatever produces the gradients) -> ``out`` [n] = sum over the ranks of the group, the same
    bits on every rank.  ``launch()`` enqueues the kernel on the current stream (capturable); every rank must call it once per step."""

    def __init__(self, n: int, device, group=None, max_spins: int = 1 << 26):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > B.GR_PEER_MAX_WORLD:
            raise ValueError(f"PeerAllReduce: at most {B.GR_PEER_MAX_WORLD} ranks")
        self.n = (int(n) + 3) // 4 * 4
        self.device = torch.device(device)
        self.buf = symm.empty(self.n, dtype=torch.float32, device=self.device)
        self.flags = symm.empty(2 * B.GR_PEER_MAX_WORLD, dtype=torch.int32, device=self.device)
        self.buf.zero_()
        self.flags.zero_()
        hb = symm.rendezvous(self.buf, group.group_name)
        hf = symm.rendezvous(self.flags, group.group_name)
        self._handles = (hb, hf)
        self._ptrs = torch.tensor([int(p) for p in hb.buffer_ptrs], dtype=torch.int64, device=self.device)
        self._fptrs = torch.tensor([int(p) for p in hf.buffer_ptrs], dtype=torch.int64, device=self.device)
        self.out = torch.zeros(self.n, device=self.device)
        self._misc = torch.zeros(4, dtype=torch.int32, device=self.device)          # epoch | block counter | error | -
        self._arg = B.GrPeerReduce(self._ptrs.data_ptr(), self._fptrs.data_ptr(), self.world, self.rank, self.n, 0, int(max_spins),
                                   self._misc.data_ptr(), self._misc.data_ptr() + 4, self._misc.data_ptr() + 8)
        self._lib = B.load()
        torch.cuda.synchronize(self.device)
        dist.barrier(group)                    # every pad is zero before anybody's first launch

    def launch(self) -> torch.Tensor:
        B.check(self._lib.gr_peer_allreduce(C.byref(self._arg), self.out.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream), "gr_peer_allreduce")
        return self.out

    def failed(self) -> bool:
        """True if a wait gave up (a rank did not show up within max_spins polls).  Reads one int from the device."""
        return bool(int(self._misc[2]))

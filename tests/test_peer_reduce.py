"""gr_peer_allreduce (csrc/peer_reduce.cu) on ONE GPU: the kernel's protocol with every "rank" mapped onto buffers of this device.  The
multi-GPU run itself (symmetric memory, NVLink reads, CUDA-graph replay, against NCCL on 2 and 8 GPUs) is tools/peer_reduce_check.py:
profiles/r2_peer_allreduce_check_{2,8}gpu.json."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _arg(B, bufs, pads, world, rank, n, max_spins, misc):
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
    fptrs = torch.tensor([p.data_ptr() for p in pads], dtype=torch.int64, device="cuda")
    arg = B.GrPeerReduce(ptrs.data_ptr(), fptrs.data_ptr(), world, rank, n, 0, max_spins, misc.data_ptr(), misc.data_ptr() + 4, misc.data_ptr() + 8)
    return arg, (ptrs, fptrs)


def test_one_rank_sum_is_a_copy_and_epochs_advance(cuda_lib):
    from generalizableracing_b200 import _lib as B
    n = 38040
    buf, out = torch.randn(n, device="cuda"), torch.zeros(n, device="cuda")
    pad = torch.zeros(2 * B.GR_PEER_MAX_WORLD, dtype=torch.int32, device="cuda")
    misc = torch.zeros(4, dtype=torch.int32, device="cuda")
    arg, keep = _arg(B, [buf], [pad], 1, 0, n, 1 << 20, misc)
    st = torch.cuda.current_stream().cuda_stream
    for k in range(1, 4):
        buf.mul_(1.5)
        B.check(cuda_lib.gr_peer_allreduce(C.byref(arg), out.data_ptr(), st), "gr_peer_allreduce")
        torch.cuda.synchronize()
        assert torch.equal(out, buf)
        assert misc.tolist()[:3] == [k, 0, 0] and int(pad[0]) == k and int(pad[B.GR_PEER_MAX_WORLD]) == k
    # argument errors come back before anything is launched
    bad = B.GrPeerReduce(arg.peer_bufs, arg.peer_flags, 1, 0, n + 1, 0, 1, arg.epoch, arg.counter, arg.error)
    assert cuda_lib.gr_peer_allreduce(C.byref(bad), out.data_ptr(), st) == -2
    assert cuda_lib.gr_peer_allreduce(None, out.data_ptr(), st) == -1


def test_two_ranks_on_one_device_and_a_missing_rank_gives_up(cuda_lib):
    """Both "ranks" launched on two streams of one device meet at the flag barriers and compute the same sum; a rank whose partner never
    shows up sets its error flag after max_spins polls instead of waiting forever."""
    from generalizableracing_b200 import _lib as B
    n = 4096
    bufs = [torch.randn(n, device="cuda") for _ in range(2)]
    pads = [torch.zeros(2 * B.GR_PEER_MAX_WORLD, dtype=torch.int32, device="cuda") for _ in range(2)]
    outs = [torch.zeros(n, device="cuda") for _ in range(2)]
    miscs = [torch.zeros(4, dtype=torch.int32, device="cuda") for _ in range(2)]
    args = [_arg(B, bufs, pads, 2, r, n, 1 << 22, miscs[r]) for r in range(2)]          # (gives up after ~0.3 s)
    streams = [torch.cuda.Stream() for _ in range(2)]
    torch.cuda.synchronize()
    for it in range(3):
        for r in range(2):
            with torch.cuda.stream(streams[r]):
                B.check(cuda_lib.gr_peer_allreduce(C.byref(args[r][0]), outs[r].data_ptr(), streams[r].cuda_stream), "gr_peer_allreduce")
        torch.cuda.synchronize()
        if it == 0 and (int(miscs[0][2]) or int(miscs[1][2])):
            pytest.skip("the kernels of two streams did not run concurrently here (serialised launches): the two-rank protocol needs co-resident kernels")
        want = bufs[0] + bufs[1]
        assert torch.equal(outs[0], want) and torch.equal(outs[1], want)
        assert miscs[0].tolist()[:3] == [it + 1, 0, 0] and miscs[1].tolist()[:3] == [it + 1, 0, 0]
        bufs[it % 2].mul_(-0.5)
    # rank 0 alone: its partner never announces
    lone_misc = torch.zeros(4, dtype=torch.int32, device="cuda")
    lone_pads = [torch.zeros_like(pads[0]), torch.zeros_like(pads[0])]
    lone, keep = _arg(B, bufs, lone_pads, 2, 0, n, 2000, lone_misc)
    B.check(cuda_lib.gr_peer_allreduce(C.byref(lone), outs[0].data_ptr(), torch.cuda.current_stream().cuda_stream), "gr_peer_allreduce")
    torch.cuda.synchronize()
    assert int(lone_misc[2]) == 1

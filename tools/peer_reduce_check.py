#!/usr/bin/env python
"""gr_peer_allreduce against NCCL on N GPUs of one node (torchrun): same sums (rank-order accumulation: compared to fp32 rounding), the
same bits on every rank, eager and replayed from a CUDA graph; device time of 20 reduces per graph for both."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from generalizableracing_b200.peer import PeerAllReduce
    rank, world = dist.get_rank(), dist.get_world_size()
    n = 38040
    pr = PeerAllReduce(n, dev)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    worst = 0.0
    for it in range(30):
        x = torch.randn(pr.n, device=dev, generator=g) * (1 + it)
        pr.buf.copy_(x)
        ref = x.clone()
        dist.all_reduce(ref)
        out = pr.launch().clone()
        worst = max(worst, float((out - ref).abs().max() / ref.abs().max()))
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        assert all(torch.equal(gathered[0], o) for o in gathered), "ranks disagree"
    assert worst < 1e-6, worst
    assert not pr.failed()
    # ---- captured: 20 reduces per replay, data refreshed by a kernel in between (as the training step does)
    s = torch.cuda.Stream(dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    x = torch.randn(pr.n, device=dev, generator=g)
    acc_p, acc_n = torch.zeros(pr.n, device=dev), torch.zeros(pr.n, device=dev)
    tmp = torch.zeros(pr.n, device=dev)

    def step_peer():
        for k in range(20):
            pr.buf.copy_(x).mul_(1.0 + 0.01 * k)
            acc_p.add_(pr.launch())

    def step_nccl():
        for k in range(20):
            tmp.copy_(x).mul_(1.0 + 0.01 * k)
            dist.all_reduce(tmp)
            acc_n.add_(tmp)
    res = {}
    for name, fn in (("peer", step_peer), ("nccl", step_nccl)):
        with torch.cuda.stream(s):
            fn()
        torch.cuda.synchronize(dev)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            fn()
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize(dev)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            gr.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) * 1e3 / (30 * 20)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name + "_us_per_reduce_incl_copy_mul_add"] = float(t)
    rel = float((acc_p - acc_n).abs().max() / acc_n.abs().max())
    assert rel < 1e-5, rel
    assert not pr.failed()
    # the three element-wise kernels alone
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        pass
    with torch.cuda.graph(gr, stream=s):
        for k in range(20):
            tmp.copy_(x).mul_(1.0 + 0.01 * k)
            acc_n.add_(tmp)
    gr.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        gr.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    res["elementwise_only_us"] = e0.elapsed_time(e1) * 1e3 / (30 * 20)
    res.update(world=world, floats=pr.n, worst_rel_err_vs_nccl=worst, graph_rel_err=rel)
    if rank == 0:
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

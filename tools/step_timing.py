#!/usr/bin/env python
"""Micro-benchmark of gr_step_fwd: block sizes x (cold single launch | graph-replayed rotation | L2-resident)."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from generalizableracing_b200 import _lib as B  # noqa: E402
from generalizableracing_b200 import build as BLD  # noqa: E402
from generalizableracing_b200.config import RacingCfg  # noqa: E402
from generalizableracing_b200.env import RacingVecEnv  # noqa: E402
from generalizableracing_b200.tracks import synthetic_track_table  # noqa: E402


def make_io(e, a, dones="i64"):
    o = e._outs[0]
    io = B.GrStepIO()
    io.action = a.data_ptr()
    io.obs, io.critic_obs, io.aux_obs = o["obs"].data_ptr(), o["critic"].data_ptr(), o["aux"].data_ptr()
    io.reward, io.terminated, io.time_out = o["reward"].data_ptr(), o["terminated"].data_ptr(), o["time_out"].data_ptr()
    if dones == "i64":
        io.dones = o["dones"].data_ptr()
    elif dones == "u8":                      # one byte per env (the int64 view of the wrapper's `.long()` only on request)
        e._dones_u8 = torch.zeros(e.num_envs, dtype=torch.uint8, device=a.device)
        io.dones_u8 = e._dones_u8.data_ptr()
    io.log_accum = e._log_accum.data_ptr()
    return io


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--sets", type=int, default=11)
    ap.add_argument("--stage", type=int, default=1)
    ap.add_argument("--blocks", default="32,64,128,256")
    ap.add_argument("--no-stats", action="store_true")
    ap.add_argument("--pdl", type=int, default=1)
    ap.add_argument("--dones", default="i64", choices=["i64", "u8", "none"], help="what the kernel writes besides terminated / time_out")
    ap.add_argument("--complex", type=int, default=0, help="1: the bench's 20 x 10 x 8 gate table instead of the synthetic one")
    ap.add_argument("--hover", type=int, default=0, help="1: near-hover actions (time-out resets only, ~0.5%/step)")
    args = ap.parse_args()
    BLD.build()
    lib = B.load()
    dev = torch.device("cuda:0")
    cfg = RacingCfg.for_stage(args.stage)
    table = synthetic_track_table()
    if args.complex:
        from generalizableracing_b200.track_gen import generate_track_table, racing_complex_cfg
        table = generate_track_table(racing_complex_cfg())
    N = args.envs
    flush = torch.empty(int(512e6) // 4, device=dev)
    res = []
    for blk in [int(b) for b in args.blocks.split(",")]:
        envs = [RacingVecEnv(cfg, table, N, device=dev, seed=1 + r, episode_stats=not args.no_stats, block_threads=blk, pdl=bool(args.pdl)) for r in range(args.sets)]
        acts = [torch.randn(N, 4, device=dev) * 0.5 for _ in envs]
        if args.hover:
            acts = [torch.randn(N, 4, device=dev) * 0.1 + torch.tensor([-0.3466, 0.0, 0.0, 0.0], device=dev) for _ in envs]
        for e in envs:
            e.reset()
            e.episode_length_buf = torch.randint(0, cfg.max_episode_length, (N,), device=dev, dtype=torch.int32)
        ios = [make_io(e, a, args.dones) for e, a in zip(envs, acts)]
        step = [0]

        def launch(k):
            e, io = envs[k], ios[k]
            rng = B.GrRandom(None, e.seed, step[0] & 0xFFFFFFFF)
            step[0] += 1
            B.check(lib.gr_step_fwd(C.byref(e._gcfg), C.byref(e._track), C.byref(e._state), C.byref(rng), C.byref(io),
                                    torch.cuda.current_stream(dev).cuda_stream), "step")

        for _ in range(3):
            for k in range(len(envs)):
                launch(k)
        torch.cuda.synchronize()
        # (1) cold single launches: flush L2, then time one launch with events
        cold = []
        for it in range(20):
            flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            launch(it % len(envs))
            e1.record()
            torch.cuda.synchronize()
            cold.append(e0.elapsed_time(e1) * 1e3)
        cold.sort()
        # (2) graph: one round over all sets, replayed
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for k in range(len(envs)):
                launch(k)
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for k in range(len(envs)):
                launch(k)
        out = {}
        for reps in (1, 10, 100):
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            out[reps] = e0.elapsed_time(e1) * 1e3 / (reps * len(envs))
        # (3) L2-resident: graph of 10 launches on ONE env set
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            for _ in range(10):
                launch(0)
        g1.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            g1.replay()
        e1.record()
        torch.cuda.synchronize()
        hot = e0.elapsed_time(e1) * 1e3 / 200
        # (4) eager back-to-back launches over the rotation (host-launch bound?)
        e0.record()
        for _ in range(10):
            for k in range(len(envs)):
                launch(k)
        e1.record()
        torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) * 1e3 / (10 * len(envs))
        r = {"dones": args.dones, "hover": args.hover, "reset_rate": float(torch.stack([e._log_accum.sum(0) for e in envs]).sum(0)[0]) / (step[0] * N), "pdl": args.pdl, "block": blk, "cold_single_us_median": cold[len(cold) // 2], "cold_single_us_min": cold[0], "graph_rot_us": out, "l2_resident_us": hot, "eager_rot_us": eager}
        print(json.dumps(r), flush=True)
        res.append(r)
        del envs, ios, acts
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

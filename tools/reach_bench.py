"""Reach-target tasks on the B200: step kernel and BPTT window timings (CUDA events, graph-replayed, HBM-cold rotation).
    python tools/reach_bench.py [--envs 65536] [--json out.json]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from generalizableracing_b200.config import ReachTargetCfg  # noqa: E402
from generalizableracing_b200.reach_env import ReachTargetVecEnv  # noqa: E402

B_FWD = 11 * 32 + 2 * 16 + 16 + 68 + 14          # bytes per env-step (reach_step.cu header)


def time_graph(fn, iters, warm=3):
    dev = torch.device("cuda:0")
    s = torch.cuda.Stream(dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream(dev).wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        fn()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def load_peak():
    try:
        return json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbps"]
    except Exception:
        return 6543.4


def bench_step(cfg, N, R=12, iters=30):
    """One gr_reach_step_fwd per env batch, R independent batches in rotation (R x 19 MB > L2): every step reads HBM."""
    envs = [ReachTargetVecEnv(cfg, N, seed=s) for s in range(R)]
    acts = [torch.randn(N, 4, device="cuda:0") * 0.4 for _ in range(R)]
    for e in envs:
        e.reset()

    def sweep():
        for e, x in zip(envs, acts):
            e.step(x)
    ms = time_graph(sweep, iters) / R
    gbps = N * B_FWD / (ms * 1e-3) / 1e9
    return {"us": ms * 1e3, "env_steps_per_s": N / (ms * 1e-3), "achieved_gbps": gbps, "frac": gbps / load_peak(), "bytes_per_env_step": B_FWD, "envs": N}


def bench_window(cfg, Nb=16384, T=48, iters=20):
    """BPTT window: T forward steps with tape + one reverse sweep (the reference's hover schedule: 48 steps)."""
    env = ReachTargetVecEnv(cfg, Nb, bptt_horizon=T)
    env._bptt.autograd = False
    env.reset()
    acts = [torch.randn(Nb, 4, device="cuda:0") * 0.4 for _ in range(T)]

    def window():
        env.detach()
        for x in acts:
            env.step(x)
        env._bptt.backward_window()
    ms = time_graph(window, iters)
    ms_b = time_graph(lambda: env._bptt.backward_window(), iters)
    acts_t = torch.stack(acts)

    def window1():                      # the same window as ONE forward launch (gr_reach_rollout_fwd) + the sweep; bit-identical results
        env.detach()
        env.rollout(acts_t)
        env._bptt.backward_window()
    ms1 = time_graph(window1, iters)
    return {"window_ms": ms, "sweep_us": ms_b * 1e3, "env_steps_per_s": Nb * T / (ms * 1e-3), "T": T, "envs": Nb,
            "sweep_gbps": Nb * T * (13 * 16 + 16) / (ms_b * 1e-3) / 1e9,
            "one_launch_window_ms": ms1, "one_launch_env_steps_per_s": Nb * T / (ms1 * 1e-3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None, help="lv | ps | ctbr: time only that step kernel (for ncu)")
    a = ap.parse_args()
    out = {"envs": a.envs, "bytes_per_env_step": B_FWD, "peak_gbps": load_peak()}
    for name, cfg in (("lv", ReachTargetCfg.lv(decimation=1, is_differentiable_physics=False)),
                      ("ps", ReachTargetCfg.ps(decimation=1, is_differentiable_physics=False)),
                      ("ctbr", ReachTargetCfg.ctbr(is_differentiable_physics=False))):
        if a.only and a.only != name:
            continue
        out[f"step_{name}"] = bench_step(cfg, a.envs)
        print(name, out[f"step_{name}"], flush=True)
    for name, cfg in () if a.only else (("lv", ReachTargetCfg.lv(decimation=1)), ("ctbr", ReachTargetCfg.ctbr())):
        out[f"bptt_{name}"] = bench_window(cfg)
        print("bptt", name, out[f"bptt_{name}"], flush=True)
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()

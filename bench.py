#!/usr/bin/env python
"""bench.py -- env-steps/s of the racing hot path on N B200s (contract: see the task prompt / DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA kernels through the C ABI)
  python bench.py --impl reference [--gpus N] ...                # reference arm: the oracle (CPU torch port of the
                                                                 # reference's own implementation) on the host cores

A "step" = ONE env.step() over one batch of 65,536 envs (BASELINE.json configs[3]: randomized tracks, per-env domain
randomisation, resets): one launch of gr_step_fwd.  `value` = device-timed throughput with inputs resident in HBM,
`e2e` = the same metric through RacingVecEnv.step() with pinned HOST actions in and obs/reward/dones out every step.
L2 hygiene: the step's working set (~40 MB at 65,536 envs) would live in the 126 MB L2, so the timed loop rotates over
R independent env batches (R x working set >= 3 x L2): every step reads its state from HBM.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

NUM_ENVS = 65536
STAGE = 1
L2_BYTES = 126e6
METRIC = "env-steps/sec (fwd)"
UNIT = "env-steps/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def ncu_traffic(kernel_key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full`
    capture (profiles/traffic.json, written by tools/ncu_summary.py from the .ncu-rep); None if no capture is committed."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.isfile(p):
        return None
    with open(p) as f:
        d = json.load(f)
    e = d.get(kernel_key)
    return None if e is None else float(e["dram_bytes_read"]) + float(e["dram_bytes_write"])


def algorithmic_bytes(cfg, stats: bool, critic: bool = True, dones64: bool = True) -> int:
    """Every persistent column read once + written once, every API output written once."""
    rd = 7 * 16 + 5 * 16 + 16                      # hot planes, cold planes, action
    wr = 7 * 16 + 64 + 4 + 1 + 1                   # hot planes, policy obs, reward, terminated, time_out
    if critic:
        wr += 64
    if dones64:
        wr += 8
    if cfg.add_cmd_noise:
        rd += 32                                   # gate-noise planes (read-only except on gate switch / reset)
    if stats:                                      # episode sums of reward terms 0..3 (terms 4, 5 ride in spare words of the hot planes)
        rd += 16
        wr += 16
    return rd + wr


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------------------
def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(x: float, world: int, device) -> float:
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def bind_to_gpu_numa(local: int):
    """Pin this rank's host threads to the CPU set nearest to its GPU (NVML ideal affinity), BEFORE any pinned buffer is allocated, so
    first-touch places the staging pages on the GPU's NUMA node.  Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus ({min(cpus)}-{max(cpus)})"
    except Exception as ex:                                 # NVML absent / restricted container: leave the scheduler alone
        return f"unbound ({type(ex).__name__})"
    return "unbound"


def reduce_vector(x, world: int, device, op="max"):
    """element-wise MAX (or all-gather) of a per-repetition timing vector over the ranks"""
    t = torch.tensor(x, dtype=torch.float64, device=device)
    if world == 1:
        return t.cpu() if op == "max" else t.cpu()[None]
    import torch.distributed as dist
    if op == "max":
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu()
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return torch.stack(out).cpu()


def med(v):
    v = sorted(float(x) for x in v)
    return v[len(v) // 2]


CONTRACT_BYTES_FWD = 478        # STAGE >= 1 forward, single-step API (454 B for STAGE 0)


def run_ours(args):
    from generalizableracing_b200 import build as BLD
    BLD.build()
    from generalizableracing_b200 import _lib as B
    from generalizableracing_b200.config import RacingCfg
    from generalizableracing_b200.env import RacingVecEnv
    from generalizableracing_b200.track_gen import generate_track_table, racing_complex_cfg

    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa = bind_to_gpu_numa(local)
    N = args.envs
    cfg = RacingCfg.for_stage(STAGE)
    table = generate_track_table(racing_complex_cfg())      # RacingComplexTerrainCfg: 20 types x 10 levels x 8 gates (the reference's seed-42 table)
    stats = not args.no_stats
    b_cols = algorithmic_bytes(cfg, stats)                   # this build's column set
    b_alg = CONTRACT_BYTES_FWD if cfg.add_cmd_noise else 454
    working_set = N * b_cols                                           # bytes touched by one step
    R = max(2, int(3 * L2_BYTES / working_set) + 1)
    envs = [RacingVecEnv(cfg, table, N, device=dev, seed=42 + r, episode_stats=stats, env_id_offset=rank * N,
                         global_num_envs=world * N, block_threads=args.block_threads) for r in range(R)]
    for e in envs:
        e.reset()
        # PPO's init_at_random_ep_len (on_policy_runner.py:118-121): staggered time-outs, ~1/200 of the envs reset per step
        e.episode_length_buf = torch.randint(0, cfg.max_episode_length, (N,), device=dev, dtype=torch.int32)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    actions = [torch.randn(N, 4, device=dev, generator=g) * 0.5 for _ in range(R)]      # resident in HBM before the timed region
    lib = B.load()

    ios = []
    for e, a in zip(envs, actions):
        o = e._outs[0]
        io = B.GrStepIO()
        io.action = a.data_ptr()
        io.obs, io.critic_obs, io.aux_obs = o["obs"].data_ptr(), o["critic"].data_ptr(), o["aux"].data_ptr()
        io.reward, io.terminated, io.time_out, io.dones = o["reward"].data_ptr(), o["terminated"].data_ptr(), o["time_out"].data_ptr(), o["dones"].data_ptr()
        io.log_accum = e._log_accum.data_ptr()
        ios.append(io)

    def launch_seq(start: int, n: int, step0: int = 0):
        """n launches of gr_step_fwd, round-robin over the R env batches beginning with batch `start % R`"""
        s = torch.cuda.current_stream(dev).cuda_stream
        for i in range(n):
            k = (start + i) % R
            e = envs[k]
            rng = B.GrRandom(None, e.seed, (step0 + start + i) & 0xFFFFFFFF)
            B.check(lib.gr_step_fwd(C.byref(e._gcfg), C.byref(e._track), C.byref(e._state), C.byref(rng), C.byref(ios[k]), s), "gr_step_fwd")

    # ---- W untimed warm-up steps (eager launches), exactly as requested
    K, W = args.steps, args.warmup
    launch_seq(0, W)
    torch.cuda.synchronize(dev)

    # ---- the K timed steps of one repetition = CUDA graphs holding exactly K launches (launch-bound otherwise); repetition r
    # starts with batch (W + r*K) % R, so a batch is never touched again before the R-1 others (>= 3 x L2 of other traffic)
    CH = R * 90
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        launch_seq(0, R, 10_000)            # warm the capture stream
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    gcache = {}

    def graph_for(start: int, n: int):
        key = (start % R, n)
        if key not in gcache:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                launch_seq(start % R, n, 20_000)
            gcache[key] = gr
        return gcache[key]

    def rep_graphs(r: int):
        s0 = (W + r * K) % R
        out = [graph_for(s0, CH)] * (K // CH)
        if K % CH:
            out.append(graph_for(s0, K % CH))
        return out

    n_distinct = R // __import__("math").gcd(K, R) if K % R else 1
    plans = [rep_graphs(r) for r in range(n_distinct)]
    for pl in plans:                        # first replay of every graph (upload) outside the timed region
        for gr in pl:
            gr.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(3):
        for gr in plans[r % n_distinct]:
            gr.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    est_rep_ms = max(e0.elapsed_time(e1) / 3, 1e-3)
    reps = args.repeats or int(min(8000, max(100, 1200.0 / est_rep_ms)))
    if world > 1:                           # every rank must time the same number of repetitions
        import torch.distributed as dist
        t = torch.tensor([reps], device=dev)
        dist.broadcast(t, 0)
        reps = int(t.item())
    # ---- the timed set as ONE graph: M back-to-back repetitions of exactly K launches with a timing event at every repetition
    # boundary.  The events are event-record NODES on a forked branch of the graph (torch.cuda.Event(external=True)): event r fires when
    # the last kernel of repetition r-1 has completed, so ev[r+1] - ev[r] is the device time of exactly the K steps of repetition r in
    # the steady state of the dependent-launch chain (the kernels' own programmatic edges stay intact; no graph-launch ramp inside a
    # repetition).  Falls back to one graph launch per repetition, events on the stream, if the capture is refused.
    def build_timed_graph(M):
        evn = [torch.cuda.Event(enable_timing=True, external=True) for _ in range(M + 1)]
        gr = torch.cuda.CUDAGraph()
        branch = torch.cuda.Stream(dev)
        with torch.cuda.graph(gr):
            main = torch.cuda.current_stream(dev)
            for r in range(M + 1):
                mark = torch.cuda.Event()
                mark.record(main)
                branch.wait_event(mark)
                evn[r].record(branch)
                if r < M:
                    launch_seq((W + r * K) % R, K, 40_000)
            main.wait_stream(branch)
        return gr, evn

    # M: as many repetitions as ~2000 kernel nodes hold, and a total launch count that leaves the next replay's first batches cold
    M = max(1, min(reps, 2000 // K if K <= 2000 else 1))
    while M > 1 and 0 < (M * K) % R < 4:
        M -= 1
    timed_graph = None
    if not args.events_on_stream:
        try:
            timed_graph, evn = build_timed_graph(M)
            for _ in range(2):                           # (the first replay of a fresh graph pays its upload)
                timed_graph.replay()
                torch.cuda.synchronize(dev)
            chk = med([evn[r].elapsed_time(evn[r + 1]) for r in range(M)])
            if not (0.4 * est_rep_ms < chk < 1.2 * est_rep_ms):
                raise RuntimeError(f"in-graph events read {chk} ms against {est_rep_ms} ms per repetition")
        except Exception as ex:                            # noqa: BLE001 -- any capture problem: the per-launch method still measures
            sys.stderr.write(f"bench.py: in-graph timing events unavailable ({ex}); timing one graph launch per repetition\n")
            timed_graph = None
    n_replays = -(-reps // M) if timed_graph is not None else 0
    if timed_graph is not None:
        reps = n_replays * M
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps if timed_graph is None else min(reps, 200))]

    # ---- clock sampler (rank 0's GPU) starts >= 0.5 s BEFORE the first timed repetition, under the same load
    clk = ClockSampler(local) if rank == 0 else None
    if clk is not None:
        clk.__enter__()
    lead_replays = 0
    t_end = time.time() + 0.6
    while time.time() < t_end:
        for gr in plans[lead_replays % n_distinct]:
            gr.replay()
        lead_replays += 1
        if lead_replays % 64 == 0:
            torch.cuda.synchronize(dev)
    # ---- the same K steps as isolated repetitions (one graph launch each, events on the stream): what a caller who launches K steps
    # and waits would see, graph-launch ramp included.  Reported beside the headline, and the headline itself in the fallback.
    torch.cuda.synchronize(dev)
    barrier(world)
    for r in range(len(evs)):
        evs[r][0].record()
        for gr in plans[r % n_distinct]:
            gr.replay()
        evs[r][1].record()
    torch.cuda.synchronize(dev)
    isolated_ms = [a.elapsed_time(b) for a, b in evs]
    for e in envs:
        e._log_accum.zero_()
    torch.cuda.synchronize(dev)
    barrier(world)
    torch.cuda.synchronize(dev)
    # ---- timed: `reps` back-to-back repetitions of EXACTLY K steps, each bracketed by CUDA events
    if timed_graph is not None:
        local_ms = []
        for _ in range(n_replays):
            timed_graph.replay()
            torch.cuda.synchronize(dev)
            local_ms += [evn[r].elapsed_time(evn[r + 1]) for r in range(M)]
    else:
        local_ms = isolated_ms
    barrier(world)
    if clk is not None:
        clk.__exit__()
    rep_ms = reduce_vector(local_ms, world, dev, "max")          # per repetition: max over ranks
    per_rank = reduce_vector([med(local_ms), min(local_ms), max(local_ms)], world, dev, "gather")
    ms = med(rep_ms)                                             # reported: the median repetition
    # emergent reset rate of the timed workload, from the kernels' own log accumulators
    tot = torch.stack([e._log_accum.sum(dim=0) for e in envs]).sum(dim=0)
    reset_rate = float(tot[0].item()) / ((reps if timed_graph is not None else len(evs)) * K * N)
    value = world * N * K / (ms * 1e-3)
    kernel_us = med(local_ms) * 1e3 / K
    peak, peak_src, _ = load_peaks()
    achieved = b_alg * N / (kernel_us * 1e-6) / 1e9
    achieved_cols = b_cols * N / (kernel_us * 1e-6) / 1e9
    timing = {"repeats": reps, "rep_ms_median": ms, "rep_ms_min": float(rep_ms.min()), "rep_ms_max": float(rep_ms.max()), "rep_ms_first": float(rep_ms[0]),
              "rep_ms_p10": float(rep_ms.kthvalue(max(1, reps // 10)).values), "rep_ms_p90": float(rep_ms.kthvalue(max(1, reps * 9 // 10)).values),
              "kernel_us_per_rank_median": [float(x) * 1e3 / K for x in per_rank[:, 0]],
              "kernel_us_per_rank_min": [float(x) * 1e3 / K for x in per_rank[:, 1]],
              "kernel_us_per_rank_max": [float(x) * 1e3 / K for x in per_rank[:, 2]],
              "isolated_rep_ms_median": med(isolated_ms), "isolated_kernel_us": med(isolated_ms) * 1e3 / K,
              "events": "in-graph" if timed_graph is not None else "on-stream",
              "how": (f"{reps} back-to-back repetitions of exactly K={K} steps = {n_replays} replays of one CUDA graph holding {M} repetitions with a timing-event node "
                      f"at every repetition boundary (fires when the repetition's last kernel completes); " if timed_graph is not None else
                      f"{reps} back-to-back repetitions of exactly K={K} steps, one graph launch each, CUDA events on the launching stream; ") +
                     f"every repetition max-reduced over ranks, value = median repetition; barrier + synchronize around the set; untimed lead-in of {lead_replays} "
                     f"repetitions for the clock sampler; isolated_* = the same K steps as one graph launch between two stream events (launch ramp included)",
              "host_numa_binding": numa}

    def timed_graphs(pls, n_reps):
        """median device time (ms) of one K-step repetition over n_reps repetitions of the given plans"""
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_reps)]
        for r in range(n_reps):
            pairs[r][0].record()
            for gr in pls[r % len(pls)]:
                gr.replay()
            pairs[r][1].record()
        torch.cuda.synchronize(dev)
        return med([a.elapsed_time(b) for a, b in pairs])

    def sum_resets():
        return float(torch.stack([e._log_accum.sum(dim=0) for e in envs]).sum(dim=0)[0].item())

    # ---- same graphs, near-hover actions (time-out resets only): the low-reset regime of a trained policy
    low = None
    sub_reps = max(20, min(reps, 400))
    if not args.no_extra:
        for a in actions:
            a.copy_(torch.randn(N, 4, device=dev, generator=g) * 0.1 + torch.tensor([-0.3466, 0.0, 0.0, 0.0], device=dev))
        timed_graphs(plans, max(4, 440 // max(K, 1)))             # let the tumbling drones of the random-action phase reset
        for e in envs:
            e._log_accum.zero_()
        us = timed_graphs(plans, sub_reps) * 1e3 / K
        low = {"kernel_us": us, "env_steps_per_s_per_gpu": N / (us * 1e-6), "frac_of_hbm_peak": b_alg * N / (us * 1e-6) / 1e9 / peak,
               "frac_of_hbm_peak_column_set": b_cols * N / (us * 1e-6) / 1e9 / peak,
               "resets_per_env_step": sum_resets() / (sub_reps * K * N), "actions": "N((-0.35,0,0,0), 0.1^2): near hover"}

    # ---- forced reset rates ("also report a forced 1 % and 10 % reset-rate case for C4"): the same launches re-captured
    # with max_episode_length = 100 / 10 and staggered episode counters, so 1 % / 10 % of the envs time out (and run the reset
    # tail: curriculum, pose / velocity / drag / noise re-draws) every step, on top of the near-hover crash rate
    forced = {}
    if not args.no_extra:
        for label, max_len in (("1pct", 100), ("10pct", 10)):
            for e in envs:
                e._gcfg.max_episode_length = max_len
                e.episode_length_buf = torch.randint(0, max_len, (N,), device=dev, dtype=torch.int32)
            torch.cuda.synchronize(dev)
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2):
                launch_seq(0, 4 * R, 30_000)
            for _ in range(10):
                g2.replay()
            for e in envs:
                e._log_accum.zero_()
            us = timed_graphs([[g2]], 50) * 1e3 / (4 * R)
            forced[label] = {"kernel_us": us, "env_steps_per_s_per_gpu": N / (us * 1e-6), "frac_of_hbm_peak": b_alg * N / (us * 1e-6) / 1e9 / peak,
                             "frac_of_hbm_peak_column_set": b_cols * N / (us * 1e-6) / 1e9 / peak,
                             "resets_per_env_step": sum_resets() / (50 * 4 * R * N), "max_episode_length": max_len}
            del g2
        for e in envs:
            e._gcfg.max_episode_length = cfg.max_episode_length
            e.episode_length_buf = torch.randint(0, cfg.max_episode_length, (N,), device=dev, dtype=torch.int32)
        torch.cuda.synchronize(dev)

    # ---- e2e: public API with HOST buffers (RacingVecEnv.step_host -> gr_host_pipe_*), H2D + D2H inside the timed region.
    # Every step copies its own actions host->device and its obs / reward / dones device->host; `depth` steps are in flight
    # so the copies of step t overlap the kernel of step t+1 (the consumer reads step t's results while t+1 runs).
    env = envs[0]
    depth = 3
    hb = env.host_buffers(depth)                     # pinned; obs | reward | dones of a set contiguous: one device->host copy per step
    for b in hb:
        b["actions"].copy_(torch.randn(N, 4) * 0.5)
    h_act, h_obs, h_rew, h_done = [b["actions"] for b in hb], [b["obs"] for b in hb], [b["reward"] for b in hb], [b["dones"] for b in hb]
    torch.cuda.synchronize(dev)

    def e2e_loop(n, d):
        tickets = []
        for t in range(n):
            k = t % d
            if t >= d:
                env.wait_host(tickets[t - d])            # results of step t-d are on the host before its buffers are reused
            tickets.append(env.step_host(h_act[k], h_obs[k], h_rew[k], h_done[k], depth=depth))
        for tk in tickets[-d:]:
            env.wait_host(tk)

    def timed_e2e(n, d, n_reps):
        """n_reps repetitions of n host-API steps; each repetition synchronised on both sides (the results are in host memory when the
        clock stops, so host time IS the metric here); per repetition max over ranks, median reported"""
        e2e_loop(max(3, W), d)
        out = []
        for _ in range(n_reps):
            barrier(world)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            e2e_loop(n, d)
            torch.cuda.synchronize(dev)
            out.append((time.perf_counter() - t0) * 1e3)
        return reduce_vector(out, world, dev, "max")

    Ke = max(1, min(K, 2000))
    e2e_reps = int(min(100, max(9, 1000.0 / (Ke * 0.105))))
    if world > 1:
        e2e_reps = min(e2e_reps, 40)
    e2e_v = timed_e2e(Ke, depth, e2e_reps)
    e2e_ms = med(e2e_v)
    e2e_value = world * N * Ke / (e2e_ms * 1e-3)
    e2e_sync_ms = med(timed_e2e(Ke, 1, max(5, e2e_reps // 4)))          # one step in flight: H2D -> kernel -> D2H strictly in sequence
    assert torch.isfinite(h_obs[0]).all() and h_done[0].min() >= 0 and h_done[0].max() <= 1 and float(h_rew[0].abs().max()) < 1e3
    packed_sets = (h_obs, h_rew, h_done)
    h_obs, h_rew = [torch.empty(N, 16).pin_memory() for _ in range(depth)], [torch.empty(N).pin_memory() for _ in range(depth)]
    h_done = [torch.empty(N, dtype=torch.int64).pin_memory() for _ in range(depth)]              # three separate buffers: three copies per step
    e2e_sep_ms = med(timed_e2e(Ke, depth, max(5, e2e_reps // 3)))
    h_done = [torch.empty(N, dtype=torch.uint8).pin_memory() for _ in range(depth)]           # the same call with one-byte dones
    e2e_u8_ms = med(timed_e2e(Ke, depth, max(5, e2e_reps // 3)))
    assert int(h_done[0].max()) <= 1
    h_obs, h_rew, h_done = packed_sets
    h2d = N * 4 * 4
    d2h = N * 16 * 4 + N * 4 + N * 8
    probe = copy_only_probe(dev, world, lib, N, h2d, d2h, Ke, packed=1)
    probe["separate_buffers"] = copy_only_probe(dev, world, lib, N, h2d, d2h, Ke, packed=0)

    # ---- C5: the one collective of the path (NCCL policy-gradient all-reduce) and a whole PPO iteration with it inside the captured
    # update graph, 65,536 envs per GPU (on_policy_runner.py:135-183)
    collective = None
    if not args.no_extra and not args.no_collective:
        for e in envs[1:]:
            e.close()
        del envs[1:]
        torch.cuda.empty_cache()
        collective = bench_collective(dev, cfg, table, N, rank, world)

    # every collective of this run is behind us: leave the process group BEFORE anything rank 0 does on its own (a module
    # built on rank 0 alone would otherwise wait for the other ranks in its parameter broadcast)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()

    # ---- extras: the other kernels of the path (BPTT C3, GAE / add_transitions C2, fused collection), single-GPU runs only
    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        gcache.clear()
        plans.clear()
        extra = bench_extras(dev, cfg, table)
    if rank == 0 and low is not None:
        extra["fwd_low_reset"] = low
        extra["fwd_forced_reset_rate"] = forced

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        for e in envs:
            e.close()
        del envs, env
        torch.cuda.empty_cache()
        extra["torch_eager_gpu_baseline"] = torch_gpu_baseline(dev)
        cpu = cpu_baseline(sample_s=args.cpu_seconds)
        extra["cpu_baseline_fwd_bwd_bptt"] = cpu_baseline_bptt()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"C4: {N} envs/GPU, STAGE {STAGE}, RacingComplexTerrainCfg gate table (20 types x 10 levels x 8 gates, the reference's seed-42 table incl. its "
                                   f"obstacle draws), per-env DR, staggered resets, in-kernel Philox, episode_stats={stats}",
                       "envs_per_gpu": N, "l2": f"inputs larger than L2: rotating {R} independent env batches ({R}x{working_set / 1e6:.0f} MB > 126 MB L2), "
                                               f"each repetition = CUDA graphs holding exactly K launches",
                       "mass_kg": cfg.mass, "actions": "N(0, 0.5^2) resident in HBM", "resets_per_env_step": reset_rate,
                       "prefetch": os.environ.get("GRACING_PREFETCH", "1")},
            "timing": timing,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "ms_per_step": e2e_ms / Ke, "repeats": e2e_reps, "rep_ms_min": float(e2e_v.min()), "rep_ms_max": float(e2e_v.max()),
                    "api": f"RacingVecEnv.step_host -> gr_host_pipe_step/wait (C ABI, pinned HOST buffers from env.host_buffers(): actions in; obs, reward, "
                           f"int64 dones out every step -- contiguous in host memory, so they cross PCIe as one copy; {depth} steps in flight, results of "
                           f"step t read while step t+1 runs)",
                    "separate_buffers": {"value": world * N * Ke / (e2e_sep_ms * 1e-3), "ms_per_step": e2e_sep_ms / Ke,
                                         "note": "same call with three separately allocated pinned output tensors (three device->host copies per step)"},
                    "sync_per_step": {"value": world * N * Ke / (e2e_sync_ms * 1e-3), "ms_per_step": e2e_sync_ms / Ke,
                                      "note": "same call, one step in flight (H2D -> kernel -> D2H in sequence)"},
                    "byte_dones": {"value": world * N * Ke / (e2e_u8_ms * 1e-3), "ms_per_step": e2e_u8_ms / Ke, "d2h_bytes_per_step": N * 16 * 4 + N * 4 + N,
                                   "note": "same call, dones handed out as uint8 (1 B per env over PCIe)"},
                    "copy_only_probe": probe},
            "gpu_launches": K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic("racing_step_fwd_kernel<1,0,1,1>"),
                         "traffic_note": "static: one `ncu --set full` capture of this kernel (profiles/traffic.json), not measured in this run; ncu flushes the caches "
                                         "per launch, so the writes that drain after the kernel's end are not in it",
                         "kernel": "racing_step_fwd_kernel<noise,nodiff,philox,stats>",
                         "bytes_per_env_step": b_alg, "bytes_rule": "SURVEY 8(d) contract: 454 B STAGE 0 + 24 B gate-noise reads = 478 B",
                         "bytes_per_env_step_column_set": b_cols, "achieved_column_set": achieved_cols, "frac_column_set": achieved_cols / peak,
                         "kernel_us": kernel_us, "peak_source": peak_src},
            "clocks": clk.summary(),
            "cpu_baseline": cpu,
            "collective": collective,
            "extra": extra,
        }
        if "bptt_fwd_bwd_c3" in extra:          # the second half of BASELINE's metric ("fwd, and fwd+bwd BPTT"), C3 = 16,384 envs x 32
            line["fwd_bwd_bptt"] = {"unit": UNIT, "config": "C3: 16384 envs, horizon 32",
                                    "dynamics_only": extra["bptt_fwd_bwd_c3"]["env_steps_per_s"],
                                    "dynamics_only_one_launch_window": extra.get("bptt_fwd_bwd_c3_one_launch_window", {}).get("env_steps_per_s"),
                                    "training_iteration_with_policy": extra.get("bptt_training_c3", {}).get("fused_kernel_backward", {}).get("env_steps_per_s"),
                                    "training_iteration_with_policy_log_every_20": extra.get("bptt_training_c3", {}).get("fused_kernel_backward_log_every_20", {}).get("env_steps_per_s"),
                                    "cpu_baseline": extra.get("cpu_baseline_fwd_bwd_bptt")}
        print(json.dumps(line))


def copy_only_probe(dev, world, lib, N, h2d, d2h, steps, packed=1):
    """The platform's ceiling for the e2e loop: the SAME transfers per step (actions in; obs, reward, int64 dones out) from pinned memory --
    obs | reward | dones as ONE copy (packed, what the pipe issues for env.host_buffers()) or one cudaMemcpyAsync per buffer -- H2D on one
    stream and D2H on another, all ranks concurrently, NO kernel and no dependencies (gr_host_copy_probe2: a C loop, so that Python's
    per-call cost is not what gets measured)."""
    import ctypes as C
    n = max(50, min(steps, 500))
    out = []
    for _ in range(7):
        barrier(world)
        torch.cuda.synchronize(dev)
        sec = C.c_double()
        rc = lib.gr_host_copy_probe2(N, n, 8, int(packed), C.byref(sec))
        if rc:
            return {"error": int(rc)}
        out.append(sec.value * 1e3)
    ms = med(reduce_vector(out, world, dev, "max"))
    return {"us_per_step": ms * 1e3 / n, "aggregate_GBps": world * (h2d + d2h) * n / (ms * 1e-3) / 1e9, "per_gpu_GBps": (h2d + d2h) * n / (ms * 1e-3) / 1e9,
            "env_steps_per_s_ceiling": world * N * n / (ms * 1e-3),
            "what": f"copies only (pinned buffers, {h2d} B H2D + {d2h} B D2H per step as {'one copy' if packed else 'three copies'}, two streams, C loop, "
                    f"{world} rank(s) concurrently): the platform limit of the e2e loop"}


def bench_collective(dev, cfg, table, N, rank, world, T: int = 24):
    """BASELINE C5: env-sharded N envs per GPU with the NCCL policy-gradient all-reduce.  (1) the all-reduce alone, on the flat buffer of
    algorithms/ppo.py (all gradients + loss / KL sums), replayed from a CUDA graph; (2) one whole PPO iteration -- fused 24-step
    collection (gr_ppo_collect), GAE with globally merged moments, 5 x 4 mini-batch steps from libgracing kernels with the all-reduce
    INSIDE the captured step graph -- device-timed, max over ranks.  Runs at world == 1 too (no collective) so that the 1 -> N curve of
    the training iteration has its base point."""
    from generalizableracing_b200.env import RacingVecEnv
    from generalizableracing_b200.runners import OnPolicyRunner
    out = {"envs_per_gpu": N, "steps_per_env": T, "world": world}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agent = {"num_steps_per_env": T, "save_interval": 10 ** 9, "empirical_normalization": False, "fused_collection": True,
             "policy": {"class_name": "ActorCritic", "init_noise_std": 1.0, "actor_hidden_dims": [128, 128], "critic_hidden_dims": [128, 128], "activation": "lrelu"},
             "algorithm": {"class_name": "PPO", "value_loss_coef": 1.0, "use_clipped_value_loss": True, "clip_param": 0.2, "entropy_coef": 0.0,
                           "num_learning_epochs": 5, "num_mini_batches": 4, "learning_rate": 5.0e-4, "schedule": "adaptive", "gamma": 0.99, "lam": 0.95,
                           "desired_kl": 0.01, "max_grad_norm": 1.0, "graphed_update": True, "kernel_update": True}}
    env = RacingVecEnv(cfg, table, N, device=dev, seed=9, env_id_offset=rank * N, global_num_envs=world * N)
    runner = OnPolicyRunner(env, agent, log_dir=None, device=str(dev))
    runner.learn(3, init_at_random_ep_len=True)          # eager first update, graph capture on the second
    iters = 10
    per_it = []
    for _ in range(3):
        barrier(world)
        torch.cuda.synchronize(dev)
        e0.record()
        runner.learn(iters)
        e1.record()
        torch.cuda.synchronize(dev)
        per_it.append(e0.elapsed_time(e1) / iters)
    ms = med(reduce_vector(per_it, world, dev, "max"))
    flat = runner.alg._graph["flat_grad"]
    peer = runner.alg._graph.get("peer_allreduce")
    out["ppo_iteration"] = {"ms_per_iteration": ms, "env_steps_per_s": world * N * T / (ms * 1e-3), "allreduces_per_iteration": 20 if world > 1 else 0,
                            "gradient_sum": "none (one rank)" if world == 1 else ("gr_peer_allreduce (own kernel over NVLink peer memory)" if peer is not None else "NCCL all_reduce"),
                            "what": "gr_ppo_collect (24 steps, tcgen05 policy) + GAE + transition records packed in mini-batch order + 20 captured mini-batch steps "
                                    "(forward head + loss + tcgen05 weight gradients in one launch, [sum of the flat gradient buffer over the ranks], clip + Adam), "
                                    "max over ranks, median of 3 x 10 iterations"}
    if peer is not None:                 # the exchange kernel alone, as the NCCL one below: 20 per graph replay
        gr = torch.cuda.CUDAGraph()
        n_in = 20
        for _ in range(3):
            peer.launch()
        torch.cuda.synchronize(dev)
        with torch.cuda.graph(gr):
            for _ in range(n_in):
                peer.launch()
        gr.replay()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(30):
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1) * 1e3 / n_in)
        us = reduce_vector(ts, world, dev, "max")
        out["peer_allreduce"] = {"us": med(us), "us_min": float(us.min()), "us_max": float(us.max()), "floats": int(peer.n), "failed": bool(peer.failed()),
                                 "what": "gr_peer_allreduce (csrc/peer_reduce.cu): every rank reads the other ranks' gradient buffers over NVLink between two flag "
                                         "barriers, rank-order sum; 20 per CUDA-graph replay, device-timed, max over ranks, median of 30 replays"}
        del gr
    if world > 1:
        import torch.distributed as dist
        buf = torch.zeros_like(flat)
        for _ in range(5):
            dist.all_reduce(buf)
        torch.cuda.synchronize(dev)
        gr = torch.cuda.CUDAGraph()
        n_in = 20
        with torch.cuda.graph(gr):
            for _ in range(n_in):
                dist.all_reduce(buf)
        gr.replay()
        torch.cuda.synchronize(dev)
        ts = []
        for _ in range(30):
            e0.record()
            gr.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            ts.append(e0.elapsed_time(e1) * 1e3 / n_in)
        us = reduce_vector(ts, world, dev, "max")
        out["allreduce"] = {"us": med(us), "us_min": float(us.min()), "us_max": float(us.max()), "floats": int(flat.numel()), "bytes": int(flat.numel()) * 4,
                            "what": "dist.all_reduce (NCCL, sum) of the flat policy-gradient buffer (all 13 parameter gradients + 16 loss / KL sums), 20 per CUDA-graph "
                                    "replay, device-timed, max over ranks, median of 30 replays"}
        del gr
    runner.alg.close()
    env.close()
    del runner, env
    torch.cuda.empty_cache()
    return out


def bench_extras(dev, cfg, table):
    """Device timings of the other §8 kernels (reported next to the headline): BPTT fwd+bwd (BASELINE C3), GAE and
    add_transitions (C2), and the headline kernel on a low-reset (near-hover) action distribution."""
    import dataclasses
    from generalizableracing_b200 import _lib as B
    from generalizableracing_b200.env import RacingVecEnv
    from generalizableracing_b200.storage import RolloutStorage
    out = {}
    peak, _, _ = load_peaks()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---- C3: 16384 envs, horizon 32: 32 forward launches (loss + tape) + ONE reverse sweep, dynamics only (actions resident)
    N, H, R = 16384, 32, 4                       # R rotating env sets: 4 x (59 MB tape + state) > L2
    dcfg = dataclasses.replace(cfg, is_differentiable_physics=True)
    envs = [RacingVecEnv(dcfg, table, N, device=dev, seed=7 + r, episode_stats=False, bptt_horizon=H) for r in range(R)]
    acts = torch.randn(H, N, 4, device=dev) * 0.5
    for e in envs:
        e.reset()
        e.episode_length_buf = torch.randint(0, cfg.max_episode_length, (N,), device=dev, dtype=torch.int32)
        e._bptt.autograd = False

    def window(e):
        e.detach()
        for t in range(H):
            e.step(acts[t])
        return e._bptt.backward_window()

    for e in envs:
        window(e)
    torch.cuda.synchronize(dev)
    reps = 5
    e0.record()
    for _ in range(reps):
        for e in envs:
            window(e)
    e1.record()
    torch.cuda.synchronize(dev)
    ms_py = e0.elapsed_time(e1) / (reps * R)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for e in envs:
            window(e)
    graph.replay()
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    ms_g = e0.elapsed_time(e1) / (reps * R)
    # bytes: fwd(diff, critic, noise, no stats) = 240 read + (112 state + 128 obs + 14 + 16 loss/terms + 112 tape) write; bwd = 112 + 16
    b_fb = 240 + 112 + 128 + 14 + 16 + 112 + 112 + 16
    out["bptt_fwd_bwd_c3"] = {"envs": N, "horizon": H, "env_steps_per_s": N * H / (ms_g * 1e-3), "ms_per_window_graph": ms_g,
                              "ms_per_window_python_driven": ms_py, "bytes_per_env_step": b_fb,
                              "achieved_GBps": b_fb * N * H / (ms_g * 1e-3) / 1e9, "frac_of_hbm_peak": b_fb * N * H / (ms_g * 1e-3) / 1e9 / peak,
                              "l2": f"rotating {R} env sets + tapes ({R}x59 MB), one CUDA graph of {R} windows x (32 fwd + memsets + 1 sweep)"}
    # ---- the same window as ONE forward launch (gr_rollout_fwd: the actions of a dynamics-only window are known in advance, so
    # the env state stays in registers over the horizon) + the same reverse sweep; results bit-identical (tests/test_rollout_window.py)
    def window1(e):
        e.detach()
        e.rollout(acts)
        return e._bptt.backward_window()

    for e in envs:
        window1(e)
    torch.cuda.synchronize(dev)
    graph1 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph1):
        for e in envs:
            window1(e)
    graph1.replay()
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(reps):
        graph1.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    ms_w = e0.elapsed_time(e1) / (reps * R)
    e0.record()
    for _ in range(reps):
        for e in envs:
            e.detach()
            e.rollout(acts)
    e1.record()
    torch.cuda.synchronize(dev)
    ms_wf = e0.elapsed_time(e1) / (reps * R)
    # bytes per env-step: action 16 + tape 112 + loss 4 + loss terms 12 + reward 4 + 3 masks written, state + observations once per
    # window ((240 + 112 + 128) / H); sweep 112 + 16
    b_w = 16 + 112 + 4 + 12 + 4 + 3 + (240 + 112 + 128) / H + 112 + 16
    out["bptt_fwd_bwd_c3_one_launch_window"] = {
        "envs": N, "horizon": H, "env_steps_per_s": N * H / (ms_w * 1e-3), "ms_per_window_graph": ms_w, "ms_forward_window_python_driven": ms_wf,
        "bytes_per_env_step": b_w, "achieved_GBps": b_w * N * H / (ms_w * 1e-3) / 1e9, "frac_of_hbm_peak": b_w * N * H / (ms_w * 1e-3) / 1e9 / peak,
        "what": "gr_rollout_fwd (32 steps, one launch, state in registers) + gr_step_bwd; same rotation of 4 env sets + tapes"}
    del graph1

    def graph_ms(fn, n_rep=10):
        """device time of fn(e) per env set, replayed from one CUDA graph over the R rotating sets"""
        for e in envs:
            fn(e)
        torch.cuda.synchronize(dev)
        gx = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gx):
            for e in envs:
                fn(e)
        gx.replay()
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(n_rep):
            gx.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / (n_rep * R)

    def fwd_only(e):
        e.detach()
        e.rollout(acts)

    ms_f = graph_ms(fwd_only)
    ms_b = graph_ms(lambda e: e._bptt.backward_window())
    out["bptt_fwd_bwd_c3_one_launch_window"]["ms_forward_window_graph"] = ms_f
    out["bptt_bwd_sweep"] = {"ms": ms_b, "achieved_GBps": N * H * 128 / (ms_b * 1e-3) / 1e9, "frac_of_hbm_peak": N * H * 128 / (ms_b * 1e-3) / 1e9 / peak,
                             "what": "gr_step_bwd over the 32-step tape (+ the memset of the carried adjoints), graph-replayed over the 4 rotating env sets"}
    del envs, graph
    torch.cuda.empty_cache()

    # ---- forward-only window at C4 size: 65,536 envs x 24 steps in one launch, observations of every step recorded (play-back of
    # recorded actions / open-loop evaluation): per env-step 16 B action + 64 B observation + 7 B reward / masks, state once per window
    Nw, Tw = 65536, 24
    wenvs = [RacingVecEnv(cfg, table, Nw, device=dev, seed=31 + r) for r in range(2)]
    wacts = torch.randn(Tw, Nw, 4, device=dev) * 0.5
    for e in wenvs:
        e.reset()
        e.episode_length_buf = torch.randint(0, cfg.max_episode_length, (Nw,), device=dev, dtype=torch.int32)
        e.rollout(wacts, record_obs=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(reps):
        for e in wenvs:
            e.rollout(wacts, record_obs=True)
    e1.record()
    torch.cuda.synchronize(dev)
    ms_fw = e0.elapsed_time(e1) / (reps * len(wenvs))
    b_fw = 16 + 64 + 7 + algorithmic_bytes(cfg, True) / Tw
    out["fwd_window_65536x24"] = {"ms_per_window": ms_fw, "us_per_step": ms_fw * 1e3 / Tw, "env_steps_per_s": Nw * Tw / (ms_fw * 1e-3),
                                  "bytes_per_env_step": b_fw, "achieved_GBps": b_fw * Nw * Tw / (ms_fw * 1e-3) / 1e9,
                                  "what": "gr_rollout_fwd, STAGE 1, stats on, obs_seq recorded; 2 env sets alternating (each window moves 174 MB > L2)"}
    for e in wenvs:
        e.close()
    del wenvs
    torch.cuda.empty_cache()

    # ---- C2: rollout storage on [24, 4096]
    T, N2 = 24, 4096
    sto = RolloutStorage("rl", N2, T, [16], [16], [4], device=dev)
    sto.rewards.normal_()
    sto.values.normal_()
    last = torch.randn(N2, 1, device=dev)
    for _ in range(3):
        sto.compute_returns(last, 0.99, 0.95)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(50):
        sto.compute_returns(last, 0.99, 0.95)
    e1.record()
    torch.cuda.synchronize(dev)
    out["gae_24x4096_us"] = e0.elapsed_time(e1) * 1e3 / 50
    gg = torch.cuda.CUDAGraph()                              # the same call replayed from a graph: device time without the host's launch cost
    with torch.cuda.graph(gg):
        for _ in range(10):
            sto.compute_returns(last, 0.99, 0.95)
    gg.replay()
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(20):
        gg.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    out["gae_24x4096_graph_us"] = e0.elapsed_time(e1) * 1e3 / 200
    del gg
    tr = sto.Transition()
    tr.observations, tr.privileged_observations, tr.actions = torch.randn(N2, 16, device=dev), torch.randn(N2, 16, device=dev), torch.randn(N2, 4, device=dev)
    tr.rewards, tr.values, tr.dones = torch.randn(N2, device=dev), torch.randn(N2, 1, device=dev), torch.zeros(N2, dtype=torch.int64, device=dev)
    tr.actions_log_prob, tr.action_mean, tr.action_sigma = torch.randn(N2, device=dev), torch.randn(N2, 4, device=dev), torch.rand(N2, 4, device=dev)
    tr.time_outs, tr.gamma = torch.zeros(N2, dtype=torch.bool, device=dev), 0.99
    for _ in range(3):
        sto.clear()
        sto.add_transitions(tr)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(20):
        sto.clear()
        for _t in range(T):
            sto.add_transitions(tr)
    e1.record()
    torch.cuda.synchronize(dev)
    out["add_transitions_4096_us"] = e0.elapsed_time(e1) * 1e3 / (20 * T)
    del sto
    out["ppo_collection"] = bench_collection(dev, cfg, table)
    out["bptt_training_c3"] = bench_bptt_training(dev, cfg, table)
    out["ppo_training_c2"] = bench_ppo_training(dev, cfg, table)
    # ---- the reach-target tasks / LV command mode (tools/reach_bench.py)
    from generalizableracing_b200.config import ReachTargetCfg
    from tools.reach_bench import bench_step, bench_window
    out["reach_target"] = {"step_lv_65536": bench_step(ReachTargetCfg.lv(decimation=1, is_differentiable_physics=False), 65536),
                           "bptt_window_ctbr_16384x48": bench_window(ReachTargetCfg.ctbr()),
                           "bptt_window_lv_16384x48": bench_window(ReachTargetCfg.lv(decimation=1))}
    out["mesh_collision"] = bench_mesh_collision(dev, cfg, table)
    return out


def bench_mesh_collision(dev, cfg, table, N: int = NUM_ENVS):
    """SURVEY 8f rank 4, last item: the terrain-mesh collision count (gr_uav_collision_ray = the reference's Warp kernel behind
    get_uav_collision_num_ray) for 65,536 UAVs x 17 lattice points x up to 6 axis rays against a box model of the bench's gate table (200
    tiles: ground slabs + four-bar gate frames), on the poses of a stepped env batch; and env.step with the term switched on."""
    from generalizableracing_b200 import mesh as M
    from generalizableracing_b200.env import RacingVecEnv
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pts, faces = M.track_table_mesh(table)
    t0 = time.perf_counter()
    tm = M.TerrainMesh(pts, faces, device=dev)
    build_s = time.perf_counter() - t0
    env = RacingVecEnv(cfg, table, N, device=dev, seed=77)
    env.reset()
    env.episode_length_buf = torch.randint(0, cfg.max_episode_length, (N,), device=dev, dtype=torch.int32)
    acts = torch.randn(N, 4, device=dev) * 0.5
    for _ in range(30):
        env.step(acts)
    sv = env.state_dict_view()
    pos, quat = sv["root_pos_w"].contiguous(), sv["root_quat_w"].contiguous()
    lat = M.LATTICE_TENSOR.to(dev)
    num = M.get_uav_collision_num_ray(tm, pos, quat, 0.09, 0.05, 1e3, lat)
    torch.cuda.synchronize(dev)
    reps = 20
    e0.record()
    for _ in range(reps):
        M.get_uav_collision_num_ray(tm, pos, quat, 0.09, 0.05, 1e3, lat)
    e1.record()
    torch.cuda.synchronize(dev)
    us = e0.elapsed_time(e1) * 1e3 / reps
    env.set_terrain_mesh(tm, -50.0)
    for _ in range(5):
        env.step(acts)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(reps):
        env.step(acts)
    e1.record()
    torch.cuda.synchronize(dev)
    us_step = e0.elapsed_time(e1) * 1e3 / reps
    env.close()
    return {"uavs": N, "faces": int(faces.shape[0]), "bvh_nodes": int(tm.num_nodes), "bvh_build_host_s": build_s, "us_per_call": us,
            "uav_checks_per_s": N / (us * 1e-6), "lattice_points_inside_mean": float(num.float().mean()), "colliding_fraction": float((num > 2).float().mean()),
            "env_step_with_term_us": us_step,
            "what": "get_uav_collision_num_ray (17 lattice points x <= 6 axis rays per UAV, BVH traversal) on the poses of a stepped 65,536-env batch; "
                    "env_step_with_term_us = RacingVecEnv.step with set_terrain_mesh (step kernel + ray casts + reward add, Python-driven)"}


def bench_ppo_training(dev, cfg, table, N: int = 4096, T: int = 24):
    """BASELINE C2 as the trainer runs it (OnPolicyRunner.learn, on_policy_runner.py:135-183): one PPO iteration = 24-step
    rollout + GAE + 5 epochs x 4 mini-batches.  step_by_step = torch policy + gr_step_fwd + gr_storage_add, eager torch update (the
    reference's structure on our env kernels); fused = gr_ppo_collect + CUDA-graph torch update; fused_kernel_update = the
    update's forward / loss / weight gradients from libgracing kernels as well."""
    from generalizableracing_b200.env import RacingVecEnv
    from generalizableracing_b200.runners import OnPolicyRunner
    agent = {"num_steps_per_env": T, "save_interval": 10 ** 9, "empirical_normalization": False,
             "policy": {"class_name": "ActorCritic", "init_noise_std": 1.0, "actor_hidden_dims": [128, 128], "critic_hidden_dims": [128, 128], "activation": "lrelu"},
             "algorithm": {"class_name": "PPO", "value_loss_coef": 1.0, "use_clipped_value_loss": True, "clip_param": 0.2, "entropy_coef": 0.0,
                           "num_learning_epochs": 5, "num_mini_batches": 4, "learning_rate": 5.0e-4, "schedule": "adaptive", "gamma": 0.99, "lam": 0.95,
                           "desired_kl": 0.01, "max_grad_norm": 1.0}}
    out = {"envs": N, "steps_per_env": T}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fused, graphed, kern, iters in (("fused_kernel_update", True, True, True, 20), ("fused_graphed_update", True, True, False, 10), ("step_by_step", False, False, False, 5)):
        a = json.loads(json.dumps(agent))
        a["fused_collection"] = fused
        a["algorithm"]["graphed_update"], a["algorithm"]["kernel_update"] = graphed, kern
        env = RacingVecEnv(cfg, table, N, device=dev, seed=9)
        runner = OnPolicyRunner(env, a, log_dir=None, device=str(dev))
        runner.learn(3, init_at_random_ep_len=True)          # eager first update, graph capture on the second
        torch.cuda.synchronize(dev)
        e0.record()
        runner.learn(iters)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / iters
        out[name] = {"ms_per_iteration": ms, "env_steps_per_s": N * T / (ms * 1e-3)}
        env.close()
        del env, runner
        torch.cuda.empty_cache()
    out["speedup_kernel_update"] = out["step_by_step"]["ms_per_iteration"] / out["fused_kernel_update"]["ms_per_iteration"]
    return out


def bench_bptt_training(dev, cfg, table, N: int = 16384, H: int = 32):
    """BASELINE C3 as the trainer runs it (AlgoRunner.learn, runner.py:107-155): one BPTT iteration = window forward WITH the
    policy in the loop (BaseModel 16->256->128->4, rsample) + reverse sweep + policy backward + AdamW step.  Fused = one
    gr_bptt_collect launch + gr_step_bwd + one batched actor backward; step-by-step = torch policy + gr_step_fwd per step."""
    import dataclasses
    from generalizableracing_b200.env import RacingVecEnv
    from generalizableracing_b200.runners import AlgoRunner
    dcfg = dataclasses.replace(cfg, is_differentiable_physics=True)
    agent = {"num_steps_per_env": H, "max_iterations": 1000, "save_interval": 10 ** 9, "empirical_normalization": False,
             "algorithm": {"class_name": "BPTT", "schedule": "CosineAnnealingLR", "optimizer": "AdamW", "learning_rate": 5e-4},
             "policy": {"class_name": "BaseModel", "actor_hidden_dims": [256, 128], "critic_hidden_dims": [256, 128], "activation": "lrelu", "init_noise_std": 0.3}}
    out = {"envs": N, "horizon": H}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fused, kbw, iters, li in (("fused_kernel_backward", True, True, 60, 1), ("fused_kernel_backward_log_every_20", True, True, 60, 20),
                                        ("fused", True, False, 30, 1), ("step_by_step", False, False, 5, 1)):
        env = RacingVecEnv(dcfg, table, N, device=dev, seed=5, bptt_horizon=H)
        runner = AlgoRunner(env, {**agent, "fused_collection": fused, "fused_backward_kernel": kbw, "log_interval": li}, log_dir=None, device=str(dev))
        runner.learn(3, init_at_random_ep_len=True)
        torch.cuda.synchronize(dev)
        e0.record()
        runner.learn(iters)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / iters
        out[name] = {"ms_per_iteration": ms, "env_steps_per_s": N * H / (ms * 1e-3)}
        env.close()
        del env, runner
        torch.cuda.empty_cache()
    out["speedup"] = out["step_by_step"]["ms_per_iteration"] / out["fused"]["ms_per_iteration"]
    out["speedup_kernel_backward"] = out["step_by_step"]["ms_per_iteration"] / out["fused_kernel_backward"]["ms_per_iteration"]
    out["note"] = ("fused = gr_bptt_collect + gr_step_bwd + one batched fp32 torch actor backward; fused_kernel_backward = the actor's "
                   "weight gradients from gr_actor_backward (tcgen05) instead; log_every_20 = the same with the runner's log records "
                   "(loss, reward, CUDA-event timings) resolved every 20 iterations instead of every iteration: no host wait in the loop")
    return out


def bench_collection(dev, cfg, table, T: int = 24):
    """PPO collection throughput INCLUDING policy inference and rollout storage (the reference's rollout loop,
    on_policy_runner.py:141-175): fused kernel (gr_ppo_collect: tcgen05 MLPs + env.step + add_transitions, one launch per
    rollout) vs the step-by-step path (torch ActorCritic + gr_step_fwd + gr_storage_add), C2 and C4 env counts."""
    from generalizableracing_b200.algorithms.ppo import PPO
    from generalizableracing_b200.collect import FusedCollector
    from generalizableracing_b200.env import RacingVecEnv
    from generalizableracing_b200.modules import ActorCritic
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for N in (4096, 65536):
        env = RacingVecEnv(cfg, table, N, device=dev, seed=3)
        env.reset()
        env.episode_length_buf = torch.randint(0, cfg.max_episode_length, (N,), device=dev, dtype=torch.int32)
        pol = ActorCritic(16, 16, 4).to(dev)
        alg = PPO(pol, device=dev, gamma=0.99)
        alg.init_storage("rl", N, T, [16], [16], [4])
        col = FusedCollector(env, pol, alg.storage, gamma=0.99)
        col.pack()

        def fused():
            alg.storage.clear()
            col.collect()

        def stepwise():
            alg.storage.clear()
            obs, ex = env.get_observations()
            critic = ex["observations"]["critic"]
            with torch.inference_mode():
                for _ in range(T):
                    a = alg.act(obs, critic)
                    obs, r, d, info = env.step(a)
                    critic = info["observations"]["critic"]
                    alg.process_env_step(r, d, info)

        row = {}
        for name, fn, reps in (("fused", fused, 20), ("step_by_step", stepwise, 3)):
            for _ in range(2):
                fn()
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / reps
            row[name] = {"ms_per_rollout": ms, "us_per_step": ms * 1e3 / T, "env_steps_per_s": N * T / (ms * 1e-3)}
        row["speedup"] = row["step_by_step"]["ms_per_rollout"] / row["fused"]["ms_per_rollout"]
        out[str(N)] = row
        env.close()
        del env, alg, col
        torch.cuda.empty_cache()
    out["note"] = "env-steps/s of a 24-step rollout incl. actor+critic inference, sampling, log-prob, time-out bootstrap and storage rows; fused = 1 launch"
    return out


# ---------------------------------------------------------------------------------------------------------------
def _oracle_env(N, threads, device="cpu"):
    """The reference arm: the oracle = CPU torch restatement of the reference's own implementation of the path
    (the reference itself needs Isaac Sim; /root/reference does not exist on the GPU box)."""
    from generalizableracing_b200 import layout as L_
    from generalizableracing_b200.config import RacingCfg
    from generalizableracing_b200.track_gen import generate_track_table, racing_complex_cfg
    from oracle import racing_oracle as RO
    torch.set_num_threads(threads)
    cfg = RacingCfg.for_stage(STAGE)
    g = torch.Generator().manual_seed(0)
    srnd = torch.rand(N, L_.SRND_STRIDE, generator=g)
    srnd[:, 12:] = torch.randn(N, 4, generator=g)
    env = RO.OracleRacingEnv(cfg, generate_track_table(racing_complex_cfg()), N, srnd.to(device), device=device)

    def draw():
        r = torch.rand(N, L_.RND_STRIDE, generator=g)
        r[:, :8] = torch.randn(N, 8, generator=g)
        return r.to(device)

    env.reset(draw())
    env.episode_length_buf[:] = torch.randint(0, cfg.max_episode_length, (N,), generator=g).to(device)
    return env, draw, g


def torch_gpu_baseline(dev, N: int = NUM_ENVS, steps: int = 20):
    """The reference's implementation style on the SAME GPU: the oracle (torch port of the reference's eager tensor code,
    ~700 aten launches per env.step) on the B200 -- the denominator of the north star's ">= 100x torch-on-GPU" target.
    A reported baseline (bench.py's baseline leg), never part of the product path."""
    env, draw, g = _oracle_env(N, os.cpu_count() or 1, device=dev)
    a = (torch.randn(N, 4, generator=g) * 0.5).to(dev)
    rs = [draw() for _ in range(steps + 3)]
    with torch.no_grad():
        for k in range(3):
            env.step(a, rs[k])
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for k in range(steps):
            env.step(a, rs[3 + k])
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
    return {"value": N * steps / dt, "unit": UNIT, "ms_per_step": dt * 1e3 / steps, "steps": steps,
            "what": f"oracle (torch eager port of the reference) on cuda, {N} envs, STAGE {STAGE}, no_grad, dense pre-drawn randoms"}


def cpu_baseline(sample_s: float = 15.0, N: int = NUM_ENVS):
    threads = os.cpu_count() or 1
    env, draw, g = _oracle_env(N, threads)
    a = torch.randn(N, 4, generator=g) * 0.5
    with torch.no_grad():
        for _ in range(2):
            env.step(a, draw())
        n, t_used = 0, 0.0
        while t_used < sample_s and n < 200:
            r = draw()
            t0 = time.perf_counter()
            env.step(a, r)
            t_used += time.perf_counter() - t0
            n += 1
    return {"value": N * n / t_used, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} steps of the same C4 workload ({N} envs, STAGE {STAGE}) on the oracle (CPU torch port of the reference), "
                      f"{t_used:.1f} s, torch.set_num_threads({threads}), no_grad"}


def cpu_baseline_bptt(N: int = 16384, H: int = 32, windows: int = 2):
    """the reference CPU path for fwd+bwd -- torch.autograd through the oracle over one C3 window (16,384 envs, horizon 32,
    loss = mean of the per-step losses), a bounded sample of `windows` windows after one warm-up window."""
    from generalizableracing_b200 import layout as L_
    from generalizableracing_b200.config import RacingCfg
    from generalizableracing_b200.track_gen import generate_track_table, racing_complex_cfg
    from oracle import racing_oracle as RO
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = RacingCfg.for_stage(STAGE, is_differentiable_physics=True)
    g = torch.Generator().manual_seed(0)
    srnd = torch.rand(N, L_.SRND_STRIDE, generator=g)
    srnd[:, 12:] = torch.randn(N, 4, generator=g)
    env = RO.OracleRacingEnv(cfg, generate_track_table(racing_complex_cfg()), N, srnd)

    def draw():
        r = torch.rand(N, L_.RND_STRIDE, generator=g)
        r[:, :8] = torch.randn(N, 8, generator=g)
        return r

    env.reset(draw())
    env.episode_length_buf[:] = torch.randint(0, cfg.max_episode_length, (N,), generator=g)
    times = []
    for w in range(windows + 1):
        env.detach()
        acts = [(torch.randn(N, 4, generator=g) * 0.5).requires_grad_(True) for _ in range(H)]
        rs = [draw() for _ in range(H)]
        t0 = time.perf_counter()
        losses = [env.step(a, r)[4]["losses"] for a, r in zip(acts, rs)]
        torch.stack(losses).mean().backward()
        if w:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return {"value": N * H / dt, "unit": UNIT, "cores": threads, "kind": "port", "ms_per_window": dt * 1e3,
            "sample": f"{windows} windows of C3 ({N} envs x {H} steps, STAGE {STAGE}) on the oracle with torch.autograd (forward with tape + "
                      f"backward of the mean loss w.r.t. the {H} action tensors), {threads} threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    N = args.envs
    env, draw, g = _oracle_env(N, threads)
    a = torch.randn(N, 4, generator=g) * 0.5
    W, K = max(1, min(args.warmup, 3)), max(1, min(args.steps, 20))
    with torch.no_grad():
        for _ in range(W):
            env.step(a, draw())
        rs = [draw() for _ in range(K)]
        t0 = time.perf_counter()
        for r in rs:
            env.step(a, r)
        dt = time.perf_counter() - t0
    value = N * K / dt
    cfg = env.cfg
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", 1)), "steps": K, "warmup": W,
        "ms_per_step": dt * 1e3 / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C4: {N} envs, STAGE {STAGE}, RacingComplexTerrainCfg gate table (20 types x 10 levels x 8 gates), per-env DR, staggered resets "
                               f"(each step = one full env.step over the {N}-env batch on the host cores)", "envs_per_gpu": N, "mass_kg": cfg.mass},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{K} timed steps, oracle = CPU torch port of the reference path (reference needs Isaac Sim), {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=NUM_ENVS, help="envs per GPU (BASELINE configs[3]: 65536)")
    ap.add_argument("--block-threads", type=int, default=0, help="threads per block of the step kernel (0 = library default)")
    ap.add_argument("--no-stats", action="store_true", help="drop the per-env episode-sum planes (extras['log'] reward terms)")
    ap.add_argument("--prefetch", default=None, choices=["0", "1", "l2"], help="GRACING_PREFETCH for this run (read-mostly planes before the grid dependency)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-collective", action="store_true", help="skip the C5 block (all-reduce + PPO iteration at 65,536 envs per GPU)")
    ap.add_argument("--events-on-stream", action="store_true", help="time one graph launch per repetition (events on the stream) instead of in-graph event nodes")
    ap.add_argument("--repeats", type=int, default=0, help="timed repetitions of the K-step region (0 = ~1.2 s worth, between 100 and 8000)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.prefetch is not None:
        os.environ["GRACING_PREFETCH"] = args.prefetch
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the racing hot path has no CPU fallback (use --impl reference for the CPU arm)")
        run_ours(args)


if __name__ == "__main__":
    main()

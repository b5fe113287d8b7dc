#!/usr/bin/env python
"""Why are back-to-back CUDA-graph replays slow?  Host time per replay vs device time, for several graph sizes."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from generalizableracing_b200 import _lib as B
from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.env import RacingVecEnv
from generalizableracing_b200.tracks import synthetic_track_table
from tools.step_timing import make_io

lib = B.load()
dev = torch.device("cuda:0")
cfg = RacingCfg.for_stage(1)
table = synthetic_track_table()
N = 65536
R = 11
envs = [RacingVecEnv(cfg, table, N, device=dev, seed=1 + r) for r in range(R)]
acts = [torch.randn(N, 4, device=dev) * 0.5 for _ in envs]
for e in envs:
    e.reset()
    e.episode_length_buf = torch.randint(0, cfg.max_episode_length, (N,), device=dev, dtype=torch.int32)
ios = [make_io(e, a) for e, a in zip(envs, acts)]
step = [0]
def launch(k, with_log=True):
    e, io = envs[k], ios[k]
    rng = B.GrRandom(None, e.seed, step[0] & 0xFFFFFFFF); step[0] += 1
    B.check(lib.gr_step_fwd(C.byref(e._gcfg), C.byref(e._track), C.byref(e._state), C.byref(rng), C.byref(io), torch.cuda.current_stream(dev).cuda_stream), "step")
for _ in range(3):
    for k in range(R): launch(k)
torch.cuda.synchronize()
# host cost of an eager launch
t0 = time.perf_counter()
for _ in range(20):
    for k in range(R): launch(k)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(json.dumps({"eager_host_us_per_launch": (t1 - t0) * 1e6 / (20 * R), "eager_total_us_per_launch": (t2 - t0) * 1e6 / (20 * R)}), flush=True)
for rounds in (1, 5, 20):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(rounds):
            for k in range(R): launch(k)
    g.replay(); torch.cuda.synchronize()
    reps = max(2, 40 // rounds)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): g.replay()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    n = reps * rounds * R
    print(json.dumps({"graph_nodes": rounds * R, "reps": reps, "host_us_per_replay": (t1 - t0) * 1e6 / reps, "device_us_per_step": e0.elapsed_time(e1) * 1e3 / n,
                      "wall_us_per_step": (t2 - t0) * 1e6 / n}), flush=True)
# does the log_accum atomic matter?  same graph without log accumulation
for io in ios: io.log_accum = None
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(20):
        for k in range(R): launch(k)
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
print(json.dumps({"no_log_accum_device_us_per_step": e0.elapsed_time(e1) * 1e3 / (2 * 20 * R)}), flush=True)
# clocks while looping
import subprocess, threading
stop = threading.Event(); rows = []
def samp():
    while not stop.is_set():
        rows.append(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()); time.sleep(0.05)
th = threading.Thread(target=samp); th.start()
t_end = time.time() + 1.5
e0.record(); n = 0
while time.time() < t_end:
    g.replay(); n += 1
e1.record(); torch.cuda.synchronize(); stop.set(); th.join()
print(json.dumps({"sustained_1.5s_device_us_per_step": e0.elapsed_time(e1) * 1e3 / (n * 20 * R), "clocks": rows[:3] + rows[-3:]}), flush=True)

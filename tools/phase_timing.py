#!/usr/bin/env python
"""Phase breakdown of gr_step_fwd from in-kernel %globaltimer stamps (debug build of the library, tool only)."""
import ctypes as C, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from generalizableracing_b200 import _lib as B, build as BLD
so = os.path.join(ROOT, "tools", "libgracing_dbg.so")
csrc = os.path.join(ROOT, "generalizableracing_b200", "csrc")
subprocess.run(["nvcc"] + BLD.NVCC_FLAGS + ["-DGR_PHASE_TIMING", "-shared", "-o", so] + [os.path.join(csrc, f) for f in BLD.SOURCES], check=True)
B.LIB_PATH = so
lib = B.load()
from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.env import RacingVecEnv
from generalizableracing_b200.tracks import synthetic_track_table
from tools.step_timing import make_io
hover = int(sys.argv[1]) if len(sys.argv) > 1 else 0
dev = torch.device("cuda:0"); cfg = RacingCfg.for_stage(1); table = synthetic_track_table(); N, R = 65536, 11
envs = [RacingVecEnv(cfg, table, N, device=dev, seed=1 + r) for r in range(R)]
acts = [torch.randn(N, 4, device=dev) * 0.5 for _ in envs]
if hover: acts = [torch.randn(N, 4, device=dev) * 0.1 + torch.tensor([-0.3466, 0., 0., 0.], device=dev) for _ in envs]
for e in envs:
    e.reset(); e.episode_length_buf = torch.randint(0, cfg.max_episode_length, (N,), device=dev, dtype=torch.int32)
ios = [make_io(e, a) for e, a in zip(envs, acts)]
NW = N // 32
times = torch.zeros(R, NW, 5, dtype=torch.int64, device=dev)
for k, io in enumerate(ios): io.phase_times = times[k].data_ptr()
step = [0]
def launch(k):
    e, io = envs[k], ios[k]
    rng = B.GrRandom(None, e.seed, step[0] & 0xFFFFFFFF); step[0] += 1
    B.check(lib.gr_step_fwd(C.byref(e._gcfg), C.byref(e._track), C.byref(e._state), C.byref(rng), C.byref(io), torch.cuda.current_stream(dev).cuda_stream), "step")
for _ in range(5):
    for k in range(R): launch(k)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for k in range(R): launch(k)
for _ in range(20): g.replay()
torch.cuda.synchronize()
t = times.cpu().double()                      # last replay
base = t[:, :, 0].min()
rows = []
for k in range(R):
    tk = t[k] - base
    rows.append(dict(k=k, entry_min=tk[:, 0].min().item(), entry_max=tk[:, 0].max().item(), wait_min=tk[:, 1].min().item(), wait_max=tk[:, 1].max().item(),
                     data_min=tk[:, 2].min().item(), data_med=tk[:, 2].median().item(), data_max=tk[:, 2].max().item(),
                     comp_med=(tk[:, 3] - tk[:, 2]).median().item(), comp_max=(tk[:, 3] - tk[:, 2]).max().item(),
                     store_med=(tk[:, 4] - tk[:, 3]).median().item(), exit_min=tk[:, 4].min().item(), exit_med=tk[:, 4].median().item(), exit_max=tk[:, 4].max().item()))
for r in rows: print(json.dumps({k: (round(v) if isinstance(v, float) else v) for k, v in r.items()}))
tk = t[5] - t[5][:, 1].min()          # one mid-graph step, relative to its wait release
d, c, st_, ex = tk[:, 2], tk[:, 3] - tk[:, 2], tk[:, 4] - tk[:, 3], tk[:, 4]
import numpy as np
q = lambda x, p: float(np.quantile(x.numpy(), p))
print("data arrival quantiles  (10,50,90,99,100):", [round(q(d, p)) for p in (0.1, 0.5, 0.9, 0.99, 1.0)])
print("compute quantiles       (10,50,90,99,100):", [round(q(c, p)) for p in (0.1, 0.5, 0.9, 0.99, 1.0)])
print("exit quantiles          (10,50,90,99,100):", [round(q(ex, p)) for p in (0.1, 0.5, 0.9, 0.99, 1.0)])
late = d >= q(d, 0.9)
print("late-data warps: compute (50,90,100):", [round(q(c[late], p)) for p in (0.5, 0.9, 1.0)], " early-data warps compute (50,90,100):", [round(q(c[d <= q(d, 0.3)], p)) for p in (0.5, 0.9, 1.0)])
last = ex >= q(ex, 0.98)
print("last 2% exits: data", [round(q(d[last], p)) for p in (0.0, 0.5, 1.0)], "compute", [round(q(c[last], p)) for p in (0.0, 0.5, 1.0)], "store", [round(q(st_[last], p)) for p in (0.0, 0.5, 1.0)])
idx = torch.arange(NW)
print("corr(warp index, data arrival) =", float(np.corrcoef(idx.numpy(), d.numpy())[0, 1]), " corr(data, compute) =", float(np.corrcoef(d.numpy(), c.numpy())[0, 1]))
blocks = d.view(-1, 64).mean(dim=1)      # 64 consecutive warps
print("mean data arrival by warp-index decile:", [round(float(x)) for x in d.view(8, -1).mean(dim=1)])
per = [(rows[k + 1]["exit_max"] - rows[k]["exit_max"]) for k in range(R - 1)]
print("exit_max to exit_max per step (ns):", [round(x) for x in per])

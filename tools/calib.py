#!/usr/bin/env python
"""Time the traffic-only calibration kernel in the same rotating-graph harness as bench.py."""
import ctypes as C, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
so = os.path.join(ROOT, "tools", "libcalib.so")
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-shared", "-Xcompiler", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "calib.cu")], check=True)
lib = C.CDLL(so)
lib.calib_launch.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
dev = torch.device("cuda:0")
N, R = 65536, 11
sets = []
for r in range(R):
    sets.append(dict(P=torch.randn(16, N, 4, device=dev), act=torch.randn(N, 4, device=dev), obs=torch.empty(N, 16, device=dev), critic=torch.empty(N, 16, device=dev),
                     rew=torch.empty(N, device=dev), term=torch.empty(N, dtype=torch.uint8, device=dev), to=torch.empty(N, dtype=torch.uint8, device=dev),
                     dones=torch.empty(N, dtype=torch.int64, device=dev)))
def launch(s, block, nload, nstore):
    rc = lib.calib_launch(s["P"].data_ptr(), N, N, s["act"].data_ptr(), s["obs"].data_ptr(), s["critic"].data_ptr(), s["rew"].data_ptr(), s["term"].data_ptr(),
                          s["to"].data_ptr(), s["dones"].data_ptr(), block, nload, nstore, torch.cuda.current_stream(dev).cuda_stream)
    assert rc == 0, rc
for block in (64, 128, 256):
    for nload, nstore in ((16, 9), (16, 0), (1, 9), (1, 0)):
        for s in sets: launch(s, block, nload, nstore)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(5):
                for s in sets: launch(s, block, nload, nstore)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (10 * 5 * R)
        byts = N * (16 + nload * 16 + nstore * 16 + 128 + 14)
        print(json.dumps({"block": block, "nload": nload, "nstore": nstore, "us_per_launch": us, "GBps": byts / us / 1e3}), flush=True)

"""Host cost of one env.step() through the direct ctypes call vs through torch.ops.gracing.step_fwd (ops.py), same kernel."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.env import RacingVecEnv
from generalizableracing_b200.tracks import synthetic_track_table


def main(N=4096, steps=2000):
    out = {}
    for name, op in (("ctypes", False), ("torch_ops", True)):
        env = RacingVecEnv(RacingCfg.for_stage(1), synthetic_track_table(), N, seed=1, op_layer=op)
        env.reset()
        a = torch.randn(N, 4, device="cuda") * 0.3
        for _ in range(50):
            env.step(a)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            env.step(a)
        torch.cuda.synchronize()
        out[name] = {"us_per_step": (time.perf_counter() - t0) / steps * 1e6}
    out["envs"] = N
    print(json.dumps(out))


if __name__ == "__main__":
    main()

"""Generates tests/golden/reach_golden.npz by executing the UNMODIFIED reference LVController / PSController
(L/controllers/controller_diff.py:172-443, loaded by oracle/ref_modules.py) in this container: seeded states / commands in,
(thrust, torque) and d(sum w*out)/d(cmd) out, over 3 chained calls (the thrust low-pass state carries).  The GPU box has no
/root/reference: there tests/test_reach_golden.py checks oracle/reach_oracle.py against these vectors.
    python tests/golden/make_reach_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from generalizableracing_b200.config import ReachTargetCfg  # noqa: E402
from oracle import ref_modules  # noqa: E402


def inputs(N, seed):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(3):
        q = torch.randn(N, 4, generator=g, dtype=torch.float64)
        q = q / q.norm(dim=-1, keepdim=True)
        st = {"pos": torch.randn(N, 3, generator=g, dtype=torch.float64), "quat": q, "lin_vel_w": torch.randn(N, 3, generator=g, dtype=torch.float64) * 2,
              "ang_vel_b": torch.randn(N, 3, generator=g, dtype=torch.float64) * 2}
        cmd = torch.randn(N, 4, generator=g, dtype=torch.float64) * torch.tensor([3.0, 6.0, 6.0, 6.0], dtype=torch.float64)
        w = torch.randn(N, 4, generator=g, dtype=torch.float64)
        out.append((st, cmd, w))
    return out


def main():
    ref = ref_modules.load()
    torch.set_default_dtype(torch.float64)
    N, blob = 64, {}
    for name, cfg in (("LVController", ReachTargetCfg.lv()), ("PSController", ReachTargetCfg.ps())):
        inertia = torch.diag(torch.tensor(cfg.inertia_diag))[None].repeat(N, 1, 1)
        ctl = getattr(ref, name)(ref_modules.outer_loop_cfg(cfg), N, "cpu", cfg.mass, inertia, cfg.step_dt)
        for k, (st, cmd, w) in enumerate(inputs(N, 1234)):
            full = dict(st, ang_vel_w=torch.zeros(N, 3), lin_vel_b=torch.zeros(N, 3))
            c = cmd.clone().requires_grad_(True)
            _, tt = ctl.compute(full, c)
            (tt * w).sum().backward()
            blob[f"{name}_out{k}"] = tt.detach().numpy()
            blob[f"{name}_grad{k}"] = c.grad.numpy()
            ctl.detach()
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "reach_golden.npz"), **blob)
    print("wrote reach_golden.npz", {k: v.shape for k, v in blob.items()})


if __name__ == "__main__":
    main()

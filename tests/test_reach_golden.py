"""oracle/reach_oracle.py against the golden vectors produced by the reference's own LVController / PSController
(tests/golden/make_reach_golden.py): runs everywhere, including the GPU box where /root/reference does not exist."""
import os

import numpy as np
import pytest
import torch

from generalizableracing_b200.config import ReachTargetCfg
from oracle.reach_oracle import OracleLVController
from tests.golden.make_reach_golden import inputs

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reach_golden.npz"))


@pytest.mark.parametrize("name", ["LVController", "PSController"])
def test_outer_loop_oracle_matches_reference_golden(name):
    cfg = ReachTargetCfg.lv() if name == "LVController" else ReachTargetCfg.ps()
    N, dtype = 64, torch.float64
    inertia = torch.diag(torch.tensor(cfg.inertia_diag, dtype=dtype))[None].repeat(N, 1, 1)
    oc = OracleLVController(cfg, N, torch.full((N,), cfg.mass, dtype=dtype), inertia, cfg.step_dt, "cpu", dtype, position_loop=name == "PSController")
    for k, (st, cmd, w) in enumerate(inputs(N, 1234)):
        c = cmd.clone().requires_grad_(True)
        tt = oc.compute(st, c)
        (tt * w).sum().backward()
        assert np.allclose(tt.detach().numpy(), GOLD[f"{name}_out{k}"], rtol=1e-11, atol=1e-11)
        assert np.allclose(c.grad.numpy(), GOLD[f"{name}_grad{k}"], rtol=1e-10, atol=1e-10)
        oc.detach()

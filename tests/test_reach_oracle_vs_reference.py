"""Pins oracle/reach_oracle.py against the UNMODIFIED reference modules that import without Isaac Sim: LVController /
PSController (L/controllers/controller_diff.py:172-443), values and autograd gradients, chained with DroneDynamics over a
few steps.  Runs only where /root/reference exists (this container); the GPU box uses tests/golden/reach_golden.npz."""
import pytest
import torch

from oracle import ref_modules
from oracle.reach_oracle import OracleLVController
from oracle.racing_oracle import OracleDroneDynamics
from generalizableracing_b200.config import ReachTargetCfg

pytestmark = pytest.mark.skipif(not ref_modules.available(), reason="reference tree not present")


def _state(N, g, dtype):
    q = torch.randn(N, 4, generator=g, dtype=dtype)
    q = q / q.norm(dim=-1, keepdim=True)
    return {"pos": torch.randn(N, 3, generator=g, dtype=dtype), "quat": q, "lin_vel_w": torch.randn(N, 3, generator=g, dtype=dtype) * 2,
            "ang_vel_w": torch.randn(N, 3, generator=g, dtype=dtype), "lin_vel_b": torch.randn(N, 3, generator=g, dtype=dtype),
            "ang_vel_b": torch.randn(N, 3, generator=g, dtype=dtype) * 2}


@pytest.mark.parametrize("name", ["LVController", "PSController"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_outer_loop_controllers_match_reference(name, dtype):
    ref = ref_modules.load()
    cfg = ReachTargetCfg.lv() if name == "LVController" else ReachTargetCfg.ps()
    N, dt = 257, cfg.step_dt
    torch.set_default_dtype(dtype)
    try:
        inertia = torch.diag(torch.tensor(cfg.inertia_diag, dtype=dtype))[None].repeat(N, 1, 1)
        # the reference multiplies mass * [N,3] (controller_diff.py:252): only a float (its type hint) broadcasts for N = 257
        rc = getattr(ref, name)(ref_modules.outer_loop_cfg(cfg), N, "cpu", cfg.mass, inertia, dt)
        oc = OracleLVController(cfg, N, torch.full((N,), cfg.mass, dtype=dtype), inertia, dt, "cpu", dtype, position_loop=name == "PSController")
        g = torch.Generator().manual_seed(0)
        for it in range(4):
            st = _state(N, g, dtype)
            scale = torch.tensor([3.0, 6.0, 6.0, 6.0], dtype=dtype)          # large commands: the accel / thrust / rate clamps fire
            c1 = (torch.randn(N, 4, generator=g, dtype=dtype) * scale).requires_grad_(True)
            c2 = c1.detach().clone().requires_grad_(True)
            _, tt_ref = rc.compute(st, c1)
            tt = oc.compute(st, c2)
            tol = 1e-4 if dtype == torch.float32 else 1e-11
            assert torch.allclose(tt, tt_ref, rtol=tol, atol=tol)
            w = torch.randn(N, 4, generator=g, dtype=dtype)
            (tt_ref * w).sum().backward()
            (tt * w).sum().backward()
            assert torch.allclose(c2.grad, c1.grad, rtol=tol * 10, atol=tol * 10)
            rc.detach(), oc.detach()
            if it == 1:
                ids = torch.tensor([0, 5, 100])
                rc.reset_idx(ids), oc.reset_idx(ids)
                assert torch.equal(rc.gross_thrust, oc.gross_thrust)
    finally:
        torch.set_default_dtype(torch.float32)


def test_outer_loop_with_dynamics_gradient_chain():
    """3 chained steps LVController -> DroneDynamics.step/align in fp64: loss value and d loss / d command agree with the reference."""
    ref = ref_modules.load()
    cfg = ReachTargetCfg.lv(random_drag=False)
    N, dt, dtype = 3, cfg.step_dt, torch.float64          # N = 3: the reference's [N] mass broadcast happens to work
    torch.set_default_dtype(dtype)
    try:
        mass = torch.full((N,), cfg.mass)
        inertia = torch.diag(torch.tensor(cfg.inertia_diag))[None].repeat(N, 1, 1)
        rd = ref.DroneDynamics(N, mass, inertia, dt, cfg.decimation, False, "cpu")
        od = OracleDroneDynamics(cfg, N, mass, inertia, dt, "cpu", dtype)
        rc = ref.LVController(ref_modules.outer_loop_cfg(cfg), N, "cpu", cfg.mass, inertia, dt)
        oc = OracleLVController(cfg, N, mass, inertia, dt, "cpu", dtype)
        g = torch.Generator().manual_seed(1)
        s0 = torch.zeros(N, 13)
        s0[:, 3] = 1.0
        s0[:, :3] = torch.randn(N, 3, generator=g)
        ids = torch.arange(N)
        rd.reset_state(s0, ids), od.reset_state(s0, ids)
        cmds = [(torch.randn(N, 4, generator=g) * 2).requires_grad_(True) for _ in range(3)]
        cmds2 = [c.detach().clone().requires_grad_(True) for c in cmds]
        from oracle import isaac_math as M
        losses = []
        for dynm, ctl, cs in ((rd, rc, cmds), (od, oc, cmds2)):
            loss = 0.0
            for c in cs:
                st = {"pos": dynm.pos.detach(), "quat": dynm.quat.detach(), "lin_vel_w": dynm.lin_vel_w.detach(), "ang_vel_w": dynm.ang_vel_w.detach(),
                      "lin_vel_b": dynm.lin_vel_b.detach(), "ang_vel_b": dynm.ang_vel_b.detach()}
                out = ctl.compute(st, c)
                tt = out[1] if isinstance(out, tuple) else out
                nom, _a = dynm.step(tt)
                al = dynm.align(nom.detach(), nom)
                loss = loss + al[:, :3].norm(dim=-1).sum() + 0.3 * (al[:, 7:10].norm(dim=-1) + 0.5 * al[:, 10:13].norm(dim=-1)).sum()
            loss.backward()
            losses.append(loss.detach())
        assert torch.allclose(losses[0], losses[1], rtol=1e-12)
        for a, b in zip(cmds, cmds2):
            assert torch.allclose(a.grad, b.grad, rtol=1e-9, atol=1e-12)
    finally:
        torch.set_default_dtype(torch.float32)

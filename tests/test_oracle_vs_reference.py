"""Pins the oracle restatement against the UNMODIFIED reference modules that import without Isaac Sim
(droneDynamics.py, controller_diff.py, rollout_storage.py), executed where they lie under /root/reference.
Skipped on the GPU box (the reference tree does not travel); tests/test_golden.py covers the same ground there
with vectors these modules generated (tests/golden/make_golden.py)."""
import math

import numpy as np
import pytest
import torch

from generalizableracing_b200.config import RacingCfg
from oracle import isaac_math as M
from oracle import racing_oracle as RO
from oracle import ref_modules, rollout_oracle

pytestmark = pytest.mark.skipif(not ref_modules.available(), reason="reference tree not present")


def _pair(N, cfg):
    ns = ref_modules.load()
    mass = torch.full((N,), cfg.mass)
    J = torch.tensor([[cfg.inertia_diag[0], 0, 0, 0, cfg.inertia_diag[1], 0, 0, 0, cfg.inertia_diag[2]]]).reshape(1, 3, 3).repeat(N, 1, 1)
    ref = (ns.DroneDynamics(N, mass, J, cfg.step_dt, cfg.decimation, True, "cpu"),
           ns.CTBRController(ref_modules.ctbr_cfg(cfg), N, "cpu", mass, J, cfg.step_dt))
    orc = (RO.OracleDroneDynamics(cfg, N, mass, J, cfg.step_dt, "cpu", torch.float32), RO.OracleCTBRController(cfg, N, J, cfg.step_dt, "cpu", torch.float32))
    return ref, orc


def test_constants_match_reference_files():
    import yaml
    cfg = RacingCfg.for_stage(1)
    y = yaml.safe_load(open(ref_modules.REF_ROOT + "/extensions/diff.lab_tasks/diff/lab_tasks/tasks/quadcopter_diff/mdp/dynamics/dynamics.yaml"))
    assert y["grad_decay_factor"] == cfg.grad_decay and y["g"] == cfg.gravity
    assert y["drag_1_coeffs"] == [cfg.drag_1] * 3 and y["drag_2_coeffs"] == [cfg.drag_2] * 3
    assert (y["drag_1_randomness"], y["drag_2_randomness"], y["z_drag_coeff"], y["z_drag_randomness"]) == (cfg.drag_1_rand, cfg.drag_2_rand, cfg.z_drag, cfg.z_drag_rand)
    src = open(ref_modules.REF_ROOT + "/extensions/diff.lab_tasks/diff/lab_tasks/tasks/quadcopter_diff/racing_ctbr_env.py").read()
    for needle in ("rate_gain_p=[35, 35, 35]", "rate_gain_d=[0.0005, 0.0005, 0.0003]", "body_rate_bound=[-6, 6]", "update_threshold=0.35",
                   "self.decimation = 3", "self.sim.dt = 0.01", '"thrust_delay_scale_factor": (0.8, 1.3)', '"pid_scale_factor": (0.9, 1.1)'):
        assert needle in src, needle
    (ref_dyn, ref_ctl), _ = _pair(2, cfg)
    assert tuple(ref_ctl.gross_thrust_bound) == cfg.gross_thrust_bound
    assert cfg.max_episode_length == math.ceil(6.0 / (0.01 * 3)) == 200 and RacingCfg.for_stage(2).max_episode_length == 267


def test_dynamics_controller_bit_exact_with_reference():
    cfg = RacingCfg.for_stage(1)
    N = 64
    (ref_dyn, ref_ctl), (o_dyn, o_ctl) = _pair(N, cfg)
    g = torch.Generator().manual_seed(0)
    st = torch.randn(N, 13, generator=g)
    st[:, 3:7] = torch.nn.functional.normalize(st[:, 3:7], dim=-1)
    for d in (ref_dyn, o_dyn):
        d.reset_state(st.clone(), torch.arange(N))
    for t in range(60):
        cmd = torch.randn(N, 4, generator=g) * torch.tensor([10.0, 8, 8, 8]) + torch.tensor([12.0, 0, 0, 0])
        now = {k: getattr(ref_dyn, k).detach() for k in ("pos", "quat", "lin_vel_w", "ang_vel_w", "lin_vel_b", "ang_vel_b")}
        now.update(lin_acc_w=torch.zeros(N, 3), ang_acc_w=torch.zeros(N, 3), lin_acc_b=torch.zeros(N, 3), ang_acc_b=torch.randn(N, 3, generator=g))
        _, tt_r = ref_ctl.compute(now, cmd)
        tt_o = o_ctl.compute(now, cmd)
        assert torch.equal(tt_r, tt_o)
        (nr, ar), (no, ao) = ref_dyn.step(tt_r), o_dyn.step(tt_o)
        assert torch.equal(nr, no) and torch.equal(ar, ao)
        assert torch.equal(ref_dyn.align(nr.detach(), nr), o_dyn.align(no.detach(), no))
        if t % 17 == 5:                                        # drag re-draw + reset on a subset (same torch.rand stream)
            idx = torch.arange(0, N, 3)
            torch.manual_seed(t)
            ref_dyn.reset_idx(idx)
            torch.manual_seed(t)
            u_z, u_d2, u_d1 = torch.rand(len(idx)), torch.rand(len(idx), 3), torch.rand(len(idx), 3)
            o_dyn.reset_idx(idx, u_z, u_d2, u_d1)
            assert torch.equal(ref_dyn.drag_coeffs, o_dyn.drag_coeffs) and torch.equal(ref_dyn.h_force_drag_coeffs, o_dyn.h_force_drag_coeffs)
            ref_ctl.reset_idx(idx)
            o_ctl.reset_idx(idx)
            assert torch.equal(ref_ctl.torque, o_ctl.torque) and torch.equal(ref_ctl.gross_thrust, o_ctl.gross_thrust)


def test_bptt_gradient_bit_exact_with_reference():
    """32-step window, autograd through the reference modules vs through the oracle (same graph => same bits)."""
    cfg = RacingCfg.for_stage(0)
    N, H = 16, 32
    (ref_dyn, ref_ctl), (o_dyn, o_ctl) = _pair(N, cfg)
    g = torch.Generator().manual_seed(3)
    st = torch.randn(N, 13, generator=g) * 0.3
    st[:, 3:7] = torch.nn.functional.normalize(torch.randn(N, 4, generator=g), dim=-1)
    grads = []
    for dyn, ctl, is_ref in ((ref_dyn, ref_ctl, True), (o_dyn, o_ctl, False)):
        gg = torch.Generator().manual_seed(4)
        dyn.reset_state(st.clone(), torch.arange(N))
        cmds = [(torch.randn(N, 4, generator=gg) * 3 + torch.tensor([10.0, 0, 0, 0])).requires_grad_(True) for _ in range(H)]
        loss = 0
        for c in cmds:
            now = {k: getattr(dyn, k).detach() for k in ("pos", "quat", "lin_vel_w", "ang_vel_w", "lin_vel_b", "ang_vel_b")}
            now.update(lin_acc_w=torch.zeros(N, 3), ang_acc_w=torch.zeros(N, 3), lin_acc_b=torch.zeros(N, 3), ang_acc_b=torch.zeros(N, 3))
            tt = ctl.compute(now, c)
            tt = tt[1] if is_ref else tt
            nom, _ = dyn.step(tt)
            al = dyn.align(nom.detach(), nom)
            loss = loss + (al[:, :3].norm(dim=-1) + 0.05 * (al[:, 7:10] ** 2).mean(-1)).mean()
        loss.backward()
        grads.append(torch.stack([c.grad for c in cmds]))
    assert torch.equal(grads[0], grads[1])
    assert float(grads[0].abs().max()) > 0


@pytest.mark.parametrize("T,N", [(24, 4096), (5, 33)])
def test_gae_bit_exact_with_reference(T, N):
    ns = ref_modules.load()
    g = torch.Generator().manual_seed(T + N)
    sto = ns.RolloutStorage("rl", N, T, [16], [16], [4], "cpu")
    sto.rewards = torch.randn(T, N, 1, generator=g)
    sto.values = torch.randn(T, N, 1, generator=g)
    sto.dones = (torch.rand(T, N, 1, generator=g) < 0.02).byte()
    last = torch.randn(N, 1, generator=g)
    sto.compute_returns(last, 0.99, 0.95)
    ret, adv = rollout_oracle.compute_returns(sto.rewards, sto.values, sto.dones, last, 0.99, 0.95)
    assert torch.equal(ret, sto.returns) and torch.equal(adv, sto.advantages)
    # mini-batch gather with the same permutation
    torch.manual_seed(1)
    ref_batches = list(sto.mini_batch_generator(4, 2))
    torch.manual_seed(1)
    idx = torch.randperm(4 * (T * N // 4))
    fields = [sto.observations, sto.privileged_observations, sto.actions, sto.values, sto.advantages, sto.returns, sto.actions_log_prob, sto.mu, sto.sigma]
    for rb, ob in zip(ref_batches, rollout_oracle.mini_batches(fields, idx, 4, 2)):
        for a, b in zip(rb[:9], ob):
            assert torch.equal(a, b)


def test_isaac_math_identities():
    """Appendix-B restatements: quat_mul == Hamilton product, rotate/rotate_inverse are inverse maps on unit quats,
    euler round trip, wrap_to_pi range; track-table euler->quat rule == scipy (terrain_generator.py:69-73)."""
    g = torch.Generator().manual_seed(0)
    q = torch.nn.functional.normalize(torch.randn(500, 4, generator=g, dtype=torch.float64), dim=-1)
    p = torch.nn.functional.normalize(torch.randn(500, 4, generator=g, dtype=torch.float64), dim=-1)
    v = torch.randn(500, 3, generator=g, dtype=torch.float64)
    w1, x1, y1, z1 = q.unbind(-1)
    w2, x2, y2, z2 = p.unbind(-1)
    ham = torch.stack([w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2, w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                       w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2, w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2], -1)
    assert torch.allclose(M.quat_mul(q, p), ham, atol=1e-13)
    assert torch.allclose(M.quat_rotate_inverse(q, M.quat_rotate(q, v)), v, atol=1e-12)
    assert torch.allclose(M.quat_rotate(q, v), torch.einsum("nij,nj->ni", M.matrix_from_quat(q), v), atol=1e-12)
    e = (torch.rand(500, 3, generator=g, dtype=torch.float64) - 0.5) * torch.tensor([6.0, 3.0, 6.0])
    r, pt, yw = M.euler_xyz_from_quat(M.quat_from_euler_xyz(e[:, 0], e[:, 1], e[:, 2]))
    assert torch.allclose(M.wrap_to_pi(r), e[:, 0], atol=1e-9) and torch.allclose(M.wrap_to_pi(pt), e[:, 1], atol=1e-9) and torch.allclose(M.wrap_to_pi(yw), e[:, 2], atol=1e-9)
    from scipy.spatial.transform import Rotation as R
    from generalizableracing_b200.tracks import gate_euler_to_quat_wxyz
    ge = np.random.default_rng(0).uniform(-180, 180, (64, 3))
    sq = (R.from_euler("YXZ", np.stack([ge[:, 0], -ge[:, 1], ge[:, 2]], 1), degrees=True) * R.from_euler("XYZ", [-90, -90, 0], degrees=True)).as_quat()
    mine = gate_euler_to_quat_wxyz(ge)
    ref = sq[:, [3, 0, 1, 2]]
    ref = np.where(ref[:, :1] < 0, -ref, ref)
    assert np.abs(mine - ref).max() < 1e-12

#!/usr/bin/env python
"""Generates the committed golden vectors.  Run in the build container (needs /root/reference):
    PYTHONPATH=. python tests/golden/make_golden.py
(a) ref_dyn_ctrl.pt  -- UNMODIFIED reference DroneDynamics + CTBRController: 32-step trajectory + autograd gradients
(b) ref_gae.pt       -- UNMODIFIED reference RolloutStorage.compute_returns / mini_batch_generator
(c) closure_c1.pt    -- the PhysX-free closure (oracle; the reference cannot run it without Isaac Sim): BASELINE C1,
                        64 envs, fixed figure-8 track, STAGE 0, 300 steps; regression pin of the restatement itself.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from generalizableracing_b200 import layout as L_  # noqa: E402
from generalizableracing_b200.config import RacingCfg  # noqa: E402
from generalizableracing_b200.tracks import figure_eight_track  # noqa: E402
from oracle import racing_oracle as RO, ref_modules  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def dyn_ctrl():
    ns = ref_modules.load()
    cfg = RacingCfg.for_stage(1)
    N, H = 16, 32
    mass = torch.full((N,), cfg.mass)
    J = torch.tensor([[cfg.inertia_diag[0], 0, 0, 0, cfg.inertia_diag[1], 0, 0, 0, cfg.inertia_diag[2]]]).reshape(1, 3, 3).repeat(N, 1, 1)
    dyn = ns.DroneDynamics(N, mass, J, cfg.step_dt, cfg.decimation, True, "cpu")
    ctl = ns.CTBRController(ref_modules.ctbr_cfg(cfg), N, "cpu", mass, J, cfg.step_dt)
    g = torch.Generator().manual_seed(0)
    st = torch.randn(N, 13, generator=g) * 0.5
    st[:, 3:7] = torch.nn.functional.normalize(torch.randn(N, 4, generator=g), dim=-1)
    dyn.reset_state(st.clone(), torch.arange(N))
    cmds = [(torch.randn(N, 4, generator=g) * 4 + torch.tensor([10.0, 0, 0, 0])).requires_grad_(True) for _ in range(H)]
    acc_b = torch.randn(H, N, 3, generator=g)
    traj, tts, loss = [], [], 0
    for t, c in enumerate(cmds):
        now = {k: getattr(dyn, k).detach() for k in ("pos", "quat", "lin_vel_w", "ang_vel_w", "lin_vel_b", "ang_vel_b")}
        now.update(lin_acc_w=torch.zeros(N, 3), ang_acc_w=torch.zeros(N, 3), lin_acc_b=torch.zeros(N, 3), ang_acc_b=acc_b[t])
        _, tt = ctl.compute(now, c)
        nom, a = dyn.step(tt)
        al = dyn.align(nom.detach(), nom)
        traj.append(torch.cat([nom.detach(), a.detach()], dim=-1))
        tts.append(tt.detach())
        loss = loss + (al[:, :3].norm(dim=-1) + 0.05 * (al[:, 7:10] ** 2).mean(-1) + 0.5 / (1 + al[:, 2] + 10 * al[:, 2] ** 2)).mean()
    loss.backward()
    torch.save({"state0": st, "cmds": torch.stack([c.detach() for c in cmds]), "ang_acc_b": acc_b, "traj": torch.stack(traj),
                "thrust_torque": torch.stack(tts), "loss": loss.detach(), "grad_cmds": torch.stack([c.grad for c in cmds])},
               os.path.join(OUT, "ref_dyn_ctrl.pt"))


def gae():
    ns = ref_modules.load()
    T, N = 24, 256
    g = torch.Generator().manual_seed(1)
    sto = ns.RolloutStorage("rl", N, T, [16], [16], [4], "cpu")
    sto.rewards = torch.randn(T, N, 1, generator=g)
    sto.values = torch.randn(T, N, 1, generator=g)
    sto.dones = (torch.rand(T, N, 1, generator=g) < 0.02).byte()
    last = torch.randn(N, 1, generator=g)
    sto.compute_returns(last, 0.99, 0.95)
    torch.save({"rewards": sto.rewards, "values": sto.values, "dones": sto.dones, "last_values": last, "gamma": 0.99, "lam": 0.95,
                "returns": sto.returns, "advantages": sto.advantages}, os.path.join(OUT, "ref_gae.pt"))


def closure_c1():
    cfg = RacingCfg.for_stage(0, is_differentiable_physics=True)
    N, T = 64, 300
    g = torch.Generator().manual_seed(0)
    srnd = torch.rand(N, L_.SRND_STRIDE, generator=g)
    srnd[:, 12:] = torch.randn(N, 4, generator=g)
    env = RO.OracleRacingEnv(cfg, figure_eight_track(), N, srnd)

    def draw():
        r = torch.rand(N, L_.RND_STRIDE, generator=g)
        r[:, :8] = torch.randn(N, 8, generator=g)
        return r

    env.reset(draw())
    rew, masks, obs_ck, loss = [], [], [], []
    with torch.no_grad():
        for t in range(T):
            a = torch.randn(N, 4, generator=g) * 0.5
            o, r, term, to, ex = env.step(a, draw())
            rew.append(r.clone())
            masks.append(torch.stack([term, to]).clone())
            obs_ck.append(torch.stack([o["policy"].sum(-1), o["critic"].sum(-1)]))
            loss.append(ex["losses"].clone())
    torch.save({"seed": 0, "N": N, "T": T, "rewards": torch.stack(rew), "masks": torch.stack(masks), "obs_checksums": torch.stack(obs_ck),
                "losses": torch.stack(loss), "final_root_state": env._root_state_w(), "final_gate_id": env.gate_id.clone(),
                "final_episode_length": env.episode_length_buf.clone(), "final_obs": o["policy"].clone()}, os.path.join(OUT, "closure_c1.pt"))


if __name__ == "__main__":
    dyn_ctrl()
    gae()
    closure_c1()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(OUT, f)))

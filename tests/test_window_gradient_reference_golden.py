"""d mean(loss) / d action over a 16-step differentiable window WITH mid-window resets, recorded from the reference's OWN autograd
graph through its unmodified env step (tests/golden/make_ref_env_golden.py::bptt_window, S/diff_rl/algorithms/bptt.py:38-44),
replayed where the reference tree is absent: through torch.autograd on the oracle (CPU, 1e-6) and through the kernels' analytic
reverse sweep (``emul`` here, ``cuda`` = libgracing.so through the C ABI on the B200; 1e-4 of the largest entry, the tolerance of
tests/test_bptt_parity.py)."""
import os

import pytest
import torch

from generalizableracing_b200 import layout as L_
from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.env import RacingVecEnv
from generalizableracing_b200.tracks import figure_eight_track
from oracle import racing_oracle as RO
from tests import parity_cases as PC
from tests.conftest import backend_params

G = os.path.join(os.path.dirname(__file__), "golden", "ref_env_closure.pt")


def _load():
    return torch.load(G)["bptt_window"], RacingCfg.for_stage(0, is_differentiable_physics=True), figure_eight_track()


def test_oracle_autograd_matches_reference_window_gradient():
    d, cfg, table = _load()
    N, H = d["N"], d["H"]
    orc = RO.OracleRacingEnv(cfg, table, N, d["startup_rnd"])
    orc.reset(d["rnd0"])
    orc.root_pos_w = d["pos0"].clone()
    orc._get_state_from_sim()
    orc.dyn.reset_state(orc.states_all, torch.arange(N))
    orc.episode_length_buf[:] = d["episode_length0"]
    orc.detach()
    acts = [d["actions"][t].clone().requires_grad_(True) for t in range(H)]
    losses = [orc.step(acts[t], d["rnd"][t])[4]["losses"] for t in range(H)]
    assert torch.equal(torch.stack(losses).detach(), d["losses"])
    torch.stack(losses).mean().backward()
    grad = torch.stack([a.grad if a.grad is not None else torch.zeros(N, 4) for a in acts])
    assert d["resets"] > 0 and float((grad - d["grad_actions"]).abs().max() / d["grad_actions"].abs().max()) < 1e-6


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_kernel_reverse_sweep_matches_reference_window_gradient(backend):
    device, lib = backend
    d, cfg, table = _load()
    N, H = d["N"], d["H"]
    env = RacingVecEnv(cfg, table, N, device=device, rng_mode="dense", startup_rnd=d["startup_rnd"], bptt_horizon=H, _lib=lib)
    env.reset(d["rnd0"].to(device))
    env.write_plane(L_.PL_POS, slice(0, 3), d["pos0"])
    env.episode_length_buf = d["episode_length0"]
    env.detach()
    worst = 0.0
    for t in range(H):
        ex = env.step(d["actions"][t].to(device), d["rnd"][t].to(device))[3]
        worst = max(worst, PC.rel_err(d["losses"][t], ex["losses"]))
    grad = env._bptt.backward_window().cpu()
    err = float((grad - d["grad_actions"]).abs().max() / d["grad_actions"].abs().max())
    print("window losses rel err", worst, "gradient err / max entry", err)
    assert worst < PC.REL_TOL_STEP and err < 1e-4


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_observations_handed_to_a_differentiated_policy_stay_valid(backend):
    """Differentiable mode: the observation returned by reset / get_observations / step(actions requiring grad) is not one of the
    ping-pong output buffers the kernels rewrite two steps later (autograd would have saved it for the policy's weight gradients)."""
    device, lib = backend
    d, cfg, table = _load()
    N = d["N"]
    env = RacingVecEnv(cfg, table, N, device=device, rng_mode="dense", startup_rnd=d["startup_rnd"], bptt_horizon=8, _lib=lib)
    obs = [env.reset(d["rnd0"].to(device))[0], env.get_observations()[0]]
    snap = [o.clone() for o in obs]
    for t in range(5):
        o = env.step(d["actions"][t].to(device).requires_grad_(True), d["rnd"][t].to(device))[0]
        obs.append(o)
        snap.append(o.clone())
    assert len({o.data_ptr() for o in obs}) == len(obs)
    for o, s0 in zip(obs, snap):
        assert torch.equal(o, s0)
    # without a differentiated policy in the loop the step keeps handing out its own buffers (no extra copy on the hot path)
    o1 = env.step(d["actions"][5].to(device), d["rnd"][5].to(device))[0]
    o2 = env.step(d["actions"][6].to(device), d["rnd"][6].to(device))[0]
    o3 = env.step(d["actions"][7].to(device), d["rnd"][7].to(device))[0]
    assert o1.data_ptr() == o3.data_ptr() != o2.data_ptr()

"""Analytic backward of the racing step vs torch.autograd on the oracle, BASELINE config C3
shape (horizon 32), resets inside the window, both entry points: the one-launch window sweep and the chained
autograd.Function path with arbitrary upstream weights."""
import pytest
import torch

from tests.conftest import backend_params
from tests import parity_cases as PC

pytestmark = pytest.mark.timeout(900)
# gradients: 1e-5 relative per step accumulated over the horizon (reported); bound 1e-4 of the largest gradient entry
GRAD_TOL = 1e-4


def _window(backend, stage, N, H, seed, weighted, use_autograd):
    cfg, table, orc, env, g = PC.make_pair(backend, stage=stage, N=N, seed=seed, diff=True, horizon=H)
    dev = env.device
    r0 = PC.draw_rnd(N, g)
    orc.reset(r0)
    env.reset(r0.to(dev))
    ep = torch.randint(0, cfg.max_episode_length, (N,), generator=g)      # staggered: time-outs fall inside the window
    orc.episode_length_buf[:] = ep
    env.episode_length_buf = ep
    for _ in range(3):                                                      # a few un-differentiated steps first
        a, r = torch.randn(N, 4, generator=g) * 0.5, PC.draw_rnd(N, g)
        with torch.no_grad():
            orc.step(a, r)
        env.step(a.to(dev), r.to(dev))
    orc.detach()
    env.detach()
    acts = [(torch.randn(N, 4, generator=g) * 0.5).requires_grad_(True) for _ in range(H)]
    acts_k = [a.detach().clone().to(dev).requires_grad_(True) for a in acts]
    ol, kl, nreset = [], [], 0
    for t in range(H):
        r = PC.draw_rnd(N, g)
        _, _, term, to, ex = orc.step(acts[t], r)
        ol.append(ex["losses"])
        nreset += int((term | to).sum())
        kl.append(env.step(acts_k[t], r.to(dev))[3]["losses"])
    loss_err = PC.rel_err(torch.stack(ol), torch.stack(kl))
    w = torch.rand(H, N, generator=g) if weighted else torch.full((H, N), 1.0 / (H * N))
    (torch.stack(ol) * w).sum().backward()
    ref = torch.stack([a.grad if a.grad is not None else torch.zeros_like(a) for a in acts])
    if use_autograd:
        tot = 0
        for t in range(H):                                                   # accumulate-as-you-go graph shape
            tot = tot + (kl[t] * w[t].to(dev)).sum()
        tot.backward()
        got = torch.stack([a.grad if a.grad is not None else torch.zeros_like(a) for a in acts_k]).cpu()
    elif weighted:
        got = env._bptt.backward_window(grad_losses=w.to(dev)).cpu()
    else:
        got = env._bptt.backward_window().cpu()                             # BPTT.update: uniform 1/(T*N)
    err = float((ref - got).abs().max() / ref.abs().max())
    return nreset, loss_err, err, ref, got


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
@pytest.mark.parametrize("stage,N,weighted", [(0, 64, False), (1, 160, True)])
def test_window_sweep_matches_autograd(backend, stage, N, weighted):
    H = 32
    nreset, loss_err, err, ref, got = _window(backend, stage, N, H, seed=21 + stage, weighted=weighted, use_autograd=False)
    print(f"stage {stage}: resets in window {nreset}, loss rel err {loss_err:.2e}, grad rel err over H={H}: {err:.2e}")
    assert nreset > 0
    assert loss_err < PC.REL_TOL_STEP * 10
    assert err < GRAD_TOL
    assert torch.all(got[-1] == 0)          # the last action of a window only acts in the next one (1-step lag)


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_chained_autograd_path(backend):
    nreset, loss_err, err, ref, got = _window(backend, 1, 96, 16, seed=33, weighted=True, use_autograd=True)
    assert nreset > 0 and err < GRAD_TOL


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_second_window_after_detach(backend):
    """env.unwrapped.detach() starts a new window: adjoints and tape restart, gradients stay exact."""
    cfg, table, orc, env, g = PC.make_pair(backend, stage=0, N=64, seed=5, diff=True, horizon=8)
    dev = env.device
    r0 = PC.draw_rnd(64, g)
    orc.reset(r0)
    env.reset(r0.to(dev))
    for w in range(3):
        orc.detach()
        env.detach()
        acts = [(torch.randn(64, 4, generator=g) * 0.5).requires_grad_(True) for _ in range(8)]
        ol = []
        for a in acts:
            r = PC.draw_rnd(64, g)
            ol.append(orc.step(a, r)[4]["losses"])
            env.step(a.detach().to(dev), r.to(dev))
        torch.stack(ol).mean().backward()
        ref = torch.stack([a.grad if a.grad is not None else torch.zeros_like(a) for a in acts])
        got = env._bptt.backward_window().cpu()
        assert float((ref - got).abs().max() / ref.abs().max()) < GRAD_TOL
    with pytest.raises(RuntimeError):
        for _ in range(9):
            env.step(torch.zeros(64, 4, device=dev), PC.draw_rnd(64, g).to(dev))


@pytest.mark.gpu
@pytest.mark.parametrize("N", [64, 130, 4096])
def test_two_lane_sweep_equals_one_lane_sweep(cuda_lib, N):
    """racing_step_bwd2_kernel (translational / rotational half of an env on two lanes, nine shuffles per step) against the one-lane
    kernel that the emulation runs: the same device functions in the same order, so the gradients agree to the last bit or two (the two
    kernels are compiled separately: fused-multiply-add contraction may differ) -- and both match autograd through the oracle."""
    H = 32
    cfg, table, orc, env, g = PC.make_pair(("cuda:0", None), stage=1, N=N, seed=40 + N, diff=True, horizon=H)
    dev = env.device
    r0 = PC.draw_rnd(N, g)
    env.reset(r0.to(dev))
    env.episode_length_buf = torch.randint(0, cfg.max_episode_length, (N,), generator=g)
    env.detach()
    env._bptt.autograd = False
    acts = (torch.randn(H, N, 4, generator=g) * 0.5).to(dev)
    rnd = torch.stack([PC.draw_rnd(N, g) for _ in range(H)]).to(dev)
    env.rollout(acts, rnd)
    w = torch.rand(H, N, generator=g).to(dev)
    out = {}
    for lanes in (1, 2):
        env._bptt.lanes = lanes
        out[lanes] = (env._bptt.backward_window(grad_losses=w).clone(), env._bptt.backward_window().clone(), env._bptt.adjoint.clone())
    for a, b in zip(out[1], out[2]):
        assert float(a.abs().max()) > 0 or a is out[1][2]
        assert float((a - b).abs().max()) <= 2e-6 * max(float(a.abs().max()), 1e-30)
    # chained launches (one step per launch, adjoints carried through HBM between them) == one sweep, on the two-lane kernel
    env._bptt.lanes = 2
    env._bptt.adjoint.zero_()
    env._bptt.grad_loss[:H].copy_(w)
    for t in range(H - 1, -1, -1):
        env._bptt._launch(t, t + 1, env._bptt.grad_loss, 0.0)
    assert float((env._bptt.grad_action[:H] - out[2][0]).abs().max()) <= 2e-6 * float(out[2][0].abs().max())

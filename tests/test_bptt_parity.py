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


"""gr_ppo_loss_grad against torch.autograd on the reference's loss expression (standalone/rsl_rl/ext/algorithms/ppo.py:143-171),
for every option (clipped / plain value loss, entropy bonus), and gr_policy_forward against the fp32 modules."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,clipped,entropy_coef", [(1000, True, 0.0), (24576, True, 0.005), (777, False, 0.01)])
def test_loss_gradients_match_autograd(cuda_lib, rows, clipped, entropy_coef):
    from generalizableracing_b200 import _lib as B
    lib = cuda_lib
    g = torch.Generator(device="cuda").manual_seed(rows)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    mu, v = rn(rows, 4).requires_grad_(True), rn(rows).requires_grad_(True)
    std = (0.5 + torch.rand(4, device="cuda", generator=g)).requires_grad_(True)
    old_mu, old_sigma = mu.detach() + 0.3 * rn(rows, 4), (std.detach() * (1 + 0.1 * rn(4))).abs().expand(rows, 4).contiguous()
    actions = old_mu + old_sigma * rn(rows, 4)
    old_logp = torch.distributions.Normal(old_mu, old_sigma).log_prob(actions).sum(-1)
    adv, ret, old_v = rn(rows), rn(rows), v.detach() + 0.3 * rn(rows)
    clip, vcoef = 0.2, 1.0
    # --- reference expression
    dist = torch.distributions.Normal(mu, std.expand_as(mu))
    logp = dist.log_prob(actions).sum(-1)
    ratio = torch.exp(logp - old_logp)
    surrogate = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 1 - clip, 1 + clip)).mean()
    if clipped:
        vc = old_v + (v - old_v).clamp(-clip, clip)
        vloss = torch.max((v - ret).pow(2), (vc - ret).pow(2)).mean()
    else:
        vloss = (ret - v).pow(2).mean()
    loss = surrogate + vcoef * vloss - entropy_coef * dist.entropy().sum(-1).mean()
    loss.backward()
    kl = torch.sum(torch.log(std.detach() / old_sigma + 1e-5) + (old_sigma ** 2 + (old_mu - mu.detach()) ** 2) / (2 * std.detach() ** 2) - 0.5, dim=-1).mean()
    # --- kernel
    gm, gv, sums = torch.zeros(rows, 4, device="cuda"), torch.zeros(rows, 4, device="cuda"), torch.zeros(16, device="cuda")
    sig = std.detach().contiguous()
    b = B.GrPpoBatch(mu.data_ptr(), v.data_ptr(), sig.data_ptr(), actions.data_ptr(), old_logp.data_ptr(), adv.data_ptr(), ret.data_ptr(), old_v.data_ptr(),
                     old_mu.data_ptr(), old_sigma.data_ptr(), clip, vcoef, entropy_coef, int(clipped))
    B.check(lib.gr_ppo_loss_grad(C.byref(b), rows, gm.data_ptr(), gv.data_ptr(), sums.data_ptr(), torch.cuda.current_stream().cuda_stream), "gr_ppo_loss_grad")
    torch.cuda.synchronize()
    assert torch.allclose(gm, mu.grad, rtol=2e-4, atol=1e-9 + 1e-5 * float(mu.grad.abs().max()))
    assert torch.allclose(gv[:, 0], v.grad, rtol=2e-4, atol=1e-9 + 1e-5 * float(v.grad.abs().max())) and float(gv[:, 1:].abs().max()) == 0.0
    assert torch.allclose(sums[3:7], std.grad, rtol=2e-3, atol=1e-5 * float(std.grad.abs().max()) + 1e-7)
    assert int(sums[7]) == rows
    assert abs(float(sums[0]) / rows - float(surrogate)) < 1e-5 + 1e-4 * abs(float(surrogate))
    assert abs(float(sums[1]) / rows - float(vloss)) < 1e-4 * abs(float(vloss))
    assert abs(float(sums[2]) / rows - float(kl)) < 1e-4 * abs(float(kl)) + 1e-6


@pytest.mark.parametrize("rows", [100, 4096, 24576 + 5])
def test_policy_forward_matches_modules(cuda_lib, rows):
    from generalizableracing_b200 import _lib as B
    from generalizableracing_b200.modules import ActorCritic
    lib = cuda_lib
    torch.manual_seed(rows)
    pol = ActorCritic(16, 16, 4).cuda()
    la, lc = [m for m in pol.actor if isinstance(m, torch.nn.Linear)], [m for m in pol.critic if isinstance(m, torch.nn.Linear)]
    mk = lambda l, out: B.GrMlp(l[0].weight.data_ptr(), l[0].bias.data_ptr(), l[1].weight.data_ptr(), l[1].bias.data_ptr(), l[2].weight.data_ptr(), l[2].bias.data_ptr(), 16, 128, 128, out)
    packed = torch.zeros(int(lib.gr_policy_packed_bytes(128, 128, 2)), dtype=torch.uint8, device="cuda")
    a, c = mk(la, 4), mk(lc, 1)
    st = torch.cuda.current_stream().cuda_stream
    B.check(lib.gr_policy_pack(C.byref(a), C.byref(c), packed.data_ptr(), st), "pack")
    sigma = torch.ones(4, device="cuda")
    p = B.GrPolicy(packed.data_ptr(), sigma.data_ptr(), 0.01)
    obs, cobs = torch.randn(rows, 16, device="cuda") * 3, torch.randn(rows, 16, device="cuda") * 3
    mu, val = torch.zeros(rows, 4, device="cuda"), torch.zeros(rows, device="cuda")
    B.check(lib.gr_policy_forward(C.byref(p), obs.data_ptr(), cobs.data_ptr(), mu.data_ptr(), val.data_ptr(), rows, st), "gr_policy_forward")
    torch.cuda.synchronize()
    with torch.no_grad():
        mu_ref, v_ref = pol.actor(obs), pol.critic(cobs)[:, 0]
    assert float((mu - mu_ref).abs().max()) < 1e-2 and float((mu - mu_ref).abs().mean()) < 1e-3
    assert float((val - v_ref).abs().max()) < 1e-2 and float((val - v_ref).abs().mean()) < 1e-3


def test_kernels_do_not_write_past_their_rows(cuda_lib):
    """Ragged row counts (not a multiple of the 128-row tile): outputs live inside sentinel-filled buffers, the sentinels
    survive (the sanitizer is not available on the GPU pool, so the bounds are checked this way)."""
    from generalizableracing_b200 import _lib as B
    from generalizableracing_b200.modules import ActorCritic
    lib = cuda_lib
    rows, pad, S = 333, 512, 12345.0
    torch.manual_seed(0)
    pol = ActorCritic(16, 16, 4).cuda()
    la, lc = [m for m in pol.actor if isinstance(m, torch.nn.Linear)], [m for m in pol.critic if isinstance(m, torch.nn.Linear)]
    mk = lambda l, out: B.GrMlp(l[0].weight.data_ptr(), l[0].bias.data_ptr(), l[1].weight.data_ptr(), l[1].bias.data_ptr(), l[2].weight.data_ptr(), l[2].bias.data_ptr(), 16, 128, 128, out)
    packed = torch.zeros(int(lib.gr_policy_packed_bytes(128, 128, 2)), dtype=torch.uint8, device="cuda")
    a, c = mk(la, 4), mk(lc, 1)
    st = torch.cuda.current_stream().cuda_stream
    B.check(lib.gr_policy_pack(C.byref(a), C.byref(c), packed.data_ptr(), st), "pack")
    sigma = torch.ones(4, device="cuda")
    p = B.GrPolicy(packed.data_ptr(), sigma.data_ptr(), 0.01)
    obs = torch.randn(rows, 16, device="cuda")
    guard = lambda n: torch.full((pad + n + pad,), S, device="cuda")
    mu_buf, v_buf = guard(rows * 4), guard(rows)
    mu, val = mu_buf[pad:pad + rows * 4], v_buf[pad:pad + rows]
    B.check(lib.gr_policy_forward(C.byref(p), obs.data_ptr(), obs.data_ptr(), mu.data_ptr(), val.data_ptr(), rows, st), "gr_policy_forward")
    # gradient buffers of the actor inside one guarded arena
    sizes = [t.numel() for t in (la[0].weight, la[0].bias, la[1].weight, la[1].bias, la[2].weight, la[2].bias)]
    arena = torch.full((pad + sum(sizes) + pad * 7,), S, device="cuda")
    ptrs, off = [], pad
    views = []
    for n in sizes:
        views.append(arena[off:off + n])
        arena[off:off + n] = 0.0
        ptrs.append(views[-1].data_ptr())
        off += n + pad
    g = torch.randn(rows, 4, device="cuda")
    scale = torch.ones(1, device="cuda")
    out = B.GrMlpGrad(*ptrs, 4, 0)
    B.check(lib.gr_actor_backward(C.byref(p), 128, 128, obs.data_ptr(), g.data_ptr(), scale.data_ptr(), rows, C.byref(out), st), "gr_actor_backward")
    torch.cuda.synchronize()
    assert bool((mu_buf[:pad] == S).all()) and bool((mu_buf[pad + rows * 4:] == S).all()) and bool((v_buf[:pad] == S).all()) and bool((v_buf[pad + rows:] == S).all())
    assert bool(torch.isfinite(mu).all()) and bool((mu != S).all())
    mask = torch.ones_like(arena, dtype=torch.bool)
    off = pad
    for n in sizes:
        mask[off:off + n] = False
        off += n + pad
    assert bool((arena[mask] == S).all())
    assert all(bool(torch.isfinite(v).all()) and float(v.abs().max()) > 0 for v in views)


def test_adam_clip_step_matches_torch(cuda_lib):
    """gr_adam_clip_step == nn.utils.clip_grad_norm_ + torch.optim.Adam.step (+ the KL-adaptive learning rate) over several steps."""
    from generalizableracing_b200 import _lib as B
    lib = cuda_lib
    torch.manual_seed(4)
    shapes = [(128, 16), (128,), (128, 128), (4, 128), (4,), (3,)]
    ref = [torch.nn.Parameter(torch.randn(*s, device="cuda")) for s in shapes]
    mine = [p.detach().clone() for p in ref]
    opt = torch.optim.Adam(ref, lr=5e-4)
    offs, off = [], 0
    for p in mine:
        offs.append(off)
        off += (p.numel() + 3) // 4 * 4
    n_flat = off + 16
    flat, m, v = (torch.zeros(n_flat, device="cuda") for _ in range(3))
    state = torch.zeros(16, device="cuda")
    state[0] = 5e-4
    ptrs = torch.tensor([p.data_ptr() for p in mine], dtype=torch.int64, device="cuda")
    so, sn = torch.tensor(offs, dtype=torch.int32, device="cuda"), torch.tensor([p.numel() for p in mine], dtype=torch.int32, device="cuda")
    kl = flat[off:off + 16]
    a = B.GrAdamStep(ptrs.data_ptr(), so.data_ptr(), sn.data_ptr(), len(mine), n_flat, flat.data_ptr(), m.data_ptr(), v.data_ptr(), state.data_ptr(), kl.data_ptr(),
                     1.0, 0.9, 0.999, 1e-8, 1.0, 0.01, 1e-5, 1e-2)
    lr = 5e-4
    for it in range(6):
        flat.zero_()
        scale = [3.0, 0.01, 1.0, 5.0, 0.2, 1.0][it]                       # some steps get clipped, some not
        for p, q, o in zip(ref, mine, offs):
            g = torch.randn_like(p) * scale
            p.grad = g.clone()
            flat[o:o + p.numel()] = g.reshape(-1)
        kl_mean = [0.05, 0.001, 0.01, 0.03, 0.0, 0.002][it]               # > 2*desired, < desired/2, in band, ...
        kl[2], kl[7], kl[0], kl[1] = kl_mean * 100, 100.0, 7.0, 9.0
        if kl_mean > 0.02:
            lr = max(1e-5, lr / 1.5)
        elif 0.0 < kl_mean < 0.005:
            lr = min(1e-2, lr * 1.5)
        for gr in opt.param_groups:
            gr["lr"] = lr
        torch.nn.utils.clip_grad_norm_(ref, 1.0)
        opt.step()
        B.check(lib.gr_adam_clip_step(C.byref(a), torch.cuda.current_stream().cuda_stream), "gr_adam_clip_step")
        torch.cuda.synchronize()
        assert abs(float(state[0]) - lr) < 1e-9 and float(state[1]) == it + 1
        for p, q in zip(ref, mine):
            assert torch.allclose(p, q, rtol=1e-5, atol=1e-6), (it, float((p - q).abs().max()))
    assert abs(float(state[5]) - 6 * 0.09) < 1e-5 and abs(float(state[6]) - 6 * 0.07) < 1e-5


@pytest.mark.parametrize("rows,pool", [(1000, 5000), (24576 + 5, 98304), (128 * 149, 400000)])
def test_gather_on_load_equals_gather_then_dense(cuda_lib, rows, pool):
    """The mini-batch gather of rollout_storage.py:179-187 done on load (`indices` of gr_policy_forward_gather, GrPpoBatch, GrBackwardJob)
    against the same kernels on rows gathered beforehand: identical arithmetic on identical values, so forward outputs and loss
    gradients are bit-identical and the weight gradients agree to accumulation order."""
    from generalizableracing_b200 import _lib as B
    from generalizableracing_b200.modules import ActorCritic
    lib = cuda_lib
    torch.manual_seed(rows)
    pol = ActorCritic(16, 16, 4).cuda()
    la, lc = [m for m in pol.actor if isinstance(m, torch.nn.Linear)], [m for m in pol.critic if isinstance(m, torch.nn.Linear)]
    mk = lambda l, out: B.GrMlp(l[0].weight.data_ptr(), l[0].bias.data_ptr(), l[1].weight.data_ptr(), l[1].bias.data_ptr(), l[2].weight.data_ptr(), l[2].bias.data_ptr(), 16, 128, 128, out)
    packed = torch.zeros(int(lib.gr_policy_packed_bytes(128, 128, 2)), dtype=torch.uint8, device="cuda")
    a, c = mk(la, 4), mk(lc, 1)
    st = torch.cuda.current_stream().cuda_stream
    B.check(lib.gr_policy_pack(C.byref(a), C.byref(c), packed.data_ptr(), st), "pack")
    sigma = torch.full((4,), 0.8, device="cuda")
    p = B.GrPolicy(packed.data_ptr(), sigma.data_ptr(), 0.01)
    rn = lambda *s: torch.randn(*s, device="cuda")
    S = dict(obs=rn(pool, 16) * 3, cobs=rn(pool, 16) * 3, actions=rn(pool, 4), logp=rn(pool), adv=rn(pool), ret=rn(pool), val=rn(pool), mu=rn(pool, 4),
             sig=(0.5 + torch.rand(pool, 4, device="cuda")))
    idx = torch.randperm(pool, device="cuda")[:rows].contiguous()
    D = {k: v[idx].contiguous() for k, v in S.items()}
    out = {}
    for mode, T, ip in (("dense", D, None), ("gather", S, idx.data_ptr())):
        mu, val = torch.zeros(rows, 4, device="cuda"), torch.zeros(rows, device="cuda")
        B.check(lib.gr_policy_forward_gather(C.byref(p), T["obs"].data_ptr(), T["cobs"].data_ptr(), ip, mu.data_ptr(), val.data_ptr(), rows, st), "fwd")
        gm, gv, sums = torch.zeros(rows, 4, device="cuda"), torch.zeros(rows, 4, device="cuda"), torch.zeros(16, device="cuda")
        b = B.GrPpoBatch(mu.data_ptr(), val.data_ptr(), sigma.data_ptr(), T["actions"].data_ptr(), T["logp"].data_ptr(), T["adv"].data_ptr(), T["ret"].data_ptr(),
                         T["val"].data_ptr(), T["mu"].data_ptr(), T["sig"].data_ptr(), 0.2, 1.0, 0.01, 1, ip)
        B.check(lib.gr_ppo_loss_grad(C.byref(b), rows, gm.data_ptr(), gv.data_ptr(), sums.data_ptr(), st), "loss")
        grads = [torch.zeros_like(t) for l in (la, lc) for m in l for t in (m.weight, m.bias)]
        ga = B.GrMlpGrad(*(t.data_ptr() for t in grads[:6]), 4, 1)
        gc = B.GrMlpGrad(*(t.data_ptr() for t in grads[6:]), 1, 1)
        pc = B.GrPolicy(packed.data_ptr() + packed.numel() // 2, sigma.data_ptr(), 0.01)
        jobs = (B.GrBackwardJob * 2)(B.GrBackwardJob(p, T["obs"].data_ptr(), gm.data_ptr(), sums.data_ptr() + 32, ga, ip),
                                     B.GrBackwardJob(pc, T["cobs"].data_ptr(), gv.data_ptr(), sums.data_ptr() + 36, gc, ip))
        B.check(lib.gr_actor_backward_jobs(jobs, 2, 128, 128, rows, st), "bwd")
        torch.cuda.synchronize()
        out[mode] = (mu, val, gm, gv, sums[:8].clone(), grads)
    d, g = out["dense"], out["gather"]
    for k in range(4):
        assert torch.equal(d[k], g[k]), k
    assert torch.allclose(d[4], g[4], rtol=1e-5, atol=1e-6)          # sums: atomics in any order
    for x, y in zip(d[5], g[5]):
        assert float((x - y).abs().max()) <= 2e-5 * float(x.abs().max()) + 1e-12
        assert float(x.abs().max()) > 0


@pytest.mark.parametrize("rows,pool,clipped,entropy_coef", [(1000, 5000, True, 0.0), (128 * 149 + 3, 400000, True, 0.01), (65536 * 6, 65536 * 24, False, 0.0)])
def test_fused_step_equals_the_three_launches(cuda_lib, rows, pool, clipped, entropy_coef):
    """gr_ppo_fused_step (forward head + loss + weight gradients in one kernel per net) against gr_policy_forward_gather -> gr_ppo_loss_grad ->
    gr_actor_backward_jobs on the same mini-batch: the loss sums agree to summation order, the weight gradients to the fp16 rounding of the
    cotangent rows (static scale vs batch-maximum scale: different bits of the same value)."""
    from generalizableracing_b200 import _lib as B
    from generalizableracing_b200.modules import ActorCritic
    lib = cuda_lib
    torch.manual_seed(rows)
    pol = ActorCritic(16, 16, 4).cuda()
    la, lc = [m for m in pol.actor if isinstance(m, torch.nn.Linear)], [m for m in pol.critic if isinstance(m, torch.nn.Linear)]
    mk = lambda l, out: B.GrMlp(l[0].weight.data_ptr(), l[0].bias.data_ptr(), l[1].weight.data_ptr(), l[1].bias.data_ptr(), l[2].weight.data_ptr(), l[2].bias.data_ptr(), 16, 128, 128, out)
    packed = torch.zeros(int(lib.gr_policy_packed_bytes(128, 128, 2)), dtype=torch.uint8, device="cuda")
    a, c = mk(la, 4), mk(lc, 1)
    st = torch.cuda.current_stream().cuda_stream
    B.check(lib.gr_policy_pack(C.byref(a), C.byref(c), packed.data_ptr(), st), "pack")
    sigma = torch.tensor([0.8, 0.6, 1.0, 0.7], device="cuda")
    p = B.GrPolicy(packed.data_ptr(), sigma.data_ptr(), 0.01)
    pc = B.GrPolicy(packed.data_ptr() + packed.numel() // 2, sigma.data_ptr(), 0.01)
    rn = lambda *s: torch.randn(*s, device="cuda")
    obs, cobs = rn(pool, 16) * 2, rn(pool, 16) * 2
    with torch.no_grad():
        old_mu = pol.actor(obs) + 0.1 * rn(pool, 4)
        old_v = pol.critic(cobs)[:, 0] + 0.2 * rn(pool)
    old_sig = (sigma * (1 + 0.05 * rn(4))).abs().expand(pool, 4).contiguous()
    actions = old_mu + old_sig * rn(pool, 4)
    logp = torch.distributions.Normal(old_mu, old_sig).log_prob(actions).sum(-1)
    adv, ret = rn(pool), old_v + 0.5 * rn(pool)
    idx = torch.randperm(pool, device="cuda")[:rows].contiguous()
    ip = idx.data_ptr()

    def grads():
        gs = [torch.zeros_like(t) for l in (la, lc) for m in l for t in (m.weight, m.bias)]
        return gs, B.GrMlpGrad(*(t.data_ptr() for t in gs[:6]), 4, 1), B.GrMlpGrad(*(t.data_ptr() for t in gs[6:]), 1, 1)

    # --- three launches
    mu, val = torch.zeros(rows, 4, device="cuda"), torch.zeros(rows, device="cuda")
    gm, gv, sums = torch.zeros(rows, 4, device="cuda"), torch.zeros(rows, 4, device="cuda"), torch.zeros(16, device="cuda")
    B.check(lib.gr_policy_forward_gather(C.byref(p), obs.data_ptr(), cobs.data_ptr(), ip, mu.data_ptr(), val.data_ptr(), rows, st), "fwd")
    b = B.GrPpoBatch(mu.data_ptr(), val.data_ptr(), sigma.data_ptr(), actions.data_ptr(), logp.data_ptr(), adv.data_ptr(), ret.data_ptr(), old_v.data_ptr(),
                     old_mu.data_ptr(), old_sig.data_ptr(), 0.2, 1.0, entropy_coef, int(clipped), ip)
    B.check(lib.gr_ppo_loss_grad(C.byref(b), rows, gm.data_ptr(), gv.data_ptr(), sums.data_ptr(), st), "loss")
    g3, ga, gc = grads()
    jobs = (B.GrBackwardJob * 2)(B.GrBackwardJob(p, obs.data_ptr(), gm.data_ptr(), sums.data_ptr() + 32, ga, ip),
                                 B.GrBackwardJob(pc, cobs.data_ptr(), gv.data_ptr(), sums.data_ptr() + 36, gc, ip))
    B.check(lib.gr_actor_backward_jobs(jobs, 2, 128, 128, rows, st), "bwd")
    # --- one launch
    g1, ga1, gc1 = grads()
    sums1 = torch.zeros(16, device="cuda")
    b1 = B.GrPpoBatch(None, None, sigma.data_ptr(), actions.data_ptr(), logp.data_ptr(), adv.data_ptr(), ret.data_ptr(), old_v.data_ptr(),
                      old_mu.data_ptr(), old_sig.data_ptr(), 0.2, 1.0, entropy_coef, int(clipped), ip)
    step = B.GrPpoStep(p, obs.data_ptr(), cobs.data_ptr(), b1, ga1, gc1, sums1.data_ptr(), 0.0)
    B.check(lib.gr_ppo_fused_step(C.byref(step), rows, st), "gr_ppo_fused_step")
    # ... and with the rows coming from transition records
    rec = torch.cat([obs, cobs, actions, old_mu, old_sig, logp[:, None], adv[:, None], ret[:, None], old_v[:, None]], dim=1).contiguous()
    g2, ga2, gc2 = grads()
    sums2 = torch.zeros(16, device="cuda")
    b2 = B.GrPpoBatch(None, None, sigma.data_ptr(), None, None, None, None, None, None, None, 0.2, 1.0, entropy_coef, int(clipped), ip, rec.data_ptr())
    step2 = B.GrPpoStep(p, None, None, b2, ga2, gc2, sums2.data_ptr(), 0.0)
    B.check(lib.gr_ppo_fused_step(C.byref(step2), rows, st), "gr_ppo_fused_step (records)")
    torch.cuda.synchronize()
    assert torch.allclose(sums2[:8], sums1[:8], rtol=2e-5, atol=1e-5 * float(sums1[:8].abs().max()))
    for x, y in zip(g1, g2):
        assert float((x - y).abs().max()) <= 2e-5 * float(x.abs().max()) + 1e-12
    assert torch.allclose(sums1[:8], sums[:8], rtol=2e-4, atol=1e-4 * float(sums[:8].abs().max())), (sums1[:8], sums[:8])
    assert int(sums1[7]) == rows
    names = [f"{net}.{n}" for net in ("actor", "critic") for n in ("w1", "b1", "w2", "b2", "w3", "b3")]
    for name, x, y in zip(names, g3, g1):
        assert bool(torch.isfinite(y).all()) and float(x.abs().max()) > 0, name
        assert float((x - y).abs().max()) <= 1e-2 * float(x.abs().max()), (name, float((x - y).abs().max() / x.abs().max()))


@pytest.mark.parametrize("rows,pool,clipped", [(1000, 5000, True), (128 * 149 + 3, 400000, False), (65536 * 6, 65536 * 24, True)])
def test_forward_with_loss_equals_forward_then_loss(cuda_lib, rows, pool, clipped):
    """gr_policy_forward_loss (the thread holding a row's mean / value evaluates the row's loss) against gr_policy_forward_gather followed by
    gr_ppo_loss_grad: same arithmetic per row -> bit-identical policy outputs and gradients; the sums agree to summation order."""
    from generalizableracing_b200 import _lib as B
    from generalizableracing_b200.modules import ActorCritic
    lib = cuda_lib
    torch.manual_seed(rows + 1)
    pol = ActorCritic(16, 16, 4).cuda()
    la, lc = [m for m in pol.actor if isinstance(m, torch.nn.Linear)], [m for m in pol.critic if isinstance(m, torch.nn.Linear)]
    mk = lambda l, out: B.GrMlp(l[0].weight.data_ptr(), l[0].bias.data_ptr(), l[1].weight.data_ptr(), l[1].bias.data_ptr(), l[2].weight.data_ptr(), l[2].bias.data_ptr(), 16, 128, 128, out)
    packed = torch.zeros(int(lib.gr_policy_packed_bytes(128, 128, 2)), dtype=torch.uint8, device="cuda")
    a, c = mk(la, 4), mk(lc, 1)
    st = torch.cuda.current_stream().cuda_stream
    B.check(lib.gr_policy_pack(C.byref(a), C.byref(c), packed.data_ptr(), st), "pack")
    sigma = torch.tensor([0.8, 0.6, 1.0, 0.7], device="cuda")
    p = B.GrPolicy(packed.data_ptr(), sigma.data_ptr(), 0.01)
    rn = lambda *s: torch.randn(*s, device="cuda")
    obs, cobs = rn(pool, 16) * 2, rn(pool, 16) * 2
    old_mu, old_v = rn(pool, 4) * 0.5, rn(pool)
    old_sig = (sigma * (1 + 0.05 * rn(4))).abs().expand(pool, 4).contiguous()
    actions = old_mu + old_sig * rn(pool, 4)
    logp = torch.distributions.Normal(old_mu, old_sig).log_prob(actions).sum(-1)
    adv, ret = rn(pool), old_v + 0.5 * rn(pool)
    idx = torch.randperm(pool, device="cuda")[:rows].contiguous()
    ip = idx.data_ptr()
    # two launches
    mu, val = torch.zeros(rows, 4, device="cuda"), torch.zeros(rows, device="cuda")
    gm, gv, sums = torch.zeros(rows, 4, device="cuda"), torch.zeros(rows, 4, device="cuda"), torch.zeros(16, device="cuda")
    B.check(lib.gr_policy_forward_gather(C.byref(p), obs.data_ptr(), cobs.data_ptr(), ip, mu.data_ptr(), val.data_ptr(), rows, st), "fwd")
    b = B.GrPpoBatch(mu.data_ptr(), val.data_ptr(), sigma.data_ptr(), actions.data_ptr(), logp.data_ptr(), adv.data_ptr(), ret.data_ptr(), old_v.data_ptr(),
                     old_mu.data_ptr(), old_sig.data_ptr(), 0.2, 1.0, 0.003, int(clipped), ip)
    B.check(lib.gr_ppo_loss_grad(C.byref(b), rows, gm.data_ptr(), gv.data_ptr(), sums.data_ptr(), st), "loss")
    # one launch (with and without the optional policy outputs)
    mu1, val1 = torch.zeros(rows, 4, device="cuda"), torch.zeros(rows, device="cuda")
    gm1, gv1, sums1 = torch.zeros(rows, 4, device="cuda"), torch.zeros(rows, 4, device="cuda"), torch.zeros(16, device="cuda")
    b1 = B.GrPpoBatch(mu1.data_ptr(), val1.data_ptr(), sigma.data_ptr(), actions.data_ptr(), logp.data_ptr(), adv.data_ptr(), ret.data_ptr(), old_v.data_ptr(),
                      old_mu.data_ptr(), old_sig.data_ptr(), 0.2, 1.0, 0.003, int(clipped), ip)
    B.check(lib.gr_policy_forward_loss(C.byref(p), obs.data_ptr(), cobs.data_ptr(), C.byref(b1), rows, gm1.data_ptr(), gv1.data_ptr(), sums1.data_ptr(), st), "fwd+loss")
    gm2, gv2, sums2 = torch.zeros(rows, 4, device="cuda"), torch.zeros(rows, 4, device="cuda"), torch.zeros(16, device="cuda")
    b2 = B.GrPpoBatch(None, None, sigma.data_ptr(), actions.data_ptr(), logp.data_ptr(), adv.data_ptr(), ret.data_ptr(), old_v.data_ptr(),
                      old_mu.data_ptr(), old_sig.data_ptr(), 0.2, 1.0, 0.003, int(clipped), ip)
    B.check(lib.gr_policy_forward_loss(C.byref(p), obs.data_ptr(), cobs.data_ptr(), C.byref(b2), rows, gm2.data_ptr(), gv2.data_ptr(), sums2.data_ptr(), st), "fwd+loss")
    # transition records: the same columns side by side, one 192-byte record per transition
    rec = torch.cat([obs, cobs, actions, old_mu, old_sig, logp[:, None], adv[:, None], ret[:, None], old_v[:, None]], dim=1).contiguous()
    assert rec.shape == (pool, B.GR_RECORD_FLOATS)
    gm3, gv3, sums3 = torch.zeros(rows, 4, device="cuda"), torch.zeros(rows, 4, device="cuda"), torch.zeros(16, device="cuda")
    b3 = B.GrPpoBatch(None, None, sigma.data_ptr(), None, None, None, None, None, None, None, 0.2, 1.0, 0.003, int(clipped), ip, rec.data_ptr())
    B.check(lib.gr_policy_forward_loss(C.byref(p), None, None, C.byref(b3), rows, gm3.data_ptr(), gv3.data_ptr(), sums3.data_ptr(), st), "fwd+loss (records)")
    # weight gradients reading the observation rows out of the records (obs_stride) against dense gathered rows
    def bwd(obs_ptr, cobs_ptr, stride, index_ptr):
        gs = [torch.zeros_like(t) for l in (la, lc) for m in l for t in (m.weight, m.bias)]
        ga, gc = B.GrMlpGrad(*(t.data_ptr() for t in gs[:6]), 4, 1), B.GrMlpGrad(*(t.data_ptr() for t in gs[6:]), 1, 1)
        pc = B.GrPolicy(packed.data_ptr() + packed.numel() // 2, sigma.data_ptr(), 0.01)
        jobs = (B.GrBackwardJob * 2)(B.GrBackwardJob(p, obs_ptr, gm.data_ptr(), sums.data_ptr() + 32, ga, index_ptr, stride),
                                     B.GrBackwardJob(pc, cobs_ptr, gv.data_ptr(), sums.data_ptr() + 36, gc, index_ptr, stride))
        B.check(lib.gr_actor_backward_jobs(jobs, 2, 128, 128, rows, st), "bwd")
        return gs
    g_dense = bwd(obs.data_ptr(), cobs.data_ptr(), 0, ip)
    g_rec = bwd(rec.data_ptr(), rec.data_ptr() + 64, B.GR_RECORD_FLOATS, ip)
    torch.cuda.synchronize()
    assert torch.equal(gm, gm3) and torch.equal(gv, gv3) and torch.equal(sums[8:10], sums3[8:10])
    assert torch.allclose(sums[:8], sums3[:8], rtol=2e-5, atol=1e-5 * float(sums[:8].abs().max()))
    for x, y in zip(g_dense, g_rec):
        assert float(x.abs().max()) > 0 and float((x - y).abs().max()) <= 2e-5 * float(x.abs().max()) + 1e-12
    assert torch.equal(mu, mu1) and torch.equal(val, val1)
    assert torch.equal(gm, gm1) and torch.equal(gv, gv1) and torch.equal(gm, gm2) and torch.equal(gv, gv2)
    assert torch.equal(sums[8:10], sums1[8:10]) and torch.equal(sums[8:10], sums2[8:10])          # maxima: exact
    assert torch.allclose(sums[:8], sums1[:8], rtol=2e-5, atol=1e-5 * float(sums[:8].abs().max()))
    assert torch.allclose(sums[:8], sums2[:8], rtol=2e-5, atol=1e-5 * float(sums[:8].abs().max()))


def test_storage_pack_records(cuda_lib):
    """gr_storage_pack_records against torch.cat of the storage's own columns."""
    from generalizableracing_b200 import _lib as B
    from generalizableracing_b200.storage import RolloutStorage
    T, N = 5, 333
    sto = RolloutStorage("rl", N, T, [16], [16], [4], device="cuda:0")
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in (sto.observations, sto.privileged_observations, sto.actions, sto.mu, sto.sigma, sto.actions_log_prob, sto.advantages, sto.returns, sto.values):
        t.copy_(torch.randn(t.shape, device="cuda", generator=g))
    rec = sto.pack_records()
    f = lambda t: t.reshape(T * N, -1)
    want = torch.cat([f(sto.observations), f(sto.privileged_observations), f(sto.actions), f(sto.mu), f(sto.sigma), f(sto.actions_log_prob), f(sto.advantages),
                      f(sto.returns), f(sto.values)], dim=1)
    assert rec.shape == (T * N, B.GR_RECORD_FLOATS) and torch.equal(rec, want)


def test_storage_pack_records_permuted(cuda_lib):
    """gr_storage_pack_records_permuted: record r = transition perm[r]; a prefix of a permutation packs only that many records."""
    from generalizableracing_b200 import _lib as B
    from generalizableracing_b200.storage import RolloutStorage
    T, N = 5, 333
    sto = RolloutStorage("rl", N, T, [16], [16], [4], device="cuda:0")
    g = torch.Generator(device="cuda").manual_seed(1)
    for t in (sto.observations, sto.privileged_observations, sto.actions, sto.mu, sto.sigma, sto.actions_log_prob, sto.advantages, sto.returns, sto.values):
        t.copy_(torch.randn(t.shape, device="cuda", generator=g))
    want = sto.pack_records().clone()
    perm = torch.randperm(T * N, device="cuda", generator=g)
    rec = sto.pack_records(perm)
    assert torch.equal(rec, want[perm])
    sto._records.fill_(-7.0)
    k = 4 * ((T * N) // 4)
    rec = sto.pack_records(perm[:k].contiguous())
    assert torch.equal(rec[:k], want[perm[:k]]) and bool((rec[k:] == -7.0).all())
    with pytest.raises(ValueError):
        sto.pack_records(perm.to(torch.int32))


@pytest.mark.parametrize("rows,pool", [(1000, 5000), (40000, 98304)])
def test_dense_records_are_the_gathered_rows(cuda_lib, rows, pool):
    """Records packed in mini-batch order and read densely (indices = NULL, records + slot offset) against the same records gathered through
    the index buffer: per-row outputs bit for bit, sums and weight gradients up to the order of the flush atomics."""
    import ctypes as C
    from generalizableracing_b200 import _lib as B
    from generalizableracing_b200.modules import ActorCritic
    lib = cuda_lib
    torch.manual_seed(rows + 5)
    pol = ActorCritic(16, 16, 4).cuda()
    la, lc = [m for m in pol.actor if isinstance(m, torch.nn.Linear)], [m for m in pol.critic if isinstance(m, torch.nn.Linear)]
    mk = lambda l, out: B.GrMlp(l[0].weight.data_ptr(), l[0].bias.data_ptr(), l[1].weight.data_ptr(), l[1].bias.data_ptr(), l[2].weight.data_ptr(), l[2].bias.data_ptr(), 16, 128, 128, out)
    packed = torch.zeros(int(lib.gr_policy_packed_bytes(128, 128, 2)), dtype=torch.uint8, device="cuda")
    a, c = mk(la, 4), mk(lc, 1)
    st = torch.cuda.current_stream().cuda_stream
    B.check(lib.gr_policy_pack(C.byref(a), C.byref(c), packed.data_ptr(), st), "pack")
    sigma = torch.tensor([0.8, 0.6, 1.0, 0.7], device="cuda")
    p = B.GrPolicy(packed.data_ptr(), sigma.data_ptr(), 0.01)
    pc = B.GrPolicy(packed.data_ptr() + packed.numel() // 2, sigma.data_ptr(), 0.01)
    rn = lambda *s: torch.randn(*s, device="cuda")
    old_mu, old_v = rn(pool, 4) * 0.5, rn(pool)
    old_sig = (sigma * (1 + 0.05 * rn(4))).abs().expand(pool, 4).contiguous()
    actions = old_mu + old_sig * rn(pool, 4)
    logp = torch.distributions.Normal(old_mu, old_sig).log_prob(actions).sum(-1)
    rec = torch.cat([rn(pool, 16) * 2, rn(pool, 16) * 2, actions, old_mu, old_sig, logp[:, None], rn(pool, 1), old_v[:, None] + 0.5 * rn(pool, 1), old_v[:, None]], dim=1).contiguous()
    perm = torch.randperm(pool, device="cuda")
    slot = 1                                                         # the second mini-batch slot: rows [rows, 2*rows) of the permutation
    idx = perm[slot * rows:(slot + 1) * rows].contiguous()
    rec_perm = rec[perm].contiguous()
    dense_ptr = rec_perm.data_ptr() + slot * rows * B.GR_RECORD_FLOATS * 4

    def run(rec_ptr, index_ptr):
        gm, gv, sums = torch.zeros(rows, 4, device="cuda"), torch.zeros(rows, 4, device="cuda"), torch.zeros(16, device="cuda")
        b = B.GrPpoBatch(None, None, sigma.data_ptr(), None, None, None, None, None, None, None, 0.2, 1.0, 0.003, 1, index_ptr, rec_ptr)
        B.check(lib.gr_policy_forward_loss(C.byref(p), None, None, C.byref(b), rows, gm.data_ptr(), gv.data_ptr(), sums.data_ptr(), st), "fwd+loss")
        gs = [torch.zeros_like(t) for l in (la, lc) for m in l for t in (m.weight, m.bias)]
        ga, gc = B.GrMlpGrad(*(t.data_ptr() for t in gs[:6]), 4, 1), B.GrMlpGrad(*(t.data_ptr() for t in gs[6:]), 1, 1)
        jobs = (B.GrBackwardJob * 2)(B.GrBackwardJob(p, rec_ptr, gm.data_ptr(), sums.data_ptr() + 32, ga, index_ptr, B.GR_RECORD_FLOATS),
                                     B.GrBackwardJob(pc, rec_ptr + 64, gv.data_ptr(), sums.data_ptr() + 36, gc, index_ptr, B.GR_RECORD_FLOATS))
        B.check(lib.gr_actor_backward_jobs(jobs, 2, 128, 128, rows, st), "bwd")
        # the one-launch step on the same rows
        gs1 = [torch.zeros_like(t) for t in gs]
        ga1, gc1 = B.GrMlpGrad(*(t.data_ptr() for t in gs1[:6]), 4, 1), B.GrMlpGrad(*(t.data_ptr() for t in gs1[6:]), 1, 1)
        sums1 = torch.zeros(16, device="cuda")
        fs = B.GrPpoStep(p, None, None, b, ga1, gc1, sums1.data_ptr(), 0.0)
        B.check(lib.gr_ppo_fused_step(C.byref(fs), rows, st), "fused step")
        torch.cuda.synchronize()
        return gm, gv, sums, gs, sums1, gs1
    gm, gv, sums, gs, sums1, gs1 = run(rec.data_ptr(), idx.data_ptr())
    dm, dv, dsums, ds, dsums1, ds1 = run(dense_ptr, None)
    assert torch.equal(gm, dm) and torch.equal(gv, dv) and torch.equal(sums[8:10], dsums[8:10])
    assert float(gm.abs().max()) > 0 and float(gv.abs().max()) > 0
    for x, y in ((sums[:8], dsums[:8]), (sums1[:8], dsums1[:8])):
        assert torch.allclose(x, y, rtol=2e-5, atol=1e-5 * float(x.abs().max()))
    for x, y in list(zip(gs, ds)) + list(zip(gs1, ds1)):
        assert float(x.abs().max()) > 0 and float((x - y).abs().max()) <= 2e-5 * float(x.abs().max()) + 1e-12

"""The ``torch.ops.gracing.*`` operator layer (generalizableracing_b200/ops.py): schemas, mutation
declarations and fake-tensor kernels (torch.library.opcheck), bit-identity with the direct ctypes path for step / reset,
the registered autograd formula against the chained autograd.Function path and against the oracle, functional use on a
cloned state, and ``gracing::gae`` against ``RolloutStorage.compute_returns``."""
import pytest
import torch

from generalizableracing_b200 import ops  # noqa: F401  (registers the operators)
from generalizableracing_b200 import layout as L
from tests import parity_cases as PC
from tests.conftest import backend_params

pytestmark = pytest.mark.timeout(600)


def _pair(backend, monkeypatch, op_layer, **kw):
    monkeypatch.setenv("GRACING_OP_LAYER", "1" if op_layer else "0")
    out = PC.make_pair(backend, **kw)
    assert (out[3]._ops is not None) == op_layer
    return out


def test_operators_are_registered_with_mutation_schemas():
    want = {
        "step_fwd": ("Tensor(a1!) planes", "Tensor(a5!) log_accum"),
        "step_fwd_tape": ("Tensor(a1!) planes", "Tensor(a6!) tape"),
        "step_loss": (),
        "step_bwd": ("Tensor(a5!) adjoint", "Tensor(a6!) grad_action"),
        "rollout_fwd": ("Tensor(a1!) planes", "Tensor(a5!) log_accum"),
        "rollout_fwd_tape": ("Tensor(a1!) planes", "Tensor(a6!) tape"),
        "reset": ("Tensor(a1!) planes",),
        "gae": (),
    }
    assert set(ops.OPERATORS) == set(want)
    for name, mutated in want.items():
        schema = str(getattr(torch.ops.gracing, name).default._schema)
        for m in mutated:
            assert m in schema, (name, schema)
        if not mutated:
            assert "!" not in schema, schema


def test_gae_refuses_cpu_tensors():
    T, N = 4, 8
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        torch.ops.gracing.gae(torch.zeros(T, N, 1), torch.zeros(T, N, 1), torch.zeros(T, N, 1, dtype=torch.uint8), torch.zeros(N, 1), 0.99, 0.95)


def test_unknown_handle_raises():
    with pytest.raises(RuntimeError, match="unknown or released env handle"):
        torch.ops.gracing.reset(987654321, torch.zeros(1, L.TILE_PLANES, L.TILE, 4), None, None, 0)


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_step_and_reset_bit_identical_to_direct_path(backend, monkeypatch):
    N, T = 130, 70
    _, _, _, a, g = _pair(backend, monkeypatch, False, stage=1, N=N, seed=5)
    _, _, _, b, _ = _pair(backend, monkeypatch, True, stage=1, N=N, seed=5)
    dev = a.device
    r0 = PC.draw_rnd(N, g).to(dev)
    oa, _ = a.reset(r0)
    ob, _ = b.reset(r0)
    assert torch.equal(oa, ob)
    resets = 0
    for t in range(T):
        act, r = (torch.randn(N, 4, generator=g) * 0.5).to(dev), PC.draw_rnd(N, g).to(dev)
        xa, xb = a.step(act, r), b.step(act, r)
        for i in range(3):
            assert torch.equal(xa[i], xb[i]), (t, i)
        for k in ("time_outs", "terminated"):
            assert torch.equal(xa[3][k], xb[3][k]), (t, k)
        for k in ("policy", "critic", "auxiliary"):
            assert torch.equal(xa[3]["observations"][k], xb[3]["observations"][k]), (t, k)
        assert torch.equal(a._last["reward_terms"], b._last["reward_terms"]) and torch.equal(a._last["gate_passed"], b._last["gate_passed"])
        resets += int(xa[2].sum())
    assert resets > 0
    assert torch.equal(a.planes, b.planes) and torch.equal(a._log_accum, b._log_accum)
    assert torch.equal(a.get_observations()[0], b.get_observations()[0])
    with torch.inference_mode():                       # the rollout loop of OnPolicyRunner.learn (on_policy_runner.py:141)
        b.step(act, r)
    obs = b.get_observations()[0]                      # outside it the observations must be usable under autograd again
    lin = torch.nn.Linear(L.OBS_DIM, 4).to(dev)
    lin(obs).sum().backward()
    assert lin.weight.grad is not None


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_functional_use_on_a_cloned_state_and_masked_reset(backend, monkeypatch):
    N = 96
    _, _, _, env, g = _pair(backend, monkeypatch, True, stage=1, N=N, seed=7)
    dev = env.device
    env.reset(PC.draw_rnd(N, g).to(dev))
    for _ in range(5):
        env.step((torch.randn(N, 4, generator=g) * 0.5).to(dev), PC.draw_rnd(N, g).to(dev))
    h = env._op_handle
    before, log_before = env.planes.clone(), env._log_accum.clone()
    act, r = (torch.randn(N, 4, generator=g) * 0.5).to(dev), PC.draw_rnd(N, g).to(dev)
    scratch, log = env.planes.clone(), env._log_accum.clone()
    out = torch.ops.gracing.step_fwd(h, scratch, act, r, 123, log, False)
    assert torch.equal(env.planes, before) and torch.equal(env._log_accum, log_before)      # only the declared arguments moved
    assert not torch.equal(scratch, before)
    assert out[7].shape == (0, L.NUM_REWARD_TERMS) and out[8].shape == (0,)
    ref = env.step(act, r)
    assert torch.equal(out[0], ref[0]) and torch.equal(out[3], ref[1]) and torch.equal(out[6], ref[2])
    assert torch.equal(scratch, env.planes)
    # masked reset: only the masked envs change, and exactly as the C entry point changes them
    mask = (torch.rand(N, generator=g) < 0.3).to(dev)
    r = PC.draw_rnd(N, g).to(dev)
    p0 = env.planes.clone()
    p1 = p0.clone()
    obs, critic, aux = torch.ops.gracing.reset(h, p1, mask, r, 7)
    rows0 = p0.permute(1, 0, 2, 3).reshape(L.TILE_PLANES, -1, 4)[:, :N]
    rows1 = p1.permute(1, 0, 2, 3).reshape(L.TILE_PLANES, -1, 4)[:, :N]
    keep = ~mask
    assert torch.equal(rows0[:, keep], rows1[:, keep])
    assert not torch.equal(rows0[:, mask], rows1[:, mask])
    assert obs.shape == (N, L.OBS_DIM) and critic.shape == (N, L.OBS_DIM) and aux.shape == (N, 1) and bool(torch.isfinite(obs).all())
    with pytest.raises(ValueError, match="planes must be"):
        torch.ops.gracing.step_fwd(h, scratch[:-1], act, r, 0, log, False)
    with pytest.raises(ValueError, match="Invalid action shape"):
        torch.ops.gracing.step_fwd(h, scratch, act[:-1].contiguous(), r, 0, log, False)


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_registered_autograd_matches_function_path_and_oracle(backend, monkeypatch):
    N, H = 64, 12
    grads = {}
    for op_layer in (False, True):
        cfg, table, orc, env, g = _pair(backend, monkeypatch, op_layer, stage=0, N=N, seed=3, diff=True, horizon=H)
        dev = env.device
        r0 = PC.draw_rnd(N, g)
        orc.reset(r0)
        env.reset(r0.to(dev))
        ep = torch.randint(0, cfg.max_episode_length, (N,), generator=g)
        orc.episode_length_buf[:] = ep
        env.episode_length_buf = ep
        orc.detach()
        env.detach()
        acts = [(torch.randn(N, 4, generator=g) * 0.5).requires_grad_(True) for _ in range(H)]
        acts_k = [a.detach().clone().to(dev).requires_grad_(True) for a in acts]
        w = torch.rand(H, N, generator=g)
        ol, kl = [], []
        for t in range(H):
            r = PC.draw_rnd(N, g)
            ol.append(orc.step(acts[t], r)[4]["losses"])
            ex = env.step(acts_k[t], r.to(dev))[3]
            assert ex["losses"].requires_grad
            kl.append(ex["losses"])
        if op_layer:
            assert "StepLoss" not in type(kl[0].grad_fn).__name__          # the formula registered on gracing::step_loss, not _StepLoss
        (torch.stack(ol) * w).sum().backward()
        ref = torch.stack([a.grad if a.grad is not None else torch.zeros_like(a) for a in acts])
        (torch.stack(kl) * w.to(dev)).sum().backward()
        got = torch.stack([a.grad if a.grad is not None else torch.zeros_like(a) for a in acts_k]).cpu()
        assert float((ref - got).abs().max() / ref.abs().max()) < 1e-4        # same bound as tests/test_bptt_parity.py
        grads[op_layer] = got
        if op_layer:       # one-launch sweep through gracing::step_bwd
            sweep = env._bptt.backward_window(grad_losses=w.to(dev)).cpu()
            assert float((ref - sweep).abs().max() / ref.abs().max()) < 1e-4
            env.detach()
            with pytest.raises(RuntimeError, match="already detached"):
                (kl[-1].sum()).backward()
    assert torch.equal(grads[False], grads[True])


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_opcheck_schema_and_fake_kernels(backend, monkeypatch):
    N = 64
    _, _, _, env, g = _pair(backend, monkeypatch, True, stage=1, N=N, seed=9)
    dev = env.device
    env.reset(PC.draw_rnd(N, g).to(dev))
    h = env._op_handle
    act, r = (torch.randn(N, 4, generator=g) * 0.5).to(dev), PC.draw_rnd(N, g).to(dev)
    utils = ("test_schema", "test_faketensor")
    torch.library.opcheck(torch.ops.gracing.step_fwd.default, (h, env.planes.clone(), act, r, 3, env._log_accum.clone(), True), test_utils=utils)
    torch.library.opcheck(torch.ops.gracing.reset.default, (h, env.planes.clone(), None, r, 4), test_utils=utils)
    _, _, _, denv, g = _pair(backend, monkeypatch, True, stage=0, N=N, seed=9, diff=True, horizon=4)
    denv.reset(PC.draw_rnd(N, g).to(dev))
    win, hd = denv._bptt, denv._op_handle
    torch.library.opcheck(torch.ops.gracing.step_fwd_tape.default,
                          (hd, denv.planes.clone(), act, r, 5, denv._log_accum.clone(), win.tape.clone(), 0, False), test_utils=utils)
    denv.step(act, r)
    torch.library.opcheck(torch.ops.gracing.step_bwd.default,
                          (hd, denv.planes, win.tape, None, 0.5, win.adjoint.clone(), win.grad_action.clone(), 0, 1), test_utils=utils)
    tok = torch.zeros(1, device=dev, requires_grad=True)
    torch.library.opcheck(torch.ops.gracing.step_loss.default, (hd, act.clone().requires_grad_(True), tok, win.loss[0].clone(), 0, win.epoch),
                          test_utils=utils + ("test_autograd_registration",))


@pytest.mark.gpu
@pytest.mark.parametrize("T,N,normalize", [(24, 4096, True), (24, 4096, False), (7, 130, True)])
def test_gae_operator_equals_storage_kernel(cuda_lib, T, N, normalize):
    from generalizableracing_b200.storage import RolloutStorage
    from oracle import rollout_oracle as RO
    g = torch.Generator().manual_seed(T * N)
    sto = RolloutStorage("rl", N, T, [16], [16], [4], device="cuda:0")
    sto.rewards.copy_(torch.randn(T, N, 1, generator=g))
    sto.values.copy_(torch.randn(T, N, 1, generator=g))
    sto.dones.copy_((torch.rand(T, N, 1, generator=g) < 0.02).byte())
    last = torch.randn(N, 1, generator=g).cuda()
    sto.compute_returns(last, 0.99, 0.95, normalize=normalize)
    ret, adv, mom = torch.ops.gracing.gae(sto.rewards, sto.values, sto.dones, last, 0.99, 0.95, normalize)
    assert torch.equal(ret, sto.returns) and torch.equal(adv, sto.advantages) and torch.equal(mom, sto.moments)
    ref_ret, ref_adv = RO.compute_returns(sto.rewards.cpu(), sto.values.cpu(), sto.dones.cpu(), last.cpu(), 0.99, 0.95)
    assert PC.rel_err(ref_ret, ret) < 1e-5                                  # GAE tolerance of tests/test_rollout_parity.py
    if normalize:
        assert PC.rel_err(ref_adv, adv) < 5e-5
    torch.library.opcheck(torch.ops.gracing.gae.default, (sto.rewards, sto.values, sto.dones, last, 0.99, 0.95, normalize),
                          test_utils=("test_schema", "test_faketensor"))


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
@pytest.mark.parametrize("diff", [False, True])
def test_rollout_window_through_the_operators(backend, monkeypatch, diff):
    N, T = 96, 20
    kw = dict(stage=0 if diff else 1, N=N, seed=13, diff=diff, horizon=T if diff else 0)
    _, _, _, a, g = _pair(backend, monkeypatch, False, **kw)
    _, _, _, b, _ = _pair(backend, monkeypatch, True, **kw)
    dev = a.device
    r0 = PC.draw_rnd(N, g).to(dev)
    a.reset(r0)
    b.reset(r0)
    acts = (torch.randn(T, N, 4, generator=g) * 0.5).to(dev)
    rnd = torch.stack([PC.draw_rnd(N, g) for _ in range(T)]).to(dev)
    oa, ob = a.rollout(acts, rnd, record_obs=True), b.rollout(acts, rnd, record_obs=True)
    assert set(oa) == set(ob)
    for k in oa:
        assert torch.equal(oa[k], ob[k]), k
    assert torch.equal(a.planes, b.planes)
    if diff:
        assert torch.equal(a._bptt.tape[:T], b._bptt.tape[:T]) and a._bptt.t == b._bptt.t == T
        assert torch.equal(a._bptt.backward_window().clone(), b._bptt.backward_window().clone())
    utils = ("test_schema", "test_faketensor")
    if diff:
        b.detach()
        torch.library.opcheck(torch.ops.gracing.rollout_fwd_tape.default,
                              (b._op_handle, b.planes.clone(), acts, rnd, 3, b._log_accum.clone(), b._bptt.tape.clone(), 0, True), test_utils=utils)
    else:
        torch.library.opcheck(torch.ops.gracing.rollout_fwd.default, (b._op_handle, b.planes.clone(), acts, rnd, 3, b._log_accum.clone(), True), test_utils=utils)

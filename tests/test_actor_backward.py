"""gr_actor_backward (tcgen05 weight gradients of the BPTT actor) against torch.autograd.

The kernel differentiates the function the fused forward evaluates (fp16 operands, activations rounded to fp16 between
layers), so the tight reference is autograd through a torch model with the same roundings (straight-through): weights agree to
a few 1e-3 of each tensor's scale.  Against the plain fp32 module the difference is that of the two FUNCTIONS: with a random
zero-mean cotangent the true gradient is a near-cancelling sum and merely rounding the observations to fp16 moves it by 1.5 %
in pure fp32 torch (measured), so that comparison uses a cotangent correlated with the input (no cancellation)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _rnd16(x):
    return x + (x.half().float() - x).detach()


def _emulated(actor, x):
    """the fused forward's arithmetic with straight-through roundings, so that autograd differentiates what the kernel does"""
    l1, _, l2, _, l3 = actor
    b1 = l1.bias.half().float() + (l1.bias - l1.bias.half().float()).half().float()
    h = _rnd16(_rnd16(x) @ _rnd16(l1.weight).T + b1)
    h = _rnd16(torch.where(h > 0, h, h * 0.01))
    h = _rnd16(_rnd16(h @ _rnd16(l2.weight).T) + l2.bias.half().float())
    h = _rnd16(torch.where(h > 0, h, h * 0.01))
    return h @ _rnd16(l3.weight).T + l3.bias


@pytest.mark.parametrize("rows,hidden", [(128, (256, 128)), (1000, (256, 128)), (4096 * 24, (256, 128)), (777, (128, 128)), (16384 * 32, (256, 128)),
                                         (65536 * 6 + 77, (128, 128)), (148 * 128 + 1, (128, 128)), (149 * 128, (128, 128))])
def test_actor_backward_matches_autograd(cuda_lib, rows, hidden):
    from generalizableracing_b200 import _lib as B
    from generalizableracing_b200.modules import BaseModel
    lib = cuda_lib
    torch.manual_seed(rows)
    pol = BaseModel(16, 16, 4, actor_hidden_dims=hidden, critic_hidden_dims=hidden, activation="lrelu").cuda()
    l1, l2, l3 = [m for m in pol.actor if isinstance(m, torch.nn.Linear)]
    x = torch.randn(rows, 16, device="cuda") * torch.tensor([3.0] * 6 + [8.0] * 6 + [5.0] * 4, device="cuda")
    g = torch.randn(rows, 4, device="cuda") * 1e-6                      # d(loss)/d(action) of a mean over ~1e6 env-steps
    packed = torch.zeros(int(lib.gr_policy_packed_bytes(hidden[0], hidden[1], 1)), dtype=torch.uint8, device="cuda")
    mlp = B.GrMlp(l1.weight.data_ptr(), l1.bias.data_ptr(), l2.weight.data_ptr(), l2.bias.data_ptr(), l3.weight.data_ptr(), l3.bias.data_ptr(), 16, hidden[0], hidden[1], 4)
    B.check(lib.gr_policy_pack(C.byref(mlp), None, packed.data_ptr(), torch.cuda.current_stream().cuda_stream), "pack")
    sigma = torch.ones(4, device="cuda")
    polc = B.GrPolicy(packed.data_ptr(), sigma.data_ptr(), 0.01)
    params = (l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias)
    grads = [torch.zeros_like(t) for t in params]
    out = B.GrMlpGrad(*(t.data_ptr() for t in grads), 4, 0)
    scale = (1024.0 / g.abs().max()).reshape(1)
    B.check(lib.gr_actor_backward(C.byref(polc), hidden[0], hidden[1], x.data_ptr(), g.data_ptr(), scale.data_ptr(), rows, C.byref(out),
                                  torch.cuda.current_stream().cuda_stream), "gr_actor_backward")
    torch.cuda.synchronize()
    torch.autograd.backward([_emulated(pol.actor, x)], [g])
    for name, t, got in zip(("w1", "b1", "w2", "b2", "w3", "b3"), params, grads):
        ref = t.grad
        tol = 2e-2 if name.startswith("w") else 8e-2          # bias gradients are plain sums of zero-mean rows: cancellation
        assert float((got - ref).abs().max() / ref.abs().max()) < tol, (name, float((got - ref).abs().max() / ref.abs().max()))
    # against the fp32 module, with a cotangent that does not cancel
    g2 = (torch.tanh(x[:, :4]) + 1.5) * 1e-6
    grads2 = [torch.zeros_like(t) for t in params]
    out2 = B.GrMlpGrad(*(t.data_ptr() for t in grads2), 4, 0)
    scale2 = (1024.0 / g2.abs().max()).reshape(1)
    B.check(lib.gr_actor_backward(C.byref(polc), hidden[0], hidden[1], x.data_ptr(), g2.data_ptr(), scale2.data_ptr(), rows, C.byref(out2),
                                  torch.cuda.current_stream().cuda_stream), "gr_actor_backward")
    for t in params:
        t.grad = None
    torch.autograd.backward([pol.actor(x)], [g2])
    for name, t, got in zip(("w1", "b1", "w2", "b2", "w3", "b3"), params, grads2):
        assert float((got - t.grad).abs().max() / t.grad.abs().max()) < 5e-2, (name, float((got - t.grad).abs().max() / t.grad.abs().max()))


@pytest.mark.parametrize("rows", [149 * 128, 3 * 148 * 128 + 5, 300])
def test_two_tiles_in_flight_equals_one(cuda_lib, rows, monkeypatch):
    """16 -> 128 -> 128: the two-group kernel (two tiles in flight per CTA, shared weight-gradient accumulators in tensor memory) against
    the one-group kernel on the same inputs: identical fp16 roundings, so they may differ by the order of the fp32 accumulations only."""
    from generalizableracing_b200 import _lib as B
    from generalizableracing_b200.modules import BaseModel
    lib = cuda_lib
    torch.manual_seed(rows)
    pol = BaseModel(16, 16, 4, actor_hidden_dims=(128, 128), critic_hidden_dims=(128, 128), activation="lrelu").cuda()
    l1, l2, l3 = [m for m in pol.actor if isinstance(m, torch.nn.Linear)]
    x = torch.randn(rows, 16, device="cuda") * 3.0
    g = torch.randn(rows, 4, device="cuda") * 1e-6
    packed = torch.zeros(int(lib.gr_policy_packed_bytes(128, 128, 1)), dtype=torch.uint8, device="cuda")
    mlp = B.GrMlp(l1.weight.data_ptr(), l1.bias.data_ptr(), l2.weight.data_ptr(), l2.bias.data_ptr(), l3.weight.data_ptr(), l3.bias.data_ptr(), 16, 128, 128, 4)
    B.check(lib.gr_policy_pack(C.byref(mlp), None, packed.data_ptr(), torch.cuda.current_stream().cuda_stream), "pack")
    sigma = torch.ones(4, device="cuda")
    polc = B.GrPolicy(packed.data_ptr(), sigma.data_ptr(), 0.01)
    params = (l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias)
    scale = (1024.0 / g.abs().max()).reshape(1)
    res = {}
    for groups in ("1", "2"):
        monkeypatch.setenv("GRACING_ACTOR_BACKWARD_GROUPS", groups)
        grads = [torch.zeros_like(t) for t in params]
        out = B.GrMlpGrad(*(t.data_ptr() for t in grads), 4, 0)
        for _ in range(2):                     # accumulates: two launches = twice the gradient
            B.check(lib.gr_actor_backward(C.byref(polc), 128, 128, x.data_ptr(), g.data_ptr(), scale.data_ptr(), rows, C.byref(out),
                                          torch.cuda.current_stream().cuda_stream), "gr_actor_backward")
        torch.cuda.synchronize()
        res[groups] = grads
    for name, a, b in zip(("w1", "b1", "w2", "b2", "w3", "b3"), res["1"], res["2"]):
        assert float((a - b).abs().max() / a.abs().max()) < 2e-5, (name, float((a - b).abs().max() / a.abs().max()))

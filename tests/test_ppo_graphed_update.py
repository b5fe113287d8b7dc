"""PPO.update() replayed from a CUDA graph (gather -> forward -> losses -> backward -> clip -> Adam with the adaptive learning
rate on the device) against the eager loop of the reference (standalone/rsl_rl/ext/algorithms/ppo.py:103-190): same
mini-batches => same parameters, learning rate and reported losses."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _make(graphed, N=512, T=24, seed=3, kernel=False):
    from generalizableracing_b200.algorithms.ppo import PPO
    from generalizableracing_b200.modules import ActorCritic
    torch.manual_seed(seed)
    pol = ActorCritic(16, 16, 4).cuda()
    alg = PPO(pol, device="cuda:0", num_learning_epochs=5, num_mini_batches=4, schedule="adaptive", learning_rate=5e-4, gamma=0.99, lam=0.95,
              desired_kl=0.01, graphed_update=graphed, kernel_update=kernel)
    alg.init_storage("rl", N, T, [16], [16], [4])
    return alg


def _fill(alg, seed):
    """A synthetic rollout whose old policy is the current policy plus a perturbation (so that KL, clipping and the
    adaptive learning rate all get exercised)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    s = alg.storage
    T, N = s.num_transitions_per_env, s.num_envs
    with torch.no_grad():
        s.observations.copy_(torch.randn(T, N, 16, device="cuda", generator=g))
        s.privileged_observations.copy_(torch.randn(T, N, 16, device="cuda", generator=g))
        mu = alg.policy.actor(s.observations) + 0.05 * torch.randn(T, N, 4, device="cuda", generator=g)
        sigma = alg.policy.std.detach().expand(T, N, 4) * 1.02
        a = mu + sigma * torch.randn(T, N, 4, device="cuda", generator=g)
        s.mu.copy_(mu); s.sigma.copy_(sigma); s.actions.copy_(a)
        s.actions_log_prob.copy_(torch.distributions.Normal(mu, sigma).log_prob(a).sum(-1, keepdim=True))
        s.values.copy_(alg.policy.critic(s.privileged_observations) + 0.1 * torch.randn(T, N, 1, device="cuda", generator=g))
        s.rewards.copy_(torch.randn(T, N, 1, device="cuda", generator=g))
        s.dones.copy_((torch.rand(T, N, 1, device="cuda", generator=g) < 0.02).byte())
        s.compute_returns(torch.randn(N, 1, device="cuda", generator=g), 0.99, 0.95)
    s.step = T


def test_graphed_update_matches_eager(cuda_lib):
    eager, graphed = _make(False), _make(True)
    graphed.policy.load_state_dict(copy.deepcopy(eager.policy.state_dict()))
    probe = torch.randn(4096, 16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    for it in range(4):          # iteration 0 is eager on both sides (optimizer state), the graph is built at iteration 1
        for alg in (eager, graphed):
            _fill(alg, 100 + it)
            torch.manual_seed(7 + it)           # same randperm
            alg.last = alg.update()
        # Adam's update m / (sqrt(v) + 1e-8) is rounding noise wherever the gradient itself is ~1e-8 (weights behind inactive
        # units): such entries random-walk by ~lr per step in ANY two correct implementations (measured: first moments agree
        # to 6e-9, those weights differ by 2e-4).  Compare what the parameters compute, and the moments, not the raw entries.
        with torch.no_grad():
            for net_e, net_g in ((eager.policy.actor, graphed.policy.actor), (eager.policy.critic, graphed.policy.critic)):
                assert float((net_e(probe) - net_g(probe)).abs().max()) < (5e-3 if it <= 1 else 5e-2), it       # (measured 0, 1.9e-3, 2.6e-3 on random probes; the reported losses agree to 1e-6)
        m_e, m_g = eager.optimizer.state_dict()["state"], graphed.optimizer.state_dict()["state"]
        assert all(float(m_e[k]["step"]) == float(m_g[k]["step"]) for k in m_e)
        assert abs(eager.learning_rate - graphed.learning_rate) < 1e-6 * eager.learning_rate, (it, eager.learning_rate, graphed.learning_rate)
        assert abs(eager.last["value_function"] - graphed.last["value_function"]) < 1e-3 * abs(eager.last["value_function"]) + 1e-6
        assert abs(eager.last["surrogate"] - graphed.last["surrogate"]) < 2e-4 + 1e-2 * abs(eager.last["surrogate"])
    assert graphed._graph is not None


def test_kernel_update_tracks_eager(cuda_lib):
    """kernel_update: forward, loss gradients and weight gradients from the libgracing kernels (fp16 operands on the tensor
    cores).  Same mini-batches => the same losses to fp16-forward accuracy, and the policies stay close as functions."""
    eager, kern = _make(False, N=2048), _make(True, N=2048, kernel=True)
    kern.policy.load_state_dict(copy.deepcopy(eager.policy.state_dict()))
    probe = torch.randn(4096, 16, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    for it in range(3):
        for alg in (eager, kern):
            _fill(alg, 200 + it)
            torch.manual_seed(17 + it)
            alg.last = alg.update()
        with torch.no_grad():
            da = float((eager.policy.actor(probe) - kern.policy.actor(probe)).abs().max())
            dc = float((eager.policy.critic(probe) - kern.policy.critic(probe)).abs().max())
        print(it, eager.last, kern.last, eager.learning_rate, kern.learning_rate, da, dc)
        assert abs(eager.last["value_function"] - kern.last["value_function"]) < 2e-2 * abs(eager.last["value_function"]) + 1e-4, it
        assert abs(eager.last["surrogate"] - kern.last["surrogate"]) < 2e-3 + 5e-2 * abs(eager.last["surrogate"]), it
        assert da < 5e-2 and dc < 5e-2, (it, da, dc)
        assert float((eager.policy.std - kern.policy.std).abs().max()) < 5e-3
    assert kern._graph is not None and "adam_state" in kern._graph

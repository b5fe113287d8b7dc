"""rsl_rl rollout storage + GAE: CUDA kernels vs the oracle restatement of
standalone/rsl_rl/ext/storage/rollout_storage.py and ppo.py:85-97, on identical seeded inputs."""
import pytest
import torch

from oracle import rollout_oracle as RO

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]
TOL = 1e-5


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / max(1.0, float(a.abs().max())))


def _fill(cuda_lib, T, N, seed, done_p=0.02, to_p=0.01, od=16, cd=16):
    from generalizableracing_b200.storage import RolloutStorage
    g = torch.Generator().manual_seed(seed)
    sto = RolloutStorage("rl", N, T, [od], [cd], [4], device="cuda:0")
    ref = dict(obs=[], critic=[], actions=[], rewards=[], dones=[], values=[], logp=[], mu=[], sigma=[])
    gamma = 0.99
    for t in range(T):
        tr = sto.Transition()
        obs, cri, act = torch.randn(N, od, generator=g), torch.randn(N, cd, generator=g), torch.randn(N, 4, generator=g)
        rew, val = torch.randn(N, generator=g), torch.randn(N, 1, generator=g)
        dones = (torch.rand(N, generator=g) < done_p)
        tos = dones & (torch.rand(N, generator=g) < 0.5)
        logp, mu, sig = torch.randn(N, generator=g), torch.randn(N, 4, generator=g), torch.rand(N, 4, generator=g)
        tr.observations, tr.privileged_observations, tr.actions = obs.cuda(), cri.cuda(), act.cuda()
        tr.rewards, tr.values, tr.dones = rew.cuda(), val.cuda(), dones.long().cuda()
        tr.actions_log_prob, tr.action_mean, tr.action_sigma = logp.cuda(), mu.cuda(), sig.cuda()
        tr.time_outs, tr.gamma = tos.cuda(), gamma
        sto.add_transitions(tr)
        ref["obs"].append(obs); ref["critic"].append(cri); ref["actions"].append(act)
        ref["rewards"].append(RO.bootstrap_rewards(rew, val, tos, gamma).view(-1, 1)); ref["dones"].append(dones.view(-1, 1).byte())
        ref["values"].append(val); ref["logp"].append(logp.view(-1, 1)); ref["mu"].append(mu); ref["sigma"].append(sig)
    ref = {k: torch.stack(v) for k, v in ref.items()}
    return sto, ref, g


@pytest.mark.parametrize("T,N,od,cd", [(24, 4096, 16, 16), (1, 1, 16, 16), (7, 130, 16, 16), (24, 65536, 16, 16),
                                       (24, 1024, 17, 17), (5, 131, 17, 5)])         # 17: the reach-target observation (scalar row path)
def test_add_transitions_and_gae(cuda_lib, T, N, od, cd):
    sto, ref, g = _fill(cuda_lib, T, N, seed=T * 1000 + N, od=od, cd=cd)
    assert torch.equal(sto.observations.cpu(), ref["obs"]) and torch.equal(sto.privileged_observations.cpu(), ref["critic"])
    assert torch.equal(sto.actions.cpu(), ref["actions"]) and torch.equal(sto.mu.cpu(), ref["mu"]) and torch.equal(sto.sigma.cpu(), ref["sigma"])
    assert torch.equal(sto.dones.cpu(), ref["dones"]) and torch.equal(sto.values.cpu(), ref["values"])
    assert torch.equal(sto.actions_log_prob.cpu(), ref["logp"])
    assert _rel(ref["rewards"], sto.rewards) < TOL
    with pytest.raises(AssertionError):
        sto.add_transitions(sto.Transition())
    last = torch.randn(N, 1, generator=g)
    ret, adv = RO.compute_returns(sto.rewards.cpu(), ref["values"], ref["dones"], last, 0.99, 0.95)
    sto.compute_returns(last.cuda(), 0.99, 0.95)
    assert _rel(ret, sto.returns) < TOL
    if T * N > 1:
        assert _rel(adv, sto.advantages) < 5 * TOL     # includes the global mean/std reduction
        a = sto.advantages.double()
        assert abs(float(a.mean())) < 1e-4 and abs(float(a.std()) - 1.0) < 1e-4      # size-independent property


def test_split_normalisation_matches_single_shot(cuda_lib):
    """normalize=False + moments + normalize_advantages (the multi-GPU path) == one-shot normalisation."""
    sto, ref, g = _fill(cuda_lib, 24, 4096, seed=5)
    last = torch.randn(4096, 1, generator=g).cuda()
    sto.compute_returns(last, 0.99, 0.95)
    one = sto.advantages.clone()
    sto.compute_returns(last, 0.99, 0.95, normalize=False)
    raw = sto.advantages.clone()
    m = sto.moments.cpu()
    assert int(m[0]) == 24 * 4096
    assert abs(float(m[1]) - float(raw.double().mean())) < 1e-9 * max(1, abs(float(m[1]))) + 1e-7
    sto.normalize_advantages()
    assert torch.equal(one, sto.advantages)


@pytest.mark.parametrize("od", [16, 17])
def test_minibatch_gather(cuda_lib, od):
    sto, ref, g = _fill(cuda_lib, 24, 4096, seed=9, od=od, cd=od)
    sto.compute_returns(torch.randn(4096, 1, generator=g).cuda(), 0.99, 0.95)
    B = 24 * 4096
    idx = torch.randperm(B, generator=g)
    fields = [sto.observations, sto.privileged_observations, sto.actions, sto.values, sto.advantages, sto.returns,
              sto.actions_log_prob, sto.mu, sto.sigma]
    exp = RO.mini_batches([f.cpu() for f in fields], idx, 4, 2)
    got = sto.mini_batch_generator(4, 2, indices=idx.cuda())
    n = 0
    for e, k in zip(exp, got):
        for a, b in zip(e, k[:9]):
            assert torch.equal(a, b.cpu())
        assert k[9] == (None, None) and k[10] is None
        n += 1
    assert n == 8

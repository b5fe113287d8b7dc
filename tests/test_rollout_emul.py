"""The rollout-storage kernels (csrc/rollout.cu: add_transitions with the fused time-out bootstrap, GAE + normalisation, mini-batch gather)
compiled by g++ (tests/emul) behind this repo's RolloutStorage class, on CPU, against the reference's OWN RolloutStorage + the bootstrap
lines of its PPO.process_env_step (standalone/rsl_rl/ext/storage/rollout_storage.py, ext/algorithms/ppo.py:89-92; unmodified, where the
reference tree exists) or against the oracle restatement (pinned to the same file) elsewhere.  The B200 counterpart is
tests/test_rollout_parity.py."""
import pytest
import torch

from generalizableracing_b200.storage import RolloutStorage
from oracle import ref_modules, rollout_oracle as RO


def _run(emul_lib, T, N, od, cd, seed, use_reference):
    g = torch.Generator().manual_seed(seed)
    gamma, lam = 0.99, 0.95
    sto = RolloutStorage("rl", N, T, [od], [cd], [4], device="cpu", _lib=emul_lib)
    ref = ref_modules.load().RolloutStorage("rl", N, T, [od], [cd], [4], device="cpu") if use_reference else None
    cols = dict(rewards=[], values=[], dones=[])
    for t in range(T):
        obs, cri, act = torch.randn(N, od, generator=g), torch.randn(N, cd, generator=g), torch.randn(N, 4, generator=g)
        rew, val = torch.randn(N, generator=g), torch.randn(N, 1, generator=g)
        dones = torch.rand(N, generator=g) < 0.05
        tos = dones & (torch.rand(N, generator=g) < 0.5)
        logp, mu, sig = torch.randn(N, generator=g), torch.randn(N, 4, generator=g), torch.rand(N, 4, generator=g)
        tr = sto.Transition()
        tr.observations, tr.privileged_observations, tr.actions, tr.rewards, tr.values, tr.dones = obs, cri, act, rew, val, dones.long()
        tr.actions_log_prob, tr.action_mean, tr.action_sigma, tr.time_outs, tr.gamma = logp, mu, sig, tos, gamma
        sto.add_transitions(tr)
        boot = rew.clone() + gamma * torch.squeeze(val * tos.unsqueeze(1), 1)            # ppo.py:89-92, literally
        if ref is not None:
            rt = ref.Transition()
            rt.observations, rt.privileged_observations, rt.actions, rt.rewards, rt.values, rt.dones = obs, cri, act, boot, val, dones.long()
            rt.actions_log_prob, rt.action_mean, rt.action_sigma = logp, mu, sig
            ref.add_transitions(rt)
        cols["rewards"].append(boot.view(-1, 1)); cols["values"].append(val); cols["dones"].append(dones.view(-1, 1).byte())
    last = torch.randn(N, 1, generator=g)
    sto.compute_returns(last, gamma, lam)
    if ref is not None:
        ref.compute_returns(last, gamma, lam)
        for name in ("observations", "privileged_observations", "actions", "mu", "sigma", "values", "actions_log_prob", "dones"):
            assert torch.equal(getattr(sto, name), getattr(ref, name)), name
        ret, adv, rew = ref.returns, ref.advantages, ref.rewards
    else:
        rew = torch.stack(cols["rewards"])
        ret, adv = RO.compute_returns(rew, torch.stack(cols["values"]), torch.stack(cols["dones"]), last, gamma, lam)
    assert float((sto.rewards - rew).abs().max()) < 1e-6
    assert float((sto.returns - ret).abs().max()) < 1e-5 * max(1.0, float(ret.abs().max()))
    if T * N > 1:
        assert float((sto.advantages - adv).abs().max()) < 5e-5 * max(1.0, float(adv.abs().max()))
    # mini-batches: same permutation on both sides (the reference draws it with torch.randperm first thing, rollout_storage.py:158)
    if ref is not None and (T * N) % 4 == 0:
        torch.manual_seed(seed)
        idx = torch.randperm(T * N)
        torch.manual_seed(seed)
        for ours, theirs in zip(sto.mini_batch_generator(4, 2, indices=idx), ref.mini_batch_generator(4, 2)):
            for k in range(9):
                if k in (4, 5):              # advantages / returns: their values agree at the GAE tolerance (above); the gather itself is exact
                    assert float((ours[k] - theirs[k]).abs().max()) < 5e-5 * max(1.0, float(theirs[k].abs().max())), k
                else:
                    assert torch.equal(ours[k], theirs[k]), k
    return sto


@pytest.mark.parametrize("T,N,od,cd", [(24, 256, 16, 16), (1, 1, 16, 16), (7, 130, 16, 16), (24, 64, 17, 17), (5, 131, 17, 5)])
def test_storage_kernels_emulated_match_oracle(emul_lib, T, N, od, cd):
    _run(emul_lib, T, N, od, cd, seed=T * 1000 + N, use_reference=False)


@pytest.mark.skipif(not ref_modules.available(), reason="reference tree not present")
@pytest.mark.parametrize("T,N,od,cd", [(24, 256, 16, 16), (7, 130, 16, 16), (24, 64, 17, 17), (8, 36, 17, 5)])
def test_storage_kernels_emulated_match_reference_storage(emul_lib, T, N, od, cd):
    _run(emul_lib, T, N, od, cd, seed=T * 1000 + N + 1, use_reference=True)


def test_gather_rows_are_the_indexed_rows(emul_lib):
    sto = _run(emul_lib, 6, 50, 16, 16, seed=9, use_reference=False)
    idx = torch.randperm(300, generator=torch.Generator().manual_seed(1))
    flat = {k: getattr(sto, k).flatten(0, 1) for k in ("observations", "privileged_observations", "actions", "values", "advantages", "returns",
                                                        "actions_log_prob", "mu", "sigma")}
    for i, batch in enumerate(sto.mini_batch_generator(3, 1, indices=idx)):
        rows = idx[i * 100:(i + 1) * 100]
        for k, name in enumerate(flat):
            assert torch.equal(batch[k], flat[name][rows]), name


def test_transition_records_plain_and_in_mini_batch_order(emul_lib):
    """gr_storage_pack_records / _permuted (csrc/rollout.cu, emulated): record r = the nine columns of transition r (or perm[r]) side by
    side, so that mini-batch i of every epoch of rollout_storage.py:165-178 is the contiguous record slice [i*mb, (i+1)*mb) -- the rows the
    reference's generator yields for that mini-batch, in its order."""
    from generalizableracing_b200 import _lib as B
    T, N, nmb = 6, 50, 3
    sto = _run(emul_lib, T, N, 16, 16, seed=11, use_reference=False)
    cols = [getattr(sto, k).flatten(0, 1) for k in ("observations", "privileged_observations", "actions", "mu", "sigma", "actions_log_prob", "advantages",
                                                    "returns", "values")]
    want = torch.cat([c.reshape(T * N, -1) for c in cols], dim=1)
    rec = sto.pack_records().clone()
    assert rec.shape == (T * N, B.GR_RECORD_FLOATS) and torch.equal(rec, want)
    perm = torch.randperm(T * N, generator=torch.Generator().manual_seed(2))
    rec_p = sto.pack_records(perm).clone()
    assert torch.equal(rec_p, want[perm])
    mb = T * N // nmb
    for i, batch in enumerate(sto.mini_batch_generator(nmb, 1, indices=perm)):          # (obs, critic, actions, values, advantages, returns, logp, mu, sigma, ...)
        sl = rec_p[i * mb:(i + 1) * mb]
        assert torch.equal(sl[:, 0:16], batch[0]) and torch.equal(sl[:, 16:32], batch[1]) and torch.equal(sl[:, 32:36], batch[2])
        assert torch.equal(sl[:, 36:40], batch[7]) and torch.equal(sl[:, 40:44], batch[8])
        assert torch.equal(sl[:, 44:45], batch[6].reshape(-1, 1)) and torch.equal(sl[:, 45:46], batch[4].reshape(-1, 1))
        assert torch.equal(sl[:, 46:47], batch[5].reshape(-1, 1)) and torch.equal(sl[:, 47:48], batch[3].reshape(-1, 1))
    # a prefix of the permutation packs only that many records
    sto._records.fill_(-7.0)
    k = 100
    rec_k = sto.pack_records(perm[:k].contiguous())
    assert torch.equal(rec_k[:k], want[perm[:k]]) and bool((rec_k[k:] == -7.0).all())

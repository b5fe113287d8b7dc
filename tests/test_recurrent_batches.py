"""Recurrent mini-batches (S/rsl_rl/ext/storage/rollout_storage.py:194-254): the trajectory kernels
(csrc/traj.cu) behind RolloutStorage.reccurent_mini_batch_generator against the oracle restatement, which is itself pinned
(here, where /root/reference exists) against the reference's UNMODIFIED generator and against the known answer of rsl_rl's
split_and_pad_trajectories docstring."""
import pytest
import torch

from oracle import ref_modules
from oracle import rollout_oracle as RO


def test_split_and_pad_known_answer():
    """The a/b example of rsl_rl.utils.split_and_pad_trajectories: [a1..a4 | a5 a6], [b1 b2 | b3 b4 b5 | b6]."""
    T, N = 6, 2
    x = torch.arange(1, T * N + 1, dtype=torch.float32).view(N, T).t().reshape(T, N, 1).contiguous()      # x[t, n] = n*T + t + 1
    dones = torch.zeros(T, N, 1, dtype=torch.uint8)
    dones[3, 0] = 1
    dones[1, 1] = 1
    dones[4, 1] = 1
    padded, masks = RO.split_and_pad_trajectories(x, dones)
    assert padded.shape == (T, 5, 1) and masks.shape == (T, 5)
    rows = padded[:, :, 0].t().tolist()
    assert rows == [[1, 2, 3, 4, 0, 0], [5, 6, 0, 0, 0, 0], [7, 8, 0, 0, 0, 0], [9, 10, 11, 0, 0, 0], [12, 0, 0, 0, 0, 0]]
    assert masks.t().sum(1).tolist() == [4, 2, 2, 3, 1]
    assert torch.equal(RO.unpad_trajectories(padded, masks), x)


def _storage_inputs(T, N, D, H, L, seed, done_p=0.1):
    g = torch.Generator().manual_seed(seed)
    obs, cri = torch.randn(T, N, D, generator=g), torch.randn(T, N, D + 3, generator=g)
    dones = (torch.rand(T, N, 1, generator=g) < done_p).to(torch.uint8)
    rows = [torch.randn(T, N, 4, generator=g), torch.randn(T, N, 1, generator=g)]
    hid_a = [torch.randn(T, L, N, H, generator=g) for _ in range(2)]        # LSTM: (h, c)
    hid_c = [torch.randn(T, L, N, H, generator=g) for _ in range(2)]
    return obs, cri, dones, rows, hid_a, hid_c


@pytest.mark.skipif(not ref_modules.available(), reason="reference tree not present")
def test_oracle_generator_matches_reference_generator():
    """The reference RolloutStorage.reccurent_mini_batch_generator, executed where it lies with the restated
    split_and_pad_trajectories injected for the absent rsl_rl import, against the oracle's restatement of the generator."""
    import sys
    ref = ref_modules.load()
    mod = sys.modules["_gr_ref_rollout_storage"]
    mod.split_and_pad_trajectories = RO.split_and_pad_trajectories
    T, N, D, H, L = 12, 24, 5, 7, 2
    obs, cri, dones, rows, hid_a, hid_c = _storage_inputs(T, N, D, H, L, seed=2)
    sto = ref.RolloutStorage("rl", N, T, [D], [D + 3], [4], "cpu")
    sto.observations, sto.privileged_observations, sto.dones = obs.clone(), cri.clone(), dones.clone()
    sto.actions, sto.values = rows[0].clone(), rows[1].clone()
    sto.saved_hidden_states_a, sto.saved_hidden_states_c = [h.clone() for h in hid_a], [h.clone() for h in hid_c]
    got = list(sto.reccurent_mini_batch_generator(3, 2))
    exp = list(RO.recurrent_mini_batches(obs, cri, rows, dones, hid_a, hid_c, 3, 2))
    assert len(got) == len(exp) == 6
    for g_, e_ in zip(got, exp):
        assert torch.equal(g_[0], e_[0]) and torch.equal(g_[1], e_[1]) and torch.equal(g_[10], e_[5])
        assert torch.equal(g_[2], e_[2][0]) and torch.equal(g_[3], e_[2][1])
        for a, b in zip(g_[9][0], e_[3]):
            assert torch.equal(a, b)
        for a, b in zip(g_[9][1], e_[4]):
            assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("T,N,D,mbs,done_p", [(24, 256, 16, 4, 0.05), (12, 24, 5, 3, 0.3), (24, 4096, 17, 4, 0.02), (7, 33, 3, 1, 0.0), (5, 8, 4, 2, 1.0)])
def test_recurrent_generator_matches_oracle(cuda_lib, T, N, D, mbs, done_p):
    from generalizableracing_b200.storage import RolloutStorage
    H, L = 8, 2
    obs, cri, dones, rows, hid_a, hid_c = _storage_inputs(T, N, D, H, L, seed=T * N, done_p=done_p)
    sto = RolloutStorage("rl", N, T, [D], [D + 3], [4], device="cuda:0")
    sto.observations.copy_(obs), sto.privileged_observations.copy_(cri), sto.dones.copy_(dones)
    sto.actions.copy_(rows[0]), sto.values.copy_(rows[1])
    sto.saved_hidden_states_a, sto.saved_hidden_states_c = [h.cuda() for h in hid_a], [h.cuda() for h in hid_c]
    got = list(sto.reccurent_mini_batch_generator(mbs, 2))
    exp = list(RO.recurrent_mini_batches(obs, cri, rows, dones, hid_a, hid_c, mbs, 2))
    assert len(got) == len(exp)
    for g_, e_ in zip(got, exp):
        assert torch.equal(g_[0].cpu(), e_[0]) and torch.equal(g_[1].cpu(), e_[1]) and torch.equal(g_[10].cpu(), e_[5])
        assert torch.equal(g_[2].cpu(), e_[2][0]) and torch.equal(g_[3].cpu(), e_[2][1])
        for a, b in zip(g_[9][0], e_[3]):
            assert torch.equal(a.cpu(), b)
        for a, b in zip(g_[9][1], e_[4]):
            assert torch.equal(a.cpu(), b)


@pytest.mark.gpu
def test_split_unpad_round_trip_and_gradient(cuda_lib):
    """unpad(split(x)) == x at full size, and the adjoint of unpad is the pad of the cotangent (checked against torch indexing)."""
    from generalizableracing_b200.trajectories import split_and_pad_trajectories, unpad_trajectories
    g = torch.Generator().manual_seed(0)
    T, N, D = 24, 16384, 16
    x = torch.randn(T, N, D, generator=g).cuda()
    dones = (torch.rand(T, N, 1, generator=g) < 0.03).cuda()
    padded, masks = split_and_pad_trajectories(x, dones)
    ref_p, ref_m = RO.split_and_pad_trajectories(x.cpu(), dones.cpu().to(torch.uint8))
    assert torch.equal(padded.cpu(), ref_p) and torch.equal(masks.cpu(), ref_m)
    assert torch.equal(unpad_trajectories(padded, masks), x)
    # gradient
    Ts, Ns = 9, 37
    xs = torch.randn(Ts, Ns, 6, generator=g).cuda()
    ds = (torch.rand(Ts, Ns, 1, generator=g) < 0.2).cuda()
    p, m = split_and_pad_trajectories(xs, ds)
    p1 = p.clone().requires_grad_(True)
    p2 = p.clone().requires_grad_(True)
    w = torch.randn(Ts, Ns, 6, generator=g).cuda()
    (unpad_trajectories(p1, m) * w).sum().backward()
    (RO.unpad_trajectories(p2, m) * w).sum().backward()
    assert torch.equal(p1.grad, p2.grad)


@pytest.mark.gpu
def test_recurrent_ppo_runs_on_the_racing_env(cuda_lib):
    """OnPolicyRunner with ActorCriticRecurrent (GRU): rollouts save hidden states, the update consumes padded trajectories."""
    from generalizableracing_b200 import make_env
    from generalizableracing_b200.runners import OnPolicyRunner
    torch.manual_seed(0)
    env = make_env(num_envs=512, stage=1)
    cfg = {"num_steps_per_env": 24, "save_interval": 1000, "empirical_normalization": False,
           "policy": {"class_name": "ActorCriticRecurrent", "init_noise_std": 1.0, "actor_hidden_dims": [128, 128], "critic_hidden_dims": [128, 128],
                      "activation": "lrelu", "rnn_type": "gru", "rnn_hidden_dim": 64, "rnn_num_layers": 1},
           "algorithm": {"class_name": "PPO", "value_loss_coef": 1.0, "use_clipped_value_loss": True, "clip_param": 0.2, "entropy_coef": 0.0,
                         "num_learning_epochs": 2, "num_mini_batches": 4, "learning_rate": 5.0e-4, "schedule": "adaptive", "gamma": 0.99, "lam": 0.95,
                         "desired_kl": 0.01, "max_grad_norm": 1.0}}
    runner = OnPolicyRunner(env, cfg, device="cuda:0")
    before = [p.detach().clone() for p in runner.alg.policy.parameters()]
    hist = runner.learn(3, init_at_random_ep_len=True)
    assert all(torch.isfinite(torch.tensor(h["Loss/value_function"])) for h in hist)
    assert runner.alg.storage.saved_hidden_states_a[0].shape == (24, 1, 512, 64)
    assert any(not torch.equal(a, b) for a, b in zip(before, runner.alg.policy.parameters()))
    for p in runner.alg.policy.parameters():
        assert torch.isfinite(p).all()

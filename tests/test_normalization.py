"""EmpiricalNormalization (rsl_rl.modules, third-party: restated) as OnPolicyRunner uses it with ``empirical_normalization``
(standalone/rsl_rl/ext/runners/on_policy_runner.py:67-73,151-155,295-332): the batched running moments equal the pooled moments
of everything seen in training mode; eval mode and the ``until`` bound freeze them; state_dict round trip; inverse."""
import pytest
import torch

from generalizableracing_b200.modules import EmpiricalNormalization


def test_running_moments_equal_pooled_moments():
    torch.manual_seed(0)
    n = EmpiricalNormalization([16], until=1.0e8)
    xs = [torch.randn(int(b), 16) * (i + 1) + i for i, b in enumerate((64, 1, 300, 17, 128))]
    for x in xs:
        y = n(x)
    allx = torch.cat(xs).double()
    assert int(n.count) == allx.shape[0]
    assert float((n.mean.double() - allx.mean(0)).abs().max()) < 1e-5
    assert float((n._var.squeeze(0).double() - allx.var(0, unbiased=False)).abs().max() / allx.var(0, unbiased=False).max()) < 1e-5
    assert torch.allclose(y, (xs[-1] - n._mean) / (n._std + n.eps))            # the batch is normalised with the UPDATED moments
    assert torch.allclose(n.inverse(n(xs[0])), xs[0], atol=1e-4)


def test_eval_mode_and_until_freeze_the_moments():
    n = EmpiricalNormalization([4], until=100)
    n(torch.randn(60, 4))
    n(torch.randn(60, 4) + 5)                  # count 120 >= until afterwards
    m, c = n.mean, int(n.count)
    n(torch.randn(60, 4) - 9)                  # beyond `until`: no update
    assert torch.equal(n.mean, m) and int(n.count) == c == 120
    n2 = EmpiricalNormalization([4])
    n2(torch.randn(8, 4))
    n2.eval()
    m2 = n2.mean
    n2(torch.randn(8, 4) + 3)
    assert torch.equal(n2.mean, m2)


def test_state_dict_round_trip():
    n = EmpiricalNormalization([3])
    n(torch.randn(40, 3) * 2 + 1)
    n2 = EmpiricalNormalization([3])
    n2.load_state_dict(n.state_dict())
    assert n2._count_host == 40 and torch.equal(n2._mean, n._mean) and torch.equal(n2._std, n._std)
    x = torch.randn(5, 3)
    n.eval(); n2.eval()
    assert torch.equal(n(x), n2(x))


@pytest.mark.gpu
def test_ppo_runner_with_empirical_normalization(cuda_lib, tmp_path):
    from generalizableracing_b200 import make_env
    from generalizableracing_b200.runners import OnPolicyRunner
    cfg = {"num_steps_per_env": 8, "save_interval": 10 ** 9, "empirical_normalization": True,
           "policy": {"class_name": "ActorCritic", "init_noise_std": 1.0, "actor_hidden_dims": [64, 64], "critic_hidden_dims": [64, 64], "activation": "elu"},
           "algorithm": {"class_name": "PPO", "value_loss_coef": 1.0, "use_clipped_value_loss": True, "clip_param": 0.2, "entropy_coef": 0.0,
                         "num_learning_epochs": 2, "num_mini_batches": 2, "learning_rate": 5.0e-4, "schedule": "adaptive", "gamma": 0.99, "lam": 0.95,
                         "desired_kl": 0.01, "max_grad_norm": 1.0}}
    torch.manual_seed(0)
    env = make_env(num_envs=512, stage=1, track="synthetic")
    runner = OnPolicyRunner(env, cfg, log_dir=None, device="cuda:0")
    hist = runner.learn(2, init_at_random_ep_len=True)
    assert len(hist) == 2 and all(torch.isfinite(torch.tensor(h["Loss/value_function"])) for h in hist)
    assert int(runner.obs_normalizer.count) == 2 * 8 * 512
    assert float(runner.obs_normalizer.std.min()) > 0 and float((runner.obs_normalizer.mean).abs().max()) > 0
    path = str(tmp_path / "m.pt")
    runner.save(path)
    r2 = OnPolicyRunner(make_env(num_envs=512, stage=1, track="synthetic"), cfg, device="cuda:0")
    r2.load(path)
    assert torch.equal(r2.obs_normalizer._mean, runner.obs_normalizer._mean) and r2.obs_normalizer._count_host == 2 * 8 * 512
    pol = r2.get_inference_policy()
    obs = env.get_observations()[0]
    assert torch.allclose(pol(obs), runner.get_inference_policy()(obs))
    with pytest.raises(ValueError, match="not with empirical_normalization"):
        OnPolicyRunner(env, {**cfg, "fused_collection": True}, device="cuda:0")

"""The drop-in callers on the GPU: OnPolicyRunner (PPO, rsl_rl storage path) and AlgoRunner (BPTT, analytic sweep) run
a few iterations through the public API; full-size properties of the step kernel (BASELINE C4 / C2 / C3 shapes)."""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

PPO_CFG = {"num_steps_per_env": 24, "save_interval": 1000, "empirical_normalization": False,
           "policy": {"class_name": "ActorCritic", "init_noise_std": 1.0, "actor_hidden_dims": [128, 128], "critic_hidden_dims": [128, 128], "activation": "lrelu"},
           "algorithm": {"class_name": "PPO", "value_loss_coef": 1.0, "use_clipped_value_loss": True, "clip_param": 0.2, "entropy_coef": 0.0,
                         "num_learning_epochs": 5, "num_mini_batches": 4, "learning_rate": 5.0e-4, "schedule": "adaptive", "gamma": 0.99, "lam": 0.95,
                         "desired_kl": 0.01, "max_grad_norm": 1.0}}
BPTT_CFG = {"num_steps_per_env": 32, "max_iterations": 50, "save_interval": 1000, "empirical_normalization": False,
            "algorithm": {"class_name": "BPTT", "schedule": "CosineAnnealingLR", "optimizer": "AdamW", "learning_rate": 5e-4},
            "policy": {"class_name": "BaseModel", "actor_hidden_dims": [256, 128], "critic_hidden_dims": [256, 128], "activation": "lrelu", "init_noise_std": 1.0}}


def test_ppo_runner_c2(cuda_lib, tmp_path):
    """BASELINE C2: 4096 envs, 24-step rollouts + GAE."""
    from generalizableracing_b200 import make_env
    from generalizableracing_b200.runners import OnPolicyRunner
    torch.manual_seed(0)
    env = make_env(num_envs=4096, stage=1)
    runner = OnPolicyRunner(env, PPO_CFG, log_dir=str(tmp_path), device="cuda:0")
    hist = runner.learn(3, init_at_random_ep_len=True)
    assert len(hist) == 3 and all(torch.isfinite(torch.tensor(h["Loss/value_function"])) for h in hist)
    assert hist[-1]["Train/episodes"] > 0 and "Episode_Reward/progress_rewards" in hist[-1]
    assert (tmp_path / "progress.jsonl").exists()
    sto = runner.alg.storage
    a = sto.advantages.double()
    assert abs(float(a.mean())) < 1e-3 and abs(float(a.std()) - 1) < 1e-3
    pol = runner.get_inference_policy()
    assert pol(env.get_observations()[0]).shape == (4096, 4)
    runner.save(str(tmp_path / "m.pt"))
    runner.load(str(tmp_path / "m.pt"))


def test_ppo_runner_checkpoint_resume_with_kernel_update(cuda_lib, tmp_path):
    """save() after captured kernel updates, load() into a FRESH runner (and into the live one): the Adam moments / step count /
    adaptive learning rate survive, and both keep training through the captured path (on_policy_runner.py:288-318)."""
    from generalizableracing_b200 import make_env
    from generalizableracing_b200.runners import OnPolicyRunner
    cfg = dict(PPO_CFG, fused_collection=True)
    cfg["algorithm"] = dict(PPO_CFG["algorithm"], graphed_update=True, kernel_update=True)
    torch.manual_seed(0)
    r1 = OnPolicyRunner(make_env(num_envs=1024, stage=1), cfg, device="cuda:0")
    r1.learn(4, init_at_random_ep_len=True)
    assert r1.alg._graph is not None
    r1.save(str(tmp_path / "ck.pt"))
    ref_params = [p.detach().clone() for p in r1.alg.policy.parameters()]
    ref_state = {i: {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
                 for i, st in r1.alg.optimizer.state_dict()["state"].items()}
    steps_done = float(ref_state[0]["step"])
    assert steps_done == 4 * 5 * 4

    r2 = OnPolicyRunner(make_env(num_envs=1024, stage=1, seed=5), cfg, device="cuda:0")
    r2.load(str(tmp_path / "ck.pt"))
    assert r2.current_learning_iteration == 3          # the reference's counter is the last iteration index
    for p, q in zip(r2.alg.policy.parameters(), ref_params):
        assert torch.equal(p, q)
    st2 = r2.alg.optimizer.state_dict()["state"]
    for i, st in ref_state.items():
        assert torch.equal(st2[i]["exp_avg"], st["exp_avg"]) and torch.equal(st2[i]["exp_avg_sq"], st["exp_avg_sq"])
    hist = r2.learn(3)
    assert r2.alg._graph is not None and r2.current_learning_iteration == 5
    assert float(r2.alg.optimizer.state_dict()["state"][0]["step"]) == steps_done + 3 * 5 * 4
    assert all(torch.isfinite(torch.tensor(h["Loss/value_function"])) for h in hist)

    r1.load(str(tmp_path / "ck.pt"))                       # into the live runner: the graph is dropped and rebuilt from the loaded state
    assert r1.alg._graph is None
    r1.learn(2)
    assert r1.alg._graph is not None
    assert float(r1.alg.optimizer.state_dict()["state"][0]["step"]) == steps_done + 2 * 5 * 4
    for p in r1.alg.policy.parameters():
        assert torch.isfinite(p).all()


def test_bptt_runner_reduces_loss(cuda_lib):
    """BPTT on the differentiable closure: the analytic gradient must be a descent direction (loss goes down)."""
    from generalizableracing_b200 import make_env
    from generalizableracing_b200.runners import AlgoRunner
    torch.manual_seed(0)
    env = make_env(num_envs=2048, stage=0, track="figure8", differentiable=True, bptt_horizon=32)
    cfg = dict(BPTT_CFG)
    cfg["policy"] = dict(cfg["policy"], init_noise_std=0.1)
    runner = AlgoRunner(env, cfg, device="cuda:0")
    hist = runner.learn(40, init_at_random_ep_len=True)
    first = sum(h["Loss/mean_total_loss"] for h in hist[:5]) / 5
    last = sum(h["Loss/mean_total_loss"] for h in hist[-5:]) / 5
    print("BPTT loss", first, "->", last)
    assert last < first


def test_bptt_runner_checkpoints(cuda_lib, tmp_path):
    """runner.py:193-199: interval checkpoints hold the weights of THEIR iteration (also with deferred logging), and the run ends with
    model_{last iteration}.pt -- the last updates are never lost."""
    import os
    from generalizableracing_b200 import make_env
    from generalizableracing_b200.runners import AlgoRunner
    torch.manual_seed(0)
    env = make_env(num_envs=256, stage=0, track="figure8", differentiable=True, bptt_horizon=8)
    cfg = dict(BPTT_CFG, num_steps_per_env=8, save_interval=4, log_interval=3)
    runner = AlgoRunner(env, cfg, log_dir=str(tmp_path), device="cuda:0")
    snaps = {}
    orig = runner.alg.update

    def update():
        out = orig()
        snaps[len(snaps)] = {k: v.clone() for k, v in runner.alg.actor_critic.state_dict().items()}
        return out
    runner.alg.update = update
    runner.learn(7)
    files = sorted(f for f in os.listdir(tmp_path) if f.startswith("model_"))
    assert files == ["model_0.pt", "model_4.pt", "model_6.pt"], files
    for it in (0, 4, 6):
        ck = torch.load(str(tmp_path / f"model_{it}.pt"), map_location="cuda:0", weights_only=False)
        assert ck["iter"] == it
        for k, v in ck["model_state_dict"].items():
            assert torch.equal(v, snaps[it][k]), (it, k)
    with pytest.raises(ValueError, match="empirical_normalization"):
        AlgoRunner(env, dict(cfg, empirical_normalization=True), device="cuda:0")


def test_bptt_algorithm_fast_path_equals_autograd_path(cuda_lib):
    """BPTT.update through backward_window == BPTT through the chained autograd Function (same policy gradient)."""
    from generalizableracing_b200 import make_env
    from generalizableracing_b200.algorithms import BPTT
    from generalizableracing_b200.modules import BaseModel
    grads = []
    for fast in (True, False):
        torch.manual_seed(1)
        env = make_env(num_envs=512, stage=0, track="figure8", differentiable=True, bptt_horizon=16, seed=3)
        model = BaseModel(16, 16, 4, actor_hidden_dims=[64, 64], critic_hidden_dims=[64, 64], activation="lrelu", init_noise_std=0.3)
        alg = BPTT(model, max_iterations=10, optimizer="SGD", learning_rate=0.0, env=env if fast else None)
        env._bptt.autograd = not fast
        obs, _ = env.reset()
        env.detach()
        for t in range(16):
            a = alg.act(obs)
            obs, rew, dones, ex = env.step(a)
            alg.process_env_step(ex["losses"], ex["losses_detached"], dones, rew, ex)
        alg.update()
        grads.append(torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None]).clone())
    assert float((grads[0] - grads[1]).abs().max() / grads[1].abs().max()) < 1e-4


def test_full_size_properties_c4(cuda_lib):
    """65,536 envs with per-env DR and resets (BASELINE C4): size-independent invariants after 300 steps."""
    from generalizableracing_b200 import make_env
    env = make_env(num_envs=65536, stage=1)
    env.reset()
    env.episode_length_buf = torch.randint(0, 200, (65536,), device="cuda:0", dtype=torch.int32)
    g = torch.Generator(device="cuda:0").manual_seed(0)
    n_done = 0
    for t in range(300):
        a = torch.randn(65536, 4, device="cuda:0", generator=g) * 0.3 + torch.tensor([-0.35, 0, 0, 0], device="cuda:0")
        obs, rew, dones, ex = env.step(a)
        n_done += int(dones.sum())
        assert torch.equal(dones.bool(), ex["time_outs"] | ex["terminated"])
    sv = env.state_dict_view()
    assert torch.isfinite(obs).all() and torch.isfinite(rew).all() and torch.isfinite(env.planes).all()
    assert float((sv["root_quat_w"].norm(dim=-1) - 1).abs().max()) < 1e-5                 # quaternions stay normalised
    assert int(sv["episode_length"].max()) < 200 and int(sv["episode_length"].min()) >= 0
    assert int(sv["gate_id"].max()) < 8 and int(sv["terrain_levels"].max()) < 10 and int(sv["terrain_types"].max()) == 19
    assert n_done >= 65536                                                                  # every env timed out at least once
    # critic rows 3..5 are a rotation-matrix row: unit norm
    crit = ex["observations"]["critic"]
    assert float((crit[:, 3:6].norm(dim=-1) - 1).abs().max()) < 1e-5
    # obs last-action block == ctbr(tanh(a_{t-1})) bounds
    assert float(obs[:, 13:].abs().max()) <= 6.0 + 1e-4 and float(obs[:, 12].min()) >= -1e-4
    log = ex["log"]
    assert float(log["Metrics/next_gate_pose/accumulate_gates"]) >= 0

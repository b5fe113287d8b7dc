"""Drop-in of the env under the reference's OWN PPO trainer code: the unmodified ``PPO`` (standalone/rsl_rl/ext/algorithms/ppo.py) and
``RolloutStorage`` (ext/storage/rollout_storage.py) run the collection loop of ext/runners/on_policy_runner.py:141-157 twice on CPU --
once on the reference's env (its own ManagerBasedDiffRLEnv.step over the closure simulator, oracle/ref_closure.py) and once on
``RacingVecEnv`` (kernel sources through the g++ emulation) -- with the same policy, action noise and env draws.  What the trainer sees
through the RslRlVecEnvWrapper surface (obs, critic obs, rewards, long dones, extras["time_outs"]) must lead to the same rollout storage
and, after update(), the same weights.  Skipped on the GPU box (needs the reference tree)."""
import copy
import importlib.util
import os

import pytest
import torch

from generalizableracing_b200 import layout as L_
from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.env import RacingVecEnv
from generalizableracing_b200.modules import ActorCritic
from generalizableracing_b200.tracks import synthetic_track_table
from oracle import ref_modules as RM
from tests import parity_cases as PC

pytestmark = pytest.mark.skipif(not RM.available(), reason="reference tree not present")


def _golden_tools():
    spec = importlib.util.spec_from_file_location("_make_ppo_golden", os.path.join(os.path.dirname(__file__), "golden", "make_ppo_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("trainer", ["reference", "repo"])
def test_reference_ppo_trainer_on_kernels_matches_reference_env(emul_lib, trainer):
    """trainer = "reference": the reference's PPO + RolloutStorage on RacingVecEnv (env drop-in); "repo": this repo's PPO with its RolloutStorage
    on the emulated storage kernels (csrc/rollout.cu) on RacingVecEnv -- the whole product stack of the PPO path, on CPU.  Both against the
    reference's PPO on the reference's env."""
    from generalizableracing_b200.algorithms.ppo import PPO as RepoPPO
    from generalizableracing_b200.storage import RolloutStorage
    from oracle import ref_closure as RC
    tools = _golden_tools()
    PPO = tools.load_reference_ppo()
    N, T, K = 64, 24, 2
    cfg, table = RacingCfg.for_stage(1), synthetic_track_table()
    g = torch.Generator().manual_seed(12)
    ref, srnd = RC.make_reference_env(cfg, table, N, PC.draw_startup(N, g), seed=6000)
    env = RacingVecEnv(cfg, table, N, device="cpu", rng_mode="dense", startup_rnd=srnd, _lib=emul_lib)
    ter = ref.scene.terrain
    torch.manual_seed(0)
    pol_r = ActorCritic(16, 16, 4, actor_hidden_dims=[128, 128], critic_hidden_dims=[128, 128], activation="lrelu", init_noise_std=1.0)
    pol_k = copy.deepcopy(pol_r)
    alg_r = PPO(pol_r, None, device="cpu", **tools.ALG)
    alg_r.init_storage("rl", N, T, [16], [16], [4])
    if trainer == "reference":
        alg_k = PPO(pol_k, None, device="cpu", **tools.ALG)
        alg_k.init_storage("rl", N, T, [16], [16], [4])
    else:
        alg_k = RepoPPO(pol_k, None, device="cpu", **tools.ALG)
        alg_k.storage = RolloutStorage("rl", N, T, [16], [16], [4], device="cpu", _lib=emul_lib)
    ids = torch.arange(N)
    rnd = torch.zeros(N, L_.RND_STRIDE)
    rnd[:, L_.RND_LEVEL] = torch.rand(N, generator=g)
    ter.pending_level_u = rnd[:, L_.RND_LEVEL].clone()
    torch.manual_seed(1)
    ref._reset_idx(ids)
    o = ref.observation_manager.compute()
    obs_r, critic_r = o["policy"], o["critic"]
    torch.manual_seed(1)
    RC.replay_reset_draws(rnd, ids, cfg.add_cmd_noise)
    RC.replay_obs_draws(rnd)
    env.reset(rnd)
    obs_k, ex_k = env.get_observations()                        # on_policy_runner.py:122-124
    critic_k = ex_k["observations"]["critic"]
    ep = torch.randint(0, cfg.max_episode_length, (N,), generator=g)      # init_at_random_ep_len (on_policy_runner.py:117-120)
    ref.episode_length_buf[:] = ep
    env.episode_length_buf = ep
    n_done = 0
    for it in range(K):
        with torch.inference_mode():                            # on_policy_runner.py:141
            for t in range(T):
                torch.manual_seed(10_000 + it * T + t)
                a_r = alg_r.act(obs_r, critic_r)
                torch.manual_seed(10_000 + it * T + t)
                a_k = alg_k.act(obs_k, critic_k)
                rnd = torch.zeros(N, L_.RND_STRIDE)
                rnd[:, L_.RND_LEVEL] = torch.rand(N, generator=g)
                ter.pending_level_u = rnd[:, L_.RND_LEVEL].clone()
                torch.manual_seed(20_000 + it * T + t)
                o, rew_r, terminated, time_outs, _ = ref.step(a_r)
                reset_ids = ref.reset_buf.nonzero(as_tuple=False).squeeze(-1)
                torch.manual_seed(20_000 + it * T + t)
                RC.replay_reset_draws(rnd, reset_ids, cfg.add_cmd_noise)
                RC.replay_pass_draws(rnd, ref.command_manager.last_achieved, cfg.add_cmd_noise)
                RC.replay_obs_draws(rnd)
                # what RslRlVecEnvWrapper.step hands the runner: obs, rew, dones (long), extras with time_outs
                obs_r, critic_r, dones_r = o["policy"], o["critic"], (terminated | time_outs).to(torch.long)
                alg_r.process_env_step(rew_r, dones_r, {"time_outs": time_outs})
                obs_k, rew_k, dones_k, ex_k = env.step(a_k, rnd)
                critic_k = ex_k["observations"]["critic"]
                alg_k.process_env_step(rew_k, dones_k, ex_k)
                assert dones_k.dtype == torch.long and torch.equal(dones_k, dones_r) and torch.equal(ex_k["time_outs"], time_outs), (it, t)
                n_done += int(dones_r.sum())
            alg_r.compute_returns(critic_r)                     # on_policy_runner.py:181
            alg_k.compute_returns(critic_k)
        sr, sk = alg_r.storage, alg_k.storage
        for name in ("observations", "privileged_observations", "actions", "rewards", "values", "returns", "advantages", "actions_log_prob"):
            assert PC.rel_err(getattr(sr, name), getattr(sk, name)) < (1e-3 if it else 1e-4), (it, name)
        assert torch.equal(sr.dones.long(), sk.dones.long()), it
        torch.manual_seed(30_000 + it)
        loss_r = alg_r.update()
        torch.manual_seed(30_000 + it)
        loss_k = alg_k.update()
        diffs = torch.cat([(p - q).abs().flatten() for p, q in zip(pol_r.parameters(), pol_k.parameters())])
        print(f"{trainer} trainer on the kernels, iteration {it}: value loss {loss_r['value_function']:.5f} / {loss_k['value_function']:.5f}, lr {alg_r.learning_rate:.3e} / {alg_k.learning_rate:.3e}, "
              f"weights differ by <= {float(diffs.max()):.2e}, > 2e-4: {int((diffs > 2e-4).sum())} of {diffs.numel()}")
        assert alg_r.learning_rate == alg_k.learning_rate, it
        assert abs(loss_r["value_function"] - loss_k["value_function"]) < 1e-3 and abs(loss_r["surrogate"] - loss_k["surrogate"]) < 1e-3, it
        if it == 0:      # Adam after 20 sign-like steps: see tests/test_ppo_reference_golden.py for the measured sensitivity
            assert float(diffs.max()) < 2e-3 and int((diffs > 2e-4).sum()) <= diffs.numel() // 200
    assert n_done > N // 2


def test_reference_ppo_trainer_on_reach_kernels_matches_reference_env(emul_lib):
    """The same drop-in check for the CTBR reach-target task (17-wide observation, no critic group: the runner falls back to the policy
    observation, on_policy_runner.py:124): the reference's PPO + RolloutStorage on the reference's env vs on ReachTargetVecEnv (emulation)."""
    from generalizableracing_b200.config import ReachTargetCfg
    from generalizableracing_b200.reach_env import ReachTargetVecEnv
    from oracle import ref_closure as RC
    tools = _golden_tools()
    PPO = tools.load_reference_ppo()
    N, T = 64, 24
    cfg = ReachTargetCfg.ctbr(episode_length_s=0.6, resampling_time=0.3, is_differentiable_physics=False)
    ref = RC.make_reference_reach_env(cfg, N, seed=8000)
    env = ReachTargetVecEnv(cfg, N, device="cpu", rng_mode="dense", _lib=emul_lib)
    cmd = ref.command_manager.get_term("desired_pos_b")
    torch.manual_seed(0)
    pol_r = ActorCritic(17, 17, 4, actor_hidden_dims=[128, 128], critic_hidden_dims=[128, 128], activation="lrelu", init_noise_std=1.0)
    pol_k = copy.deepcopy(pol_r)
    alg_r, alg_k = PPO(pol_r, None, device="cpu", **tools.ALG), PPO(pol_k, None, device="cpu", **tools.ALG)
    for alg in (alg_r, alg_k):
        alg.init_storage("rl", N, T, [17], [17], [4])
    ids = torch.arange(N)
    rnd = torch.zeros(N, L_.REACH_RND_STRIDE)
    torch.manual_seed(1)
    ref._reset_idx(ids)
    cmd._update_command()                                       # reach_oracle R.5
    obs_r = ref.observation_manager.compute()["policy"]
    torch.manual_seed(1)
    RC.replay_reach_reset_draws(rnd, ids, cfg.random_drag)
    env.reset(rnd)
    obs_k, ex_k = env.get_observations()
    assert "critic" not in ex_k["observations"]
    n_done = 0
    with torch.inference_mode():
        for t in range(T):
            torch.manual_seed(10_000 + t)
            a_r = alg_r.act(obs_r, obs_r)
            torch.manual_seed(10_000 + t)
            a_k = alg_k.act(obs_k, ex_k["observations"].get("critic", obs_k))
            rnd = torch.zeros(N, L_.REACH_RND_STRIDE)
            torch.manual_seed(20_000 + t)
            o, rew_r, terminated, time_outs, _ = ref.step(a_r)
            reset_ids = ref.reset_buf.nonzero(as_tuple=False).squeeze(-1)
            torch.manual_seed(20_000 + t)
            RC.replay_reach_reset_draws(rnd, reset_ids, cfg.random_drag)
            RC.replay_reach_command_draws(rnd, ref.command_manager.last_timer_ids, L_.REACH_RND_CMD_TIMER)
            obs_r, dones_r = o["policy"], (terminated | time_outs).to(torch.long)
            alg_r.process_env_step(rew_r, dones_r, {"time_outs": time_outs})
            obs_k, rew_k, dones_k, ex_k = env.step(a_k, rnd)
            alg_k.process_env_step(rew_k, dones_k, ex_k)
            assert dones_k.dtype == torch.long and torch.equal(dones_k, dones_r) and torch.equal(ex_k["time_outs"], time_outs), t
            n_done += int(dones_r.sum())
        alg_r.compute_returns(obs_r)
        alg_k.compute_returns(obs_k)
    for name in ("observations", "actions", "rewards", "values", "returns", "advantages", "actions_log_prob"):
        assert PC.rel_err(getattr(alg_r.storage, name), getattr(alg_k.storage, name)) < 2e-4, name
    torch.manual_seed(30_000)
    alg_r.update()
    torch.manual_seed(30_000)
    alg_k.update()
    diffs = torch.cat([(p - q).abs().flatten() for p, q in zip(pol_r.parameters(), pol_k.parameters())])
    print(f"reach: lr {alg_r.learning_rate:.3e} / {alg_k.learning_rate:.3e}, weights differ by <= {float(diffs.max()):.2e}")
    assert alg_r.learning_rate == alg_k.learning_rate and n_done > N // 2
    assert float(diffs.max()) < 2e-3 and int((diffs > 2e-4).sum()) <= diffs.numel() // 200

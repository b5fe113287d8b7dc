#!/usr/bin/env python
"""Generates tests/golden/ref_training_curves.json: learning curves of the reference's OWN trainers on the reference's OWN env step.

PPO  -- the unmodified ``PPO`` + ``RolloutStorage`` (standalone/rsl_rl/ext/algorithms/ppo.py, ext/storage/rollout_storage.py) running the
        collection / update loop of ext/runners/on_policy_runner.py:135-183 on ``ManagerBasedDiffRLEnv.step`` over the closure simulator
        (oracle/ref_closure.py), STAGE 1, hyper-parameters of QD/agents/rsl_rl_ppo_cfg.py:35-48, 256 envs x 24 steps x 150 iterations,
        init_at_random_ep_len, three seeds.
BPTT -- the unmodified ``BPTT`` (standalone/diff_rl/algorithms/bptt.py) on the reference's CTBR reach-target env, 48-step windows
        (QD/agents/diff_rl_naive_cfg.py:9-32), 256 envs x 60 iterations, three seeds.

Per iteration: mean step reward, resets per env-step, gates passed per env-step (PPO) / mean loss (BPTT).  tests/test_training_curves.py runs
this repo's trainers on the kernels with the same schedules and compares the curves statistically (band of the reference seeds).
Run in the build container (needs /root/reference; about ten minutes on 8 cores):
    python tests/golden/make_training_curves.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from generalizableracing_b200.config import RacingCfg, ReachTargetCfg  # noqa: E402
from generalizableracing_b200.modules import ActorCritic, BaseModel  # noqa: E402
from generalizableracing_b200.tracks import synthetic_track_table  # noqa: E402
from oracle import ref_closure as RC  # noqa: E402
from tests import parity_cases as PC  # noqa: E402
from tests.golden.make_ppo_golden import ALG, load_reference_ppo  # noqa: E402

PPO_SCHEDULE = dict(num_envs=256, steps=24, iterations=150, stage=1)
BPTT_SCHEDULE = dict(num_envs=256, steps=48, iterations=60, learning_rate=5e-4)


def ppo_curve(seed, num_envs, steps, iterations, stage):
    PPO = load_reference_ppo()
    cfg, table = RacingCfg.for_stage(stage), synthetic_track_table()
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    env, _ = RC.make_reference_env(cfg, table, num_envs, PC.draw_startup(num_envs, g), seed=1000 + seed)
    env.scene.terrain.pending_level_u = torch.rand(num_envs, generator=g)
    policy = ActorCritic(16, 16, 4, actor_hidden_dims=[128, 128], critic_hidden_dims=[128, 128], activation="lrelu", init_noise_std=1.0)
    alg = PPO(policy, None, device="cpu", **ALG)
    alg.init_storage("rl", num_envs, steps, [16], [16], [4])
    env._reset_idx(torch.arange(num_envs))
    obs = env.observation_manager.compute()
    obs, critic = obs["policy"], obs["critic"]
    env.episode_length_buf[:] = torch.randint(0, cfg.max_episode_length, (num_envs,), generator=g)      # on_policy_runner.py:118-121
    curve = []
    for it in range(iterations):
        rew_sum, n_done, n_gate = 0.0, 0, 0
        with torch.inference_mode():
            for t in range(steps):
                actions = alg.act(obs, critic)
                env.scene.terrain.pending_level_u = torch.rand(num_envs, generator=g)
                o, rew, terminated, time_outs, _ = env.step(actions)
                obs, critic = o["policy"], o["critic"]
                dones = (terminated | time_outs).to(torch.long)
                alg.process_env_step(rew, dones, {"time_outs": time_outs})
                rew_sum += float(rew.mean())
                n_done += int(dones.sum())
                n_gate += int(env.command_manager.last_achieved.sum())
            alg.compute_returns(critic)
        loss = alg.update()
        curve.append({"mean_step_reward": rew_sum / steps, "resets_per_env_step": n_done / (steps * num_envs), "gates_per_env_step": n_gate / (steps * num_envs),
                      "value_loss": loss["value_function"], "learning_rate": alg.learning_rate})
        if it % 10 == 0:
            print(f"ppo seed {seed} it {it}: {curve[-1]}", flush=True)
    return curve


def bptt_curve(seed, num_envs, steps, iterations, learning_rate):
    """runner.py:107-155 of standalone/diff_rl with the reference's BPTT class on its CTBR reach-target env"""
    import importlib.util
    from oracle import ref_modules as RM
    spec = importlib.util.spec_from_file_location("_gr_ref_bptt_tools", os.path.join(ROOT, "tests", "test_bptt_trainer_vs_reference.py"))
    tools = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tools)
    BPTT = tools._load_reference_bptt()
    cfg = ReachTargetCfg.ctbr()
    torch.manual_seed(seed)
    env = RC.make_reference_reach_env(cfg, num_envs, seed=2000 + seed)
    term = env.action_manager.get_term("force_torque")
    dyn = term.drone_dynamics

    def fresh_copies(env_ids):        # the reference rewrites these in place after autograd saved them: same values, new tensors
        term.thr_est_error = term.thr_est_error.clone()
        dyn.drag_coeffs, dyn.h_force_drag_coeffs = dyn.drag_coeffs.clone(), dyn.h_force_drag_coeffs.clone()
    env.recorder_manager.pre_reset_hook = fresh_copies
    model = BaseModel(17, 17, 4, actor_hidden_dims=[256, 128], critic_hidden_dims=[256, 128], activation="lrelu", init_noise_std=0.3)
    alg = BPTT(actor_critic=model, max_iterations=iterations, device="cpu", schedule="CosineAnnealingLR", optimizer="AdamW", learning_rate=learning_rate)
    env._reset_idx(torch.arange(num_envs))
    env.command_manager.get_term("desired_pos_b")._update_command()
    obs = env.observation_manager.compute()["policy"]
    curve = []
    for it in range(iterations):
        env.detach()
        term.action_buffer = [a.detach() for a in term.action_buffer]
        rew_sum, n_done = 0.0, 0
        for t in range(steps):
            actions = alg.act(obs)
            o, rew, terminated, time_outs, ex = env.step(actions)
            obs = o["policy"]
            dones = (terminated | time_outs).to(torch.long)
            alg.process_env_step(ex["losses"], ex["losses_detached"], dones, rew, ex)
            rew_sum += float(rew.mean())
            n_done += int(dones.sum())
        _, loss = alg.update()
        curve.append({"mean_loss": float(loss.detach()), "mean_step_reward": rew_sum / steps, "resets_per_env_step": n_done / (steps * num_envs)})
        if it % 10 == 0:
            print(f"bptt seed {seed} it {it}: {curve[-1]}", flush=True)
    return curve


def main():
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    which = sys.argv[1:] or ["ppo", "bptt"]
    path = os.path.join(ROOT, "tests", "golden", "ref_training_curves.json")
    out = json.load(open(path)) if os.path.isfile(path) else {}
    if "ppo" in which:
        out["ppo"] = {"schedule": PPO_SCHEDULE, "alg": ALG, "seeds": {str(s): ppo_curve(s, **PPO_SCHEDULE) for s in (0, 1, 2)}}
    if "bptt" in which:
        out["bptt"] = {"schedule": BPTT_SCHEDULE, "seeds": {str(s): bptt_curve(s, **BPTT_SCHEDULE) for s in (0, 1, 2)}}
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Collection throughput (env-steps/s INCLUDING policy inference and rollout storage) of one PPO rollout of T = 24 steps:
fused (gr_ppo_collect, one launch) vs step-by-step (torch ActorCritic + gr_step_fwd + gr_storage_add), at the C2 and C4
env counts.  Device time by CUDA events; the step-by-step path is also timed under a CUDA graph (its best case)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from generalizableracing_b200.algorithms.ppo import PPO  # noqa: E402
from generalizableracing_b200.collect import FusedCollector  # noqa: E402
from generalizableracing_b200.config import RacingCfg  # noqa: E402
from generalizableracing_b200.env import RacingVecEnv  # noqa: E402
from generalizableracing_b200.modules import ActorCritic  # noqa: E402
from generalizableracing_b200.tracks import synthetic_track_table  # noqa: E402


def timeit(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n      # ms


def main():
    T = 24
    out = {}
    groups = [int(x) for x in os.environ.get("TILE_GROUPS", "0").split(",")]
    for N in [int(x) for x in os.environ.get("ENVS", "4096,16384,65536").split(",")]:
        cfg, table = RacingCfg.for_stage(1), synthetic_track_table()
        env = RacingVecEnv(cfg, table, N)
        env.reset()
        env.episode_length_buf = torch.randint(0, cfg.max_episode_length, (N,), dtype=torch.int32)
        pol = ActorCritic(16, 16, 4).cuda()
        alg = PPO(pol, device="cuda:0", gamma=0.99)
        alg.init_storage("rl", N, T, [16], [16], [4])
        row = {}
        for G in groups:
            col = FusedCollector(env, pol, alg.storage, gamma=0.99, groups_per_cta=G)
            col.pack()

            def fused():
                alg.storage.clear()
                col.collect()

            reps = int(os.environ.get("REPS", "1"))
            ms = sorted(timeit(fused, 20 if reps == 1 else 100) for _ in range(reps))[reps // 2]          # REPS > 1: median of REPS x 100 rollouts
            row[f"fused_G{G}"] = {"ms_per_rollout": ms, "us_per_step": ms * 1e3 / T, "env_steps_per_s": N * T / (ms * 1e-3)}

        if os.environ.get("SKIP_EAGER"):
            out[str(N)] = row
            env.close()
            continue

        def unfused():
            alg.storage.clear()
            obs, ex = env.get_observations()
            critic = ex["observations"]["critic"]
            with torch.inference_mode():
                for _ in range(T):
                    a = alg.act(obs, critic)
                    obs, r, d, info = env.step(a)
                    critic = info["observations"]["critic"]
                    alg.process_env_step(r, d, info)

        ms = timeit(unfused, 5)
        row["step_by_step_eager"] = {"ms_per_rollout": ms, "us_per_step": ms * 1e3 / T, "env_steps_per_s": N * T / (ms * 1e-3)}
        out[str(N)] = row
        env.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

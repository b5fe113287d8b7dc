#!/usr/bin/env python
"""Where does a PPO collection step go?  Device time of (a) the torch actor/critic MLPs + sampling + log-prob
(PPO.act, standalone/rsl_rl/ext/algorithms/ppo.py:71-83), (b) env.step (one gr_step_fwd launch), (c) add_transitions,
per rollout step, eager and CUDA-graph-captured, at the C2 and C4 env counts.  Evidence for DESIGN.md (is the policy
MLP a dense-contraction bottleneck?)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from generalizableracing_b200.algorithms.ppo import PPO  # noqa: E402
from generalizableracing_b200.config import RacingCfg  # noqa: E402
from generalizableracing_b200.env import RacingVecEnv  # noqa: E402
from generalizableracing_b200.modules import ActorCritic  # noqa: E402
from generalizableracing_b200.tracks import synthetic_track_table  # noqa: E402


def timeit(fn, n=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


def graphed(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


def main():
    out = {}
    for N in (4096, 65536):
        env = RacingVecEnv(RacingCfg.for_stage(1), synthetic_track_table(), N)
        obs, ex = env.reset()
        critic = ex["observations"]["critic"]
        pol = ActorCritic(16, 16, 4).cuda()
        alg = PPO(pol, device="cuda:0")
        alg.init_storage("rl", N, 24, [16], [16], [4])
        acts = torch.randn(N, 4, device="cuda") * 0.5

        def policy():
            with torch.inference_mode():
                alg.act(obs, critic)

        def envstep():
            env.step(acts)

        def both():
            with torch.inference_mode():
                a = alg.act(obs, critic)
                o, r, d, info = env.step(a)
                alg.storage.step = 0
                alg.process_env_step(r, d, info)

        def policy_nosync():      # same math without torch.normal's host-synchronising std check (graph-capturable)
            with torch.inference_mode():
                mean = pol.actor(obs)
                a = mean + pol.std * torch.randn_like(mean)
                v = pol.critic(critic)
                lp = (-((a - mean) ** 2) / (2 * pol.std ** 2) - pol.std.log() - 0.9189385332046727).sum(-1)
                return a, v, lp

        row = {"policy_act_eager_us": timeit(policy), "env_step_eager_us": timeit(envstep), "collect_step_eager_us": timeit(both)}
        row["policy_act_graph_us"] = timeit(graphed(policy_nosync))
        row["env_step_graph_us"] = timeit(graphed(envstep))
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            row[f"policy_act_graph_tf32_{tf32}_us"] = timeit(graphed(policy_nosync))
        torch.backends.cuda.matmul.allow_tf32 = False
        row["mlp_flop_per_step"] = N * 2 * ((16 * 128 + 128 * 128 + 128 * 4) + (16 * 128 + 128 * 128 + 128))
        out[str(N)] = row
        env.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

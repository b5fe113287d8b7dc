#!/usr/bin/env python
"""gr_rollout_fwd (RacingVecEnv.rollout): window timings (forward-only with / without recorded observations, with tape)."""
import dataclasses
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from generalizableracing_b200 import _lib as B  # noqa: E402
from generalizableracing_b200.config import RacingCfg  # noqa: E402
from generalizableracing_b200.env import RacingVecEnv  # noqa: E402
from generalizableracing_b200.track_gen import generate_track_table, racing_complex_cfg  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    table = generate_track_table(racing_complex_cfg())
    cfg = RacingCfg.for_stage(1)
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    only = sys.argv[1] if len(sys.argv) > 1 else None
    for name, N, T, diff, rec in (("fwd_65536x24_obs", 65536, 24, False, True), ("fwd_65536x24", 65536, 24, False, False),
                                  ("diff_16384x32", 16384, 32, True, False), ("diff_65536x32", 65536, 32, True, False)):
        if only and name != only:
            continue
        c = dataclasses.replace(cfg, is_differentiable_physics=diff)
        acts = torch.randn(T, N, 4, device=dev) * 0.5
        for variant, flag in (("wide", 0),):
            envs = [RacingVecEnv(c, table, N, device=dev, seed=3 + r, episode_stats=not diff, bptt_horizon=T if diff else 0) for r in range(3)]
            for e in envs:
                e._launch_flags |= flag
                e._state.launch_flags |= flag
                e.reset()
                e.episode_length_buf = torch.randint(0, cfg.max_episode_length, (N,), device=dev, dtype=torch.int32)
                if diff:
                    e._bptt.autograd = False

            def window(e):
                e.detach()
                e.rollout(acts, record_obs=rec)

            for e in envs:
                window(e)
            torch.cuda.synchronize()
            reps = 10
            e0.record()
            for _ in range(reps):
                for e in envs:
                    window(e)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / (reps * len(envs))
            out[f"{name}_{variant}"] = {"ms_per_window": ms, "us_per_step": ms * 1e3 / T, "env_steps_per_s": N * T / (ms * 1e-3)}
            for e in envs:
                e.close()
            del envs
            torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()

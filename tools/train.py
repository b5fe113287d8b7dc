#!/usr/bin/env python
"""Training driver mirroring standalone/rsl_rl/train.py and standalone/diff_rl/train.py of the reference for the state-only
racing task:  python tools/train.py ppo|bptt [--num_envs N] [--iters K] [--stage S] [--log_dir D]   (torchrun for multi-GPU)"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

PPO_CFG = {"num_steps_per_env": 24, "save_interval": 500, "empirical_normalization": False,          # QD/agents/rsl_rl_ppo_cfg.py:16-41
           "policy": {"class_name": "ActorCritic", "init_noise_std": 1.0, "actor_hidden_dims": [128, 128], "critic_hidden_dims": [128, 128], "activation": "lrelu"},
           "algorithm": {"class_name": "PPO", "value_loss_coef": 1.0, "use_clipped_value_loss": True, "clip_param": 0.2, "entropy_coef": 0.0,
                         "num_learning_epochs": 5, "num_mini_batches": 4, "learning_rate": 5.0e-4, "schedule": "adaptive", "gamma": 0.99, "lam": 0.95,
                         "desired_kl": 0.01, "max_grad_norm": 1.0}}
BPTT_CFG = {"num_steps_per_env": 48, "max_iterations": 2000, "save_interval": 200, "empirical_normalization": False,   # QD/agents/diff_rl_naive_cfg.py:9-32
            "algorithm": {"class_name": "BPTT", "schedule": "CosineAnnealingLR", "optimizer": "AdamW", "learning_rate": 5e-4},
            "policy": {"class_name": "BaseModel", "actor_hidden_dims": [256, 128], "critic_hidden_dims": [256, 128], "activation": "lrelu", "init_noise_std": 1.0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("algo", choices=["ppo", "bptt"])
    ap.add_argument("--num_envs", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--stage", type=int, default=None)
    ap.add_argument("--track", default="complex")
    ap.add_argument("--log_dir", default=None)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--noise_std", type=float, default=None)
    ap.add_argument("--fused_backward", action="store_true", help="BPTT with --fused: actor weight gradients from gr_actor_backward (tcgen05)")
    ap.add_argument("--graphed", action="store_true", help="PPO: replay the mini-batch update from a CUDA graph")
    ap.add_argument("--kernel_update", action="store_true", help="PPO: graphed update with forward / loss / weight gradients from libgracing kernels")
    ap.add_argument("--fused", action="store_true", help="PPO: collect each rollout with the fused kernel (gr_ppo_collect)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    from generalizableracing_b200 import make_env
    from generalizableracing_b200.runners import AlgoRunner, OnPolicyRunner
    torch.manual_seed(args.seed + int(os.environ.get("RANK", 0)))
    dev = f"cuda:{local}"
    if args.algo == "ppo":
        env = make_env(num_envs=args.num_envs, device=dev, stage=args.stage, track=args.track, seed=args.seed)
        cfg = json.loads(json.dumps(PPO_CFG))
        if args.noise_std is not None:
            cfg["policy"]["init_noise_std"] = args.noise_std
        cfg["fused_collection"] = bool(args.fused)
        cfg["algorithm"]["graphed_update"] = bool(args.graphed)
        cfg["algorithm"]["kernel_update"] = bool(args.kernel_update)
        runner = OnPolicyRunner(env, cfg, log_dir=args.log_dir, device=dev)
    else:
        cfg = json.loads(json.dumps(BPTT_CFG))
        cfg["max_iterations"] = args.iters
        if args.noise_std is not None:
            cfg["policy"]["init_noise_std"] = args.noise_std
        cfg["fused_collection"] = bool(args.fused)
        cfg["fused_backward_kernel"] = bool(args.fused_backward)
        env = make_env(num_envs=args.num_envs, device=dev, stage=args.stage, track=args.track, seed=args.seed, differentiable=True,
                       bptt_horizon=cfg["num_steps_per_env"])
        runner = AlgoRunner(env, cfg, log_dir=args.log_dir, device=dev)
    hist = runner.learn(args.iters, init_at_random_ep_len=True)
    if int(os.environ.get("RANK", 0)) == 0:
        keys = [k for k in ("Train/mean_reward", "Train/mean_episode_length", "Loss/mean_total_loss", "Train/mean_step_reward",
                            "Metrics/next_gate_pose/accumulate_gates", "Curriculum/terrain_levels", "Perf/total_fps", "Perf/collection time", "Perf/learning_time") if k in hist[-1]]
        for i in sorted(set(list(range(0, len(hist), max(1, len(hist) // 15))) + [len(hist) - 1])):
            print(i, {k: (round(hist[i][k], 4) if isinstance(hist[i][k], float) else hist[i][k]) for k in keys}, flush=True)
    if hasattr(runner.alg, "close"):
        runner.alg.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

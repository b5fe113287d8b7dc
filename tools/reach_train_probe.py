import sys, torch
sys.path.insert(0, "/root/repo")
from generalizableracing_b200.config import ReachTargetCfg
from generalizableracing_b200.reach_env import ReachTargetVecEnv
from generalizableracing_b200.runners import AlgoRunner
def run(name, cfg, std, iters=150, T=48, lr=5e-4):
    torch.manual_seed(0)
    env = ReachTargetVecEnv(cfg, 4096, bptt_horizon=T)
    run_cfg = {"num_steps_per_env": T, "max_iterations": iters, "save_interval": 100000, "empirical_normalization": False,
               "algorithm": {"class_name": "BPTT", "schedule": "CosineAnnealingLR", "optimizer": "AdamW", "learning_rate": lr},
               "policy": {"class_name": "BaseModel", "actor_hidden_dims": [256, 128], "critic_hidden_dims": [256, 128], "activation": "lrelu", "init_noise_std": std}}
    r = AlgoRunner(env, run_cfg, device="cuda:0")
    h = r.learn(iters, init_at_random_ep_len=True)
    L = [x["Loss/mean_total_loss"] for x in h]
    R = [x["Train/mean_step_reward"] for x in h]
    print(name, "loss", [round(sum(L[i:i+10])/10, 3) for i in range(0, iters, 30)], "rew", [round(sum(R[i:i+10])/10, 4) for i in range(0, iters, 30)], flush=True)
run("ctbr", ReachTargetCfg.ctbr(), 0.1)
run("ctbr_std1", ReachTargetCfg.ctbr(), 1.0)
run("lv_std.01", ReachTargetCfg.lv(decimation=1, episode_length_s=1.5), 0.01)
run("lv_std.1", ReachTargetCfg.lv(decimation=1, episode_length_s=1.5), 0.1)
run("ps_std.01", ReachTargetCfg.ps(decimation=1, episode_length_s=1.5), 0.01)
run("lv_dec2_std.01", ReachTargetCfg.lv(decimation=2, episode_length_s=3.0, rate_gain=(90.,90.,100.)), 0.01)

"""Reference-matching training curves: the SAME PPO (torch, on the GPU) trains once against the oracle env (CPU torch port of
the reference's env) and once against the CUDA env, with identical seeds, startup randomisation and per-step random
numbers.  The two closed loops differ only by the env's fp32 rounding (~1e-6 per step), which PPO amplifies slowly, so
the learning curves must coincide at first and stay close afterwards."""
import pytest
import torch

from generalizableracing_b200 import layout as L_
from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.tracks import synthetic_track_table
from tests import parity_cases as PC

pytestmark = pytest.mark.gpu

PPO_CFG = {"num_steps_per_env": 24, "save_interval": 10 ** 9, "empirical_normalization": False,
           "policy": {"class_name": "ActorCritic", "init_noise_std": 1.0, "actor_hidden_dims": [128, 128], "critic_hidden_dims": [128, 128], "activation": "lrelu"},
           "algorithm": {"class_name": "PPO", "value_loss_coef": 1.0, "use_clipped_value_loss": True, "clip_param": 0.2, "entropy_coef": 0.0,
                         "num_learning_epochs": 5, "num_mini_batches": 4, "learning_rate": 5.0e-4, "schedule": "adaptive", "gamma": 0.99, "lam": 0.95,
                         "desired_kl": 0.01, "max_grad_norm": 1.0}}


class _DenseCudaEnv:
    """RacingVecEnv in parity mode: every reset / step consumes the next pre-drawn random tensor."""

    def __init__(self, env, gen):
        self.env, self.gen = env, gen
        self.num_envs, self.num_actions, self.num_obs = env.num_envs, env.num_actions, env.num_obs
        self.device, self.cfg, self.max_episode_length = env.device, env.cfg, env.max_episode_length
        self.env.reset(PC.draw_rnd(self.num_envs, gen).cuda())

    episode_length_buf = property(lambda self: self.env.episode_length_buf, lambda self, v: setattr(self.env, "episode_length_buf", v))

    def get_observations(self):
        return self.env.get_observations()

    def step(self, actions):
        return self.env.step(actions, PC.draw_rnd(self.num_envs, self.gen).cuda())


class _OracleEnv:
    """The oracle behind the same surface (tensors cross to the GPU for the policy)."""

    def __init__(self, orc, gen):
        self.orc, self.gen = orc, gen
        self.num_envs, self.num_actions, self.num_obs = orc.num_envs, 4, 16
        self.device, self.cfg, self.max_episode_length = torch.device("cuda:0"), orc.cfg, orc.max_episode_length
        with torch.no_grad():
            self._obs, _ = orc.reset(PC.draw_rnd(self.num_envs, gen))

    @property
    def episode_length_buf(self):
        return self.orc.episode_length_buf.to("cuda:0", torch.int32)

    @episode_length_buf.setter
    def episode_length_buf(self, v):
        self.orc.episode_length_buf[:] = v.cpu().long()

    def _pack(self, obs):
        return {k: v.cuda() for k, v in obs.items()}

    def get_observations(self):
        o = self._pack(self._obs)
        return o["policy"], {"observations": o}

    def step(self, actions):
        with torch.no_grad():
            obs, rew, term, to, _ = self.orc.step(actions.detach().cpu(), PC.draw_rnd(self.num_envs, self.gen))
        self._obs = obs
        o = self._pack(obs)
        return o["policy"], rew.cuda(), (term | to).long().cuda(), {"observations": o, "time_outs": to.cuda()}


def _train(kind, iters, N=256, seed=5):
    from generalizableracing_b200.env import RacingVecEnv
    from generalizableracing_b200.runners import OnPolicyRunner
    from oracle import racing_oracle as RO
    cfg, table = RacingCfg.for_stage(1), synthetic_track_table()
    g = torch.Generator().manual_seed(seed)
    srnd = PC.draw_startup(N, g)
    if kind == "cuda":
        env = _DenseCudaEnv(RacingVecEnv(cfg, table, N, rng_mode="dense", startup_rnd=srnd), g)
    else:
        env = _OracleEnv(RO.OracleRacingEnv(cfg, table, N, srnd), g)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    runner = OnPolicyRunner(env, PPO_CFG, log_dir=None, device="cuda:0")
    return runner.learn(iters, init_at_random_ep_len=True)


def test_ppo_learning_curve_matches_the_oracle_env(cuda_lib):
    iters = 12
    ref, got = _train("oracle", iters), _train("cuda", iters)
    r = torch.tensor([[h["Train/mean_reward"], h["Train/mean_episode_length"], h["Loss/value_function"]] for h in ref])
    k = torch.tensor([[h["Train/mean_reward"], h["Train/mean_episode_length"], h["Loss/value_function"]] for h in got])
    rel = ((r - k).abs() / r.abs().clamp(min=1e-3))
    print("oracle:", r[:, 0].tolist())
    print("cuda  :", k[:, 0].tolist())
    print("rel   :", rel.max(dim=1).values.tolist())
    assert ref[0]["Train/episodes"] == got[0]["Train/episodes"]                # the same episodes end in the same steps of the first rollout
    assert float(rel[:1].max()) < 1e-5                                           # ... which coincides to fp32 round-off (same policy, same draws)
    # From the first update on, the 1e-7 differences are amplified -- by the tumbling drones' own dynamics within an episode and by the
    # 5 x 4 optimiser steps per iteration -- into different sample paths of the same learning process: which iteration sees the first
    # episode end one step apart depends on the last bit of the kernel build (measured on two builds: 4e-7, 7e-4, 1e-2, ... and 5e-3,
    # 4e-3, 2e-2, ...).  Compare the curves as curves: 256 envs => a few % of sampling noise per point early on, and the same
    # dip-and-recover shape (-5 -> -17 -> about -6) afterwards.  tests/test_training_curves.py does this comparison over seeds.
    assert float(rel[:5, :2].max()) < 0.05
    for curve in (r[:, 0], k[:, 0]):
        assert float(curve[2:7].min()) < -15.0 and float(curve[-1]) > -8.0 and int(curve.argmin()) in (2, 3, 4, 5)
    assert abs(float(r[:, 0].min() - k[:, 0].min())) < 0.1 * abs(float(r[:, 0].min()))

"""BASELINE.json configs[0] -- 64 envs, fixed figure-8 gate track, STAGE 0, 1000-step rollout -- recorded from the reference's OWN env
step (tests/golden/make_c1_golden.py; unmodified ManagerBasedDiffRLEnv.step + MDP term modules over the closure simulator) and replayed
where the reference tree is absent: bit-exact through the oracle (CPU), and through the kernels (``emul`` here, ``cuda`` = libgracing.so
through the C ABI on the B200) with masks / gate ids bit-exact and fp32 columns within 10x the per-step 1e-5 (free-running episodes).
Inputs are regenerated from the recorded seeds: actions from a seeded generator, the reference's random draws by repeating its
global-generator calls (oracle/ref_closure.py::replay_*; same torch build on the GPU box)."""
import importlib.util
import os

import pytest
import torch

from generalizableracing_b200 import layout as L_
from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.env import RacingVecEnv
from generalizableracing_b200.tracks import figure_eight_track
from oracle import racing_oracle as RO
from oracle import ref_closure as RC
from tests import parity_cases as PC
from tests.conftest import backend_params

HERE = os.path.dirname(__file__)


def _setup():
    d = torch.load(os.path.join(HERE, "golden", "ref_c1_rollout.pt"))
    spec = importlib.util.spec_from_file_location("_make_c1_golden", os.path.join(HERE, "golden", "make_c1_golden.py"))
    tools = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tools)
    g, _ = tools.inputs(d["seed"], d["N"])
    return d, RacingCfg.for_stage(0), g


def _rnd(d, t, reset_ids):
    """The rnd[N, 52] rows of step t (t = -1: the initial reset): the reference's draws after torch.manual_seed(seed + 1 + t)."""
    rnd = torch.zeros(d["N"], L_.RND_STRIDE)
    torch.manual_seed(d["seed"] + 1 + t)
    RC.replay_reset_draws(rnd, reset_ids, add_noise=False)
    RC.replay_obs_draws(rnd)
    return rnd


def _reset_ids(d, t):
    return (d["terminated"][t] | d["time_out"][t]).nonzero(as_tuple=False).squeeze(-1)


def test_oracle_replays_reference_c1_rollout_bit_exact():
    d, cfg, g = _setup()
    N, T = d["N"], d["T"]
    orc = RO.OracleRacingEnv(cfg, figure_eight_track(), N, d["startup_rnd"])
    obs, _ = orc.reset(_rnd(d, -1, torch.arange(N)))
    assert torch.equal(obs["policy"], d["policy0"])
    with torch.no_grad():
        for t in range(T):
            a = torch.randn(N, 4, generator=g) * 0.5
            obs, rew, term, to, _ = orc.step(a, _rnd(d, t, _reset_ids(d, t)))
            assert torch.equal(term, d["terminated"][t]) and torch.equal(to, d["time_out"][t]), t
            assert torch.equal(rew, d["reward"][t]) and torch.equal(obs["policy"].sum(-1), d["policy_sum"][t]), t
            assert torch.equal(obs["critic"].sum(-1), d["critic_sum"][t]) and torch.equal(orc.root_pos_w.sum(-1), d["pos_sum"][t]), t
            assert torch.equal(orc.gate_id.to(torch.int8), d["gate_id"][t]), t
    assert torch.equal(orc._root_state_w(), d["final_root_state"]) and torch.equal(obs["policy"], d["final_obs"])
    assert torch.equal(orc.episode_length_buf, d["final_episode_length"])
    assert int((d["terminated"] | d["time_out"]).sum()) > 4 * N


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_kernels_replay_reference_c1_rollout(backend):
    device, lib = backend
    d, cfg, g = _setup()
    N, T = d["N"], d["T"]
    env = RacingVecEnv(cfg, figure_eight_track(), N, device=device, rng_mode="dense", startup_rnd=d["startup_rnd"], _lib=lib)
    obs, _ = env.reset(_rnd(d, -1, torch.arange(N)).to(device))
    assert PC.rel_err(d["policy0"], obs) < PC.REL_TOL_STEP
    worst = dict(reward=0.0, policy_sum=0.0, critic_sum=0.0, pos_sum=0.0)
    for t in range(T):
        a = torch.randn(N, 4, generator=g) * 0.5
        obs, rew, dones, ex = env.step(a.to(device), _rnd(d, t, _reset_ids(d, t)).to(device))
        assert torch.equal(ex["terminated"].cpu().bool(), d["terminated"][t]) and torch.equal(ex["time_outs"].cpu().bool(), d["time_out"][t]), t
        sv = env.state_dict_view()
        assert torch.equal(sv["gate_id"].cpu().to(torch.int8), d["gate_id"][t]), t
        worst["reward"] = max(worst["reward"], PC.rel_err(d["reward"][t], rew))
        worst["policy_sum"] = max(worst["policy_sum"], PC.rel_err(d["policy_sum"][t], obs.sum(-1)))
        worst["critic_sum"] = max(worst["critic_sum"], PC.rel_err(d["critic_sum"][t], ex["observations"]["critic"].sum(-1)))
        worst["pos_sum"] = max(worst["pos_sum"], PC.rel_err(d["pos_sum"][t], sv["root_pos_w"].sum(-1)))
    print("C1 vs the reference's own rollout:", worst)
    for k, v in worst.items():
        assert v < 10 * PC.REL_TOL_STEP, (k, v)
    assert PC.rel_err(d["final_obs"], obs) < 10 * PC.REL_TOL_STEP
    assert torch.equal(sv["episode_length"].cpu().long(), d["final_episode_length"])

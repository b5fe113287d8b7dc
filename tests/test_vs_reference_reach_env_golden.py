"""Replays tests/golden/reach_env_closure.pt -- inputs and outputs of the reference's OWN env step for the CTBR reach-target task
(normalised and sim2real inputs), recorded by tests/golden/make_reach_env_golden.py from the unmodified reference modules -- where
the reference tree is absent: bit-exact through the oracle (CPU); through the kernels (``emul`` here, ``cuda`` = libgracing.so through
the C ABI on the B200) masks and counters bit-exact, fp32 columns within the tolerances of tests/reach_cases.py."""
import os

import pytest
import torch

from generalizableracing_b200.config import ReachTargetCfg
from generalizableracing_b200.reach_env import ReachTargetVecEnv
from oracle.reach_oracle import OracleReachEnv
from tests.conftest import backend_params
from tests.reach_cases import _close

G = os.path.join(os.path.dirname(__file__), "golden", "reach_env_closure.pt")
SCENARIOS = ("ctbr", "ctbr_sim2real")


def _load(name):
    d = torch.load(G)[name]
    return d, ReachTargetCfg.ctbr(sim2real_test=d["sim2real"], **d["cfg_kw"])


@pytest.mark.parametrize("name", SCENARIOS)
def test_oracle_replays_reference_reach_env_golden_bit_exact(name):
    d, cfg = _load(name)
    N, T = d["N"], d["T"]
    orc = OracleReachEnv(cfg, N)
    obs, _ = orc.reset(d["rnd0"])
    assert torch.equal(obs["policy"], d["policy0"])
    with torch.no_grad():
        for t in range(T):
            obs, rew, term, to, ex = orc.step(d["actions"][t], d["rnd"][t])
            root = torch.hstack([orc.root_pos_w, orc.root_quat_w, orc.root_lin_vel_w, orc.root_ang_vel_w])
            for k, v in (("policy", obs["policy"]), ("reward", rew), ("reward_terms", orc.step_reward), ("terminated", term), ("time_out", to),
                         ("losses", ex["losses"]), ("loss_terms", ex["loss_terms"]), ("root_state", root), ("pose_command_w", orc.pose_command_w),
                         ("time_left", orc.time_left), ("episode_length", orc.episode_length_buf)):
                assert torch.equal(v, d[k][t]), (name, t, k)
    assert int((d["terminated"] | d["time_out"]).sum()) >= 2 * N


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
@pytest.mark.parametrize("name", SCENARIOS)
def test_reach_kernels_replay_reference_env_golden(backend, name):
    device, lib = backend
    tol = 2e-4 if lib is not None else 5e-4          # tests/reach_cases.py: the sm_100a build uses the fast division / square root
    d, cfg = _load(name)
    N, T = d["N"], d["T"]
    env = ReachTargetVecEnv(cfg, N, device=device, rng_mode="dense", bptt_horizon=17, _lib=lib)
    env.export_reward_terms = True
    obs, _ = env.reset(d["rnd0"])
    _close(obs, d["policy0"], tol, "reset obs")
    for t in range(T):
        obs, rew, dones, ex = env.step(d["actions"][t].to(device), d["rnd"][t])
        assert torch.equal(ex["terminated"].cpu().bool(), d["terminated"][t]) and torch.equal(ex["time_outs"].cpu().bool(), d["time_out"][t]), (name, t)
        assert torch.equal(dones.cpu() != 0, d["terminated"][t] | d["time_out"][t]), (name, t)
        _close(obs, d["policy"][t], tol, f"obs step {t}")
        _close(rew, d["reward"][t], tol, f"reward step {t}")
        _close(env._outs[env._flip ^ 1]["reward_terms"], d["reward_terms"][t], tol, f"reward terms step {t}")
        _close(ex["losses"], d["losses"][t], tol, f"loss step {t}")
        _close(ex["loss_terms"], d["loss_terms"][t], tol, f"loss terms step {t}")
        if t % 16 == 15:
            env.detach()
    sv = env.state_dict_view()
    _close(sv["root_pos_w"], d["root_state"][-1][:, :3], tol, "pos")
    _close(sv["root_quat_w"], d["root_state"][-1][:, 3:7], tol, "quat")
    _close(sv["pose_command_w"], d["pose_command_w"][-1][:, :3], tol, "target")
    _close(sv["time_left"], d["time_left"][-1], 1e-5, "time_left")
    assert torch.equal(sv["episode_length"].cpu().long(), d["episode_length"][-1])

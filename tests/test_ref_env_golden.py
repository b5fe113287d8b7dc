"""Replays tests/golden/ref_env_closure.pt -- inputs and outputs of the reference's OWN env step, recorded by
tests/golden/make_ref_env_golden.py from the unmodified reference modules -- where the reference tree is absent:
 * through the oracle (CPU): bit-exact, so the pin of tests/test_oracle_vs_reference_env.py travels with the repo;
 * through the kernels (``emul`` here, ``cuda`` = libgracing.so through the C ABI on the B200): ids / counters / masks
   bit-exact, fp32 columns within the north star's 1e-5 relative per step (free-running, bounded at 1e-4)."""
import os

import pytest
import torch

from generalizableracing_b200 import layout as L_
from generalizableracing_b200.config import RacingCfg
from generalizableracing_b200.env import RacingVecEnv
from generalizableracing_b200.tracks import figure_eight_track, synthetic_track_table
from oracle import isaac_math as M
from oracle import racing_oracle as RO
from tests import parity_cases as PC
from tests.conftest import backend_params

G = os.path.join(os.path.dirname(__file__), "golden", "ref_env_closure.pt")
SCENARIOS = ("stage0_figure8", "stage1_table")


def _load(name):
    d = torch.load(G)[name]
    cfg = RacingCfg.for_stage(d["stage"], is_differentiable_physics=True)
    table = figure_eight_track() if d["stage"] == 0 else synthetic_track_table()
    return d, cfg, table


@pytest.mark.parametrize("name", SCENARIOS)
def test_oracle_replays_reference_env_golden_bit_exact(name):
    d, cfg, table = _load(name)
    N, T = d["N"], d["T"]
    orc = RO.OracleRacingEnv(cfg, table, N, d["startup_rnd"])
    obs, _ = orc.reset(d["rnd0"])
    assert torch.equal(obs["policy"], d["policy0"]) and torch.equal(obs["critic"], d["critic0"])
    orc.episode_length_buf[:] = d["episode_length0"]
    with torch.no_grad():
        for t in range(T):
            if t in d["teleports"]:
                orc.root_pos_w = d["teleports"][t].clone()
                orc._get_state_from_sim()
                orc.dyn.reset_state(orc.states_all, torch.arange(N))
            obs, rew, term, to, ex = orc.step(d["actions"][t], d["rnd"][t])
            for k, v in (("policy", obs["policy"]), ("critic", obs["critic"]), ("aux", obs["auxiliary"]), ("reward", rew), ("terminated", term),
                         ("time_out", to), ("achieved", orc.last_achieved), ("losses", ex["losses"]), ("root_state", orc._root_state_w()),
                         ("gate_id", orc.gate_id), ("accumulate_gates", orc.accumulate_gates), ("terrain_levels", orc.terrain_levels),
                         ("episode_length", orc.episode_length_buf), ("gross_thrust", orc.ctrl.gross_thrust[:, 0]), ("torque", orc.ctrl.torque)):
                assert torch.equal(v, d[k][t]), (name, t, k)
    assert int((d["terminated"] | d["time_out"]).sum()) >= N and int(d["achieved"].sum()) > 50


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
@pytest.mark.parametrize("name", SCENARIOS)
def test_kernels_replay_reference_env_golden(backend, name):
    device, lib = backend
    d, cfg, table = _load(name)
    N, T = d["N"], d["T"]
    env = RacingVecEnv(cfg, table, N, device=device, rng_mode="dense", startup_rnd=d["startup_rnd"], bptt_horizon=T, _lib=lib)
    env.export_gate_passed = True
    obs, ex = env.reset(d["rnd0"].to(device))
    assert PC.rel_err(d["policy0"], obs) < PC.REL_TOL_STEP and PC.rel_err(d["critic0"], ex["observations"]["critic"]) < PC.REL_TOL_STEP
    env.episode_length_buf = d["episode_length0"]
    worst = dict(policy=0.0, critic=0.0, reward=0.0, losses=0.0, state=0.0, filters=0.0)
    for t in range(T):
        if t in d["teleports"]:
            env.write_plane(L_.PL_POS, slice(0, 3), d["teleports"][t])
        obs, rew, dones, ex = env.step(d["actions"][t].to(device), d["rnd"][t].to(device))
        where = (name, t)
        # masks, ids and counters: bit-exact
        assert torch.equal(ex["terminated"].cpu().bool(), d["terminated"][t]) and torch.equal(ex["time_outs"].cpu().bool(), d["time_out"][t]), where
        assert torch.equal(dones.cpu(), (d["terminated"][t] | d["time_out"][t]).long()), where
        assert torch.equal(env._last["gate_passed"].cpu().bool(), d["achieved"][t]), where
        assert torch.equal(ex["observations"]["auxiliary"].cpu(), d["aux"][t]), where
        sv = env.state_dict_view()
        assert torch.equal(sv["gate_id"].cpu().long(), d["gate_id"][t]), where
        assert torch.equal(sv["accumulate_gates"].cpu().float(), d["accumulate_gates"][t]), where
        assert torch.equal(sv["terrain_levels"].cpu().long(), d["terrain_levels"][t]), where
        assert torch.equal(sv["episode_length"].cpu().long(), d["episode_length"][t]), where
        # fp32 columns
        worst["policy"] = max(worst["policy"], PC.rel_err(d["policy"][t], obs))
        worst["critic"] = max(worst["critic"], PC.rel_err(d["critic"][t], ex["observations"]["critic"]))
        worst["reward"] = max(worst["reward"], PC.rel_err(d["reward"][t], rew))
        worst["losses"] = max(worst["losses"], PC.rel_err(d["losses"][t], ex["losses"]))
        quat = sv["root_quat_w"].cpu()
        root = torch.cat([sv["root_pos_w"].cpu(), quat, sv["root_lin_vel_w"].cpu(), M.quat_rotate(quat, sv["root_ang_vel_b"].cpu())], dim=-1)
        worst["state"] = max(worst["state"], PC.rel_err(d["root_state"][t], root))
        worst["filters"] = max(worst["filters"], PC.rel_err(d["gross_thrust"][t], sv["gross_thrust"]), PC.rel_err(d["torque"][t], sv["torque"]))
    print(name, worst)
    for k, v in worst.items():
        assert v < 10 * PC.REL_TOL_STEP, (k, v)

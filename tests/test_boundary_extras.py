"""The env surface beyond (obs, reward, dones): what VERDICT r1 Missing #1 / #3 listed.

* the loop body of the reference's minimal BPTT trainer (standalone/diff_rl/naive_train.py:161-193) runs on RacingVecEnv as written --
  ``extras["losses"]``, ``extras["log_losses"]`` (lazily resolved: 0-dim device tensors), ``env.unwrapped.detach()`` -- and moves the
  policy exactly as the same loop on the oracle env does;
* ``extras["aligned_states"]`` / ``["nominal_states"]`` / ``["acc"]`` (L/envs/manager_based_diff_rl_env.py:205-212), opt-in export;
* ``get_observations(fresh_noise=True)``: the reference's ObservationManager.compute() recomputation with a new noise draw."""
import pytest
import torch

from tests import parity_cases as PC
from tests.conftest import backend_params

pytestmark = pytest.mark.timeout(900)


class _OracleAsWrapped:
    """RslRlVecEnvWrapper view of the oracle env (Isaac Lab: dones = terminated | truncated, obs = obs_dict["policy"])."""

    def __init__(self, orc, draws):
        self.orc, self.draws, self.unwrapped = orc, iter(draws), orc

    def step(self, actions):
        obs, rew, term, to, ex = self.orc.step(actions, next(self.draws))
        return obs["policy"], rew, (term | to).long(), ex


class _KernelWithDraws:
    def __init__(self, env, draws):
        self.env, self.draws, self.unwrapped = env, iter(draws), env

    def step(self, actions):
        return self.env.step(actions, next(self.draws).to(self.env.device))


def _naive_train_iterations(env, model, obs, iters, T):
    """standalone/diff_rl/naive_train.py:161-193, verbatim control flow (tqdm / writer / checkpoint lines dropped)"""
    optim = torch.optim.AdamW(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(optim, 10, 1e-3 * 0.01)
    logs = []
    for i in range(iters):
        dones_history = []
        loss_history = []
        log_loss_history = {}
        # detach env
        env.unwrapped.detach()
        # start rollout
        for t in range(T):
            actions = model(obs)
            obs, rews, dones, extras = env.step(actions)
            dones_history.append(dones)
            loss_history.append(extras["losses"])
            for item in extras["log_losses"]:
                if item[0] not in log_loss_history:
                    log_loss_history[item[0]] = []
                log_loss_history[item[0]].append(item[1])
        loss_history = torch.stack(loss_history)
        for k, v in log_loss_history.items():
            log_loss_history[k] = sum(v) / len(v)
        loss = loss_history.mean()
        optim.zero_grad()
        loss.backward()
        optim.step()
        sched.step()
        logs.append({"loss": loss.cpu().item(), **{k: float(v) for k, v in log_loss_history.items()}, "dones": int(torch.stack(dones_history).sum())})
    return logs


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_naive_train_loop_runs_unmodified_and_tracks_the_oracle(backend):
    N, T, iters = 96, 8, 3
    cfg, table, orc, env, g = PC.make_pair(backend, stage=1, N=N, seed=12, diff=True, horizon=T)
    dev = env.device
    r0 = PC.draw_rnd(N, g)
    o_obs, _ = orc.reset(r0)
    k_obs, _ = env.reset(r0.to(dev))
    ep = torch.randint(cfg.max_episode_length - 15, cfg.max_episode_length, (N,), generator=g)       # time-outs inside every window
    orc.episode_length_buf[:] = ep
    env.episode_length_buf = ep
    draws = [PC.draw_rnd(N, g) for _ in range(T * iters)]
    torch.manual_seed(0)
    m_o = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    m_k = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4)).to(dev)
    m_k.load_state_dict(m_o.state_dict())
    logs_o = _naive_train_iterations(_OracleAsWrapped(orc, draws), m_o, o_obs["policy"], iters, T)
    logs_k = _naive_train_iterations(_KernelWithDraws(env, draws), m_k, k_obs, iters, T)
    assert sum(l["dones"] for l in logs_o) > N // 2
    for lo, lk in zip(logs_o, logs_k):
        assert lo["dones"] == lk["dones"]
        for key in lo:
            assert abs(lo[key] - lk[key]) < 1e-4 * max(1.0, abs(lo[key])), (key, lo[key], lk[key])
    for po, pk in zip(m_o.parameters(), m_k.parameters()):
        assert float((po.detach() - pk.detach().cpu()).abs().max()) < 2e-4
    # log_losses entries are resolved lazily: device scalars, no host read inside the loop
    name, value = env.extras["log_losses"][0]
    assert name == "move_towards_goal" and torch.is_tensor(value) and value.dim() == 0 and value.device.type == dev.type


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
@pytest.mark.parametrize("stage", [0, 1])
def test_aligned_states_export_matches_the_oracle(backend, stage):
    N, T = 64, 12
    cfg, table, orc, env, g = PC.make_pair(backend, stage=stage, N=N, seed=3, diff=True, horizon=T)
    dev = env.device
    env.export_aligned_states = True
    r0 = PC.draw_rnd(N, g)
    orc.reset(r0)
    env.reset(r0.to(dev))
    ep = torch.randint(cfg.max_episode_length - 8, cfg.max_episode_length, (N,), generator=g)
    orc.episode_length_buf[:] = ep
    env.episode_length_buf = ep
    orc.detach()
    env.detach()
    resets = 0
    for t in range(T):
        a, r = torch.randn(N, 4, generator=g) * 0.5, PC.draw_rnd(N, g)
        with torch.no_grad():
            _, _, term, to, oex = orc.step(a, r)
        _, _, _, kex = env.step(a.to(dev), r.to(dev))
        resets += int((term | to).sum())
        # the aligned state of the step BEFORE the reset, for the envs that reset as well
        for key in ("aligned_states", "nominal_states", "acc"):
            assert PC.rel_err(oex[key], kex[key]) < 10 * PC.REL_TOL_STEP, (t, key)          # free-running rollout: the bound of test_env_parity.py
        assert kex["aligned_states"].shape == (N, 13) and kex["acc"].shape == (N, 3)
    assert resets > N // 2


@pytest.mark.parametrize("backend", backend_params(), indirect=True)
def test_get_observations_with_a_fresh_noise_draw(backend):
    """ObservationManager.compute() (manager_based_diff_rl_env.py:264): recomputed from the current state with NEW noise; the last-action
    columns keep showing the lagged action the last step applied (QD/mdp/observation.py:55-63), not the action still waiting in the FIFO."""
    N = 64
    cfg, table, orc, env, g = PC.make_pair(backend, stage=1, N=N, seed=4)
    dev = env.device
    r0 = PC.draw_rnd(N, g)
    orc.reset(r0)
    env.reset(r0.to(dev))
    for t in range(5):
        a, r = torch.randn(N, 4, generator=g) * 0.8, PC.draw_rnd(N, g)
        with torch.no_grad():
            orc.step(a, r)
        step_obs = env.step(a.to(dev), r.to(dev))[0].clone()
    again, ex = env.get_observations()
    assert torch.equal(again, step_obs)                                           # default: the last observation again
    r = PC.draw_rnd(N, g)
    fresh, ex = env.get_observations(fresh_noise=True, rnd=r.to(dev))
    ref = orc.compute_observations(r)
    assert PC.rel_err(ref["policy"], fresh) < PC.REL_TOL_STEP
    assert PC.rel_err(ref["critic"], ex["observations"]["critic"]) < PC.REL_TOL_STEP
    assert torch.equal(ref["auxiliary"], ex["observations"]["auxiliary"].cpu())
    assert not torch.equal(fresh[:, :6], step_obs[:, :6]) and torch.equal(fresh[:, 12:], step_obs[:, 12:])
    # stepping continues from the same state (the recomputation touched no state; dense mode draws nothing by itself)
    a, r = torch.randn(N, 4, generator=g) * 0.8, PC.draw_rnd(N, g)
    with torch.no_grad():
        oo = orc.step(a, r)[0]
    ko = env.step(a.to(dev), r.to(dev))[0]
    assert PC.rel_err(oo["policy"], ko) < PC.REL_TOL_STEP
    if backend[0] != "cpu":                                                       # in-kernel Philox: a stream no step uses
        from generalizableracing_b200.env import RacingVecEnv
        e1, e2 = (RacingVecEnv(cfg, table, N, seed=5) for _ in range(2))
        for e in (e1, e2):
            e.reset()
        act = torch.randn(N, 4, device=dev)
        e1.step(act); e2.step(act)
        f1 = e1.get_observations(fresh_noise=True)[0].clone()
        f2 = e1.get_observations(fresh_noise=True)[0].clone()
        assert not torch.equal(f1[:, :3], f2[:, :3])                               # every call draws anew
        assert torch.equal(e1.step(act)[0], e2.step(act)[0])                       # ... and the step stream did not move

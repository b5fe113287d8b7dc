"""world_size-2 `gloo` tests of the env-sharded data-parallel host logic: shard ranges, global advantage
moments, policy-gradient all-reduce, and partition invariance of the env shards (Philox keyed by the GLOBAL env id)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.timeout(600)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, world=2):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    res = [q.get() for _ in range(world)]
    for r in res:
        assert r[1] == "ok", r
    return res


def _entry(fn, rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    try:
        torch.set_num_threads(2)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        globals()[fn](rank, world)
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "fail: " + traceback.format_exc()))


def _w_moments(rank, world):
    from generalizableracing_b200 import dist_utils as D
    assert D.world() == (rank, world)
    assert D.shard_range(131072, rank, world) == (rank * 65536, 65536)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(24 * 512, generator=g, dtype=torch.float64) * (1 + rank) + 3 * rank
    m = torch.stack([torch.tensor(float(x.numel()), dtype=torch.float64), x.mean(), ((x - x.mean()) ** 2).sum()])
    merged = D.merge_moments(m)
    allx = [torch.zeros_like(x) for _ in range(world)]
    dist.all_gather(allx, x)
    allx = torch.cat(allx)
    assert abs(float(merged[0]) - allx.numel()) < 1e-9
    assert abs(float(merged[1]) - float(allx.mean())) < 1e-12
    assert abs(float(merged[2]) - float(((allx - allx.mean()) ** 2).sum())) < 1e-7
    # the normalisation every rank applies is the global one (rollout_storage.py:127 on the global batch)
    std = torch.sqrt(merged[2] / (merged[0] - 1))
    assert abs(float(std) - float(allx.std())) < 1e-12


def _w_grads(rank, world):
    from generalizableracing_b200 import dist_utils as D
    from generalizableracing_b200.modules import ActorCritic
    torch.manual_seed(rank)                       # different init per rank: broadcast must fix it
    net = ActorCritic(16, 16, 4)
    D.broadcast_module(net)
    g = torch.Generator().manual_seed(7)
    obs = torch.randn(world * 64, 16, generator=g)
    tgt = torch.randn(world * 64, 4, generator=g)
    sl = slice(rank * 64, (rank + 1) * 64)
    ((net.actor(obs[sl]) - tgt[sl]) ** 2).mean().backward()
    D.allreduce_mean_grads(net.parameters())
    ref = ActorCritic(16, 16, 4)
    ref.load_state_dict(net.state_dict())
    ((ref.actor(obs) - tgt) ** 2).mean().backward()
    for (n, p), (_, r) in zip(net.named_parameters(), ref.named_parameters()):
        if r.grad is not None:
            assert torch.allclose(p.grad, r.grad, atol=1e-6), n


def _w_env_shards(rank, world):
    """two ranks x 96 envs == one env with 192 envs: same global ids -> same tracks, same Philox draws, same results."""
    from generalizableracing_b200.config import RacingCfg
    from generalizableracing_b200.env import RacingVecEnv
    from generalizableracing_b200.tracks import synthetic_track_table
    from tests.emul import EmulLib
    lib = EmulLib()
    cfg, table, n = RacingCfg.for_stage(1), synthetic_track_table(), 96
    shard = RacingVecEnv(cfg, table, n, device="cpu", seed=5, env_id_offset=rank * n, global_num_envs=world * n, _lib=lib)
    obs = [shard.reset()[0].clone()]
    g = torch.Generator().manual_seed(3)
    acts = torch.randn(12, world * n, 4, generator=g) * 0.5
    rews = []
    for t in range(12):
        o, r, d, ex = shard.step(acts[t, rank * n:(rank + 1) * n].contiguous())
        obs.append(o.clone())
        rews.append(r.clone())
    mine = torch.cat([torch.stack(obs).flatten(1), torch.stack(rews)], dim=0 if False else 1) if False else torch.stack(obs)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        full = RacingVecEnv(cfg, table, world * n, device="cpu", seed=5, _lib=lib)
        fobs = [full.reset()[0].clone()]
        for t in range(12):
            fobs.append(full.step(acts[t])[0].clone())
        fobs = torch.stack(fobs)
        assert torch.equal(torch.cat(gathered, dim=1), fobs)


def _w_reach_shards(rank, world):
    """reach-target task: two ranks x 64 envs == one env with 128 envs (Philox keyed by the global env id)."""
    from generalizableracing_b200.config import ReachTargetCfg
    from generalizableracing_b200.reach_env import ReachTargetVecEnv
    from tests.emul import EmulLib
    lib = EmulLib()
    cfg, n = ReachTargetCfg.lv(decimation=1, episode_length_s=0.04, is_differentiable_physics=False), 64
    shard = ReachTargetVecEnv(cfg, n, device="cpu", seed=5, env_id_offset=rank * n, _lib=lib)
    obs = [shard.reset()[0].clone()]
    g = torch.Generator().manual_seed(3)
    acts = torch.randn(20, world * n, 4, generator=g) * 0.5
    for t in range(20):
        obs.append(shard.step(acts[t, rank * n:(rank + 1) * n].contiguous())[0].clone())
    mine = torch.stack(obs)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        full = ReachTargetVecEnv(cfg, world * n, device="cpu", seed=5, _lib=lib)
        fobs = [full.reset()[0].clone()]
        for t in range(20):
            fobs.append(full.step(acts[t])[0].clone())
        assert torch.equal(torch.cat(gathered, dim=1), torch.stack(fobs))


def _w_storage_shards(rank, world):
    """Env-sharded PPO data path: every rank's RolloutStorage (storage kernels through the g++ emulation) runs GAE on its envs
    with ``normalize=False``, the (count, mean, M2) moments are merged over the ranks and each shard normalises with the GLOBAL mean / std
    (``PPO.compute_returns``) -- the concatenated shards must equal one storage holding all envs."""
    from generalizableracing_b200 import dist_utils as D
    from generalizableracing_b200.storage import RolloutStorage
    from tests.emul import EmulLib
    lib = EmulLib()
    T, n = 12, 48
    N = world * n
    g = torch.Generator().manual_seed(7)                      # same stream on every rank: the global batch
    data = dict(obs=torch.randn(T, N, 16, generator=g), act=torch.randn(T, N, 4, generator=g), rew=torch.randn(T, N, generator=g),
                val=torch.randn(T, N, 1, generator=g), done=torch.rand(T, N, generator=g) < 0.05, to=torch.rand(T, N, generator=g) < 0.02,
                logp=torch.randn(T, N, generator=g), mu=torch.randn(T, N, 4, generator=g), sig=torch.rand(T, N, 4, generator=g),
                last=torch.randn(N, 1, generator=g))

    def fill(lo, hi):
        sto = RolloutStorage("rl", hi - lo, T, [16], [16], [4], device="cpu", _lib=lib)
        for t in range(T):
            tr = sto.Transition()
            tr.observations = tr.privileged_observations = data["obs"][t, lo:hi].contiguous()
            tr.actions, tr.rewards, tr.values = data["act"][t, lo:hi].contiguous(), data["rew"][t, lo:hi].contiguous(), data["val"][t, lo:hi].contiguous()
            tr.dones, tr.time_outs, tr.gamma = data["done"][t, lo:hi].long(), (data["to"][t, lo:hi] & data["done"][t, lo:hi]).contiguous(), 0.99
            tr.actions_log_prob, tr.action_mean, tr.action_sigma = data["logp"][t, lo:hi].contiguous(), data["mu"][t, lo:hi].contiguous(), data["sig"][t, lo:hi].contiguous()
            sto.add_transitions(tr)
        return sto
    lo, cnt = D.shard_range(N, rank, world)
    shard = fill(lo, lo + cnt)
    shard.compute_returns(data["last"][lo:lo + cnt].contiguous(), 0.99, 0.95, normalize=False)
    shard.normalize_advantages(D.merge_moments(shard.moments))
    parts = [torch.zeros_like(shard.advantages) for _ in range(world)]
    dist.all_gather(parts, shard.advantages.contiguous())
    rets = [torch.zeros_like(shard.returns) for _ in range(world)]
    dist.all_gather(rets, shard.returns.contiguous())
    if rank == 0:
        full = fill(0, N)
        full.compute_returns(data["last"], 0.99, 0.95)
        assert torch.equal(torch.cat(rets, dim=1), full.returns)
        err = float((torch.cat(parts, dim=1) - full.advantages).abs().max())
        assert err < 1e-6, err
        a = torch.cat(parts, dim=1).double()
        assert abs(float(a.mean())) < 1e-6 and abs(float(a.std()) - 1.0) < 1e-6


def _w_ppo_update(rank, world):
    """One env-sharded PPO iteration (this repo's PPO: storage kernels through the emulation, global advantage moments, KL and policy-gradient
    all-reduce over gloo) equals the same iteration in one process holding all envs: identical learning-rate decisions, weights to fp32 round-off."""
    import copy
    from generalizableracing_b200 import dist_utils as D
    from generalizableracing_b200.algorithms.ppo import PPO
    from generalizableracing_b200.modules import ActorCritic
    from generalizableracing_b200.storage import RolloutStorage
    from tests.emul import EmulLib
    lib = EmulLib()
    T, n = 8, 32
    N = world * n
    hp = dict(num_learning_epochs=3, num_mini_batches=1, clip_param=0.2, gamma=0.99, lam=0.95, value_loss_coef=1.0, entropy_coef=0.01,
              learning_rate=5e-4, max_grad_norm=1.0, use_clipped_value_loss=True, schedule="adaptive", desired_kl=0.01)
    torch.manual_seed(0)
    policy0 = ActorCritic(16, 16, 4, actor_hidden_dims=[32, 32], critic_hidden_dims=[32, 32], activation="elu", init_noise_std=1.0)
    g = torch.Generator().manual_seed(11)
    obs, rew = torch.randn(T, N, 16, generator=g), torch.randn(T, N, generator=g)
    done, last_obs = (torch.rand(T, N, generator=g) < 0.05).long(), torch.randn(N, 16, generator=g)
    noise = torch.randn(T, N, 4, generator=g)

    def iterate(lo, hi, sharded):
        pol = copy.deepcopy(policy0)
        if not sharded:                       # the single-process reference run must not talk to the other rank
            real, D.world = D.world, lambda: (0, 1)
        try:
            alg = PPO(pol, None, device="cpu", **hp)
            alg.storage = RolloutStorage("rl", hi - lo, T, [16], [16], [4], device="cpu", _lib=lib)
            with torch.no_grad():
                for t in range(T):
                    o = obs[t, lo:hi].contiguous()
                    pol.update_distribution(o)
                    tr = alg.transition
                    tr.actions = (pol.action_mean + pol.action_std * noise[t, lo:hi]).contiguous()
                    tr.values, tr.actions_log_prob = pol.evaluate(o), pol.get_actions_log_prob(tr.actions)
                    tr.action_mean, tr.action_sigma, tr.observations, tr.privileged_observations = pol.action_mean, pol.action_std, o, o
                    alg.process_env_step(rew[t, lo:hi].contiguous(), done[t, lo:hi].contiguous(), {})
                alg.compute_returns(last_obs[lo:hi].contiguous())
            idx = torch.arange((hi - lo) * T)
            real_perm = torch.randperm
            torch.randperm = lambda m, **kw: idx
            try:
                alg.update()
            finally:
                torch.randperm = real_perm
            return pol, alg.learning_rate
        finally:
            if not sharded:
                D.world = real
    lo, cnt = D.shard_range(N, rank, world)
    pol_s, lr_s = iterate(lo, lo + cnt, sharded=True)
    flat = torch.cat([p.detach().flatten() for p in pol_s.parameters()])
    both = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    assert torch.equal(both[0], both[1])                      # the ranks stay in lock step
    if rank == 0:
        pol_1, lr_1 = iterate(0, N, sharded=False)
        one = torch.cat([p.detach().flatten() for p in pol_1.parameters()])
        moved = float((one - torch.cat([p.detach().flatten() for p in policy0.parameters()])).abs().max())
        assert lr_s == lr_1, (lr_s, lr_1)
        assert float((flat - one).abs().max()) < 1e-3 * moved, (float((flat - one).abs().max()), moved)


@pytest.mark.parametrize("fn", ["_w_moments", "_w_grads", "_w_env_shards", "_w_reach_shards", "_w_storage_shards", "_w_ppo_update"])
def test_world_size_2_gloo(fn):
    _run(fn)

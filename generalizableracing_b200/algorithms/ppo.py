"""PPO of the reference (standalone/rsl_rl/ext/algorithms/ppo.py:14-190) on the CUDA rollout storage.

Same public methods and hyper-parameters.  Differences, all on the rollout side: ``process_env_step`` hands the
time-out mask to the storage so the bootstrap ``r += gamma*V*time_out`` (ppo.py:89-92) is fused into the single
add_transitions launch; ``compute_returns`` is two launches; mini-batches are gathered by one launch each.
Multi-GPU (env-sharded): advantages are normalised with the moments of the GLOBAL batch and the policy gradient is
all-reduced (mean) before the optimiser step, so R ranks x N envs train like one process with R*N envs.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.optim as optim

from .. import dist_utils as D
from ..storage import RolloutStorage


class PPO:
    def __init__(self, policy, env=None, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95,
                 value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True,
                 schedule="fixed", desired_kl=0.01, device="cuda:0", graphed_update=False, kernel_update=False, **kwargs):
        self.env = env
        self.device = device
        self.desired_kl = desired_kl
        self.schedule = schedule
        self.learning_rate = learning_rate
        self.policy = policy
        self.policy.to(self.device)
        D.broadcast_module(self.policy)
        self.storage = None
        self.optimizer = optim.Adam(self.policy.parameters(), lr=learning_rate)
        self.transition = RolloutStorage.Transition()
        self.clip_param = clip_param
        self.num_learning_epochs = num_learning_epochs
        self.num_mini_batches = num_mini_batches
        self.value_loss_coef = value_loss_coef
        self.entropy_coef = entropy_coef
        self.gamma = gamma
        self.lam = lam
        self.max_grad_norm = max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        # opt-in: one mini-batch step (gather -> forward -> losses -> backward -> clip -> Adam, adaptive LR on the device) captured
        # once in a CUDA graph and replayed num_learning_epochs * num_mini_batches times per iteration (single-GPU runs)
        self.graphed_update = bool(graphed_update) or bool(kernel_update)
        # opt-in on top: the step's forward, loss gradients and weight gradients from libgracing kernels (tensor cores, fp16
        # operands) instead of torch autograd; clipping and Adam stay torch, everything still replayed from one CUDA graph
        self.kernel_update = bool(kernel_update)
        self._graph = None

    def init_storage(self, training_type, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape):
        self.storage = RolloutStorage(training_type, num_envs, num_transitions_per_env, actor_obs_shape, critic_obs_shape, action_shape, self.device)

    def test_mode(self):
        self.policy.eval()

    def train_mode(self):
        self.policy.train()

    def act(self, obs, critic_obs):
        # ppo.py:71-83
        if self.policy.is_recurrent:
            self.transition.hidden_states = self.policy.get_hidden_states()
        self.transition.actions = self.policy.act(obs).detach()
        self.transition.values = self.policy.evaluate(critic_obs).detach()
        self.transition.actions_log_prob = self.policy.get_actions_log_prob(self.transition.actions).detach()
        self.transition.action_mean = self.policy.action_mean.detach()
        self.transition.action_sigma = self.policy.action_std.detach()
        self.transition.observations = obs
        self.transition.privileged_observations = critic_obs
        return self.transition.actions

    def process_env_step(self, rewards, dones, infos):
        # ppo.py:85-97; the reward bootstrap on time-outs happens inside storage.add_transitions
        self.transition.rewards = rewards
        self.transition.dones = dones
        if "time_outs" in infos:
            self.transition.time_outs = infos["time_outs"]
            self.transition.gamma = self.gamma
        self.storage.add_transitions(self.transition)
        self.transition.clear()
        self.policy.reset(dones)

    def compute_returns(self, last_critic_obs, last_values=None):
        # ppo.py:99-101 (last_values: already evaluated by the fused collection kernel)
        if last_values is None:
            last_values = self.policy.evaluate(last_critic_obs).detach()
        _, w = D.world()
        if w == 1:
            self.storage.compute_returns(last_values, self.gamma, self.lam)
        else:
            self.storage.compute_returns(last_values, self.gamma, self.lam, normalize=False)
            self.storage.normalize_advantages(D.merge_moments(self.storage.moments))

    # ------------------------------------------------------------------ CUDA-graph update
    def _mini_batch_losses(self, obs_batch, critic_obs_batch, actions_batch, target_values_batch, advantages_batch, returns_batch,
                           old_actions_log_prob_batch, sample: bool):
        """ppo.py:118-171 for one mini-batch -> (loss, surrogate, value_loss, mu, sigma)"""
        if sample:
            self.policy.act(obs_batch)
        else:
            self.policy.update_distribution(obs_batch)        # same distribution; the reference's (discarded) sample draws nothing we need
        actions_log_prob_batch = self.policy.get_actions_log_prob(actions_batch)
        value_batch = self.policy.evaluate(critic_obs_batch)
        mu_batch, sigma_batch, entropy_batch = self.policy.action_mean, self.policy.action_std, self.policy.entropy
        ratio = torch.exp(actions_log_prob_batch - torch.squeeze(old_actions_log_prob_batch))
        surrogate = -torch.squeeze(advantages_batch) * ratio
        surrogate_clipped = -torch.squeeze(advantages_batch) * torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param)
        surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
        if self.use_clipped_value_loss:
            value_clipped = target_values_batch + (value_batch - target_values_batch).clamp(-self.clip_param, self.clip_param)
            value_loss = torch.max((value_batch - returns_batch).pow(2), (value_clipped - returns_batch).pow(2)).mean()
        else:
            value_loss = (returns_batch - value_batch).pow(2).mean()
        loss = surrogate_loss + self.value_loss_coef * value_loss - self.entropy_coef * entropy_batch.mean()
        return loss, surrogate_loss, value_loss, mu_batch, sigma_batch

    def _kernel_step_factory(self, g, mbs, desc, lr, mb):
        """The captured step of `kernel_update` (ppo.py:118-178 of the reference without autograd), eight launches:
        pack -> sigma -> zero(flat) -> gr_policy_forward_gather -> gr_ppo_loss_grad -> gr_actor_backward_jobs (actor + critic) -> [all-reduce] ->
        gr_adam_clip_step (2).  One flat fp32 buffer holds every gradient (the parameters' .grad are views of it) and, in its
        last 16 floats, the loss kernel's sums (losses, KL, d/d std, loss scales): a multi-GPU run all-reduces exactly that."""
        import ctypes as C
        from .. import _lib as B
        from ..collect import _mlp_layers
        dev, sto, lib, pol = self.device, self.storage, self.storage._lib, self.policy
        (a1, a2, a3), slope = _mlp_layers(pol.actor)
        (c1, c2, c3), slope_c = _mlp_layers(pol.critic)
        if slope != slope_c or pol.noise_std_type != "scalar":
            raise ValueError("kernel_update needs one activation for both nets and a scalar action std")
        for l1, l2, l3, out in ((a1, a2, a3, 4), (c1, c2, c3, 1)):
            if (l1.in_features, l1.out_features, l2.out_features, l3.out_features) != (16, 128, 128, out):
                raise ValueError("kernel_update is built for 16 -> 128 -> 128 -> 4 / 1 MLPs")
        if not isinstance(self.optimizer, optim.Adam) or self.optimizer.param_groups[0].get("weight_decay", 0) or self.optimizer.param_groups[0].get("amsgrad", False):
            raise ValueError("kernel_update implements plain Adam")
        world = D.world()[1]
        net_bytes = int(lib.gr_policy_packed_bytes(128, 128, 1))
        g["packed"] = torch.zeros(2 * net_bytes, dtype=torch.uint8, device=dev)
        g["sigma4"] = torch.ones(4, device=dev)
        g["mu_new"], g["v_new"] = torch.zeros(mb, 4, device=dev), torch.zeros(mb, device=dev)
        g["grad_mu"], g["grad_v"] = torch.zeros(mb, 4, device=dev), torch.zeros(mb, 4, device=dev)
        # ---- flat gradient / moment buffers: every tensor starts on a 16-byte boundary; pol.std's gradient IS sums[3:7] of the tail
        params = [p for p in pol.parameters() if p is not pol.std]
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        tail = off                                   # 16 floats: gr_ppo_loss_grad's sums
        n_flat = tail + 16
        flat, flat_m, flat_v = (torch.zeros(n_flat, device=dev) for _ in range(3))
        # env-sharded runs: the sum of `flat` over the ranks, once per step.  Default: our own kernel over NVLink peer memory
        # (peer.py / csrc/peer_reduce.cu: `flat` lives in symmetric memory, every rank reads the others' buffers between two flag barriers
        # and writes the sum to a local buffer the optimiser kernels read); GRACING_PEER_ALLREDUCE=0, or no symmetric memory: NCCL's all-reduce in place.
        peer = None
        if world > 1 and os.environ.get("GRACING_PEER_ALLREDUCE", "1") != "0":
            ok = torch.ones(1, device=dev)
            try:
                from ..peer import PeerAllReduce
                peer = PeerAllReduce(n_flat, dev)
            except Exception as exc:           # no symmetric memory on this platform (no NVLink / no fabric handles): NCCL carries the sum
                ok.zero_()
                if D.world()[0] == 0:
                    print(f"[gracing] peer all-reduce unavailable ({type(exc).__name__}: {exc}); using NCCL", flush=True)
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)          # every rank takes the same path
            if float(ok) < 1.0:
                peer = None
        if peer is not None:
            flat = peer.buf[:n_flat]
        g["peer_allreduce"] = peer
        reduced = peer.out[:n_flat] if peer is not None else flat            # what the optimiser kernels read
        ksums = flat[tail:tail + 16]
        segs = [(p, o, p.numel()) for p, o in zip(params, offs)] + [(pol.std, tail + 3, 4)]
        old_state = self.optimizer.state
        for p, o, n in segs:                         # adopt the moments of the eager iterations; from here on torch's optimizer state aliases ours
            st = old_state.get(p, {})
            if "exp_avg" in st:
                flat_m[o:o + n].copy_(st["exp_avg"].reshape(-1))
                flat_v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            p.grad = flat[o:o + n].view_as(p)
        step0 = max([float(st["step"]) for st in old_state.values() if "step" in st] or [0.0])
        state = torch.zeros(16, device=dev)
        state[0], state[1] = float(self.learning_rate), step0
        for p, o, n in segs:
            old_state[p] = {"step": state[1], "exp_avg": flat_m[o:o + n].view_as(p), "exp_avg_sq": flat_v[o:o + n].view_as(p)}
        g["lr"] = state[0:1]
        g["flat_grad"], g["flat_m"], g["flat_v"], g["adam_state"] = flat, flat_m, flat_v, state
        ptrs = torch.tensor([p.data_ptr() for p, _, _ in segs], dtype=torch.int64, device=dev)
        seg_off = torch.tensor([o for _, o, _ in segs], dtype=torch.int32, device=dev)
        seg_n = torch.tensor([n for _, _, n in segs], dtype=torch.int32, device=dev)
        order = torch.argsort(seg_off)               # the apply kernel walks the segments in offset order
        ptrs, seg_off, seg_n = ptrs[order].contiguous(), seg_off[order].contiguous(), seg_n[order].contiguous()
        b1, b2 = self.optimizer.param_groups[0]["betas"]
        adaptive = self.desired_kl is not None and self.schedule == "adaptive"
        adam = B.GrAdamStep(ptrs.data_ptr(), seg_off.data_ptr(), seg_n.data_ptr(), len(segs), n_flat, reduced.data_ptr(), flat_m.data_ptr(), flat_v.data_ptr(),
                            state.data_ptr(), reduced[tail:tail + 16].data_ptr() if adaptive else None, 1.0 / world, float(b1), float(b2), float(self.optimizer.param_groups[0]["eps"]),
                            float(self.max_grad_norm), float(self.desired_kl or 0.0), 1e-5, 1e-2)
        mk = lambda l1, l2, l3, out: B.GrMlp(l1.weight.data_ptr(), l1.bias.data_ptr(), l2.weight.data_ptr(), l2.bias.data_ptr(), l3.weight.data_ptr(),
                                             l3.bias.data_ptr(), 16, 128, 128, out)
        mlp_a, mlp_c = mk(a1, a2, a3, 4), mk(c1, c2, c3, 1)
        gr_a = B.GrMlpGrad(a1.weight.grad.data_ptr(), a1.bias.grad.data_ptr(), a2.weight.grad.data_ptr(), a2.bias.grad.data_ptr(), a3.weight.grad.data_ptr(),
                           a3.bias.grad.data_ptr(), 4, 1)
        gr_c = B.GrMlpGrad(c1.weight.grad.data_ptr(), c1.bias.grad.data_ptr(), c2.weight.grad.data_ptr(), c2.bias.grad.data_ptr(), c3.weight.grad.data_ptr(),
                           c3.bias.grad.data_ptr(), 1, 1)
        pol_both = B.GrPolicy(g["packed"].data_ptr(), g["sigma4"].data_ptr(), slope)
        pol_c = B.GrPolicy(g["packed"].data_ptr() + net_bytes, g["sigma4"].data_ptr(), slope)
        # the mini-batch gather (rollout_storage.py:179-187) happens ON LOAD: every kernel reads the storage's own columns at row idx[r]
        # (measured at 65,536 envs: the separate gather launch was 84 us of a 440 us step, and its output was read once)
        idx_ptr = g["idx"].data_ptr()
        s_obs, s_critic = desc.obs, desc.critic_obs or desc.obs
        batch = B.GrPpoBatch(g["mu_new"].data_ptr(), g["v_new"].data_ptr(), g["sigma4"].data_ptr(), desc.actions, desc.log_prob, desc.advantages, desc.returns,
                             desc.values, desc.mu, desc.sigma, float(self.clip_param), float(self.value_loss_coef), float(self.entropy_coef),
                             int(self.use_clipped_value_loss), idx_ptr)
        # transition records (storage.pack_records, once per iteration): one scattered 192-byte record per sampled row instead of nine scattered
        # columns (measured at 393,216 rows: the scattered 4 / 16-byte column reads were what the fused kernels waited for)
        use_rec = os.environ.get("GRACING_PPO_RECORDS", "1") != "0" and sto.obs_shape[0] == 16 and (sto.privileged_obs_shape[0] or 16) == 16
        rec_ptr = sto.pack_records().data_ptr() if use_rec else None
        g["records"] = use_rec
        # dense records (default with records): the iteration's permutation is applied ONCE, while packing (storage.pack_records(indices)), so
        # mini-batch i of every epoch is the contiguous record slice [i*mb, (i+1)*mb) -- the update kernels stream their rows (indices = NULL)
        # instead of gathering one scattered record per row behind a scattered index read; one captured step per mini-batch slot.
        dense = use_rec and os.environ.get("GRACING_PPO_DENSE_RECORDS", "1") != "0"
        g["dense_records"] = dense
        # forward + loss + weight gradients of both nets as ONE launch (gr_ppo_fused_step) or as two (gr_policy_forward_loss ->
        # gr_actor_backward_jobs).  With gathered rows the one launch lost above ~131 k rows (393,216 rows: 7.3 vs 6.4 ms of update -- its tiles
        # waited for their scattered rows with nothing left to overlap); with dense records it streams them and wins at every
        # size measured (gpurun r4j, update per iteration: 16,384 envs 2.0 vs 2.2 ms, 32,768 envs 3.3 vs 3.5 ms, 65,536 envs 5.6 vs 5.7 ms).
        # GRACING_PPO_FUSED_STEP=0|1 forces.
        fused = {"0": False, "1": True}.get(os.environ.get("GRACING_PPO_FUSED_STEP", ""), dense or mb <= 131072)
        p_max_mu, p_max_v = ksums.data_ptr() + 8 * 4, ksums.data_ptr() + 9 * 4

        def slot_args(i):
            ip = None if dense else idx_ptr
            rp = rec_ptr + i * mb * B.GR_RECORD_FLOATS * 4 if dense else rec_ptr
            batch_fl = B.GrPpoBatch(None, None, g["sigma4"].data_ptr(), desc.actions, desc.log_prob, desc.advantages, desc.returns, desc.values, desc.mu, desc.sigma,
                                    float(self.clip_param), float(self.value_loss_coef), float(self.entropy_coef), int(self.use_clipped_value_loss), ip, rp)
            if use_rec:
                jobs = (B.GrBackwardJob * 2)(B.GrBackwardJob(pol_both, rp, g["grad_mu"].data_ptr(), p_max_mu, gr_a, ip, B.GR_RECORD_FLOATS),
                                             B.GrBackwardJob(pol_c, rp + 16 * 4, g["grad_v"].data_ptr(), p_max_v, gr_c, ip, B.GR_RECORD_FLOATS))
            else:
                jobs = (B.GrBackwardJob * 2)(B.GrBackwardJob(pol_both, s_obs, g["grad_mu"].data_ptr(), p_max_mu, gr_a, idx_ptr, 0),
                                             B.GrBackwardJob(pol_c, s_critic, g["grad_v"].data_ptr(), p_max_v, gr_c, idx_ptr, 0))
            fstep = B.GrPpoStep(pol_both, s_obs, s_critic, batch_fl if use_rec else batch, gr_a, gr_c, ksums.data_ptr(), 0.0)
            return batch_fl, jobs, fstep
        slots = [slot_args(i) for i in range(self.num_mini_batches if dense else 1)]
        g["keep_k"] = (mlp_a, mlp_c, gr_a, gr_c, pol_both, pol_c, batch, adam, ptrs, seg_off, seg_n, slots)
        g["kernel_sums"] = True                      # the running loss sums live in adam_state[5:7]

        def step(slot=0):
            batch_fl, jobs, fstep = slots[slot]
            st = torch.cuda.current_stream(dev).cuda_stream
            B.check(lib.gr_policy_pack(C.byref(mlp_a), C.byref(mlp_c), g["packed"].data_ptr(), st), "gr_policy_pack")
            with torch.no_grad():
                g["sigma4"].copy_(pol.std)
                flat.zero_()
            if fused:
                B.check(lib.gr_ppo_fused_step(C.byref(fstep), mb, st), "gr_ppo_fused_step")
            else:
                # forward + loss in one launch (the thread holding a row's mean / value evaluates its loss), then the weight gradients of
                # actor (d/d mu) and critic (d/d v) in one launch
                B.check(lib.gr_policy_forward_loss(C.byref(pol_both), s_obs, s_critic, C.byref(batch_fl), mb, g["grad_mu"].data_ptr(), g["grad_v"].data_ptr(),
                                                   ksums.data_ptr(), st), "gr_policy_forward_loss")
                B.check(lib.gr_actor_backward_jobs(jobs, 2, 128, 128, mb, st), "gr_actor_backward_jobs")
            if peer is not None:       # env-sharded data parallelism: ONE sum per step carries the gradients and the loss / KL sums
                peer.launch()
            elif world > 1:
                torch.distributed.all_reduce(flat)
            B.check(lib.gr_adam_clip_step(C.byref(adam), st), "gr_adam_clip_step")
        return step

    def _build_graph(self):
        """Static mini-batch buffers + one captured optimisation step.  Called after an eager update() (optimizer state exists)."""
        import ctypes as C
        from .. import _lib as B
        sto, dev = self.storage, self.device
        mb = sto.num_envs * sto.num_transitions_per_env // self.num_mini_batches
        od, ad, cd = sto.obs_shape[0], sto.actions_shape[0], sto.privileged_obs_shape[0]
        g = {"idx": torch.zeros(mb, dtype=torch.int64, device=dev)}
        for name, w in (("obs", od), ("critic_obs", cd), ("actions", ad), ("values", 1), ("advantages", 1), ("returns", 1), ("log_prob", 1), ("mu", ad), ("sigma", ad)):
            g[name] = torch.empty(mb, w, device=dev)
        mbs = B.GrMiniBatch()
        for name in ("obs", "critic_obs", "actions", "values", "advantages", "returns", "log_prob", "mu", "sigma"):
            setattr(mbs, name, g[name].data_ptr())
        g["sums"] = torch.zeros(2, device=dev)                               # running (value loss, surrogate loss)
        # Adam with a device-side learning rate: same update rule, graph-capturable
        state = self.optimizer.state_dict()
        lr = torch.tensor(float(self.learning_rate), device=dev)
        self.optimizer = optim.Adam(self.policy.parameters(), lr=lr, capturable=True)
        for st in state["state"].values():
            st["step"] = st["step"].to(dev) if torch.is_tensor(st["step"]) else torch.tensor(float(st["step"]), device=dev)
        state["param_groups"][0]["lr"] = lr
        state["param_groups"][0]["capturable"] = True
        self.optimizer.load_state_dict(state)
        self.optimizer.param_groups[0]["lr"] = lr
        g["lr"] = lr
        desc = sto._desc()
        lib = sto._lib
        adaptive = self.desired_kl is not None and self.schedule == "adaptive"

        def step():
            B.check(lib.gr_storage_gather(C.byref(desc), g["idx"].data_ptr(), mb, C.byref(mbs), torch.cuda.current_stream(dev).cuda_stream), "gr_storage_gather")
            loss, surrogate_loss, value_loss, mu_batch, sigma_batch = self._mini_batch_losses(
                g["obs"], g["critic_obs"], g["actions"], g["values"], g["advantages"], g["returns"], g["log_prob"], sample=False)
            if adaptive:          # ppo.py:124-141 without the host round trip
                with torch.no_grad():
                    kl = torch.sum(torch.log(sigma_batch / g["sigma"] + 1.0e-5) + (torch.square(g["sigma"]) + torch.square(g["mu"] - mu_batch))
                                   / (2.0 * torch.square(sigma_batch)) - 0.5, axis=-1)
                    kl_mean = torch.mean(kl)
                    down = (lr / 1.5).clamp(min=1e-5)
                    up = (lr * 1.5).clamp(max=1e-2)
                    lr.copy_(torch.where(kl_mean > self.desired_kl * 2.0, down, torch.where((kl_mean < self.desired_kl / 2.0) & (kl_mean > 0.0), up, lr)))
            self.optimizer.zero_grad(set_to_none=True)
            loss.backward()
            nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
            self.optimizer.step()
            g["sums"] += torch.stack([value_loss.detach(), surrogate_loss.detach()])

        if self.kernel_update:
            step = self._kernel_step_factory(g, mbs, desc, lr, mb)
        g["keep"] = (mbs, desc)
        # No autograd graph of an earlier (default-stream) step may survive into the capture: its AccumulateGrad nodes are bound
        # to the stream they were created on and would make the engine synchronise across streams while capturing.
        self.policy.distribution = None
        if not self.kernel_update:             # (the kernel step owns static .grad views)
            self.optimizer.zero_grad(set_to_none=True)
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):          # warm-up on the capture stream (allocations, lazy init) before the capture
            for _ in range(3):
                step()
                self.policy.distribution = None
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        if not self.kernel_update:
            self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph, stream=side):
            step()
        g["graph"] = graph
        if g.get("dense_records"):             # one captured step per mini-batch slot: the same launches on the slot's contiguous record slice
            g["graphs"] = [graph]
            for i in range(1, self.num_mini_batches):
                gi = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gi, stream=side, pool=graph.pool()):
                    step(i)
                g["graphs"].append(gi)
        return g

    def _update_graphed(self):
        sto = self.storage
        batch = sto.num_envs * sto.num_transitions_per_env
        mb = batch // self.num_mini_batches
        if self._graph is None:
            # the warm-up steps before the capture advance the parameters and the Adam moments: snapshot, build, restore in place
            import copy
            snap = [p.detach().clone() for p in self.policy.parameters()]
            opt_snap = copy.deepcopy(self.optimizer.state_dict()["state"])
            self._graph = self._build_graph()
            with torch.no_grad():
                for p, q in zip(self.policy.parameters(), snap):
                    p.copy_(q)
                live = self.optimizer.state_dict()["state"]
                for k, st in opt_snap.items():
                    for name, v in st.items():
                        if torch.is_tensor(live[k][name]):
                            live[k][name].copy_(v if torch.is_tensor(v) else torch.tensor(float(v)))
                self._graph["lr"].fill_(float(self.learning_rate))
        g = self._graph
        g["sums"].zero_()
        if g.get("kernel_sums"):
            g["adam_state"][5:7].zero_()
        indices = torch.randperm(self.num_mini_batches * mb, requires_grad=False, device=self.device)
        if g.get("dense_records"):
            # this iteration's transitions -> the static record buffer, IN MINI-BATCH ORDER: record r = transition indices[r] (the reference draws
            # one permutation per update and reuses it in every epoch, rollout_storage.py:165-178), so slot i's captured step streams records
            # [i*mb, (i+1)*mb) -- no index copy, no gather
            sto.pack_records(indices)
            for _ in range(self.num_learning_epochs):
                for i in range(self.num_mini_batches):
                    g["graphs"][i].replay()
        else:
            if g.get("records"):
                sto.pack_records()                   # this iteration's transitions -> the static record buffer the captured step reads
            for _ in range(self.num_learning_epochs):
                for i in range(self.num_mini_batches):
                    g["idx"].copy_(indices[i * mb:(i + 1) * mb])
                    g["graph"].replay()
        num_updates = self.num_learning_epochs * self.num_mini_batches
        sums = g["adam_state"][5:7] if g.get("kernel_sums") else g["sums"]
        out = torch.cat([sums / num_updates, g["lr"].reshape(1)]).tolist()          # the iteration's only host read
        self.learning_rate = out[2]
        if g.get("peer_allreduce") is not None and g["peer_allreduce"].failed():
            raise RuntimeError("gr_peer_allreduce: a rank did not reach the gradient exchange (flag wait gave up); the update of this iteration is invalid")
        self.storage.clear()
        return {"value_function": out[0], "surrogate": out[1]}

    def close(self):
        """Drop the captured update graph (it holds NCCL work when the run is multi-GPU: release it before the process group)."""
        if self._graph is not None:
            torch.cuda.synchronize(self.device)
            self._graph.pop("graph", None)
            self._graph.pop("graphs", None)
            self._graph = None
            import gc
            gc.collect()
            torch.cuda.synchronize(self.device)

    def update(self):
        # ppo.py:103-190
        if self.graphed_update and not self.policy.is_recurrent and (D.world()[1] == 1 or self.kernel_update) and getattr(self, "_eager_updates", 0) >= 1:
            return self._update_graphed()
        self._eager_updates = getattr(self, "_eager_updates", 0) + 1
        mean_value_loss = torch.zeros((), device=self.device)
        mean_surrogate_loss = torch.zeros((), device=self.device)
        if self.policy.is_recurrent:             # ppo.py:106-109
            generator = self.storage.reccurent_mini_batch_generator(self.num_mini_batches, self.num_learning_epochs)
        else:
            generator = self.storage.mini_batch_generator(self.num_mini_batches, self.num_learning_epochs)
        for (obs_batch, critic_obs_batch, actions_batch, target_values_batch, advantages_batch, returns_batch,
             old_actions_log_prob_batch, old_mu_batch, old_sigma_batch, hid_states_batch, masks_batch) in generator:
            if self.policy.is_recurrent:
                self.policy.act(obs_batch, masks=masks_batch, hidden_states=hid_states_batch[0])
                actions_log_prob_batch = self.policy.get_actions_log_prob(actions_batch)
                value_batch = self.policy.evaluate(critic_obs_batch, masks=masks_batch, hidden_states=hid_states_batch[1])
            else:
                self.policy.act(obs_batch)
                actions_log_prob_batch = self.policy.get_actions_log_prob(actions_batch)
                value_batch = self.policy.evaluate(critic_obs_batch)
            mu_batch = self.policy.action_mean
            sigma_batch = self.policy.action_std
            entropy_batch = self.policy.entropy

            if self.desired_kl is not None and self.schedule == "adaptive":
                with torch.inference_mode():
                    kl = torch.sum(torch.log(sigma_batch / old_sigma_batch + 1.0e-5)
                                   + (torch.square(old_sigma_batch) + torch.square(old_mu_batch - mu_batch)) / (2.0 * torch.square(sigma_batch)) - 0.5, axis=-1)
                    kl_mean = torch.mean(kl)
                    _, w = D.world()
                    if w > 1:
                        torch.distributed.all_reduce(kl_mean)
                        kl_mean /= w
                    if kl_mean > self.desired_kl * 2.0:
                        self.learning_rate = max(1e-5, self.learning_rate / 1.5)
                    elif kl_mean < self.desired_kl / 2.0 and kl_mean > 0.0:
                        self.learning_rate = min(1e-2, self.learning_rate * 1.5)
                    for param_group in self.optimizer.param_groups:
                        param_group["lr"] = self.learning_rate

            ratio = torch.exp(actions_log_prob_batch - torch.squeeze(old_actions_log_prob_batch))
            surrogate = -torch.squeeze(advantages_batch) * ratio
            surrogate_clipped = -torch.squeeze(advantages_batch) * torch.clamp(ratio, 1.0 - self.clip_param, 1.0 + self.clip_param)
            surrogate_loss = torch.max(surrogate, surrogate_clipped).mean()
            if self.use_clipped_value_loss:
                value_clipped = target_values_batch + (value_batch - target_values_batch).clamp(-self.clip_param, self.clip_param)
                value_losses = (value_batch - returns_batch).pow(2)
                value_losses_clipped = (value_clipped - returns_batch).pow(2)
                value_loss = torch.max(value_losses, value_losses_clipped).mean()
            else:
                value_loss = (returns_batch - value_batch).pow(2).mean()
            loss = surrogate_loss + self.value_loss_coef * value_loss - self.entropy_coef * entropy_batch.mean()

            self.optimizer.zero_grad()
            loss.backward()
            D.allreduce_mean_grads(self.policy.parameters())          # env-sharded data parallelism (NCCL over NVLink)
            nn.utils.clip_grad_norm_(self.policy.parameters(), self.max_grad_norm)
            self.optimizer.step()
            mean_value_loss += value_loss.detach()
            mean_surrogate_loss += surrogate_loss.detach()

        num_updates = self.num_learning_epochs * self.num_mini_batches
        self.storage.clear()
        return {"value_function": (mean_value_loss / num_updates).item(), "surrogate": (mean_surrogate_loss / num_updates).item()}

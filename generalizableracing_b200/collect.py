"""FusedCollector -- the PPO collection phase (``OnPolicyRunner.learn``'s rollout loop, standalone/rsl_rl/ext/runners/
on_policy_runner.py:141-175 of the reference) as ONE launch of ``gr_ppo_collect``: the actor / critic MLPs run on the
tensor cores inside the kernel that steps the envs and fills the rollout storage (csrc/ppo_collect.cu).

Opt-in (``train_cfg["fused_collection"] = True``): the env / storage side is bit-identical to the step-by-step path given
the same actions, the policy inference runs with fp16 operands and fp32 accumulation (the optimiser still sees the fp32
torch modules; the PPO ratio is formed against the log-probs stored here).  Works for the racing task's state-only
``ActorCritic`` (16 -> 128 -> 128 -> 4 / 1, LeakyReLU or ReLU, QD/agents/rsl_rl_ppo_cfg.py:22-27); anything else raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn

from . import _lib as B
from . import layout as L


def _mlp_layers(seq: nn.Sequential):
    lin = [m for m in seq if isinstance(m, nn.Linear)]
    act = [m for m in seq if not isinstance(m, nn.Linear)]
    if len(lin) != 3 or len(act) != 2:
        raise ValueError("fused collection needs MLPs with exactly two hidden layers")
    slopes = set()
    for a in act:
        if isinstance(a, nn.LeakyReLU):
            slopes.add(float(a.negative_slope))
        elif isinstance(a, nn.ReLU):
            slopes.add(0.0)
        else:
            raise ValueError(f"fused collection supports LeakyReLU / ReLU activations, got {type(a).__name__}")
    if len(slopes) != 1:
        raise ValueError("fused collection needs one activation slope for the whole net")
    return lin, slopes.pop()


class FusedCollector:
    def __init__(self, env, policy, storage, gamma: float, groups_per_cta: int = 0):
        if env._bptt is not None or env.rng_mode != "philox":
            raise ValueError("fused collection needs a non-differentiable env drawing in-kernel (rng_mode='philox')")
        if policy.is_recurrent:
            raise ValueError("recurrent policies are out of scope")
        self.env, self.policy, self.storage = env, policy, storage
        self._lib = env._lib
        dev = env.device
        (a1, a2, a3), sa = _mlp_layers(policy.actor)
        (c1, c2, c3), sc = _mlp_layers(policy.critic)
        if sa != sc:
            raise ValueError("actor and critic must share the activation")
        for l1, l2, l3, out in ((a1, a2, a3, L.NUM_ACTIONS), (c1, c2, c3, 1)):
            if (l1.in_features, l1.out_features, l2.in_features, l2.out_features, l3.in_features, l3.out_features) != (L.OBS_DIM, 128, 128, 128, 128, out):
                raise ValueError("fused collection is built for 16 -> 128 -> 128 -> 4 / 1 MLPs (QD/agents/rsl_rl_ppo_cfg.py:22-27)")
            for l in (l1, l2, l3):
                if l.bias is None or l.weight.dtype != torch.float32 or not l.weight.is_contiguous() or l.weight.device != dev:
                    raise ValueError("fused collection needs contiguous fp32 Linear layers with bias on the env's device")
        if storage.num_envs != env.num_envs or storage.privileged_observations is None:
            raise ValueError("storage does not match the env")
        self._layers = ((a1, a2, a3, L.NUM_ACTIONS), (c1, c2, c3, 1))
        self.slope = sa
        self.packed = torch.zeros(int(self._lib.gr_policy_packed_bytes(128, 128, 2)), dtype=torch.uint8, device=dev)
        self.sigma = torch.ones(4, device=dev)
        self.last_values = torch.zeros(env.num_envs, 1, device=dev)
        self.episode_acc = torch.zeros(env.num_envs, 2, device=dev)
        self.episode_log = torch.zeros(B.GR_LOG_SHARDS, 4, device=dev)
        self.gamma = float(gamma)
        self.groups_per_cta = int(groups_per_cta)
        self.coop_reset_columns = int(os.environ.get("GRACING_COLLECT_COOP_COLUMNS", "0"))      # 0 = as many as fit (<= 4), -1 = off (A/B switch)
        self._pol = B.GrPolicy(self.packed.data_ptr(), self.sigma.data_ptr(), self.slope)

    def _mlp(self, l1, l2, l3, out) -> B.GrMlp:
        return B.GrMlp(l1.weight.data_ptr(), l1.bias.data_ptr(), l2.weight.data_ptr(), l2.bias.data_ptr(), l3.weight.data_ptr(), l3.bias.data_ptr(),
                       L.OBS_DIM, 128, 128, out)

    def pack(self):
        """fp32 torch parameters -> packed fp16 operands + action std (call after every optimiser update; two tiny launches)."""
        a, c = self._mlp(*self._layers[0]), self._mlp(*self._layers[1])
        B.check(self._lib.gr_policy_pack(C.byref(a), C.byref(c), self.packed.data_ptr(), self.env._stream()), "gr_policy_pack")
        with torch.no_grad():
            p = self.policy
            self.sigma.copy_(p.std if p.noise_std_type == "scalar" else torch.exp(p.log_std))

    def collect(self):
        """One rollout of ``storage.num_transitions_per_env`` steps.  Returns (obs, critic_obs, last_values) after the last step;
        the storage is full (``storage.step == T``), ``env.extras`` / ``env.get_observations()`` reflect the last step."""
        env, sto = self.env, self.storage
        if env._needs_reset:
            env.reset()
        T = sto.num_transitions_per_env
        src = env._last
        k = env._flip
        dst = env._outs[k]
        if dst is src:
            k ^= 1
            dst = env._outs[k]
        env._flip = k ^ 1
        io = B.GrCollectIO(src["obs"].data_ptr(), src["critic"].data_ptr(), dst["obs"].data_ptr(), dst["critic"].data_ptr(), dst["aux"].data_ptr(),
                           self.last_values.data_ptr(), self.episode_acc.data_ptr(), env._log_accum.data_ptr(), self.episode_log.data_ptr(), self.gamma, self.groups_per_cta,
                           self.coop_reset_columns)
        rng = env._rng
        rng.rnd = None
        rng.step = env._step_count & 0xFFFFFFFF
        env._step_count += T
        desc = sto._desc()
        B.check(self._lib.gr_ppo_collect(env._p_cfg, env._p_track, env._p_state, C.byref(rng), C.byref(self._pol), C.byref(desc), C.byref(io),
                                         env._stream()), "gr_ppo_collect")
        sto.step = T
        # envs that reset anywhere in the rollout rewrote their read-mostly planes: one single-step launch without the pre-dependency prefetch
        env._state.launch_flags = env._launch_flags & ~B.GR_LAUNCH_PREFETCH
        env._params_edited = True
        env._last = dst
        ex = env.extras
        dict.pop(ex, "log", None)
        ex["observations"] = env._obs_dict(dst)
        return dst["obs"], dst["critic"], self.last_values

    def episode_stats(self) -> torch.Tensor:
        """(sum of episode rewards, sum of episode lengths, finished episodes) since the last call (device tensor, no sync)."""
        acc = self.episode_log.sum(dim=0)[:3]
        self.episode_log.zero_()
        return acc


class FusedBpttCollector:
    """The forward half of a BPTT window (``AlgoRunner.learn``'s rollout loop, standalone/diff_rl/algorithms/runner.py:110-126
    of the reference) as ONE launch of ``gr_bptt_collect``: actor MLP on the tensor cores + rsample + differentiable env.step
    with tape and losses, T times.  The backward is ``BpttWindow.backward_window()`` (one ``gr_step_bwd`` launch) followed by
    ONE batched torch forward/backward of the actor over the recorded ``[T*N]`` (observation, noise) rows
    (:meth:`policy_backward`) -- the same function ``sum_t actor(obs_t) + std * eps_t`` the reference differentiates step by
    step.  Actor widths (128, 128) or (256, 128), LeakyReLU / ReLU.  Opt-in (``train_cfg["fused_collection"]``)."""

    def __init__(self, env, policy, horizon: int, groups_per_cta: int = 0, backward_tf32: bool = False, backward_kernel: bool = False):
        self.backward_tf32 = bool(backward_tf32)
        self.backward_kernel = bool(backward_kernel)
        if env._bptt is None or env.rng_mode != "philox":
            raise ValueError("fused BPTT collection needs a differentiable env drawing in-kernel (is_differentiable_physics, rng_mode='philox')")
        (l1, l2, l3), slope = _mlp_layers(policy.actor)
        dims = (l1.in_features, l1.out_features, l2.in_features, l2.out_features, l3.in_features, l3.out_features)
        if dims not in ((L.OBS_DIM, 128, 128, 128, 128, L.NUM_ACTIONS), (L.OBS_DIM, 256, 256, 128, 128, L.NUM_ACTIONS)):
            raise ValueError("fused BPTT collection is built for 16 -> 128 -> 128 -> 4 and 16 -> 256 -> 128 -> 4 actors")
        for l in (l1, l2, l3):
            if l.bias is None or l.weight.dtype != torch.float32 or not l.weight.is_contiguous() or l.weight.device != env.device:
                raise ValueError("fused collection needs contiguous fp32 Linear layers with bias on the env's device")
        if horizon > env._bptt.capacity:
            raise ValueError("horizon exceeds the env's tape capacity (bptt_horizon)")
        self.env, self.policy, self.T = env, policy, int(horizon)
        self._lib, self._layers, self.slope = env._lib, (l1, l2, l3), slope
        self.h1, self.h2 = l1.out_features, l2.out_features
        dev, N, T = env.device, env.num_envs, self.T
        self.packed = torch.zeros(int(self._lib.gr_policy_packed_bytes(self.h1, self.h2, 1)), dtype=torch.uint8, device=dev)
        self.sigma = torch.ones(4, device=dev)
        self.obs_seq = torch.zeros(T, N, L.OBS_DIM, device=dev)
        self.eps_seq = torch.zeros(T, N, L.NUM_ACTIONS, device=dev)
        self.actions = torch.zeros(T, N, L.NUM_ACTIONS, device=dev)
        self.rewards = torch.zeros(T, N, device=dev)
        self.dones = torch.zeros(T, N, dtype=torch.uint8, device=dev)
        self.groups_per_cta = int(groups_per_cta)
        self._pol = B.GrPolicy(self.packed.data_ptr(), self.sigma.data_ptr(), self.slope)

    def pack(self):
        l1, l2, l3 = self._layers
        a = B.GrMlp(l1.weight.data_ptr(), l1.bias.data_ptr(), l2.weight.data_ptr(), l2.bias.data_ptr(), l3.weight.data_ptr(), l3.bias.data_ptr(),
                    L.OBS_DIM, self.h1, self.h2, L.NUM_ACTIONS)
        B.check(self._lib.gr_policy_pack(C.byref(a), None, self.packed.data_ptr(), self.env._stream()), "gr_policy_pack")
        with torch.no_grad():
            p = self.policy
            self.sigma.copy_(p.std if p.noise_std_type == "scalar" else torch.exp(p.log_std))

    def collect(self):
        """One window of ``horizon`` steps (call ``env.unwrapped.detach()`` first, as the runner does).  Afterwards the env's BPTT
        window holds the tape and the losses of the T steps; returns (obs, critic_obs) after the last step."""
        env, win, T = self.env, self.env._bptt, self.T
        if env._needs_reset:
            env.reset()
        if win.t != 0:
            raise RuntimeError("fused BPTT collection starts a window: call env.unwrapped.detach() first")
        src = env._last
        k = env._flip
        dst = env._outs[k]
        if dst is src:
            k ^= 1
            dst = env._outs[k]
        env._flip = k ^ 1
        io = B.GrBpttCollectIO(src["obs"].data_ptr(), dst["obs"].data_ptr(), dst["critic"].data_ptr(), dst["aux"].data_ptr(), self.obs_seq.data_ptr(),
                               self.eps_seq.data_ptr(), self.actions.data_ptr(), win.loss.data_ptr(), win.loss_terms.data_ptr(), self.rewards.data_ptr(),
                               self.dones.data_ptr(), win.tape.data_ptr(), env._stride, env._log_accum.data_ptr(), T, self.groups_per_cta)
        rng = env._rng
        rng.rnd = None
        rng.step = env._step_count & 0xFFFFFFFF
        env._step_count += T
        B.check(self._lib.gr_bptt_collect(env._p_cfg, env._p_track, env._p_state, C.byref(rng), C.byref(self._pol), self.h1, self.h2, C.byref(io),
                                          env._stream()), "gr_bptt_collect")
        win.t = T
        env._state.launch_flags = env._launch_flags & ~B.GR_LAUNCH_PREFETCH
        env._params_edited = True
        env._last = dst
        ex = env.extras
        dict.pop(ex, "log", None)
        ex["observations"] = env._obs_dict(dst)
        return dst["obs"], dst["critic"]

    def policy_backward_kernel(self, grad_actions: torch.Tensor):
        """The same gradients from ONE launch of ``gr_actor_backward`` (csrc/actor_backward.cu): activations recomputed and all
        weight-gradient GEMMs on the tensor cores with fp16 operands (loss-scaled) and fp32 accumulation in tensor memory."""
        T, N = self.T, self.env.num_envs
        p = self.policy
        g = grad_actions.reshape(T * N, L.NUM_ACTIONS)
        if not g.is_contiguous():
            g = g.contiguous()
        scale = (1024.0 / g.abs().max().clamp_min(1e-30)).reshape(1).float()          # device scalar, no host sync
        l1, l2, l3 = self._layers
        grads = [torch.zeros_like(t) for t in (l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias)]
        out = B.GrMlpGrad(*(t.data_ptr() for t in grads), L.NUM_ACTIONS, 0)
        B.check(self._lib.gr_actor_backward(C.byref(self._pol), self.h1, self.h2, self.obs_seq.data_ptr(), g.data_ptr(), scale.data_ptr(), T * N,
                                            C.byref(out), self.env._stream()), "gr_actor_backward")
        for t, gt in zip((l1.weight, l1.bias, l2.weight, l2.bias, l3.weight, l3.bias), grads):
            t.grad = gt if t.grad is None else t.grad + gt
        std = p.std if p.noise_std_type == "scalar" else p.log_std
        gs = (g * self.eps_seq.reshape(T * N, L.NUM_ACTIONS)).sum(dim=0)
        if p.noise_std_type != "scalar":
            gs = gs * torch.exp(p.log_std.detach())
        std.grad = gs if std.grad is None else std.grad + gs

    def policy_backward(self, grad_actions: torch.Tensor, tf32: bool = False):
        """Accumulate d(loss)/d(policy parameters) from the sweep's ``grad_actions`` [T,N,4]: one batched actor pass over
        [T*N] rows (fp32 like the reference; ``tf32=True`` lets cuBLAS use TF32 tensor cores for it: 4.9 -> 2.2 ms at 16,384 x 32)."""
        T, N = self.T, self.env.num_envs
        p = self.policy
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
        try:
            mu = p.actor(self.obs_seq.reshape(T * N, L.OBS_DIM))
            std = p.std if p.noise_std_type == "scalar" else torch.exp(p.log_std)
            a = mu + std * self.eps_seq.reshape(T * N, L.NUM_ACTIONS)
            torch.autograd.backward([a], [grad_actions.reshape(T * N, L.NUM_ACTIONS)])
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev

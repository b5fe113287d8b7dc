"""``torch.ops.gracing.*`` -- the operator layer over the C ABI (SURVEY.md §8b "Operator layer").

Thin ``torch.library.custom_op`` wrappers around ``libgracing.so`` (include/gracing.h): every tensor a kernel reads or
writes is an explicit operator argument (mutations are declared in the schema), the POD configuration structs
(``GrConfig`` / ``GrTrack`` / launch geometry) travel as an integer handle of the env that owns them
(:func:`register_env`).  Operators:

===============================  =================================================================================
``gracing::step_fwd``            one ``env.step()`` (manager_based_diff_rl_env.py:160-267): mutates the state planes
``gracing::step_fwd_tape``       the same with loss + tape slice (differentiable physics)
``gracing::step_loss``           functional ``loss_t(action_t)``; ``register_autograd`` wires the analytic reverse step,
                                 so ``extras["losses"].mean().backward()`` of the reference trainers
                                 (standalone/diff_rl/algorithms/bptt.py:38-44) keeps working
``gracing::step_bwd``            reverse sweep over tape steps ``[t_begin, t_end)`` -> ``grad_action``
``gracing::rollout_fwd[_tape]``  T steps in one launch for actions known in advance (gr_rollout_fwd)
``gracing::reset``               ``_reset_idx`` (+ observations) of all / masked envs (:362-410)
``gracing::gae``                 ``RolloutStorage.compute_returns`` (rollout_storage.py:113-127), functional
===============================  =================================================================================

The kernels only exist for CUDA devices: the operators launch through the library the env was built with
(``_lib.load()`` raises when ``libgracing.so`` is missing) and ``gracing::gae`` refuses non-CUDA tensors.  Fake-tensor
implementations (shapes / dtypes only) are registered so the operators can be traced.  ``RacingVecEnv(op_layer=True)``
routes ``step`` / ``reset`` / the BPTT autograd through these operators instead of calling ctypes directly -- same
kernels, bit-identical results (tests/test_ops.py); the direct path stays the default because a Python custom-op
dispatch costs more host time than the 9 us kernel it launches.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib as B
from . import layout as L

_ENVS: "weakref.WeakValueDictionary[int, object]" = weakref.WeakValueDictionary()
_NEXT = [1]


def register_env(env) -> int:
    """Handle under which the operators find the env's configuration structs (weakly held)."""
    h = getattr(env, "_op_handle", None)
    if h is None:
        h = _NEXT[0]
        _NEXT[0] += 1
        env._op_handle = h
        _ENVS[h] = env
    return h


def _env(h: int):
    env = _ENVS.get(int(h))
    if env is None:
        raise RuntimeError(f"gracing: unknown or released env handle {h} (register_env(env) first and keep the env alive)")
    return env


def _stream(t: Tensor):
    return torch.cuda.current_stream(t.device).cuda_stream if t.device.type == "cuda" else None


def _f32c(t: Tensor, what: str) -> Tensor:
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError(f"gracing: {what} must be a contiguous float32 tensor")
    return t


def _state_for(env, planes: Tensor) -> B.GrState:
    """The env's GrState, re-pointed at ``planes`` (any tensor with the env's tile layout: functional use on a clone)."""
    if planes.shape != env.planes.shape or planes.dtype != torch.float32 or not planes.is_contiguous() or planes.device != env.planes.device:
        raise ValueError(f"gracing: planes must be a contiguous float32 {tuple(env.planes.shape)} tensor on {env.planes.device}")
    s = env._state
    if planes.data_ptr() == s.planes:
        return s
    return B.GrState(planes.data_ptr(), s.plane_stride, s.num_envs, s.num_planes, s.env_id_offset, s.max_types_per_block,
                     s.block_threads, s.launch_flags & ~B.GR_LAUNCH_PREFETCH, s.chunk_types)


def _rng(env, rnd: Optional[Tensor], step: int, keep: list) -> B.GrRandom:
    if rnd is None:
        if env.rng_mode == "dense":
            raise ValueError("gracing: rng_mode='dense' needs an explicit rnd tensor every call")
        return B.GrRandom(None, env.seed, step & 0xFFFFFFFF)
    if tuple(rnd.shape) != (env.num_envs, L.RND_STRIDE):
        raise ValueError(f"gracing: rnd must be [{env.num_envs}, {L.RND_STRIDE}]")
    keep.append(_f32c(rnd, "rnd"))
    return B.GrRandom(rnd.data_ptr(), env.seed, step & 0xFFFFFFFF)


def _step_outputs(N: int, dev, export_terms: bool):
    f = dict(device=dev, dtype=torch.float32)
    return (torch.empty(N, L.OBS_DIM, **f), torch.empty(N, L.OBS_DIM, **f), torch.empty(N, 1, **f), torch.empty(N, **f),
            torch.empty(N, dtype=torch.uint8, device=dev), torch.empty(N, dtype=torch.uint8, device=dev),
            torch.empty(N, dtype=torch.int64, device=dev),
            torch.empty(N if export_terms else 0, L.NUM_REWARD_TERMS, **f),
            torch.empty(N if export_terms else 0, dtype=torch.uint8, device=dev))


def _launch_step(env, planes, action, rnd, step, log_accum, export_terms, loss, loss_terms, tape_t):
    N = env.num_envs
    if tuple(action.shape) != (N, L.NUM_ACTIONS):
        raise ValueError(f"Invalid action shape, expected: ({N}, {L.NUM_ACTIONS}), received: {tuple(action.shape)}.")
    _f32c(action, "action")
    outs = _step_outputs(N, planes.device, export_terms)
    obs, critic, aux, reward, terminated, time_out, dones, terms, passed = outs
    io = B.GrStepIO()
    io.action = action.data_ptr()
    io.obs, io.critic_obs, io.aux_obs = obs.data_ptr(), critic.data_ptr(), aux.data_ptr()
    io.reward, io.terminated, io.time_out, io.dones = reward.data_ptr(), terminated.data_ptr(), time_out.data_ptr(), dones.data_ptr()
    if export_terms:
        io.reward_terms, io.gate_passed = terms.data_ptr(), passed.data_ptr()
    io.log_accum = log_accum.data_ptr()
    if tape_t is not None:
        io.loss, io.loss_terms, io.tape = loss.data_ptr(), loss_terms.data_ptr(), tape_t.data_ptr()
        io.tape_stride = env._stride
    keep: list = []
    st = _state_for(env, planes)
    rng = _rng(env, rnd, step, keep)
    B.check(env._lib.gr_step_fwd(C.byref(env._gcfg), C.byref(env._track), C.byref(st), C.byref(rng), C.byref(io), _stream(planes)), "gr_step_fwd")
    return outs


_T9 = Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]
_T11 = Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]


@torch.library.custom_op("gracing::step_fwd", mutates_args=("planes", "log_accum"))
def step_fwd(env: int, planes: Tensor, action: Tensor, rnd: Optional[Tensor], step: int, log_accum: Tensor,
             export_terms: bool = False) -> _T9:
    """-> (obs[N,16], critic_obs[N,16], aux[N,1], reward[N], terminated[N] u8, time_out[N] u8, dones[N] i64,
    reward_terms[N,6] | [0,6], gate_passed[N] u8 | [0])."""
    e = _env(env)
    if e._bptt is not None:
        raise RuntimeError("gracing::step_fwd on a differentiable env: use gracing::step_fwd_tape (the loss / tape outputs are part of the step)")
    return _launch_step(e, planes, action, rnd, step, log_accum, export_terms, None, None, None)


@step_fwd.register_fake
def _(env, planes, action, rnd, step, log_accum, export_terms=False):
    return _step_outputs(action.shape[0], action.device, export_terms)


@torch.library.custom_op("gracing::step_fwd_tape", mutates_args=("planes", "log_accum", "tape"))
def step_fwd_tape(env: int, planes: Tensor, action: Tensor, rnd: Optional[Tensor], step: int, log_accum: Tensor,
                  tape: Tensor, t: int, export_terms: bool = False) -> _T11:
    """Step ``t`` of the current BPTT window: the nine outputs of ``step_fwd`` + (loss[N], loss_terms[N,3]) and the tape
    slice ``tape[t]`` the reverse sweep reads."""
    e = _env(env)
    if e._bptt is None:
        raise RuntimeError("gracing::step_fwd_tape needs cfg.is_differentiable_physics")
    if not 0 <= t < tape.shape[0]:
        raise RuntimeError(f"BPTT horizon exceeded the tape capacity ({tape.shape[0]} steps): call env.unwrapped.detach() "
                           "between windows or construct the env with a larger bptt_horizon")
    N = e.num_envs
    loss = torch.empty(N, device=planes.device)
    loss_terms = torch.empty(N, e._bptt._terms, device=planes.device)
    outs = _launch_step(e, planes, action, rnd, step, log_accum, export_terms, loss, loss_terms, tape[t])
    return outs + (loss, loss_terms)


@step_fwd_tape.register_fake
def _(env, planes, action, rnd, step, log_accum, tape, t, export_terms=False):
    N = action.shape[0]
    terms = _env(env)._bptt._terms if env in _ENVS and _env(env)._bptt is not None else 3         # 3 racing loss terms, 4 on the reach-target env
    return _step_outputs(N, action.device, export_terms) + (action.new_empty(N), action.new_empty(N, terms))


@torch.library.custom_op("gracing::step_bwd", mutates_args=("adjoint", "grad_action"))
def step_bwd(env: int, planes: Tensor, tape: Tensor, grad_loss: Optional[Tensor], grad_scale: float, adjoint: Tensor,
             grad_action: Tensor, t_begin: int, t_end: int) -> None:
    """Reverse sweep over tape steps ``[t_begin, t_end)``: ``grad_loss`` [T,N] = dL/d(loss) (or the uniform ``grad_scale``);
    ``adjoint`` [5,stride,4] carries dL/d(state) between calls; writes ``grad_action[t-1]`` for every step t >= 1."""
    e = _env(env)
    if not 0 <= t_begin <= t_end <= tape.shape[0] or grad_action.shape[0] < t_end or (grad_loss is not None and grad_loss.shape[0] < t_end):
        raise ValueError("gracing::step_bwd: [t_begin, t_end) outside the tape / gradient buffers")
    io = B.GrBwdIO()
    io.tape, io.tape_stride = _f32c(tape, "tape").data_ptr(), e._stride
    io.t_begin, io.t_end = t_begin, t_end
    io.grad_loss = None if grad_loss is None else _f32c(grad_loss, "grad_loss").data_ptr()
    io.grad_scale = float(grad_scale)
    io.adjoint, io.adj_stride = _f32c(adjoint, "adjoint").data_ptr(), e._stride
    io.grad_action = _f32c(grad_action, "grad_action").data_ptr()
    fn = getattr(e, "_bwd_fn", None) or e._lib.gr_step_bwd
    st = _state_for(e, planes) if isinstance(e._state, B.GrState) else e._state          # (the reach-target env keeps its own state struct)
    B.check(fn(C.byref(e._gcfg), C.byref(st), C.byref(io), _stream(tape)), "gr_step_bwd")


@torch.library.custom_op("gracing::step_loss", mutates_args=())
def step_loss(env: int, action: Tensor, token: Tensor, loss: Tensor, t: int, epoch: int) -> Tuple[Tensor, Tensor]:
    """The differentiable face of step ``t``: ``loss_t`` as a function of ``action_t`` and the hidden env state (functional:
    returns a copy of the loss ``step_fwd_tape`` wrote, plus the token that chains step t -> t+1).  torch refuses autograd
    formulas on mutating operators, so the launch (``step_fwd_tape``) and its derivative (this operator, whose backward is
    ``step_bwd`` over ``[t, t+1)``) are two operators."""
    return loss.clone(), token.new_zeros(1)


@step_loss.register_fake
def _(env, action, token, loss, t, epoch):
    return torch.empty_like(loss), token.new_empty(1)


def _loss_setup(ctx, inputs, output):
    ctx.env, ctx.t, ctx.epoch = inputs[0], inputs[4], inputs[5]
    ctx.set_materialize_grads(False)


def _loss_backward(ctx, g_loss, g_token):
    e = _env(ctx.env)
    win, t = e._bptt, ctx.t
    if ctx.epoch != win.epoch:
        raise RuntimeError("backward through a BPTT window that was already detached (env.detach() started a new window)")
    if g_loss is not None:
        win.grad_loss[t].copy_(g_loss)
    else:
        win.grad_loss[t].zero_()
    torch.ops.gracing.step_bwd(ctx.env, e.planes, win.tape, win.grad_loss, 0.0, win.adjoint, win.grad_action, t, t + 1)
    return None, win.grad_action[t].clone(), None if g_token is None else torch.zeros_like(g_token), None, None, None


step_loss.register_autograd(_loss_backward, setup_context=_loss_setup)


def _launch_rollout(e, planes, actions, rnd, step, log_accum, record_obs, tape, t0):
    N = e.num_envs
    if actions.dim() != 3 or tuple(actions.shape[1:]) != (N, L.NUM_ACTIONS) or actions.shape[0] < 1:
        raise ValueError(f"Invalid actions shape, expected: (T, {N}, {L.NUM_ACTIONS}), received: {tuple(actions.shape)}.")
    _f32c(actions, "actions")
    T, dev = actions.shape[0], planes.device
    f = dict(device=dev, dtype=torch.float32)
    u8 = dict(device=dev, dtype=torch.uint8)
    obs, critic, aux = torch.empty(N, L.OBS_DIM, **f), torch.empty(N, L.OBS_DIM, **f), torch.empty(N, 1, **f)
    reward, dones, terminated, time_out = torch.empty(T, N, **f), torch.empty(T, N, **u8), torch.empty(T, N, **u8), torch.empty(T, N, **u8)
    obs_seq = torch.empty(T if record_obs else 0, N, L.OBS_DIM, **f)
    io = B.GrRolloutIO()
    io.actions, io.T = actions.data_ptr(), T
    io.obs_out, io.critic_obs_out, io.aux_out = obs.data_ptr(), critic.data_ptr(), aux.data_ptr()
    io.reward, io.dones, io.terminated, io.time_out = reward.data_ptr(), dones.data_ptr(), terminated.data_ptr(), time_out.data_ptr()
    if record_obs:
        io.obs_seq = obs_seq.data_ptr()
    io.log_accum = log_accum.data_ptr()
    outs = (obs, critic, aux, reward, dones, terminated, time_out, obs_seq)
    if tape is not None:
        if not 0 <= t0 <= t0 + T <= tape.shape[0]:
            raise RuntimeError(f"BPTT horizon exceeded the tape capacity ({tape.shape[0]} steps): call env.unwrapped.detach() "
                               "between windows or construct the env with a larger bptt_horizon")
        loss, loss_terms = torch.empty(T, N, **f), torch.empty(T, N, e._bptt._terms, **f)
        io.loss, io.loss_terms, io.tape, io.tape_stride = loss.data_ptr(), loss_terms.data_ptr(), tape[t0].data_ptr(), e._stride
        outs = outs + (loss, loss_terms)
    rng = B.GrRandom(None, e.seed, step & 0xFFFFFFFF)
    if rnd is not None:
        if tuple(rnd.shape) != (T, N, L.RND_STRIDE):
            raise ValueError(f"gracing: rnd must be [{T}, {N}, {L.RND_STRIDE}]")
        rng.rnd = _f32c(rnd, "rnd").data_ptr()
    elif e.rng_mode == "dense":
        raise ValueError("gracing: rng_mode='dense' needs an explicit rnd tensor every call")
    st = _state_for(e, planes)
    B.check(e._lib.gr_rollout_fwd(C.byref(e._gcfg), C.byref(e._track), C.byref(st), C.byref(rng), C.byref(io), _stream(planes)), "gr_rollout_fwd")
    return outs


_T8 = Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]
_T10 = Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]


@torch.library.custom_op("gracing::rollout_fwd", mutates_args=("planes", "log_accum"))
def rollout_fwd(env: int, planes: Tensor, actions: Tensor, rnd: Optional[Tensor], step: int, log_accum: Tensor, record_obs: bool = False) -> _T8:
    """T steps in one launch for ``actions`` [T,N,4] known in advance -> (obs[N,16], critic_obs[N,16], aux[N,1] after the last step,
    reward[T,N], dones / terminated / time_out [T,N] u8, obs_seq[T,N,16] | [0,N,16])."""
    e = _env(env)
    if e._bptt is not None:
        raise RuntimeError("gracing::rollout_fwd on a differentiable env: use gracing::rollout_fwd_tape")
    return _launch_rollout(e, planes, actions, rnd, step, log_accum, record_obs, None, 0)


@rollout_fwd.register_fake
def _(env, planes, actions, rnd, step, log_accum, record_obs=False):
    T, N = actions.shape[0], actions.shape[1]
    f, u = actions.new_empty, lambda *sh: actions.new_empty(*sh, dtype=torch.uint8)
    return (f(N, L.OBS_DIM), f(N, L.OBS_DIM), f(N, 1), f(T, N), u(T, N), u(T, N), u(T, N), f(T if record_obs else 0, N, L.OBS_DIM))


@torch.library.custom_op("gracing::rollout_fwd_tape", mutates_args=("planes", "log_accum", "tape"))
def rollout_fwd_tape(env: int, planes: Tensor, actions: Tensor, rnd: Optional[Tensor], step: int, log_accum: Tensor, tape: Tensor, t0: int,
                     record_obs: bool = False) -> _T10:
    """The same with differentiable physics: tape steps ``[t0, t0 + T)`` are written, + (loss[T,N], loss_terms[T,N,3])."""
    e = _env(env)
    if e._bptt is None:
        raise RuntimeError("gracing::rollout_fwd_tape needs cfg.is_differentiable_physics")
    return _launch_rollout(e, planes, actions, rnd, step, log_accum, record_obs, tape, t0)


@rollout_fwd_tape.register_fake
def _(env, planes, actions, rnd, step, log_accum, tape, t0, record_obs=False):
    T, N = actions.shape[0], actions.shape[1]
    f, u = actions.new_empty, lambda *sh: actions.new_empty(*sh, dtype=torch.uint8)
    return (f(N, L.OBS_DIM), f(N, L.OBS_DIM), f(N, 1), f(T, N), u(T, N), u(T, N), u(T, N), f(T if record_obs else 0, N, L.OBS_DIM), f(T, N), f(T, N, 3))


@torch.library.custom_op("gracing::reset", mutates_args=("planes",))
def reset(env: int, planes: Tensor, mask: Optional[Tensor], rnd: Optional[Tensor], step: int) -> Tuple[Tensor, Tensor, Tensor]:
    """``_reset_idx`` of the envs with ``mask != 0`` (None: all) followed by the observation pass ->
    (obs[N,16], critic_obs[N,16], aux[N,1])."""
    e = _env(env)
    N, dev = e.num_envs, planes.device
    keep: list = []
    if mask is not None:
        if mask.dtype == torch.bool:
            mask = mask.view(torch.uint8)
        if mask.dtype != torch.uint8 or tuple(mask.shape) != (N,) or not mask.is_contiguous():
            raise ValueError(f"gracing::reset: mask must be a contiguous bool / uint8 [{N}] tensor")
        keep.append(mask)
    obs, critic, aux = torch.empty(N, L.OBS_DIM, device=dev), torch.empty(N, L.OBS_DIM, device=dev), torch.empty(N, 1, device=dev)
    st = _state_for(e, planes)
    rng = _rng(e, rnd, step, keep)
    B.check(e._lib.gr_env_reset(C.byref(e._gcfg), C.byref(e._track), C.byref(st), C.byref(rng), B.ptr(mask),
                                obs.data_ptr(), critic.data_ptr(), aux.data_ptr(), _stream(planes)), "gr_env_reset")
    return obs, critic, aux


@reset.register_fake
def _(env, planes, mask, rnd, step):
    N = _env(env).num_envs
    return planes.new_empty(N, L.OBS_DIM), planes.new_empty(N, L.OBS_DIM), planes.new_empty(N, 1)


@torch.library.custom_op("gracing::gae", mutates_args=())
def gae(rewards: Tensor, values: Tensor, dones: Tensor, last_values: Tensor, gamma: float, lam: float,
        normalize: bool = True) -> Tuple[Tensor, Tensor, Tensor]:
    """rollout_storage.py:113-127 on ``[T,N,1]`` rewards / values (float32) and dones (uint8) + ``last_values`` [N,1] ->
    (returns[T,N,1], advantages[T,N,1], moments[3] f64 = (count, mean, M2) of the raw advantages)."""
    if rewards.device.type != "cuda":
        raise RuntimeError("gracing::gae runs only on a CUDA device (sm_100a); there is no CPU fallback")
    lib = B.load()
    if rewards.dim() < 2:
        raise ValueError("gracing::gae: rewards must be [T,N] or [T,N,1]")
    T, N = rewards.shape[0], rewards.shape[1]
    for x, name, n in ((rewards, "rewards", T * N), (values, "values", T * N), (last_values, "last_values", N)):
        _f32c(x, name)
        if x.numel() != n:
            raise ValueError(f"gracing::gae: {name} has {x.numel()} elements, expected {n}")
    if dones.dtype == torch.bool:
        dones = dones.view(torch.uint8)
    if dones.dtype != torch.uint8 or dones.numel() != T * N or not dones.is_contiguous():
        raise ValueError("gracing::gae: dones must be a contiguous uint8 / bool [T,N,1] tensor")
    dev = rewards.device
    returns, adv = torch.empty_like(rewards), torch.empty_like(rewards)
    scratch = torch.empty(int(lib.gr_gae_scratch_bytes(N)) // 8 + 1, dtype=torch.float64, device=dev)
    moments = torch.empty(3, dtype=torch.float64, device=dev)
    s = B.GrStorage()
    s.rewards, s.dones, s.values, s.returns, s.advantages = rewards.data_ptr(), dones.data_ptr(), values.data_ptr(), returns.data_ptr(), adv.data_ptr()
    s.T, s.N = T, N
    B.check(lib.gr_compute_returns(C.byref(s), last_values.data_ptr(), float(gamma), float(lam), scratch.data_ptr(), moments.data_ptr(),
                                   int(normalize), _stream(rewards)), "gr_compute_returns")
    return returns, adv, moments


@gae.register_fake
def _(rewards, values, dones, last_values, gamma, lam, normalize=True):
    return torch.empty_like(rewards), torch.empty_like(rewards), rewards.new_empty(3, dtype=torch.float64)


OPERATORS = ("step_fwd", "step_fwd_tape", "step_loss", "step_bwd", "rollout_fwd", "rollout_fwd_tape", "reset", "gae")

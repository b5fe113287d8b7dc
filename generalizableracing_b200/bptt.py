"""BPTT window: tape buffers + autograd wiring of the analytic reverse sweep (gr_step_bwd).

The reference builds a torch.autograd tape through DroneDynamics/CTBRController over the horizon
(standalone/diff_rl/algorithms/runner.py:107-155, bptt.py:38-44).  Here every forward step appends
7 float4 planes per env to a tape in HBM and the backward is ONE kernel over the window
(:meth:`BpttWindow.backward_window`, used by :class:`..algorithms.BPTT`), or -- for code that calls
``extras["losses"]...backward()`` itself -- one launch per step through a chained autograd.Function that
reproduces the same gradients for arbitrary upstream weights.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as B
from . import layout as L


class _StepLoss(torch.autograd.Function):
    """loss_t = f(action_t, hidden env state).  The token chains step t+1 -> t so the engine runs the reverse
    sweep in time order; the carried adjoints live in the window, not in autograd."""

    @staticmethod
    def forward(ctx, action, token, win, t, epoch):
        ctx.win, ctx.t, ctx.epoch = win, t, epoch
        return win.loss[t].clone(), token.new_zeros(1)

    @staticmethod
    def backward(ctx, g_loss, g_token):
        win, t = ctx.win, ctx.t
        if ctx.epoch != win.epoch:
            raise RuntimeError("backward through a BPTT window that was already detached (env.detach() started a new window)")
        if g_loss is not None:
            win.grad_loss[t].copy_(g_loss)
        else:
            win.grad_loss[t].zero_()
        win._launch(t, t + 1, win.grad_loss, 0.0)
        return win.grad_action[t].clone(), torch.zeros_like(g_token) if g_token is not None else None, None, None, None


class BpttWindow:
    def __init__(self, env, capacity: int):
        self.env = env
        self.capacity = 0
        self.epoch = 0
        self.t = 0
        self.autograd = True
        self._alloc(capacity)
        dev, N = env.device, env.num_envs
        self.adjoint = torch.zeros(5, env._stride, 4, device=dev)
        self.losses_detached = torch.zeros(N, device=dev)
        self._token = None

    def _alloc(self, capacity: int):
        env = self.env
        dev, N = env.device, env.num_envs
        self._terms = getattr(env, "_num_loss_terms", 3)               # (the reach-target env: 13 tape planes, 4 loss terms)
        self.tape = torch.zeros(capacity, env.num_tiles, getattr(env, "_tape_planes", L.TAPE_PLANES), L.TILE, 4, device=dev)     # per-warp tiles, like the state
        self.loss = torch.zeros(capacity, N, device=dev)
        self.loss_terms = torch.zeros(capacity, N, self._terms, device=dev)
        self.grad_loss = torch.zeros(capacity, N, device=dev)
        self.grad_action = torch.zeros(capacity, N, L.NUM_ACTIONS, device=dev)
        # reward / dones of every step of the window, written by the step kernel itself: a differentiable-physics loop keeps what a
        # step returns for the whole window (naive_train.py:170-172), so each step gets its own row instead of a copy of a ping-pong buffer
        self.reward = torch.zeros(capacity, N, device=dev)
        self.dones = torch.zeros(capacity, N, dtype=torch.int64, device=dev)
        self._p_reward, self._p_dones = self.reward.data_ptr(), self.dones.data_ptr()
        self.capacity = capacity
        self._p_loss, self._p_terms, self._p_tape = self.loss.data_ptr(), self.loss_terms.data_ptr(), self.tape.data_ptr()
        self._tape_step_bytes = self.tape[0].numel() * 4

    # -- called by the env ------------------------------------------------------------------
    def start_window(self):
        self.t = 0
        self.epoch += 1
        self.adjoint.zero_()
        self.grad_action.zero_()
        self._token = None

    def bind_step(self, io):
        if self.t >= self.capacity:
            raise RuntimeError(f"BPTT horizon exceeded the tape capacity ({self.capacity} steps): call env.unwrapped.detach() "
                               "between windows or construct the env with a larger bptt_horizon")
        t, n = self.t, self.env.num_envs
        io.loss = self._p_loss + t * n * 4
        io.loss_terms = self._p_terms + t * n * 4 * self._terms
        io.tape = self._p_tape + t * self._tape_step_bytes
        io.tape_stride = self.env._stride

    def bind_outputs(self, io):
        """Point the step's reward / dones at this step's rows of the window (call after bind_step, before the launch)."""
        t, n = self.t, self.env.num_envs
        io.reward = self._p_reward + t * n * 4
        io.dones = self._p_dones + t * n * 8
        return self.reward[t], self.dones[t]

    def after_step(self, actions: torch.Tensor, extras: dict):
        t = self.t
        self.t += 1
        if self.autograd and actions.requires_grad:
            if self._token is None:
                self._token = torch.zeros(1, device=self.env.device, requires_grad=True)
            loss, self._token = _StepLoss.apply(actions, self._token, self, t, self.epoch)
        else:
            loss = self.loss[t]
        extras["losses"] = loss
        extras["losses_detached"] = self.losses_detached
        extras["loss_terms"] = self.loss_terms[t]

    # -- reverse sweep -------------------------------------------------------------------------
    def _launch(self, t_begin: int, t_end: int, grad_loss, grad_scale: float):
        env = self.env
        if getattr(env, "_ops", None) is not None:          # operator layer (ops.py): the same launch through torch.ops.gracing.step_bwd
            torch.ops.gracing.step_bwd(env._op_handle, env.planes, self.tape, grad_loss, float(grad_scale), self.adjoint, self.grad_action,
                                       t_begin, t_end)
            return
        io = B.GrBwdIO()
        io.tape = self.tape.data_ptr()
        io.tape_stride = env._stride
        io.t_begin, io.t_end = t_begin, t_end
        io.grad_loss = None if grad_loss is None else grad_loss.data_ptr()
        io.grad_scale = float(grad_scale)
        io.adjoint = self.adjoint.data_ptr()
        io.adj_stride = env._stride
        io.grad_action = self.grad_action.data_ptr()
        fn = getattr(env, "_bwd_fn", None) or env._lib.gr_step_bwd
        B.check(fn(C.byref(env._gcfg), C.byref(env._state), C.byref(io), env._stream()), "gr_step_bwd")

    def backward_window(self, grad_scale: float = None, grad_losses: torch.Tensor = None) -> torch.Tensor:
        """One reverse sweep over the whole window.  ``grad_losses`` [T,N] = dL/d(extras["losses"]) or a uniform
        ``grad_scale`` (BPTT.update: 1/(T*N)).  Returns dL/d(action) [T,N,4]; row T-1 is zero (its effect lies in the
        next window because of the 1-step action lag)."""
        T = self.t
        if T == 0:
            raise RuntimeError("empty BPTT window")
        self.adjoint.zero_()
        if grad_losses is not None:
            self.grad_loss[:T].copy_(grad_losses)
            self._launch(0, T, self.grad_loss, 0.0)
        else:
            self._launch(0, T, None, 1.0 / (T * self.env.num_envs) if grad_scale is None else grad_scale)
        return self.grad_action[:T]

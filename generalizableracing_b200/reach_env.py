"""ReachTargetVecEnv -- the reference's reach-target tasks (DiffLab-Quadcopter-LV-ReachTarget-v0 /
DiffLab-Quadcopter-CTBR-ReachTarget-v0, QD/__init__.py:21-46) on top of libgracing.so, with the same wrapper surface as
:class:`RacingVecEnv`: ``get_observations()``, ``step(actions) -> (obs, rew, dones, extras)``, ``extras["time_outs"]``,
``extras["log"]`` and, with ``cfg.is_differentiable_physics``, ``extras["losses"]`` through the analytic reverse sweep.
The command mode is ``cfg.controller``: "LVController", "PSController" (L/controllers/controller_diff.py:172-443) or
"CTBRController".  One kernel launch per step (gr_reach_step_fwd); no host synchronisation; no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as B
from . import layout as L
from .config import CONTROLLERS, REACH_REWARD_TERM_NAMES, ReachTargetCfg
from .env import _Extras


def make_gr_reach_config(cfg: ReachTargetCfg) -> B.GrReachConfig:
    if cfg.controller not in CONTROLLERS:
        raise ValueError(f"controller must be one of {CONTROLLERS} (the ThrustController mode of QD/mdp/diff_action.py:189-201 is "
                         "marked buggy by the reference and not built)")
    if cfg.action_lag != 1:
        raise ValueError("only action_lag == 1 is built")
    g = B.GrReachConfig()
    g.controller = CONTROLLERS.index(cfg.controller)
    g.sim2real_test, g.last_action_modified, g.random_drag = int(cfg.sim2real_test), int(cfg.last_action_modified), int(cfg.random_drag)
    g.dt, g.max_episode_length = cfg.step_dt, cfg.max_episode_length
    g.gravity, g.grad_decay, g.mass = cfg.gravity, cfg.grad_decay, cfg.mass
    g.inertia[:] = cfg.inertia_diag
    g.action_scale[:] = cfg.action_scale
    g.action_offset[:] = cfg.action_offset
    g.thrust_lo, g.thrust_hi = cfg.gross_thrust_bound
    g.body_rate_bound = cfg.body_rate_bound
    g.kp[:], g.kd[:] = cfg.rate_gain_p, cfg.rate_gain_d
    g.thrust_delay = cfg.thrust_ctrl_delay
    g.torque_delay[:] = cfg.torque_ctrl_delay
    g.speed_gain[:], g.pose_gain[:], g.rate_gain[:], g.pos_gain[:] = cfg.speed_gain, cfg.pose_gain, cfg.rate_gain, cfg.pos_gain
    g.max_feedback_accel = cfg.max_feedback_accel
    g.drag1, g.drag1_rand, g.drag2, g.drag2_rand = cfg.drag_1, cfg.drag_1_rand, cfg.drag_2, cfg.drag_2_rand
    g.z_drag, g.z_drag_rand = cfg.z_drag, cfg.z_drag_rand
    g.thr_err_reset_std = cfg.thr_est_error_reset_std
    g.default_pos[:] = cfg.default_root_pos
    g.reset_lo[:], g.reset_hi[:] = cfg.reset_lo, cfg.reset_hi
    g.cmd_lo[:], g.cmd_hi[:] = cfg.cmd_lo, cfg.cmd_hi
    g.resample_time = cfg.resampling_time
    g.term_oob, g.oob_lo, g.oob_hi = int(cfg.term_out_of_bound), cfg.oob_lo, cfg.oob_hi
    g.w_reward[:] = cfg.w_reward
    g.move_in_dir_thr, g.reach_thr, g.hover_thr, g.hover_ratio = cfg.move_in_dir_threshold, cfg.reach_threshold, cfg.hover_threshold, cfg.hover_ratio
    g.w_loss[:] = cfg.w_loss
    g.loss_dir_thr, g.loss_smooth_ratio = cfg.loss_dir_threshold, cfg.loss_smooth_ratio
    return g


class ReachTargetVecEnv:
    num_actions = L.NUM_ACTIONS
    num_obs = L.REACH_OBS_DIM
    num_privileged_obs = L.REACH_OBS_DIM
    _tape_planes = L.REACH_TAPE_PLANES
    _num_loss_terms = L.REACH_NUM_LOSS_TERMS

    def __init__(self, cfg: ReachTargetCfg, num_envs: int, device="cuda:0", seed: int = 42, rng_mode: str = "philox", env_id_offset: int = 0,
                 bptt_horizon: int = 0, _lib=None):
        self.cfg = cfg
        self.num_envs = N = int(num_envs)
        self.device = torch.device(device)
        if _lib is None:
            if self.device.type != "cuda":
                raise RuntimeError("ReachTargetVecEnv runs only on a CUDA device (sm_100a); there is no CPU fallback")
            _lib = B.load()
        self._lib = _lib
        if rng_mode not in ("philox", "dense"):
            raise ValueError("rng_mode must be 'philox' or 'dense'")
        self.rng_mode, self.seed = rng_mode, int(seed)
        self.max_episode_length, self.step_dt = cfg.max_episode_length, cfg.step_dt
        self._gcfg = make_gr_reach_config(cfg)
        dev = self.device
        self.num_tiles = (N + L.TILE - 1) // L.TILE
        self._stride = self.num_tiles * L.TILE
        self.planes = torch.zeros(self.num_tiles, L.REACH_PLANES, L.TILE, 4, dtype=torch.float32, device=dev)
        self._state = B.GrReachState(self.planes.data_ptr(), self._stride, N, int(env_id_offset))
        self._rng = B.GrRandom(None, self.seed, 0)
        self._step_count = 0

        def outs():
            return dict(obs=torch.zeros(N, L.REACH_OBS_DIM, device=dev), reward=torch.zeros(N, device=dev),
                        terminated=torch.zeros(N, dtype=torch.uint8, device=dev), time_out=torch.zeros(N, dtype=torch.uint8, device=dev),
                        dones=torch.zeros(N, dtype=torch.int64, device=dev), reward_terms=torch.zeros(N, L.REACH_NUM_REWARD_TERMS, device=dev))
        self._outs = [outs(), outs()]
        self._flip = 0
        self._log_accum = torch.zeros(B.GR_LOG_SHARDS, B.GR_LOG_SLOTS, device=dev)
        self.extras = _Extras(self)
        self.export_reward_terms = False
        self._bwd_fn = self._lib.gr_reach_step_bwd
        self._bptt = None
        if cfg.is_differentiable_physics:
            from .bptt import BpttWindow
            self._bptt = BpttWindow(self, bptt_horizon or 64)
        self._last = self._outs[0]
        self._needs_reset = True
        self._ios = None

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None

    def _rand(self, rnd):
        if rnd is not None:
            rnd = rnd.to(self.device, torch.float32).contiguous()
            if rnd.shape != (self.num_envs, L.REACH_RND_STRIDE):
                raise ValueError(f"rnd must be [{self.num_envs}, {L.REACH_RND_STRIDE}]")
            self._rnd_keepalive = rnd
            self._rng.rnd = rnd.data_ptr()
        else:
            if self.rng_mode == "dense":
                raise ValueError("rng_mode='dense' needs an explicit rnd tensor every call")
            self._rng.rnd = None
        self._rng.step = self._step_count & 0xFFFFFFFF
        self._step_count += 1
        return self._rng

    @property
    def unwrapped(self):
        return self

    def read_plane(self, pl: int) -> torch.Tensor:
        return self.planes[:, pl].reshape(-1, 4)[: self.num_envs]

    @property
    def episode_length_buf(self) -> torch.Tensor:
        return self.planes[:, L.RPL_LINVEL, :, 3].reshape(-1)[: self.num_envs].view(torch.int32)

    @episode_length_buf.setter
    def episode_length_buf(self, value: torch.Tensor):
        buf = self.planes[:, L.RPL_LINVEL, :, 3].reshape(-1).view(torch.int32).clone()
        buf[: self.num_envs] = value.to(self.device, torch.int32).reshape(-1)
        self.planes[:, L.RPL_LINVEL, :, 3].view(torch.int32).copy_(buf.view(self.num_tiles, L.TILE))

    def state_dict_view(self) -> dict:
        R = self.read_plane
        q, pos, lin, ang, tq, aa, ff, tg = (R(p) for p in (L.RPL_QUAT, L.RPL_POS, L.RPL_LINVEL, L.RPL_ANGVEL, L.RPL_TORQUE, L.RPL_ANGACC, L.RPL_FIFO, L.RPL_TARGET))
        e0, e1, e2, d2, d1 = (R(p) for p in (L.RPL_EPSUM0, L.RPL_EPSUM1, L.RPL_EPSUM2, L.RPL_DRAG2, L.RPL_DRAG1))
        return {"root_quat_w": q, "root_pos_w": pos[:, :3], "gross_thrust": pos[:, 3], "root_lin_vel_w": lin[:, :3],
                "episode_length": lin[:, 3].contiguous().view(torch.int32), "root_ang_vel_b": ang[:, :3], "time_left": ang[:, 3],
                "torque": tq[:, :3], "ang_acc_b": aa[:, :3], "fresh": aa[:, 3] == 1, "action_fifo": ff, "pose_command_w": tg[:, :3],
                "raw_actions": torch.stack([tg[:, 3], e2[:, 2], e2[:, 3], tq[:, 3]], dim=-1),
                "episode_sums": torch.cat([e0, e1, e2[:, :2]], dim=-1), "drag_coeffs": d2[:, :3], "h_force_drag_coeffs": d1[:, :3],
                "thr_est_error": d1[:, 3]}

    def _build_log(self) -> dict:
        acc = self._log_accum.sum(dim=0)
        n = acc[B.GR_REACH_LOG_NUM_RESET].clamp(min=1.0)
        log = {}
        for k, name in enumerate(REACH_REWARD_TERM_NAMES):
            if self.cfg.w_reward[k] != 0.0:
                log["Episode_Reward/" + name] = acc[B.GR_REACH_LOG_SUM_EPSUM + k] / n / self.cfg.episode_length_s
        log["Metrics/desired_pos_b/position_error"] = acc[B.GR_REACH_LOG_SUM_POS_ERR] / n
        log["Episode_Termination/time_out"] = acc[B.GR_REACH_LOG_NUM_TIMEOUT].clone()
        log["Episode_Termination/base_contact"] = acc[B.GR_REACH_LOG_NUM_TERMINATED].clone()
        self._log_accum = torch.zeros_like(self._log_accum)
        return log

    # ------------------------------------------------------------------ API
    def reset(self, rnd: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None):
        o = self._outs[self._flip]
        self._flip ^= 1
        m = None
        if mask is not None:
            m = mask.to(self.device, torch.uint8).contiguous()
        B.check(self._lib.gr_reach_reset(C.byref(self._gcfg), C.byref(self._state), C.byref(self._rand(rnd)), B.ptr(m), o["obs"].data_ptr(), self._stream()),
                "gr_reach_reset")
        self._last = o
        self._needs_reset = False
        if self._bptt is not None:
            self._bptt.start_window()
        self.extras["observations"] = {"policy": o["obs"]}
        return self._grad_safe_obs(o["obs"]), self.extras

    def _grad_safe_obs(self, obs: torch.Tensor, needed: bool = True) -> torch.Tensor:
        """See RacingVecEnv._grad_safe_obs: the ping-pong output buffers are rewritten through raw pointers two steps later, which
        autograd cannot see; a policy differentiated across the steps of a window gets a fresh copy."""
        return obs.clone() if (self._bptt is not None and needed) else obs

    def get_observations(self):
        if self._needs_reset:
            self.reset()
        o = self._last
        return self._grad_safe_obs(o["obs"]), {"observations": {"policy": o["obs"]}}

    def detach(self):
        if self._bptt is not None:
            self._bptt.start_window()

    def close(self):
        pass

    def _make_io(self, o) -> B.GrReachStepIO:
        io = B.GrReachStepIO()
        io.obs, io.reward, io.terminated, io.time_out, io.dones = (o[k].data_ptr() for k in ("obs", "reward", "terminated", "time_out", "dones"))
        return io

    def step(self, actions: torch.Tensor, rnd: Optional[torch.Tensor] = None):
        if self._needs_reset:
            self.reset()
        k = self._flip
        o = self._outs[k]
        self._flip = k ^ 1
        act = actions.detach() if actions.requires_grad else actions
        if act.dtype != torch.float32 or not act.is_contiguous() or act.device != self.device:
            act = act.to(self.device, torch.float32).contiguous()
        if act.shape != (self.num_envs, L.NUM_ACTIONS):
            raise ValueError(f"Invalid action shape, expected: ({self.num_envs}, {L.NUM_ACTIONS}), received: {tuple(act.shape)}.")
        if self._ios is None:
            self._ios = [self._make_io(self._outs[0]), self._make_io(self._outs[1])]
            self._views = [(x["time_out"].view(torch.bool), x["terminated"].view(torch.bool)) for x in self._outs]
        io = self._ios[k]
        io.action = act.data_ptr()
        io.reward_terms = o["reward_terms"].data_ptr() if self.export_reward_terms else None
        io.log_accum = self._log_accum.data_ptr()
        reward, dones = o["reward"], o["dones"]
        if self._bptt is not None:
            self._bptt.bind_step(io)
            reward, dones = self._bptt.bind_outputs(io)     # this step's rows of the window: valid until the window is detached
        rc = self._lib.gr_reach_step_fwd(C.byref(self._gcfg), C.byref(self._state), C.byref(self._rand(rnd)), C.byref(io), self._stream())
        if rc:
            B.check(rc, "gr_reach_step_fwd")
        self._last = o if self._bptt is None else dict(o, reward=reward, dones=dones)
        ex = self.extras
        dict.pop(ex, "log", None)
        to, term = self._views[k]
        ex["observations"] = {"policy": o["obs"]}
        ex["time_outs"] = to
        ex["terminated"] = term
        if self._bptt is not None:
            self._bptt.after_step(actions, ex)
        return self._grad_safe_obs(o["obs"], actions.requires_grad), reward, dones, ex

    def rollout(self, actions: torch.Tensor, rnd: Optional[torch.Tensor] = None, record_obs: bool = False) -> dict:
        """``for t in range(T): env.step(actions[t])`` in ONE launch (gr_reach_rollout_fwd) for actions known in advance
        (``actions`` [T,N,4]; dense mode: ``rnd`` [T,N,24]); same contract as :meth:`RacingVecEnv.rollout`."""
        if self._needs_reset:
            self.reset()
        dev, N = self.device, self.num_envs
        act = actions.detach()
        if act.dtype != torch.float32 or not act.is_contiguous() or act.device != dev:
            act = act.to(dev, torch.float32).contiguous()
        if act.dim() != 3 or tuple(act.shape[1:]) != (N, L.NUM_ACTIONS) or act.shape[0] < 1:
            raise ValueError(f"Invalid actions shape, expected: (T, {N}, {L.NUM_ACTIONS}), received: {tuple(act.shape)}.")
        T = act.shape[0]
        rng = B.GrRandom(None, self.seed, self._step_count & 0xFFFFFFFF)
        if rnd is not None:
            rnd = rnd.to(dev, torch.float32).contiguous()
            if tuple(rnd.shape) != (T, N, L.REACH_RND_STRIDE):
                raise ValueError(f"rnd must be [{T}, {N}, {L.REACH_RND_STRIDE}]")
            rng.rnd = rnd.data_ptr()
        elif self.rng_mode == "dense":
            raise ValueError("rng_mode='dense' needs an explicit rnd tensor every call")
        self._step_count += T
        k = self._flip
        o = self._outs[k]
        self._flip = k ^ 1
        out = {"reward": torch.empty(T, N, device=dev), "dones": torch.empty(T, N, dtype=torch.uint8, device=dev),
               "terminated": torch.empty(T, N, dtype=torch.uint8, device=dev), "time_outs": torch.empty(T, N, dtype=torch.uint8, device=dev)}
        io = B.GrReachRolloutIO()
        io.actions, io.T, io.obs_out = act.data_ptr(), T, o["obs"].data_ptr()
        io.reward, io.dones, io.terminated, io.time_out = (out[n].data_ptr() for n in ("reward", "dones", "terminated", "time_outs"))
        if record_obs:
            out["obs_seq"] = torch.empty(T, N, L.REACH_OBS_DIM, device=dev)
            io.obs_seq = out["obs_seq"].data_ptr()
        io.log_accum = self._log_accum.data_ptr()
        win = self._bptt
        if win is not None:
            if win.t + T > win.capacity:
                raise RuntimeError(f"BPTT horizon exceeded the tape capacity ({win.capacity} steps): call env.unwrapped.detach() "
                                   "between windows or construct the env with a larger bptt_horizon")
            t0 = win.t
            io.loss, io.loss_terms, io.tape = win.loss[t0].data_ptr(), win.loss_terms[t0].data_ptr(), win.tape[t0].data_ptr()
            io.tape_stride = self._stride
        B.check(self._lib.gr_reach_rollout_fwd(C.byref(self._gcfg), C.byref(self._state), C.byref(rng), C.byref(io), self._stream()), "gr_reach_rollout_fwd")
        self._last = o
        ex = self.extras
        dict.pop(ex, "log", None)
        for n in ("dones", "terminated", "time_outs"):
            out[n] = out[n].view(torch.bool)
        ex["observations"] = {"policy": o["obs"]}
        ex["time_outs"], ex["terminated"] = out["time_outs"][-1], out["terminated"][-1]
        out["obs"] = o["obs"]
        if win is not None:
            out["losses"], out["loss_terms"] = win.loss[t0:t0 + T], win.loss_terms[t0:t0 + T]
            win.t += T
        return out


_TASKS = {"DiffLab-Quadcopter-LV-ReachTarget-v0": ReachTargetCfg.lv, "DiffLab-Quadcopter-CTBR-ReachTarget-v0": ReachTargetCfg.ctbr,
          "DiffLab-Quadcopter-PS-ReachTarget-v0": ReachTargetCfg.ps}


def make_reach_env(task: str = "DiffLab-Quadcopter-LV-ReachTarget-v0", num_envs: int = 2048, device="cuda:0", cfg: Optional[ReachTargetCfg] = None,
                   **kwargs) -> ReachTargetVecEnv:
    """``gym.make`` of the reach-target tasks (QD/__init__.py:21-46; the PS id is ours: the reference registers no task for
    PSController).  ``cfg`` overrides the task default; cfg keyword overrides go through ``ReachTargetCfg.lv(**kw)`` etc.
    Under torchrun the envs are sharded (Philox keyed by the global env id)."""
    from . import dist_utils as D
    if task not in _TASKS:
        raise ValueError(f"unknown task {task!r}; built: {sorted(_TASKS)}")
    rank, _world = D.world()
    return ReachTargetVecEnv(cfg or _TASKS[task](), num_envs, device=device, env_id_offset=rank * num_envs, **kwargs)

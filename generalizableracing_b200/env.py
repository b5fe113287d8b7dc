"""RacingVecEnv -- the reference's vectorised racing env surface on top of libgracing.so.

Mirrors what ``OnPolicyRunner`` / ``AlgoRunner`` see of the reference env through Isaac Lab's
``RslRlVecEnvWrapper``: ``get_observations() -> (obs, {"observations": {...}})``,
``step(actions) -> (obs, rew, dones, extras)`` with ``extras["observations"]["critic"]``,
``extras["time_outs"]``, ``extras["log"]`` and, when ``cfg.is_differentiable_physics``,
``extras["losses"]`` (requires grad) / ``["losses_detached"]`` / ``["loss_terms"]``;
attributes ``num_envs, num_actions, num_obs, device, cfg, episode_length_buf`` (assignable),
``max_episode_length``, ``unwrapped.detach()``, ``reset()``, ``close()``.
Reference: extensions/diff.lab/diff/lab/envs/manager_based_diff_rl_env.py:160-267,362-416.

Every step is ONE kernel launch (gr_step_fwd); there is no host synchronisation and no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from . import _lib as B
from . import layout as L
from .config import RacingCfg
from .tracks import GateTable


def make_gr_config(cfg: RacingCfg) -> B.GrConfig:
    """RacingCfg -> GrConfig; derived constants are formed in fp32 in the order torch forms them."""
    f32 = np.float32
    g = B.GrConfig()
    g.dt = cfg.step_dt
    g.max_episode_length = cfg.max_episode_length
    g.gravity = cfg.gravity
    g.grad_decay = cfg.grad_decay
    g.inertia[:] = cfg.inertia_diag
    # QD/mdp/diff_action.py:56,261: weight = mass*|g| ; scale = weight*ratio/(1-(-1))   (fp32 tensor ops)
    weight = f32(cfg.mass) * f32(abs(-cfg.gravity))
    g.action_scale0 = float(f32(f32(weight * f32(cfg.max_thrust_weight_ratio)) / f32(2.0)))
    g.body_rate_bound = cfg.body_rate_bound
    g.thrust_lo, g.thrust_hi = cfg.gross_thrust_bound
    g.update_threshold = cfg.update_threshold
    g.drag1, g.drag1_rand, g.drag2, g.drag2_rand = cfg.drag_1, cfg.drag_1_rand, cfg.drag_2, cfg.drag_2_rand
    g.z_drag, g.z_drag_rand = cfg.z_drag, cfg.z_drag_rand
    g.random_drag = int(cfg.random_drag)
    g.thr_err_reset_std, g.thr_err_init_std = cfg.thr_est_error_reset_std, cfg.thr_est_error_init_std
    g.default_pos[:] = cfg.default_root_pos
    g.reset_pos, g.reset_roll_pitch, g.reset_yaw, g.reset_vel = cfg.reset_pos, cfg.reset_roll_pitch, cfg.reset_yaw, cfg.reset_vel
    g.kp[:] = cfg.rate_gain_p
    g.kd[:] = cfg.rate_gain_d
    g.thrust_delay = cfg.thrust_ctrl_delay
    g.torque_delay[:] = cfg.torque_ctrl_delay
    g.pid_scale_lo, g.pid_scale_span = cfg.pid_scale[0], cfg.pid_scale[1] - cfg.pid_scale[0]
    g.delay_scale_lo, g.delay_scale_span = cfg.delay_scale[0], cfg.delay_scale[1] - cfg.delay_scale[0]
    g.mass = cfg.mass
    g.max_init_level = cfg.max_init_terrain_level
    g.term_oob, g.term_bad_pose = int(cfg.term_out_of_bound), int(cfg.term_bad_pose)
    g.oob_lo, g.oob_hi = cfg.oob_lo, cfg.oob_hi
    g.w_reward[:] = [cfg.w_progress, cfg.w_bodyrate, cfg.w_action_rate, cfg.w_perception, cfg.w_success, cfg.w_bad_pose]
    g.add_cmd_noise = int(cfg.add_cmd_noise)
    g.cmd_noise_pos, g.cmd_noise_yaw = cfg.cmd_noise_pos, cfg.cmd_noise_yaw
    g.level_up_gates, g.level_down_gates = cfg.level_up_gates, cfg.level_down_gates
    g.noise_curriculum = int(cfg.noise_curriculum and cfg.add_cmd_noise)
    g.noise_up_gates, g.noise_down_gates = cfg.noise_up_gates, cfg.noise_down_gates
    g.noise_up, g.noise_down = 1.0 + cfg.noise_up, 1.0 - cfg.noise_down      # the factors torch.where() selects
    g.w_loss[:] = [cfg.w_loss_target, cfg.w_loss_vel, cfg.w_loss_fall]
    g.obs_vel_noise, g.obs_euler_noise = cfg.obs_vel_noise, cfg.obs_euler_noise
    if cfg.action_lag != 1:
        raise ValueError("only action_lag == 1 (QD/racing_ctbr_env.py:136) is built")
    return g


def pack_track_rows(table: GateTable) -> np.ndarray:
    """GateTable -> [types*levels*(G+1), 4] float32 rows (layout documented on GrTrack in include/gracing.h)."""
    t, l, g = table.num_types, table.num_levels, table.num_gates
    rows = np.zeros((t, l, g + 1, 4), dtype=np.float32)
    rows[:, :, 0, :3] = np.transpose(table.terrain_origins, (1, 0, 2))
    rows[:, :, 0, 3] = table.next_gate_id.astype(np.int32).view(np.float32)
    rows[:, :, 1:, :3] = table.gate_pose[..., :3]
    return rows.reshape(-1, 4)


def default_terrain_types(num_envs: int, num_types: int, env_id_offset: int = 0, global_num_envs: Optional[int] = None) -> torch.Tensor:
    """Isaac Lab TerrainImporter: ``torch.div(arange(N), N / num_cols, rounding_mode='floor')`` on the GLOBAL env ids."""
    n_glob = global_num_envs or num_envs
    ids = torch.arange(env_id_offset, env_id_offset + num_envs)
    return torch.div(ids, (n_glob / num_types), rounding_mode="floor").to(torch.int32).clamp_(max=num_types - 1)


class _Extras(dict):
    """extras dict whose "log" entry is built on demand from the device-side accumulators (no host sync on the
    hot loop: the reference builds it eagerly inside _reset_idx, manager_based_diff_rl_env.py:380-407).  With differentiable
    physics "log_losses" (:257, ``LossManager.log_all_active_terms``: [(term name, mean over envs)]) is built on demand as well,
    its values 0-dim DEVICE tensors: ``sum(v) / len(v)`` and loggers work on them, nothing waits for the GPU inside the loop."""

    def __init__(self, env):
        super().__init__()
        self._env = env

    def _lazy(self, k):
        return k == "log" or (k == "log_losses" and super().__contains__("loss_terms"))

    def __contains__(self, k):
        return self._lazy(k) or super().__contains__(k)

    def __getitem__(self, k):
        if k == "log" and not super().__contains__("log"):
            return self._env._build_log()
        if k == "log_losses" and super().__contains__("loss_terms"):
            m = super().__getitem__("loss_terms").detach().mean(dim=0)
            return [(name, m[i]) for i, name in enumerate(self._env._loss_term_names)]
        return super().__getitem__(k)

    def get(self, k, default=None):
        return self[k] if k in self else default


class RacingVecEnv:
    num_actions = L.NUM_ACTIONS
    num_obs = L.OBS_DIM
    num_privileged_obs = L.OBS_DIM

    def __init__(self, cfg: RacingCfg, table: GateTable, num_envs: int, device="cuda:0", seed: int = 42,
                 rng_mode: str = "philox", episode_stats: bool = True, env_id_offset: int = 0,
                 global_num_envs: Optional[int] = None, terrain_types: Optional[torch.Tensor] = None,
                 startup_rnd: Optional[torch.Tensor] = None, bptt_horizon: int = 0, block_threads: int = 0, pdl: Optional[bool] = None,
                 op_layer: Optional[bool] = None, _lib=None):
        self.cfg = cfg
        self.table = table
        self.num_envs = N = int(num_envs)
        self.device = torch.device(device)
        if _lib is None:
            if self.device.type != "cuda":
                raise RuntimeError("RacingVecEnv runs only on a CUDA device (sm_100a); there is no CPU fallback")
            _lib = B.load()
        self._lib = _lib
        if rng_mode not in ("philox", "dense"):
            raise ValueError("rng_mode must be 'philox' or 'dense'")
        self.rng_mode = rng_mode
        self.seed = int(seed)
        self.max_episode_length = cfg.max_episode_length
        self.step_dt = cfg.step_dt
        self._gcfg = make_gr_config(cfg)
        dev = self.device
        # ---- track
        self._rows = torch.from_numpy(pack_track_rows(table)).to(dev).contiguous()
        self._track = B.GrTrack(self._rows.data_ptr(), table.num_types, table.num_levels, table.num_gates)
        # ---- state planes
        self.num_planes = L.NUM_PLANES_WITH_STATS if episode_stats else L.NUM_PLANES
        self.num_tiles = (N + L.TILE - 1) // L.TILE
        self._stride = self.num_tiles * L.TILE               # env capacity
        # [tiles, 16 planes, 32 lanes, 4]: one contiguous 8 KB block per warp (layout.py)
        self.planes = torch.zeros(self.num_tiles, L.TILE_PLANES, L.TILE, 4, dtype=torch.float32, device=dev)
        if terrain_types is None:
            terrain_types = default_terrain_types(N, table.num_types, env_id_offset, global_num_envs)
        tt = terrain_types.to(torch.int32).cpu()
        if tt.numel() != N or (N > 1 and bool((tt[1:] < tt[:-1]).any())) or int(tt.min()) < 0 or int(tt.max()) >= table.num_types:
            raise ValueError("terrain_types must be N non-decreasing ids in [0, num_types)")
        # distinct types inside any 256-aligned env span bound the shared-memory slice of one thread block
        spans = [int(tt[s:s + 256].max() - tt[s:s + 256].min()) + 1 for s in range(0, N, 256)]
        self._terrain_types = tt.to(dev)
        self._chunk_types = torch.zeros(((N + 63) // 64) * 2, dtype=torch.int32, device=dev)
        if pdl is None:
            pdl = os.environ.get("GRACING_PDL", "1") != "0"
        # read-mostly planes before the grid dependency: "1" = into registers with the stale-flag protocol, "l2" = into L2 only, "0" = off
        pf = {"0": 0, "1": B.GR_LAUNCH_PREFETCH, "l2": B.GR_LAUNCH_PREFETCH_L2}[os.environ.get("GRACING_PREFETCH", "1").lower()]
        flags = (B.GR_LAUNCH_PDL | pf) if (pdl and self.device.type == "cuda") else 0
        # warp-cooperative Philox draws for resetting envs: measured on the B200 -0.25 .. -0.4 us per 65,536-env step at a 4.5 % reset
        # rate (gpurun r2g / r2h / r2i A/B pairs), -0.2 us per step in the window kernel; bit-identical draws (tests/test_philox_chain.py)
        if os.environ.get("GRACING_COOP_RESET", "1") == "1" and self.device.type == "cuda":
            flags |= B.GR_LAUNCH_COOP_RESET
        if os.environ.get("GRACING_EARLY_STORE", "1") == "1":       # measured on the B200: -0.06 .. -0.15 us per 65,536-env step
            flags |= B.GR_LAUNCH_EARLY_STORE
        self._launch_flags = flags
        self._state = B.GrState(self.planes.data_ptr(), self._stride, N, self.num_planes, int(env_id_offset), max(spans),
                                int(block_threads), flags, self._chunk_types.data_ptr())
        self._rng = B.GrRandom(None, self.seed, 0)
        self._step_count = 0
        # ---- outputs (ping-pong so that the tensors returned by step t stay valid during step t+1).  The kernels write them through raw
        #      pointers, which autograd's version counters cannot see: where a policy is differentiated across steps (differentiable
        #      physics), the policy observation is handed out as a fresh copy (_grad_safe_obs), as the reference's recomputed one is.
        def outs():
            return dict(obs=torch.zeros(N, L.OBS_DIM, device=dev), critic=torch.zeros(N, L.OBS_DIM, device=dev),
                        aux=torch.zeros(N, 1, device=dev), reward=torch.zeros(N, device=dev),
                        terminated=torch.zeros(N, dtype=torch.uint8, device=dev), time_out=torch.zeros(N, dtype=torch.uint8, device=dev),
                        dones=torch.zeros(N, dtype=torch.int64, device=dev), reward_terms=torch.zeros(N, L.NUM_REWARD_TERMS, device=dev),
                        gate_passed=torch.zeros(N, dtype=torch.uint8, device=dev))
        self._outs = [outs(), outs()]
        self._flip = 0
        self._log_accum = torch.zeros(B.GR_LOG_SHARDS, B.GR_LOG_SLOTS, device=dev)
        self._log_total = torch.zeros(B.GR_LOG_SLOTS, device=dev)
        self.extras = _Extras(self)
        self.export_reward_terms = False
        self.export_gate_passed = False
        # opt-in (differentiable physics): extras["aligned_states"] / ["nominal_states"] [N,13] and ["acc"] [N,3] of every step
        # (manager_based_diff_rl_env.py:205-212; detached values -- nothing outside the reference's own LossManager consumes them)
        self.export_aligned_states = False
        self._loss_term_names = L.LOSS_TERM_NAMES
        # opt-in: the STAGE-0 reward term collision_penalty_custom (QD/mdp/rewards.py:226-242, weight -50: QD/racing_ctbr_env.py:299-303)
        # against a terrain mesh -- see set_terrain_mesh()
        self._terrain_mesh = None
        self._w_collision = 0.0
        # ---- BPTT window (cfg.is_differentiable_physics)
        self._bptt = None
        if cfg.is_differentiable_physics:
            from .bptt import BpttWindow
            self._bptt = BpttWindow(self, bptt_horizon or 64)
        # ---- startup DR
        srnd_ptr = None
        if startup_rnd is not None:
            self._startup_rnd = startup_rnd.to(dev, torch.float32).contiguous()
            assert self._startup_rnd.shape == (N, L.SRND_STRIDE)
            srnd_ptr = self._startup_rnd.data_ptr()
        B.check(self._lib.gr_env_startup(C.byref(self._gcfg), C.byref(self._track), C.byref(self._state), self._terrain_types.data_ptr(),
                                         self._chunk_types.data_ptr(), srnd_ptr, self.seed, self._stream()), "gr_env_startup")
        self._last = self._outs[0]
        self._needs_reset = True
        self._ios = None
        self._pipe = None
        self._params_edited = False
        self._p_cfg, self._p_track, self._p_state = C.byref(self._gcfg), C.byref(self._track), C.byref(self._state)
        # ---- operator layer (ops.py): step / reset / BPTT autograd through torch.ops.gracing.* instead of direct ctypes calls
        if op_layer is None:
            op_layer = os.environ.get("GRACING_OP_LAYER", "0") == "1"
        self._op_handle = None
        self._ops = None
        if op_layer:
            from . import ops
            self._ops = ops
            ops.register_env(self)

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else None

    def _rand(self, rnd):
        if rnd is not None:
            rnd = rnd.to(self.device, torch.float32).contiguous()
            if rnd.shape != (self.num_envs, L.RND_STRIDE):
                raise ValueError(f"rnd must be [{self.num_envs}, {L.RND_STRIDE}]")
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

    def plane(self, pl: int) -> torch.Tensor:
        """Live [tiles, 32, 4] view of one state plane (env i = [i // 32, i % 32])."""
        return self.planes[:, pl]

    def read_plane(self, pl: int) -> torch.Tensor:
        """[N, 4] copy of one state plane."""
        return self.planes[:, pl].reshape(-1, 4)[: self.num_envs]

    def write_plane(self, pl: int, cols: slice, value: torch.Tensor):
        """Write value[N, k] into columns `cols` of a state plane."""
        v = value.to(self.device)
        buf = torch.zeros(self._stride, v.shape[-1], dtype=v.dtype, device=self.device)
        buf[: self.num_envs] = v
        if self._stride != self.num_envs:        # keep the padding lanes of the last tile untouched
            buf[self.num_envs:] = self.planes[:, pl].reshape(-1, 4)[self.num_envs:, cols].view(v.dtype) if v.dtype != torch.float32 \
                else self.planes[:, pl].reshape(-1, 4)[self.num_envs:, cols]
        if pl >= L.PL_DRAG2:
            self._state.launch_flags = self._launch_flags & ~B.GR_LAUNCH_PREFETCH      # one step without the pre-dependency prefetch
            self._params_edited = True
        if v.dtype == torch.float32:
            self.planes[:, pl, :, cols] = buf.view(self.num_tiles, L.TILE, -1)
        else:
            self.planes[:, pl, :, cols].view(v.dtype).copy_(buf.view(self.num_tiles, L.TILE, -1))

    @property
    def episode_length_buf(self) -> torch.Tensor:
        """int32 [N] copy of the per-env episode step counters; assign to write them (on_policy_runner.py:118-121)."""
        return self.planes[:, L.PL_LINVEL, :, 3].reshape(-1)[: self.num_envs].view(torch.int32) & L.EPLEN_MASK

    @episode_length_buf.setter
    def episode_length_buf(self, value: torch.Tensor):
        word = self.planes[:, L.PL_LINVEL, :, 3].view(torch.int32)                # the counter shares its word with flags (layout.py)
        v = torch.zeros(self._stride, dtype=torch.int32, device=self.device)
        v[: self.num_envs] = value.to(self.device, torch.int32).reshape(-1)
        if int(self.max_episode_length) > L.EPLEN_MASK:
            raise ValueError(f"max_episode_length must be <= {L.EPLEN_MASK}")
        word.copy_((word & ~L.EPLEN_MASK) | (v.view(self.num_tiles, L.TILE) & L.EPLEN_MASK))

    def state_dict_view(self) -> dict:
        """Named [N, ...] copies of the env state (diagnostics / tests); world-frame root state like robot.data."""
        R = self.read_plane
        q, pos, lin, ang, tq, aa, ff = R(L.PL_QUAT), R(L.PL_POS), R(L.PL_LINVEL), R(L.PL_ANGVEL), R(L.PL_TORQUE), R(L.PL_ANGACC), R(L.PL_FIFO)
        d2, d1, kp, kd, et, n0, n1 = R(L.PL_DRAG2), R(L.PL_DRAG1), R(L.PL_KP), R(L.PL_KD), R(L.PL_ETAU), R(L.PL_NOISE0), R(L.PL_NOISE1)
        pk = ang[:, 3].contiguous().view(torch.int32)
        ew = lin[:, 3].contiguous().view(torch.int32)
        field = (ew >> L.EPLEN_METRIC_SHIFT) & L.EPLEN_METRIC_MASK
        return {
            "root_quat_w": q, "root_pos_w": pos[:, :3], "gross_thrust": pos[:, 3],
            "root_lin_vel_w": lin[:, :3], "episode_length": ew & L.EPLEN_MASK,
            "cross_obs": (ew & L.EPLEN_AUX_BIT) != 0,
            "metric_action_rate": torch.where(field == 0, torch.zeros_like(field), (field + L.EPLEN_METRIC_BIAS) << 12).view(torch.float32),
            "root_ang_vel_b": ang[:, :3], "torque": tq[:, :3], "ang_acc_b": aa[:, :3],
            "action_fifo_tanh": ff, "drag_coeffs": d2[:, :3], "mass": d2[:, 3],
            "h_force_drag_coeffs": d1[:, :3], "exp_thrust_delay": d1[:, 3],
            "rate_gain_p": kp[:, :3], "thr_est_error": kp[:, 3], "rate_gain_d": kd[:, :3],
            "exp_torque_delay": et[:, :3],
            "gate_noise": torch.cat([n0[:, :3], n0[:, 3:], n1[:, :2]], dim=-1),
            "noise_pos_hi": n1[:, 2], "noise_level": n1[:, 3],
            "episode_sums": torch.cat([R(L.PL_EPSUM0), tq[:, 3:], aa[:, 3:]], dim=-1),
            "loss_episode_sums": R(L.PL_LOSSSUM)[:, :3],
            "gate_id": (pk >> L.PK_GATE_SHIFT) & 0xFF, "accumulate_gates": (pk >> L.PK_ACC_SHIFT) & 0xFFF,
            "terrain_levels": (pk >> L.PK_LEVEL_SHIFT) & 0x3F, "terrain_types": (pk >> L.PK_TYPE_SHIFT) & 0x1F,
            "fresh": (pk >> L.PK_FRESH_SHIFT) & 0x1,
        }

    def _obs_dict(self, o):
        return {"policy": o["obs"], "critic": o["critic"], "auxiliary": o["aux"]}

    def _build_log(self) -> dict:
        """extras["log"] of the reference (_reset_idx, manager_based_diff_rl_env.py:380-407): means over the envs reset
        since the last read, as 0-dim device tensors (the runner's logger accepts tensors)."""
        acc = self._log_accum.sum(dim=0)
        n = acc[B_LOG_NUM_RESET].clamp(min=1.0)
        log = {}
        if self.num_planes == L.NUM_PLANES_WITH_STATS:
            w = [self.cfg.w_progress, self.cfg.w_bodyrate, self.cfg.w_action_rate, self.cfg.w_perception, self.cfg.w_success, self.cfg.w_bad_pose]
            for k, name in enumerate(L.REWARD_TERM_NAMES):
                if w[k] != 0.0:
                    log["Episode_Reward/" + name] = acc[B_LOG_SUM_EPSUM + k] / n / self.cfg.episode_length_s
            # LossManager.reset (L/managers/loss_manager.py:71-78): logged in every mode; the sums only move with differentiable physics
            for k, name in enumerate(L.LOSS_TERM_NAMES):
                log["Episode_Loss/" + name] = acc[B.GR_LOG_SUM_LOSS + k] / n / self.cfg.episode_length_s
            # CommandTerm.reset: every metric of RacingCommand (QD/mdp/commands.py:257-260) as the mean over the reset envs
            log["Metrics/next_gate_pose/action_rate"] = acc[B.GR_LOG_SUM_ACTION_RATE] / n
            log["Metrics/next_gate_pose/avg_lin_spd"] = acc[B.GR_LOG_SUM_LIN_SPD] / n
            log["Metrics/next_gate_pose/avg_ang_spd"] = acc[B.GR_LOG_SUM_ANG_SPD] / n
        log["Metrics/next_gate_pose/accumulate_gates"] = acc[B_LOG_SUM_GATES] / n
        log["Episode_Termination/time_out"] = acc[B_LOG_NUM_TIMEOUT].clone()
        log["Episode_Termination/terminated"] = acc[B_LOG_NUM_TERMINATED].clone()
        sv = self.state_dict_view()
        log["Curriculum/terrain_levels"] = sv["terrain_levels"].float().mean()
        if self.cfg.noise_curriculum and self.cfg.add_cmd_noise:
            log["Curriculum/command_noise_level"] = sv["noise_level"].mean()
        self._log_total += acc
        self._log_accum = torch.zeros_like(self._log_accum)
        return log

    # ------------------------------------------------------------------ API
    def reset(self, rnd: Optional[torch.Tensor] = None):
        """ManagerBasedRLEnv.reset(): reset every env, return (obs, extras)."""
        if self._ops is not None:
            return self._reset_ops(rnd)
        o = self._outs[self._flip]
        self._flip ^= 1
        B.check(self._lib.gr_env_reset(C.byref(self._gcfg), C.byref(self._track), C.byref(self._state), C.byref(self._rand(rnd)), None,
                                       o["obs"].data_ptr(), o["critic"].data_ptr(), o["aux"].data_ptr(), self._stream()), "gr_env_reset")
        self._last = o
        self._needs_reset = False
        if self._bptt is not None:
            self._bptt.start_window()
        self.extras["observations"] = self._obs_dict(o)
        return self._grad_safe_obs(o["obs"]), self.extras

    def _grad_safe_obs(self, obs: torch.Tensor, needed: bool = True) -> torch.Tensor:
        """A policy differentiated over a BPTT window saves its input for the weight gradients; the ping-pong output buffer
        would be overwritten two steps later behind autograd's back."""
        return obs.clone() if (self._bptt is not None and needed) else obs

    def get_observations(self, fresh_noise: bool = False, rnd: Optional[torch.Tensor] = None):
        """RslRlVecEnvWrapper.get_observations: (policy obs, {"observations": obs_dict}).  By default the observation of the last
        reset / step is handed out again.  ``fresh_noise=True`` is the reference's literal behaviour (``ObservationManager.compute()``,
        manager_based_diff_rl_env.py:264 / QD/mdp/observation.py:22-63): the observations are recomputed from the current state with a
        NEW observation-noise draw (a Philox stream no step uses -- counter word with the top bit set -- so the env's step stream
        does not move; ``rnd`` [N,52] in dense mode), the last-action columns still showing the lagged action the last step applied."""
        if self._needs_reset:
            self.reset()
        o = self._last
        if fresh_noise or getattr(self, "_host_stale", False):
            self._host_stale = False
            return self._observe_fresh(rnd)
        if self._ops is not None and o["obs"].is_inference() and not torch.is_inference_mode_enabled():
            # operator outputs allocated under the runner's torch.inference_mode() (on_policy_runner.py:141): hand normal tensors
            # to a caller outside it, as the reference's recomputed observations are
            o.update({k: o[k].clone() for k in ("obs", "critic", "aux")})
        return self._grad_safe_obs(o["obs"], self._ops is None), {"observations": self._obs_dict(o)}

    def _observe_fresh(self, rnd):
        last = self._last
        k = self._flip
        o = self._outs[k] if self._ops is None else dict(obs=torch.zeros_like(last["obs"]), critic=torch.zeros_like(last["critic"]), aux=torch.zeros_like(last["aux"]))
        if self._ops is None:
            self._flip = k ^ 1
        self._observe_count = getattr(self, "_observe_count", 0) + 1
        rng = B.GrRandom(None, self.seed, 0x80000000 | ((self._step_count * 977 + self._observe_count) & 0x7FFFFFFF))       # a stream no step uses
        if rnd is not None:
            rnd = rnd.to(self.device, torch.float32).contiguous()
            if rnd.shape != (self.num_envs, L.RND_STRIDE):
                raise ValueError(f"rnd must be [{self.num_envs}, {L.RND_STRIDE}]")
            rng.rnd = rnd.data_ptr()
        elif self.rng_mode == "dense":
            raise ValueError("rng_mode='dense' needs an explicit rnd tensor for a fresh noise draw")
        B.check(self._lib.gr_env_observe(C.byref(self._gcfg), C.byref(self._track), C.byref(self._state), C.byref(rng),
                                         o["obs"].data_ptr(), o["critic"].data_ptr(), o["aux"].data_ptr(), self._stream()), "gr_env_observe")
        # modified_last_action (QD/mdp/observation.py:55-63) shows raw_actions = the lagged action of the last process_actions call;
        # the FIFO the kernel reads already holds a_t, so those four columns come from the last step's own observation
        o["obs"][:, 12:].copy_(last["obs"][:, 12:])
        o["critic"][:, 12:].copy_(last["critic"][:, 12:])
        for name in ("reward", "terminated", "time_out", "dones"):
            if name in last and o is not last and name in o:
                o[name].copy_(last[name])
        self._last = o
        self.extras["observations"] = self._obs_dict(o)
        return self._grad_safe_obs(o["obs"]), {"observations": self._obs_dict(o)}

    def detach(self):
        """env.unwrapped.detach() (manager_based_diff_rl_env.py:412-416): start a new BPTT window."""
        if self._bptt is not None:
            self._bptt.start_window()

    def close(self):
        if self._pipe is not None:
            self._lib.gr_host_pipe_destroy(self._pipe)
            self._pipe = None

    def set_terrain_mesh(self, mesh, weight: float = -50.0):
        """Add the reference's mesh-collision reward term to ``step``: ``reward += weight * dt * (lattice points inside the mesh > 2)``
        (QD/mdp/rewards.py:226-242; weight -50 in STAGE 0, QD/racing_ctbr_env.py:299-303), evaluated like every reward term on the
        pose after the step's physics and BEFORE any reset (the step kernel exports that pose, ``GrStepIO.pre_reset_pos / _quat``;
        the ray casts are one more launch, ``gr_uav_collision_ray``).  ``mesh``: a :class:`..mesh.TerrainMesh` in the world frame of
        ``terrain_origins``; None switches the term off.  The term's value of the last step is ``extras["collision_penalty"]``; it is not
        part of the in-kernel episode sums (``Episode_Reward/*``)."""
        self._terrain_mesh, self._w_collision = mesh, float(weight)
        self._ios = None
        if mesh is not None and not hasattr(self, "_pre_pose"):
            self._pre_pose = (torch.zeros(self.num_envs, 3, device=self.device), torch.zeros(self.num_envs, 4, device=self.device))

    def _make_io(self, o) -> B.GrStepIO:
        io = B.GrStepIO()
        io.obs, io.critic_obs, io.aux_obs = o["obs"].data_ptr(), o["critic"].data_ptr(), o["aux"].data_ptr()
        io.reward, io.terminated, io.time_out, io.dones = o["reward"].data_ptr(), o["terminated"].data_ptr(), o["time_out"].data_ptr(), o["dones"].data_ptr()
        if self._terrain_mesh is not None:
            io.pre_reset_pos, io.pre_reset_quat = self._pre_pose[0].data_ptr(), self._pre_pose[1].data_ptr()
        return io

    def step(self, actions: torch.Tensor, rnd: Optional[torch.Tensor] = None):
        if self._needs_reset:
            self.reset()
        if self._ops is not None:
            return self._step_ops(actions, rnd)
        k = self._flip
        o = self._outs[k]
        self._flip = k ^ 1
        act = actions.detach() if actions.requires_grad else actions
        if act.dtype != torch.float32 or not act.is_contiguous() or act.device != self.device:
            act = act.to(self.device, torch.float32).contiguous()
        if act.shape != (self.num_envs, L.NUM_ACTIONS):
            raise ValueError(f"Invalid action shape, expected: ({self.num_envs}, {L.NUM_ACTIONS}), received: {tuple(act.shape)}.")
        ios = self._ios
        if ios is None:                       # the argument structs are built once per output set; only pointers that move are patched
            ios = self._ios = [self._make_io(self._outs[0]), self._make_io(self._outs[1])]
            self._views = [(x["time_out"].view(torch.bool), x["terminated"].view(torch.bool), self._obs_dict(x)) for x in self._outs]
        io = ios[k]
        io.action = act.data_ptr()
        io.reward_terms = o["reward_terms"].data_ptr() if self.export_reward_terms else None
        io.gate_passed = o["gate_passed"].data_ptr() if self.export_gate_passed else None
        io.log_accum = self._log_accum.data_ptr()
        if self._bptt is not None:
            self._bptt.bind_step(io)
            w_reward, w_dones = self._bptt.bind_outputs(io)
            if self.export_aligned_states:
                if "aligned" not in o:
                    o["aligned"], o["acc"] = torch.zeros(self.num_envs, 13, device=self.device), torch.zeros(self.num_envs, 3, device=self.device)
                io.aligned_states, io.acc = o["aligned"].data_ptr(), o["acc"].data_ptr()
        rc = self._lib.gr_step_fwd(self._p_cfg, self._p_track, self._p_state, C.byref(self._rand(rnd)), C.byref(io), self._stream())
        if rc:
            B.check(rc, "gr_step_fwd")
        if self._params_edited:
            self._state.launch_flags = self._launch_flags
            self._params_edited = False
        self._last = o
        ex = self.extras
        if self._terrain_mesh is not None:
            from .mesh import collision_penalty_custom
            pen = collision_penalty_custom(self._terrain_mesh, self._pre_pose[0], self._pre_pose[1])
            (w_reward if self._bptt is not None else o["reward"]).add_(pen, alpha=self._w_collision * self.step_dt)
            ex["collision_penalty"] = pen
        dict.pop(ex, "log", None)
        to, term, obs_dict = self._views[k]
        ex["observations"] = obs_dict
        ex["time_outs"] = to
        ex["terminated"] = term
        if self._bptt is not None:
            self._bptt.after_step(actions, ex)
            if self.export_aligned_states:
                ex["aligned_states"] = ex["nominal_states"] = o["aligned"]
                ex["acc"] = o["acc"]
            # a differentiable-physics loop keeps what a step returns for the whole window (naive_train.py:170-172 appends dones and
            # losses to lists): the kernel wrote reward / dones into this step's rows of the window (BpttWindow.bind_outputs), which no
            # later step of the window overwrites -- no copy launches on this path
            self._last = dict(o, reward=w_reward, dones=w_dones)
            return self._grad_safe_obs(o["obs"], actions.requires_grad), w_reward, w_dones, ex
        return o["obs"], o["reward"], o["dones"], ex

    def rollout(self, actions: torch.Tensor, rnd: Optional[torch.Tensor] = None, record_obs: bool = False) -> dict:
        """``for t in range(T): env.step(actions[t])`` in ONE launch (gr_rollout_fwd) for actions known in advance
        (``actions`` [T,N,4]; dense mode: ``rnd`` [T,N,52]): the env state stays in registers over the window.  Bit-identical
        to the T single steps.  Returns ``{"obs" [N,16] (after the last step), "reward" [T,N], "dones" / "terminated" /
        "time_outs" [T,N] (bool), "obs_seq" [T,N,16] if record_obs}`` and, with differentiable physics, ``"losses"`` [T,N] /
        ``"loss_terms"`` [T,N,3] while the window's tape is extended by T steps (``env._bptt.backward_window()`` then returns
        dL/d(actions) for the whole window).  ``extras`` / ``get_observations()`` afterwards are those of the last step."""
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
            if tuple(rnd.shape) != (T, N, L.RND_STRIDE):
                raise ValueError(f"rnd must be [{T}, {N}, {L.RND_STRIDE}]")
            rng.rnd = rnd.data_ptr()
        elif self.rng_mode == "dense":
            raise ValueError("rng_mode='dense' needs an explicit rnd tensor every call")
        step0 = self._step_count & 0xFFFFFFFF
        self._step_count += T
        if self._ops is not None:
            return self._rollout_ops(act, rnd, step0, record_obs)
        k = self._flip
        o = self._outs[k]
        self._flip = k ^ 1
        out = {"reward": torch.empty(T, N, device=dev), "dones": torch.empty(T, N, dtype=torch.uint8, device=dev),
               "terminated": torch.empty(T, N, dtype=torch.uint8, device=dev), "time_outs": torch.empty(T, N, dtype=torch.uint8, device=dev)}
        io = B.GrRolloutIO()
        io.actions, io.T = act.data_ptr(), T
        io.obs_out, io.critic_obs_out, io.aux_out = o["obs"].data_ptr(), o["critic"].data_ptr(), o["aux"].data_ptr()
        io.reward, io.dones, io.terminated, io.time_out = (out[n].data_ptr() for n in ("reward", "dones", "terminated", "time_outs"))
        if record_obs:
            out["obs_seq"] = torch.empty(T, N, L.OBS_DIM, device=dev)
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
        B.check(self._lib.gr_rollout_fwd(self._p_cfg, self._p_track, self._p_state, C.byref(rng), C.byref(io), self._stream()), "gr_rollout_fwd")
        # the window kernel rewrites read-mostly planes without the per-step flags a pre-dependency prefetch relies on
        self._state.launch_flags = self._launch_flags & ~B.GR_LAUNCH_PREFETCH
        self._params_edited = True
        self._last = o
        ex = self.extras
        dict.pop(ex, "log", None)
        for n in ("dones", "terminated", "time_outs"):
            out[n] = out[n].view(torch.bool)
        ex["observations"] = self._obs_dict(o)
        ex["time_outs"], ex["terminated"] = out["time_outs"][-1], out["terminated"][-1]
        out["obs"] = o["obs"]
        if win is not None:
            out["losses"], out["loss_terms"] = win.loss[t0:t0 + T], win.loss_terms[t0:t0 + T]
            win.t += T
        return out

    def _rollout_ops(self, act, rnd, step0, record_obs):
        win, T = self._bptt, act.shape[0]
        if win is None:
            r = torch.ops.gracing.rollout_fwd(self._op_handle, self.planes, act, rnd, step0, self._log_accum, record_obs)
        else:
            r = torch.ops.gracing.rollout_fwd_tape(self._op_handle, self.planes, act, rnd, step0, self._log_accum, win.tape, win.t, record_obs)
        obs, critic, aux, reward, dones, terminated, time_out, obs_seq = r[:8]
        self._state.launch_flags = self._launch_flags & ~B.GR_LAUNCH_PREFETCH
        self._params_edited = True
        self._last = o = dict(obs=obs, critic=critic, aux=aux)
        out = {"reward": reward, "dones": dones.view(torch.bool), "terminated": terminated.view(torch.bool), "time_outs": time_out.view(torch.bool), "obs": obs}
        if record_obs:
            out["obs_seq"] = obs_seq
        ex = self.extras
        dict.pop(ex, "log", None)
        ex["observations"] = self._obs_dict(o)
        ex["time_outs"], ex["terminated"] = out["time_outs"][-1], out["terminated"][-1]
        if win is not None:
            t0 = win.t
            win.loss[t0:t0 + T].copy_(r[8])
            win.loss_terms[t0:t0 + T].copy_(r[9])
            out["losses"], out["loss_terms"] = win.loss[t0:t0 + T], win.loss_terms[t0:t0 + T]
            win.t += T
        return out

    # ------------------------------------------------------------------ the same two calls through torch.ops.gracing.* (ops.py)
    def _rnd_step(self, rnd):
        if rnd is not None:
            rnd = rnd.to(self.device, torch.float32).contiguous()
        elif self.rng_mode == "dense":
            raise ValueError("rng_mode='dense' needs an explicit rnd tensor every call")
        step = self._step_count & 0xFFFFFFFF
        self._step_count += 1
        return rnd, step

    def _reset_ops(self, rnd):
        rnd, step = self._rnd_step(rnd)
        obs, critic, aux = torch.ops.gracing.reset(self._op_handle, self.planes, None, rnd, step)
        self._last = o = dict(obs=obs, critic=critic, aux=aux)
        self._needs_reset = False
        if self._bptt is not None:
            self._bptt.start_window()
        self.extras["observations"] = self._obs_dict(o)
        return obs, self.extras

    def _step_ops(self, actions, rnd):
        win = self._bptt
        act = actions.detach()
        if act.dtype != torch.float32 or not act.is_contiguous() or act.device != self.device:
            act = act.to(self.device, torch.float32).contiguous()
        rnd, step = self._rnd_step(rnd)
        export = bool(self.export_reward_terms or self.export_gate_passed)
        if win is None:
            out = torch.ops.gracing.step_fwd(self._op_handle, self.planes, act, rnd, step, self._log_accum, export)
        else:
            out = torch.ops.gracing.step_fwd_tape(self._op_handle, self.planes, act, rnd, step, self._log_accum, win.tape, win.t, export)
        obs, critic, aux, reward, terminated, time_out, dones, terms, passed = out[:9]
        if self._params_edited:
            self._state.launch_flags = self._launch_flags
            self._params_edited = False
        self._last = o = dict(obs=obs, critic=critic, aux=aux, reward=reward, terminated=terminated, time_out=time_out, dones=dones,
                              reward_terms=terms, gate_passed=passed)
        ex = self.extras
        dict.pop(ex, "log", None)
        ex["observations"] = self._obs_dict(o)
        ex["time_outs"] = time_out.view(torch.bool)
        ex["terminated"] = terminated.view(torch.bool)
        if win is not None:
            loss, loss_terms = out[9:]
            win.loss[win.t].copy_(loss)
            win.loss_terms[win.t].copy_(loss_terms)
            if win.autograd and actions.requires_grad:
                if win._token is None:
                    win._token = torch.zeros(1, device=self.device, requires_grad=True)
                loss, win._token = torch.ops.gracing.step_loss(self._op_handle, actions, win._token, loss, win.t, win.epoch)
            win.t += 1
            ex["losses"] = loss
            ex["losses_detached"] = win.losses_detached
            ex["loss_terms"] = loss_terms
        return obs, reward, dones, ex

    # ------------------------------------------------------------------ host-buffer API (gr_host_pipe_*, include/gracing.h)
    def host_buffers(self, sets: int = 1):
        """``sets`` sets of pinned host tensors for :meth:`step_host`: ``{"actions" [N,4], "obs" [N,16], "reward" [N], "dones" [N] int64}``.
        obs | reward | dones of a set are views of ONE pinned block, back to back -- the layout of the pipe's device slots -- so that a step's
        results cross PCIe as one device->host copy instead of three (measured on the B200 box: 108.9 -> 101.1 us per 65,536-env step)."""
        N = self.num_envs
        out = []
        for _ in range(sets):
            block = torch.empty(N * 76 + 8, dtype=torch.uint8).pin_memory()
            d0 = (N * 68 + 7) & ~7
            out.append({"actions": torch.zeros(N, L.NUM_ACTIONS).pin_memory(), "obs": block[: N * 64].view(torch.float32).view(N, L.OBS_DIM),
                        "reward": block[N * 64: N * 68].view(torch.float32), "dones": block[d0: d0 + N * 8].view(torch.int64), "_block": block})
        return out

    def step_host(self, actions: torch.Tensor, obs: torch.Tensor, reward: torch.Tensor, dones: Optional[torch.Tensor] = None,
                  critic_obs: Optional[torch.Tensor] = None, time_outs: Optional[torch.Tensor] = None, depth: int = 2) -> int:
        """env.step() for a caller whose tensors live in (pinned) HOST memory: enqueue H2D(actions) -> step kernel ->
        D2H(obs, reward, dones[, critic_obs, time_outs]) and return a ticket at once; the output tensors hold the step's
        results after ``wait_host(ticket)``.  Up to ``depth`` steps are in flight (copies overlap the next kernel).  ``dones`` may be an
        int64 tensor (the wrapper's ``.long()``) or a uint8 / bool tensor (one byte per env across PCIe).  With obs, reward and int64
        dones back to back in host memory (:meth:`host_buffers`) the three results travel as ONE copy.

        Rules of the asynchronous call: ``actions`` must not be rewritten before ``wait_host(ticket)`` returned for this step (the
        host->device copy reads it later); all host steps of an env use the ``depth`` and the CUDA stream of the first one (the pipe
        is created there: a different value raises); ``get_observations()`` after host steps recomputes the observation from the
        state, because the last observation lives in the caller's host buffer."""
        if self._needs_reset:
            self.reset()
        if self._bptt is not None:
            raise RuntimeError("step_host is a forward-only path; BPTT windows need device tensors (step)")
        if self._pipe is None:
            pipe = C.c_void_p()
            B.check(self._lib.gr_host_pipe_create(self.num_envs, int(depth), self._stream(), C.byref(pipe)), "gr_host_pipe_create")
            self._pipe = pipe
            self._pipe_key = (int(depth), self._stream())
        elif self._pipe_key != (int(depth), self._stream()):
            raise RuntimeError(f"step_host: the host pipe of this env was created with (depth, stream) = {self._pipe_key}; call close() before "
                               f"changing them (got {(int(depth), self._stream())})")
        dones_u8 = None
        if dones is not None and dones.dtype in (torch.uint8, torch.bool):        # one byte per env over PCIe instead of the int64 of `.long()`
            dones, dones_u8 = None, dones
        for t, shape, dt in ((actions, (self.num_envs, 4), torch.float32), (obs, (self.num_envs, L.OBS_DIM), torch.float32),
                             (reward, (self.num_envs,), torch.float32), (dones, (self.num_envs,), torch.int64),
                             (critic_obs, (self.num_envs, L.OBS_DIM), torch.float32), (time_outs, (self.num_envs,), torch.bool),
                             (dones_u8, (self.num_envs,), None)):
            if t is not None and (t.device.type != "cpu" or tuple(t.shape) != shape or (dt is not None and t.dtype != dt) or not t.is_contiguous()):
                raise ValueError(f"step_host: expected a contiguous host tensor of shape {shape} and dtype {dt or 'uint8 / bool'}")
        # obs | reward | dones as views of ONE host allocation, back to back (host_buffers()): one device->host copy instead of three
        one_block = (dones is not None and obs.untyped_storage().data_ptr() == reward.untyped_storage().data_ptr() == dones.untyped_storage().data_ptr()
                     and reward.data_ptr() == obs.data_ptr() + self.num_envs * 64 and dones.data_ptr() == obs.data_ptr() + self.num_envs * 68)
        hs = B.GrHostStep(actions.data_ptr(), obs.data_ptr(), reward.data_ptr(), B.ptr(dones), B.ptr(critic_obs), B.ptr(time_outs), B.ptr(dones_u8),
                          int(one_block))
        ticket = C.c_int64()
        B.check(self._lib.gr_host_pipe_step(self._pipe, self._p_cfg, self._p_track, self._p_state, C.byref(self._rand(None)), C.byref(hs),
                                            self._log_accum.data_ptr(), C.byref(ticket)), "gr_host_pipe_step")
        if self._params_edited:
            self._state.launch_flags = self._launch_flags
            self._params_edited = False
        self._host_stale = True                    # the device-side copy of "the last observation" is no longer the env's last one
        return ticket.value

    def wait_host(self, ticket: int) -> None:
        B.check(self._lib.gr_host_pipe_wait(self._pipe, int(ticket)), "gr_host_pipe_wait")

    def __del__(self):
        pipe = getattr(self, "_pipe", None)
        if pipe is not None:
            try:
                self._lib.gr_host_pipe_destroy(pipe)
            except Exception:
                pass


B_LOG_NUM_RESET, B_LOG_SUM_GATES, B_LOG_SUM_EPSUM, B_LOG_NUM_TIMEOUT, B_LOG_NUM_TERMINATED = 0, 1, 2, 8, 9


def make_env(task: str = "DiffLab-Quadcopter-CTBR-Racing-v0", num_envs: int = 2048, device="cuda:0", stage=None, track="complex",
             differentiable: bool = False, **kwargs) -> RacingVecEnv:
    """``gym.make(task, cfg=env_cfg)`` + ``RslRlVecEnvWrapper`` of the reference launch scripts (standalone/rsl_rl/train.py:102-120)
    for the one registered racing task (QD/__init__.py:48-73).  ``track``: "complex" = the 20x10 curriculum table of
    RacingComplexTerrainCfg built by the restated family generators (track_gen.py; the reference's own seed-42 table: its obstacle
    draws are replayed, no obstacle geometry exists), "synthetic" = the
    simplified centre-line table of tracks.py, "figure8" = RacingTestTerrainCfg.  Under torchrun the envs are sharded."""
    from . import dist_utils as D
    from .tracks import figure_eight_track, synthetic_track_table
    if task not in ("DiffLab-Quadcopter-CTBR-Racing-v0", "DiffLab-Quadcopter-CTBR-Racing-Play-v0"):
        raise ValueError(f"unknown task {task!r}: only the CTBR racing task is built")
    cfg = RacingCfg.from_env() if stage is None else RacingCfg.for_stage(stage)
    if differentiable:
        cfg.is_differentiable_physics = True
    if track == "figure8":
        table = figure_eight_track()
    elif track == "synthetic":
        table = synthetic_track_table()
    elif track == "complex":
        from .track_gen import generate_track_table, racing_complex_cfg
        table = generate_track_table(racing_complex_cfg(), name="RacingComplexTerrainCfg")
    else:
        raise ValueError(f"unknown track {track!r}")
    rank, world = D.world()
    return RacingVecEnv(cfg, table, num_envs, device=device, env_id_offset=rank * num_envs, global_num_envs=world * num_envs, **kwargs)

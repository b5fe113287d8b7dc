"""Task constants of the racing hot path (host-side mirror of the reference cfg).

Every number here is a restatement of a reference configuration value; the
citation next to it is relative to /root/reference.  QD = extensions/
diff.lab_tasks/diff/lab_tasks/tasks/quadcopter_diff, L = extensions/diff.lab/
diff/lab.

The reference selects reward weights / terminations / command noise at import
time through the TRAINING_STAGE environment variable (QD/racing_ctbr_env.py:39);
here the stage is an explicit constructor argument and `RacingCfg.from_env()`
keeps the environment-variable behaviour.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field, asdict
from typing import Tuple


# thrust map and motor speed range: L/controllers/controller_diff_cfg.py:28-33
_THRUSTMAP = (1.3298253500372892e-06, 0.0038360810526746033, -1.7689986848125325)
_MOTOR_OMEGA = (150.0, 3000.0)


def _rotor_thrust(omega: float) -> float:
    # L/controllers/controller_diff.py:96-97
    return _THRUSTMAP[0] * omega ** 2 + _THRUSTMAP[1] * omega + _THRUSTMAP[2]


@dataclass
class RacingCfg:
    """Frozen constants of DiffLab-Quadcopter-CTBR-Racing (PhysX-free closure)."""

    stage: int = 1                      # QD/racing_ctbr_env.py:39 (default TRAINING_STAGE=1)
    # --- time ----------------------------------------------------------------
    sim_dt: float = 0.01                # QD/racing_ctbr_env.py:380
    decimation: int = 3                 # QD/racing_ctbr_env.py:375
    episode_length_s: float = 6.0       # QD/racing_ctbr_env.py:376 (8.0 for stage 2)
    # --- drone -----------------------------------------------------------------
    mass: float = 0.8                   # NOT in the reference repo (read from the USD through PhysX,
                                        # QD/mdp/diff_action.py:55); synthetic value, stated in every report
    inertia_diag: Tuple[float, float, float] = (0.0015, 0.002, 0.004)   # QD/mdp/diff_action.py:59
    gravity: float = 9.81               # QD/mdp/dynamics/dynamics.yaml:12
    grad_decay: float = 0.92            # dynamics.yaml:1
    drag_1: float = 0.18                # dynamics.yaml:4 (linear "h_force" drag, x mass)
    drag_1_rand: float = 0.1            # dynamics.yaml:5
    drag_2: float = 0.01                # dynamics.yaml:7 (quadratic drag, x mass)
    drag_2_rand: float = 0.005          # dynamics.yaml:8
    z_drag: float = 4.0                 # dynamics.yaml:10
    z_drag_rand: float = 0.4            # dynamics.yaml:11
    random_drag: bool = True            # QD/racing_ctbr_env.py:135
    # --- action term / controller ----------------------------------------------
    max_thrust_weight_ratio: float = 3.0   # QD/mdp/diff_action_cfg.py:47
    body_rate_bound: float = 6.0        # QD/racing_ctbr_env.py:131
    rate_gain_p: Tuple[float, float, float] = (35.0, 35.0, 35.0)          # :128
    rate_gain_d: Tuple[float, float, float] = (0.0005, 0.0005, 0.0003)    # :130
    thrust_ctrl_delay: float = 0.03     # :132
    torque_ctrl_delay: Tuple[float, float, float] = (0.03, 0.03, 0.03)    # :133
    action_lag: int = 1                 # :136 (only lag 1 is built)
    thr_est_error_init_std: float = 0.02    # QD/mdp/diff_action.py:86
    thr_est_error_reset_std: float = 0.01   # QD/mdp/diff_action.py:233
    # startup DR (QD/racing_ctbr_env.py:211-219)
    pid_scale: Tuple[float, float] = (0.9, 1.1)
    delay_scale: Tuple[float, float] = (0.8, 1.3)
    # --- command ----------------------------------------------------------------
    update_threshold: float = 0.35      # QD/racing_ctbr_env.py:120
    cmd_noise_pos: float = 0.1          # :105-107 (0.5 for stage 2, :112-114); symmetric +-
    cmd_noise_yaw: float = 0.1          # :110 (dead data, kept for the random stream only)
    add_cmd_noise: bool = True          # :119  (STAGE != 0)
    # --- reset sampler (QD/racing_ctbr_env.py:177-197) ---------------------------
    default_root_pos: Tuple[float, float, float] = (0.0, 0.0, 0.5)  # diff.lab_assets quadcopter.py:34
    reset_pos: float = 0.5
    reset_roll_pitch: float = 0.2
    reset_yaw: float = 0.7
    reset_vel: float = 0.1
    # --- terminations (QD/racing_ctbr_env.py:247-260) -----------------------------
    term_out_of_bound: bool = False     # stage 0 only, bounds (0, 10)
    oob_lo: float = 0.0
    oob_hi: float = 10.0
    term_bad_pose: bool = True          # stage >= 1
    # --- rewards (QD/racing_ctbr_env.py:280-328) -----------------------------------
    w_progress: float = 1.0
    w_bodyrate: float = -0.1
    w_action_rate: float = -0.05
    w_perception: float = 0.1
    w_success: float = 20.0
    w_bad_pose: float = -30.0           # stage 1 only (0 => term absent)
    # --- curricula (QD/racing_ctbr_env.py:264-278) -----------------------------------
    level_up_gates: int = 3
    level_down_gates: int = 2
    noise_curriculum: bool = True       # stage 1 only
    noise_up_gates: int = 4
    noise_down_gates: int = 3
    noise_up: float = 0.02
    noise_down: float = 0.03
    max_init_terrain_level: int = 5     # QD/racing_ctbr_env.py:47
    # --- BPTT losses (QD/racing_ctbr_env.py:330-353) ------------------------------------
    is_differentiable_physics: bool = False   # :372 (flip to enable extras["losses"])
    w_loss_target: float = 1.0
    w_loss_vel: float = 0.05
    w_loss_fall: float = 0.5
    # --- observation noise (QD/mdp/observation.py:27,52) -----------------------------
    obs_vel_noise: float = 0.03
    obs_euler_noise: float = 0.05

    # ---------------------------------------------------------------------------
    @staticmethod
    def for_stage(stage: int, **overrides) -> "RacingCfg":
        """Constants the reference selects for TRAINING_STAGE = 0 | 1 | 2."""
        if stage not in (0, 1, 2):
            raise ValueError(f"TRAINING_STAGE must be 0, 1 or 2, got {stage}")
        kw = dict(stage=stage)
        if stage == 0:
            kw.update(add_cmd_noise=False, term_out_of_bound=True, term_bad_pose=False,
                      w_bodyrate=-0.02, w_action_rate=-0.01, w_success=10.0, w_bad_pose=0.0,
                      noise_curriculum=False)
        elif stage == 2:
            kw.update(episode_length_s=8.0, cmd_noise_pos=0.5, cmd_noise_yaw=0.5,
                      w_bad_pose=0.0, noise_curriculum=False)
        kw.update(overrides)
        return RacingCfg(**kw)

    @staticmethod
    def from_env(**overrides) -> "RacingCfg":
        return RacingCfg.for_stage(int(os.environ.get("TRAINING_STAGE", 1)), **overrides)

    # derived -----------------------------------------------------------------
    @property
    def step_dt(self) -> float:
        return self.sim_dt * self.decimation

    @property
    def max_episode_length(self) -> int:
        # L/envs/manager_based_diff_rl_env.py:100-102
        return math.ceil(self.episode_length_s / self.step_dt)

    @property
    def gross_thrust_bound(self) -> Tuple[float, float]:
        # L/controllers/controller_diff.py:96-99 (the lower bound is negative: -4.65 N)
        return (_rotor_thrust(_MOTOR_OMEGA[0]) * 4, _rotor_thrust(_MOTOR_OMEGA[1]) * 4)

    @property
    def weight(self) -> float:
        return self.mass * self.gravity

    def to_dict(self):
        return asdict(self)


# =====================================================================================================================
# Reach-target tasks: the reference's original differentiable tasks, sharing DroneDynamics with
# the racing path but driven by the other command modes (LVController / PSController) or CTBR, with their own command,
# reward, observation and loss terms.  QD/reach_target_lv_env.py, QD/reach_target_ctbr_env.py.
# =====================================================================================================================
REACH_REWARD_TERM_NAMES = ("move_towards", "orientation_reward", "move_in_dir", "action_rate", "reach_target", "smooth_ang_vel",
                           "smooth_lin_acc", "smooth_ang_acc", "early_termination", "hover_state")     # QD/reach_target_lv_env.py:129-189
REACH_LOSS_TERM_NAMES = ("move_towards_goal", "orientation_tracking", "move_in_dir", "smooth_vel")       # :191-219
CONTROLLERS = ("CTBRController", "LVController", "PSController")


@dataclass
class ReachTargetCfg:
    """Frozen constants of DiffLab-Quadcopter-{LV,CTBR}-ReachTarget (PhysX-free closure; DESIGN.md §4e lists the substitutions)."""

    controller: str = "LVController"    # DiffActionCfg.command_type, QD/reach_target_lv_env.py:80 / reach_target_ctbr_env.py:81-93
    sim2real_test: bool = False         # QD/mdp/diff_action.py:168-171: actions are (a_zb, body rates), no tanh, no gradient
    # --- time ----------------------------------------------------------------
    sim_dt: float = 0.005               # QD/reach_target_lv_env.py:251
    decimation: int = 4                 # :246 (6 in the CTBR env, reach_target_ctbr_env.py:258)
    episode_length_s: float = 6.0       # :247
    # --- drone (same synthetic mass as the racing closure) -----------------------
    mass: float = 0.8
    inertia_diag: Tuple[float, float, float] = (0.0015, 0.002, 0.004)   # QD/mdp/diff_action.py:59
    gravity: float = 9.81
    grad_decay: float = 0.92
    drag_1: float = 0.18
    drag_1_rand: float = 0.1
    drag_2: float = 0.01
    drag_2_rand: float = 0.005
    z_drag: float = 4.0
    z_drag_rand: float = 0.4
    random_drag: bool = True            # QD/mdp/diff_action_cfg.py:38 (False in the CTBR env, :91)
    # --- action map (QD/mdp/diff_action.py:247-283) --------------------------------
    max_thrust_weight_ratio: float = 3.0    # CTBR
    lin_vel_bound: float = 5.0          # LV   (QD/mdp/diff_action_cfg.py:46)
    pos_bound: float = 1.0              # PS   (:45)
    action_lag: int = 1
    thr_est_error_init_std: float = 0.02
    thr_est_error_reset_std: float = 0.01
    # --- controllers (L/controllers/controller_diff_cfg.py:43-79) -------------------
    body_rate_bound: float = 12.0       # 6 in the CTBR env (reach_target_ctbr_env.py:88)
    rate_gain_p: Tuple[float, float, float] = (50.0, 50.0, 50.0)      # CTBR (35 in the CTBR env)
    rate_gain_d: Tuple[float, float, float] = (0.0, 0.0, 0.0)         # CTBR (5e-4, 5e-4, 3e-4 in the CTBR env)
    thrust_ctrl_delay: float = 0.03
    torque_ctrl_delay: Tuple[float, float, float] = (0.02, 0.02, 0.02)  # CTBR only (0.03 in the CTBR env)
    max_feedback_accel: float = 20.0    # LV / PS
    speed_gain: Tuple[float, float, float] = (10.0, 10.0, 20.0)
    pose_gain: Tuple[float, float, float] = (18.0, 18.0, 20.0)
    rate_gain: Tuple[float, float, float] = (180.0, 180.0, 200.0)
    pos_gain: Tuple[float, float, float] = (3.0, 3.0, 3.0)            # PS
    # --- command: UniformWorldPoseCommand (QD/reach_target_lv_env.py:66-75, QD/mdp/commands.py:113-134) ----
    cmd_lo: Tuple[float, float, float] = (-2.0, -2.0, 0.5)
    cmd_hi: Tuple[float, float, float] = (2.0, 2.0, 2.5)
    resampling_time: float = 10.0
    # --- reset: reset_root_state_uniform (QD/reach_target_lv_env.py:106-122) -----------
    default_root_pos: Tuple[float, float, float] = (0.0, 0.0, 0.5)
    reset_lo: Tuple[float, ...] = (-0.1, -0.1, 1.0, -0.5, -0.5, -3.14)   # x y z roll pitch yaw
    reset_hi: Tuple[float, ...] = (0.1, 0.1, 2.0, 0.5, 0.5, 3.14)
    # --- terminations: time_out + illegal_contact (:221-228).  Contacts need PhysX; the closure uses the reference's own
    #     commented-out alternative `mdp.out_of_bound` (QD/mdp/observation.py:76-84) with the ground plane as lower bound.
    term_out_of_bound: bool = True
    oob_lo: float = 0.0
    oob_hi: float = 10.0
    # --- rewards (QD/reach_target_lv_env.py:129-189) -------------------------------------
    w_reward: Tuple[float, ...] = (1.0, 0.5, 1.0, -0.001, 10.0, -0.001, -0.001, -0.0001, -200.0, 1.0)
    move_in_dir_threshold: float = 0.4
    reach_threshold: float = 0.1
    hover_threshold: float = 0.2
    hover_ratio: float = 0.2
    # --- losses (:191-219) ------------------------------------------------------------------
    is_differentiable_physics: bool = True
    w_loss: Tuple[float, ...] = (1.0, 0.0, 0.0, 0.3)     # target, orientation, move_in_dir, smooth_vel
    loss_dir_threshold: float = 0.1
    loss_smooth_ratio: float = 0.5
    # --- observations (:83-104): last_action = action_manager.action (LV env) or modified_last_action (CTBR env) ----
    last_action_modified: bool = False

    @staticmethod
    def lv(**overrides) -> "ReachTargetCfg":
        """QuadcopterReachTargetLVEnvCfg."""
        return ReachTargetCfg(**overrides)

    @staticmethod
    def ps(**overrides) -> "ReachTargetCfg":
        """The LV env with command_type="PSController" (PSControllerCfg defaults, controller_diff_cfg.py:66-79)."""
        kw = dict(controller="PSController", speed_gain=(5.0, 5.0, 5.0), pose_gain=(20.0, 20.0, 20.0), rate_gain=(150.0, 150.0, 150.0))
        kw.update(overrides)
        return ReachTargetCfg(**kw)

    @staticmethod
    def ctbr(sim2real_test: bool = False, **overrides) -> "ReachTargetCfg":
        """QuadcopterReachTargetCTBREnvCfg; the shipped file has SIM2REAL_TEST=True (60 s episodes, raw a_zb / body-rate inputs)."""
        kw = dict(controller="CTBRController", sim2real_test=sim2real_test, decimation=6, episode_length_s=60.0 if sim2real_test else 6.0,
                  random_drag=False, body_rate_bound=6.0, rate_gain_p=(35.0, 35.0, 35.0), rate_gain_d=(0.0005, 0.0005, 0.0003),
                  torque_ctrl_delay=(0.03, 0.03, 0.03), reset_lo=(-0.1, -0.1, 1.0, 0.0, 0.0, -3.14), reset_hi=(0.1, 0.1, 2.0, 0.0, 0.0, 3.14),
                  w_loss=(1.0, 1.0, 1.0, 0.1), last_action_modified=True)
        kw.update(overrides)
        return ReachTargetCfg(**kw)

    @property
    def step_dt(self) -> float:
        return self.sim_dt * self.decimation

    @property
    def max_episode_length(self) -> int:
        return math.ceil(self.episode_length_s / self.step_dt)

    @property
    def gross_thrust_bound(self) -> Tuple[float, float]:
        return (_rotor_thrust(_MOTOR_OMEGA[0]) * 4, _rotor_thrust(_MOTOR_OMEGA[1]) * 4)

    @property
    def action_scale(self) -> Tuple[float, float, float, float]:
        # QD/mdp/diff_action.py:257-275 ("medium")
        if self.controller == "CTBRController":
            half = self._half_max_thrust()
            return (half, self.body_rate_bound, self.body_rate_bound, self.body_rate_bound)
        b = self.lin_vel_bound if self.controller == "LVController" else self.pos_bound
        return (3.1415926, b, b, b)

    @property
    def action_offset(self) -> Tuple[float, float, float, float]:
        if self.controller == "CTBRController":
            return (self._half_max_thrust(), 0.0, 0.0, 0.0)
        return (0.0, 0.0, 0.0, 0.0)

    def _half_max_thrust(self) -> float:
        """m g ratio / 2 with the reference's fp32 roundings (QD/mdp/diff_action.py:56,262: an fp32 mass tensor times python floats)."""
        import numpy as np
        weight = np.float32(self.mass) * np.float32(abs(self.gravity))
        return float(np.float32(weight * np.float32(self.max_thrust_weight_ratio)) / np.float32(2.0))

    def to_dict(self):
        return asdict(self)

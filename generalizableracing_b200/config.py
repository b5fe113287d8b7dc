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

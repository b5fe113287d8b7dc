"""ctypes binding of libgracing.so (include/gracing.h).  There is NO CPU fallback: importing
:func:`load` without the built CUDA library raises, and every entry point needs device pointers.
"""
from __future__ import annotations

import ctypes as C
import os

from . import layout as L

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgracing.so")

c_f = C.c_float
c_i = C.c_int32
c_p = C.c_void_p


class GrConfig(C.Structure):
    _fields_ = [
        ("dt", c_f), ("max_episode_length", c_i), ("gravity", c_f), ("grad_decay", c_f), ("inertia", c_f * 3),
        ("action_scale0", c_f), ("body_rate_bound", c_f), ("thrust_lo", c_f), ("thrust_hi", c_f), ("update_threshold", c_f),
        ("drag1", c_f), ("drag1_rand", c_f), ("drag2", c_f), ("drag2_rand", c_f), ("z_drag", c_f), ("z_drag_rand", c_f),
        ("random_drag", c_i), ("thr_err_reset_std", c_f), ("thr_err_init_std", c_f),
        ("default_pos", c_f * 3), ("reset_pos", c_f), ("reset_roll_pitch", c_f), ("reset_yaw", c_f), ("reset_vel", c_f),
        ("kp", c_f * 3), ("kd", c_f * 3), ("thrust_delay", c_f), ("torque_delay", c_f * 3),
        ("pid_scale_lo", c_f), ("pid_scale_span", c_f), ("delay_scale_lo", c_f), ("delay_scale_span", c_f),
        ("mass", c_f), ("max_init_level", c_i),
        ("term_oob", c_i), ("term_bad_pose", c_i), ("oob_lo", c_f), ("oob_hi", c_f),
        ("w_reward", c_f * L.NUM_REWARD_TERMS),
        ("add_cmd_noise", c_i), ("cmd_noise_pos", c_f), ("cmd_noise_yaw", c_f),
        ("level_up_gates", c_i), ("level_down_gates", c_i),
        ("noise_curriculum", c_i), ("noise_up_gates", c_i), ("noise_down_gates", c_i), ("noise_up", c_f), ("noise_down", c_f),
        ("w_loss", c_f * 3), ("obs_vel_noise", c_f), ("obs_euler_noise", c_f),
    ]


class GrTrack(C.Structure):
    _fields_ = [("rows", c_p), ("types", c_i), ("levels", c_i), ("gates", c_i)]


class GrState(C.Structure):
    _fields_ = [("planes", c_p), ("plane_stride", C.c_int64), ("num_envs", c_i), ("num_planes", c_i),
                ("env_id_offset", c_i), ("max_types_per_block", c_i), ("block_threads", c_i), ("launch_flags", c_i), ("chunk_types", c_p)]


class GrRandom(C.Structure):
    _fields_ = [("rnd", c_p), ("seed", C.c_uint64), ("step", C.c_uint32)]


class GrStepIO(C.Structure):
    _fields_ = [("action", c_p), ("obs", c_p), ("critic_obs", c_p), ("aux_obs", c_p), ("reward", c_p),
                ("terminated", c_p), ("time_out", c_p), ("dones", c_p), ("reward_terms", c_p), ("gate_passed", c_p),
                ("loss", c_p), ("loss_terms", c_p), ("tape", c_p), ("tape_stride", C.c_int64), ("phase_times", c_p), ("log_accum", c_p),
                ("aligned_states", c_p), ("acc", c_p), ("dones_u8", c_p),
                ("pre_reset_pos", c_p), ("pre_reset_quat", c_p)]


class GrBwdIO(C.Structure):
    _fields_ = [("tape", c_p), ("tape_stride", C.c_int64), ("t_begin", c_i), ("t_end", c_i), ("grad_loss", c_p),
                ("grad_scale", c_f), ("adjoint", c_p), ("adj_stride", C.c_int64), ("grad_action", c_p)]


class GrRolloutIO(C.Structure):
    _fields_ = [("actions", c_p), ("obs_out", c_p), ("critic_obs_out", c_p), ("aux_out", c_p), ("obs_seq", c_p), ("reward", c_p), ("dones", c_p),
                ("terminated", c_p), ("time_out", c_p), ("loss", c_p), ("loss_terms", c_p), ("tape", c_p), ("tape_stride", C.c_int64),
                ("log_accum", c_p), ("T", c_i)]


class GrTransition(C.Structure):
    _fields_ = [("obs", c_p), ("critic_obs", c_p), ("actions", c_p), ("rewards", c_p), ("dones", c_p), ("dones_is_int64", c_i),
                ("values", c_p), ("log_prob", c_p), ("mu", c_p), ("sigma", c_p), ("time_outs", c_p), ("gamma", c_f)]


class GrStorage(C.Structure):
    _fields_ = [("obs", c_p), ("critic_obs", c_p), ("actions", c_p), ("rewards", c_p), ("dones", c_p), ("values", c_p),
                ("log_prob", c_p), ("mu", c_p), ("sigma", c_p), ("returns", c_p), ("advantages", c_p),
                ("T", c_i), ("N", c_i), ("obs_dim", c_i), ("critic_dim", c_i), ("act_dim", c_i)]


class GrMiniBatch(C.Structure):
    _fields_ = [("obs", c_p), ("critic_obs", c_p), ("actions", c_p), ("values", c_p), ("advantages", c_p), ("returns", c_p),
                ("log_prob", c_p), ("mu", c_p), ("sigma", c_p)]


class GrMlp(C.Structure):
    _fields_ = [("w1", c_p), ("b1", c_p), ("w2", c_p), ("b2", c_p), ("w3", c_p), ("b3", c_p), ("in_dim", c_i), ("hidden", c_i), ("hidden2", c_i), ("out_dim", c_i)]


class GrPolicy(C.Structure):
    _fields_ = [("packed", c_p), ("sigma", c_p), ("negative_slope", c_f)]


class GrCollectIO(C.Structure):
    _fields_ = [("obs0", c_p), ("critic_obs0", c_p), ("obs_out", c_p), ("critic_obs_out", c_p), ("aux_out", c_p), ("last_values", c_p),
                ("episode_acc", c_p), ("log_accum", c_p), ("episode_log", c_p), ("gamma", c_f), ("groups_per_cta", c_i), ("coop_reset_columns", c_i)]


class GrBpttCollectIO(C.Structure):
    _fields_ = [("obs0", c_p), ("obs_out", c_p), ("critic_obs_out", c_p), ("aux_out", c_p), ("obs_seq", c_p), ("eps_seq", c_p), ("actions", c_p),
                ("loss", c_p), ("loss_terms", c_p), ("reward", c_p), ("dones", c_p), ("tape", c_p), ("tape_stride", C.c_int64), ("log_accum", c_p),
                ("T", c_i), ("groups_per_cta", c_i)]


class GrMlpGrad(C.Structure):
    _fields_ = [("w1", c_p), ("b1", c_p), ("w2", c_p), ("b2", c_p), ("w3", c_p), ("b3", c_p), ("out_dim", c_i), ("scale_is_maxabs", c_i)]


class GrBackwardJob(C.Structure):
    _fields_ = [("policy", GrPolicy), ("obs", c_p), ("grad_actions", c_p), ("scale", c_p), ("out", GrMlpGrad), ("indices", c_p), ("obs_stride", c_i)]


class GrAdamStep(C.Structure):
    _fields_ = [("param_ptrs", c_p), ("seg_offsets", c_p), ("seg_sizes", c_p), ("n_seg", c_i), ("n_flat", c_i), ("grad", c_p), ("exp_avg", c_p),
                ("exp_avg_sq", c_p), ("state", c_p), ("kl_stats", c_p), ("grad_scale", c_f), ("beta1", c_f), ("beta2", c_f), ("eps", c_f),
                ("max_grad_norm", c_f), ("desired_kl", c_f), ("lr_min", c_f), ("lr_max", c_f)]


class GrPpoBatch(C.Structure):
    _fields_ = [("mu", c_p), ("value", c_p), ("sigma", c_p), ("actions", c_p), ("old_log_prob", c_p), ("advantages", c_p), ("returns", c_p),
                ("old_values", c_p), ("old_mu", c_p), ("old_sigma", c_p), ("clip_param", c_f), ("value_loss_coef", c_f), ("entropy_coef", c_f),
                ("use_clipped_value_loss", c_i), ("indices", c_p), ("records", c_p)]


class GrPpoStep(C.Structure):
    _fields_ = [("policy", GrPolicy), ("obs", c_p), ("critic_obs", c_p), ("batch", GrPpoBatch), ("actor_grad", GrMlpGrad), ("critic_grad", GrMlpGrad),
                ("sums", c_p), ("cotangent_scale", c_f)]


class GrPeerReduce(C.Structure):
    _fields_ = [("peer_bufs", c_p), ("peer_flags", c_p), ("world", c_i), ("rank", c_i), ("n", c_i), ("reserved", c_i), ("max_spins", C.c_int64),
                ("epoch", c_p), ("counter", c_p), ("error", c_p)]


GR_PEER_MAX_WORLD = 16


class GrHostStep(C.Structure):
    _fields_ = [("action", c_p), ("obs", c_p), ("reward", c_p), ("dones", c_p), ("critic_obs", c_p), ("time_out", c_p), ("dones_u8", c_p),
                ("outputs_contiguous", c_i)]


GR_HOST_PIPE_MAX_DEPTH = 4
GR_LAUNCH_PDL = 1
GR_LAUNCH_PREFETCH = 2
GR_LAUNCH_PREFETCH_L2 = 4
GR_LAUNCH_EARLY_STORE = 8
GR_LAUNCH_COOP_RESET = 16
GR_LOG_SLOTS = 16
GR_LOG_NUM_RESET, GR_LOG_SUM_GATES, GR_LOG_SUM_EPSUM, GR_LOG_NUM_TIMEOUT, GR_LOG_NUM_TERMINATED = 0, 1, 2, 8, 9
GR_LOG_SUM_ACTION_RATE, GR_LOG_SUM_LIN_SPD, GR_LOG_SUM_ANG_SPD, GR_LOG_SUM_LOSS = 10, 11, 12, 13
class GrReachConfig(C.Structure):
    _fields_ = [
        ("controller", c_i), ("sim2real_test", c_i), ("last_action_modified", c_i), ("random_drag", c_i),
        ("dt", c_f), ("max_episode_length", c_i), ("gravity", c_f), ("grad_decay", c_f), ("mass", c_f), ("inertia", c_f * 3),
        ("action_scale", c_f * 4), ("action_offset", c_f * 4), ("thrust_lo", c_f), ("thrust_hi", c_f), ("body_rate_bound", c_f),
        ("kp", c_f * 3), ("kd", c_f * 3), ("thrust_delay", c_f), ("torque_delay", c_f * 3),
        ("speed_gain", c_f * 3), ("pose_gain", c_f * 3), ("rate_gain", c_f * 3), ("pos_gain", c_f * 3), ("max_feedback_accel", c_f),
        ("drag1", c_f), ("drag1_rand", c_f), ("drag2", c_f), ("drag2_rand", c_f), ("z_drag", c_f), ("z_drag_rand", c_f),
        ("thr_err_reset_std", c_f),
        ("default_pos", c_f * 3), ("reset_lo", c_f * 6), ("reset_hi", c_f * 6),
        ("cmd_lo", c_f * 3), ("cmd_hi", c_f * 3), ("resample_time", c_f),
        ("term_oob", c_i), ("oob_lo", c_f), ("oob_hi", c_f),
        ("w_reward", c_f * L.REACH_NUM_REWARD_TERMS), ("move_in_dir_thr", c_f), ("reach_thr", c_f), ("hover_thr", c_f), ("hover_ratio", c_f),
        ("w_loss", c_f * L.REACH_NUM_LOSS_TERMS), ("loss_dir_thr", c_f), ("loss_smooth_ratio", c_f),
    ]


class GrReachState(C.Structure):
    _fields_ = [("planes", c_p), ("plane_stride", C.c_int64), ("num_envs", c_i), ("env_id_offset", c_i)]


class GrReachStepIO(C.Structure):
    _fields_ = [("action", c_p), ("obs", c_p), ("reward", c_p), ("terminated", c_p), ("time_out", c_p), ("dones", c_p), ("reward_terms", c_p),
                ("loss", c_p), ("loss_terms", c_p), ("tape", c_p), ("tape_stride", C.c_int64), ("log_accum", c_p)]


class GrReachRolloutIO(C.Structure):
    _fields_ = [("actions", c_p), ("obs_out", c_p), ("obs_seq", c_p), ("reward", c_p), ("dones", c_p), ("terminated", c_p), ("time_out", c_p),
                ("loss", c_p), ("loss_terms", c_p), ("tape", c_p), ("tape_stride", C.c_int64), ("log_accum", c_p), ("T", c_i)]


class GrMesh(C.Structure):
    _fields_ = [("nodes", c_p), ("tris", c_p), ("num_nodes", c_i), ("num_faces", c_i)]


GR_REACH_LOG_NUM_RESET, GR_REACH_LOG_SUM_POS_ERR, GR_REACH_LOG_SUM_EPSUM, GR_REACH_LOG_NUM_TIMEOUT, GR_REACH_LOG_NUM_TERMINATED = 0, 1, 2, 12, 13
GR_LOG_SHARDS = 256
GR_RECORD_FLOATS = 48
GR_PHILOX_CALL_ACTION = 16
STATUS = {0: "GR_OK", -1: "GR_ERR_NULL", -2: "GR_ERR_SIZE", -3: "GR_ERR_ALIGN", -4: "GR_ERR_CONFIG", -5: "GR_ERR_SMEM"}

# symbol -> (restype, argtypes); every symbol include/gracing.h declares
PROTOTYPES = {
    "gr_abi_version": (C.c_int, []),
    "gr_env_startup": (C.c_int, [C.POINTER(GrConfig), C.POINTER(GrTrack), C.POINTER(GrState), c_p, c_p, c_p, C.c_uint64, c_p]),
    "gr_env_reset": (C.c_int, [C.POINTER(GrConfig), C.POINTER(GrTrack), C.POINTER(GrState), C.POINTER(GrRandom), c_p, c_p, c_p, c_p, c_p]),
    "gr_env_observe": (C.c_int, [C.POINTER(GrConfig), C.POINTER(GrTrack), C.POINTER(GrState), C.POINTER(GrRandom), c_p, c_p, c_p, c_p]),
    "gr_step_fwd": (C.c_int, [C.POINTER(GrConfig), C.POINTER(GrTrack), C.POINTER(GrState), C.POINTER(GrRandom), C.POINTER(GrStepIO), c_p]),
    "gr_step_bwd": (C.c_int, [C.POINTER(GrConfig), C.POINTER(GrState), C.POINTER(GrBwdIO), c_p]),
    "gr_rollout_fwd": (C.c_int, [C.POINTER(GrConfig), C.POINTER(GrTrack), C.POINTER(GrState), C.POINTER(GrRandom), C.POINTER(GrRolloutIO), c_p]),
    "gr_fill_rand": (C.c_int, [c_p, c_i, c_i, C.c_uint64, C.c_uint32, c_p]),
    "gr_fill_startup_rand": (C.c_int, [c_p, c_i, c_i, C.c_uint64, c_p]),
    "gr_selftest_sqrt_rn": (C.c_int, [c_p, c_p, C.c_int64, c_p]),
    "gr_storage_add": (C.c_int, [C.POINTER(GrStorage), C.POINTER(GrTransition), c_i, c_p]),
    "gr_gae_scratch_bytes": (C.c_int64, [c_i]),
    "gr_compute_returns": (C.c_int, [C.POINTER(GrStorage), c_p, c_f, c_f, c_p, c_p, c_i, c_p]),
    "gr_advantage_normalize": (C.c_int, [C.POINTER(GrStorage), c_p, c_p]),
    "gr_storage_pack_records": (C.c_int, [C.POINTER(GrStorage), c_p, c_p]),
    "gr_storage_pack_records_permuted": (C.c_int, [C.POINTER(GrStorage), c_p, C.c_int64, c_p, c_p]),
    "gr_storage_gather": (C.c_int, [C.POINTER(GrStorage), c_p, c_i, C.POINTER(GrMiniBatch), c_p]),
    "gr_policy_packed_bytes": (C.c_int64, [c_i, c_i, c_i]),
    "gr_bptt_collect": (C.c_int, [C.POINTER(GrConfig), C.POINTER(GrTrack), C.POINTER(GrState), C.POINTER(GrRandom), C.POINTER(GrPolicy), c_i, c_i,
                                  C.POINTER(GrBpttCollectIO), c_p]),
    "gr_policy_pack": (C.c_int, [C.POINTER(GrMlp), C.POINTER(GrMlp), c_p, c_p]),
    "gr_ppo_collect": (C.c_int, [C.POINTER(GrConfig), C.POINTER(GrTrack), C.POINTER(GrState), C.POINTER(GrRandom), C.POINTER(GrPolicy),
                                 C.POINTER(GrStorage), C.POINTER(GrCollectIO), c_p]),
    "gr_actor_backward": (C.c_int, [C.POINTER(GrPolicy), c_i, c_i, c_p, c_p, c_p, C.c_int64, C.POINTER(GrMlpGrad), c_p]),
    "gr_policy_forward": (C.c_int, [C.POINTER(GrPolicy), c_p, c_p, c_p, c_p, C.c_int64, c_p]),
    "gr_policy_forward_gather": (C.c_int, [C.POINTER(GrPolicy), c_p, c_p, c_p, c_p, c_p, C.c_int64, c_p]),
    "gr_ppo_loss_grad": (C.c_int, [C.POINTER(GrPpoBatch), C.c_int64, c_p, c_p, c_p, c_p]),
    "gr_policy_forward_loss": (C.c_int, [C.POINTER(GrPolicy), c_p, c_p, C.POINTER(GrPpoBatch), C.c_int64, c_p, c_p, c_p, c_p]),
    "gr_ppo_fused_step": (C.c_int, [C.POINTER(GrPpoStep), C.c_int64, c_p]),
    "gr_adam_clip_step": (C.c_int, [C.POINTER(GrAdamStep), c_p]),
    "gr_peer_allreduce": (C.c_int, [C.POINTER(GrPeerReduce), c_p, c_p]),
    "gr_actor_backward_jobs": (C.c_int, [C.POINTER(GrBackwardJob), c_i, c_i, c_i, C.c_int64, c_p]),
    "gr_reach_reset": (C.c_int, [C.POINTER(GrReachConfig), C.POINTER(GrReachState), C.POINTER(GrRandom), c_p, c_p, c_p]),
    "gr_reach_observe": (C.c_int, [C.POINTER(GrReachConfig), C.POINTER(GrReachState), c_p, c_p]),
    "gr_reach_step_fwd": (C.c_int, [C.POINTER(GrReachConfig), C.POINTER(GrReachState), C.POINTER(GrRandom), C.POINTER(GrReachStepIO), c_p]),
    "gr_reach_rollout_fwd": (C.c_int, [C.POINTER(GrReachConfig), C.POINTER(GrReachState), C.POINTER(GrRandom), C.POINTER(GrReachRolloutIO), c_p]),
    "gr_reach_step_bwd": (C.c_int, [C.POINTER(GrReachConfig), C.POINTER(GrReachState), C.POINTER(GrBwdIO), c_p]),
    "gr_reach_fill_rand": (C.c_int, [c_p, c_i, c_i, C.c_uint64, C.c_uint32, c_p]),
    "gr_traj_index": (C.c_int, [c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_p]),
    "gr_traj_pad": (C.c_int, [c_p, c_i, c_i, c_i, c_p, c_p, c_p, c_i, c_i, c_p, c_p, c_p]),
    "gr_traj_unpad": (C.c_int, [c_p, c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p]),
    "gr_traj_hidden": (C.c_int, [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_i, c_i, c_p, c_p]),
    "gr_host_pipe_create": (C.c_int, [c_i, c_i, c_p, C.POINTER(c_p)]),
    "gr_host_pipe_destroy": (C.c_int, [c_p]),
    "gr_host_pipe_step": (C.c_int, [c_p, C.POINTER(GrConfig), C.POINTER(GrTrack), C.POINTER(GrState), C.POINTER(GrRandom),
                                    C.POINTER(GrHostStep), c_p, C.POINTER(C.c_int64)]),
    "gr_host_pipe_wait": (C.c_int, [c_p, C.c_int64]),
    "gr_host_copy_probe": (C.c_int, [c_i, c_i, c_i, C.POINTER(C.c_double)]),
    "gr_host_copy_probe2": (C.c_int, [c_i, c_i, c_i, c_i, C.POINTER(C.c_double)]),
    "gr_mesh_bvh_max_nodes": (C.c_int64, [c_i]),
    "gr_mesh_build_bvh": (C.c_int, [c_p, c_p, c_i, c_i, c_p, C.c_int64, c_p, c_p, C.POINTER(c_i)]),
    "gr_uav_collision_ray": (C.c_int, [C.POINTER(GrMesh), c_p, c_p, c_i, c_p, c_i, c_f, c_f, c_f, c_p, c_p]),
    "gr_mesh_query_rays": (C.c_int, [C.POINTER(GrMesh), c_p, c_p, C.c_int64, c_f, c_p, c_p, c_p]),
}


class GracingError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise GracingError(f"{what}: {STATUS.get(rc, rc)}")
    raise GracingError(f"{what}: CUDA error {rc} at launch")


_lib = None


def load() -> C.CDLL:
    """Load libgracing.so and bind every prototype.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the racing hot path has no CPU fallback. Build the sm_100a library with "
            "`python -m generalizableracing_b200.build` (or __graft_entry__.build()).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.gr_abi_version() != 4:
        raise ImportError("libgracing.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None stays NULL)."""
    return None if t is None else t.data_ptr()

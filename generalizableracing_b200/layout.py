"""Shared layout constants of the drop-in boundary (mirrored in include/gracing.h).

Random-number slots: the reference draws with ``torch.rand/randn/uniform_`` over
*compacted* index sets (e.g. QD/mdp/commands.py:286-293, QD/mdp/dynamics/
droneDynamics.py:53-56, QD/mdp/events.py:153), so the owner of the k-th number
depends on which envs reset.  The boundary therefore takes the draws as a dense
per-step tensor ``rnd[N, RND_STRIDE]`` ("parity mode") or generates exactly the same
slots in-kernel from Philox4x32-10 keyed by (seed, global env id, step, slot/4)
("throughput mode").  U = uniform [0,1), N = standard normal.
"""

# --- per-step slots -------------------------------------------------------------
RND_OBS_VEL = 0        # 3 N : policy obs lin-vel noise        (QD/mdp/observation.py:52)
RND_OBS_EUL = 3        # 3 N : policy obs attitude noise       (QD/mdp/observation.py:27)
RND_THR_ERR = 6        # 1 N : thr_est_error re-draw on reset  (QD/mdp/diff_action.py:233)
RND_SPARE_N = 7        # 1 N : unused (second output of the Box-Muller pair)
# --- consumed only by envs that reset this step -----------------------------------
RND_RESET_POSE = 8     # 6 U : x y z roll pitch yaw            (QD/mdp/events.py:153)
RND_RESET_VEL = 14     # 6 U : linear + angular velocity       (QD/mdp/events.py:171)
RND_Z_DRAG = 20        # 1 U                                   (QD/mdp/dynamics/droneDynamics.py:53)
RND_DRAG2 = 21         # 3 U : quadratic drag                  (droneDynamics.py:54)
RND_DRAG1 = 24         # 3 U : linear drag                     (droneDynamics.py:56)
RND_LEVEL = 27         # 1 U : terrain level re-draw when the curriculum tops out
# gate noise on resample (QD/mdp/commands.py:286-306): 6 U for the current gate + 6 U for the next gate, ordered
# (x y z roll pitch yaw).  Only the positional components are live, so they sit first: the in-kernel
# Philox path never generates the calls that hold only dead draws.
RND_RESET_GATE = (28, 29, 30, 34, 35, 36)
RND_RESET_NEXT = (31, 32, 33, 37, 38, 39)
# --- consumed only by envs that passed a gate this step (commands.py:330-350), same ordering ----
RND_PASS_GATE = (40, 41, 42, 46, 47, 48)
RND_PASS_NEXT = (43, 44, 45, 49, 50, 51)
RND_STRIDE = 52        # floats per env per step (13 Philox calls)

# --- startup slots (one draw per env at construction) ---------------------------------
SRND_KP = 0            # 3 U : rate_gain_p scale               (QD/mdp/events.py:116-117)
SRND_KD = 3            # 3 U : rate_gain_d scale               (events.py:126-127)
SRND_THRUST_DELAY = 6  # 1 U                                   (events.py:131-132)
SRND_TORQUE_DELAY = 7  # 3 U                                   (events.py:136-137)
SRND_LEVEL = 10        # 1 U : initial terrain level
SRND_SPARE = 11
SRND_THR_ERR = 12      # 1 N : initial thr_est_error           (QD/mdp/diff_action.py:86)
SRND_STRIDE = 16

# Philox stream ids (counter word 2): per-step draws use the step counter, startup uses this.
PHILOX_STARTUP_STREAM = 0xFFFFFFFF

# --- observation layout (QD/racing_ctbr_env.py:139-174) ---------------------------------
OBS_DIM = 16           # lin_vel_b 3 | R(q)[2,:] 3 | command 6 | last action (a_z, omega) 4
NUM_ACTIONS = 4
NUM_REWARD_TERMS = 6   # progress, bodyrate, action_rate, perception, success_cross, bad_pose
REWARD_TERM_NAMES = ("progress_rewards", "command_bodyrate_penalty", "action_rate",
                     "perception_reward", "success_cross", "bad_pose_penalty")

# --- env state: array of 32-env tiles, each tile = 16 planes x 32 lanes x float4 (8 KB contiguous) ------------
# torch view: planes[num_tiles, TILE_PLANES, TILE, 4]; plane p of env i lives at planes[i // 32, p, i % 32].
# A warp owns one tile, so its loads are one contiguous 8 KB block (see gr_common.cuh for why).
TILE = 32
TILE_PLANES = 16
# planes written every step first (one contiguous write-back per tile) ...
PL_QUAT = 0      # q.w q.x q.y q.z
PL_POS = 1       # world pos x y z | thrust filter state f
PL_LINVEL = 2    # v_w x y z       | bits: episode_length [0:12) | cross_obs flag [12] | noise-planes-rewritten flag [13] |
                 #                   "action_rate" command metric as a 16-bit float [14:30) (5-bit exponent biased at 2^-21, 11-bit
                 #                   mantissa) | 0 [30] | metrics-are-zero flag [31]
PL_ANGVEL = 3    # omega_b x y z   | packed ints: gate_id | acc_gates<<8 | level<<20 | type<<26 | fresh<<31
PL_TORQUE = 4    # torque filter state x y z | episode sum of reward term 4 (success_cross)
PL_ANGACC = 5    # alpha_b x y z   | episode sum of reward term 5 (bad_pose)
PL_FIFO = 6      # action-lag FIFO (tanh(a_{t-1}))
NUM_HOT_PLANES = 7
PL_EPSUM0 = 7    # episode sums of reward terms 0..3          (touched only with episode_stats)
PL_LOSSSUM = 8   # LossManager episode sums of the 3 loss terms | spare   (episode_stats + differentiable physics)
EPLEN_MASK = 0xFFF          # PL_LINVEL.w bit fields
EPLEN_AUX_BIT, EPLEN_NOISE_DIRTY_BIT, EPLEN_METRIC_SHIFT, EPLEN_METRIC_MASK, EPLEN_METRIC_BIAS = 1 << 12, 1 << 13, 14, 0xFFFF, 106 << 11
LOSS_TERM_NAMES = ("move_towards_goal", "falling", "falling_speed")      # QD/racing_ctbr_env.py:331-353
# ... then planes read every step and written on reset / startup
PL_DRAG2 = 9     # quadratic drag x y z(*z_drag) | mass
PL_DRAG1 = 10    # linear drag x y z(*z_drag)    | exp(-dt/thrust_delay)
PL_KP = 11       # rate_gain_p x y z | thr_est_error
PL_KD = 12       # rate_gain_d x y z | spare
PL_ETAU = 13     # exp(-dt/torque_delay) x y z | spare
PL_NOISE0 = 14   # delta_cur x y z | delta_next x            (touched only with add_cmd_noise)
PL_NOISE1 = 15   # delta_next y z | noise_pos_hi | noise_level
NUM_PLANES = 14             # GrState.num_planes value meaning "no episode sums"
NUM_PLANES_WITH_STATS = 16  # GrState.num_planes value meaning "episode sums maintained"

# packed-int field positions in PL_ANGVEL.w
PK_GATE_BITS, PK_GATE_SHIFT = 8, 0
PK_ACC_BITS, PK_ACC_SHIFT = 12, 8
PK_LEVEL_BITS, PK_LEVEL_SHIFT = 6, 20
PK_TYPE_BITS, PK_TYPE_SHIFT = 5, 26
PK_FRESH_SHIFT = 31    # 1 = env was reset at the previous step (action latches are zero)

# --- BPTT tape: 7 float4 planes per env-step ------------------------------------------------
TAPE_PLANES = 7

# =====================================================================================================================
# Reach-target tasks (mirrored in include/gracing.h, GR_REACH_*)
# =====================================================================================================================
REACH_OBS_DIM = 17          # lin_vel_b 3 | ang_vel_b 3 | last action 4 | root quat 4 | desired_pos_b 3   (QD/reach_target_lv_env.py:83-104)
REACH_NUM_REWARD_TERMS = 10
REACH_NUM_LOSS_TERMS = 4
# per-step random slots, consumed only on reset (U uniform [0,1), N standard normal) ...
REACH_RND_RESET_POSE = 0    # 6 U : x y z roll pitch yaw            (Isaac Lab mdp.reset_root_state_uniform)
REACH_RND_Z_DRAG = 6        # 1 U                                   (QD/mdp/dynamics/droneDynamics.py:53)
REACH_RND_DRAG2 = 7         # 3 U
REACH_RND_DRAG1 = 10        # 3 U
REACH_RND_THR_ERR = 13      # 1 N : thr_est_error re-draw on reset  (QD/mdp/diff_action.py:233)
REACH_RND_CMD = 14          # 3 U : new target, on reset            (QD/mdp/commands.py:113-121)
REACH_RND_SPARE = 17
# ... or when the command timer runs out (commands resample every `resampling_time` seconds)
REACH_RND_CMD_TIMER = 18    # 3 U  (18, 19, 20)
REACH_RND_STRIDE = 24       # 6 Philox calls
# env state: 32-env tiles x REACH_PLANES planes x float4 (csrc/reach_core.cuh ReachPlane)
REACH_PLANES = 13
RPL_QUAT = 0      # q.w q.x q.y q.z
RPL_POS = 1       # pos x y z       | thrust filter state f
RPL_LINVEL = 2    # v_w x y z       | episode_length (int32 bits)
RPL_ANGVEL = 3    # omega_b x y z   | command time_left
RPL_TORQUE = 4    # CTBR torque filter state x y z | raw_actions.w
RPL_ANGACC = 5    # alpha_b x y z   | fresh flag (1.0 = reset at the previous step: action latches are zero)
RPL_FIFO = 6      # action-lag FIFO (raw a_{t-1})
RPL_TARGET = 7    # pose_command_w xyz | raw_actions.x
RPL_EPSUM0 = 8    # episode sums of reward terms 0..3
RPL_EPSUM1 = 9    # 4..7
RPL_EPSUM2 = 10   # 8..9 | raw_actions.y | raw_actions.z     (raw_actions = the lagged action the last step applied)
RPL_DRAG2 = 11    # quadratic drag x y z(*z_drag) | spare           -- planes 11, 12 are rewritten only on reset
RPL_DRAG1 = 12    # linear drag x y z(*z_drag)    | thr_est_error
REACH_TAPE_PLANES = 13

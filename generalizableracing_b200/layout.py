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
RND_RESET_GATE = 28    # 6 U : noise of the current gate on resample (QD/mdp/commands.py:286-295)
RND_RESET_NEXT = 34    # 6 U : noise of the next gate on resample    (commands.py:297-306)
# --- consumed only by envs that passed a gate this step ------------------------------
RND_PASS_GATE = 40     # 6 U                                   (commands.py:330-339)
RND_PASS_NEXT = 46     # 6 U                                   (commands.py:341-350)
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

# --- SoA state planes: each plane is a [N] array of float4 (16 B per env) -----------------
# hot planes (read + written every step)
PL_QUAT = 0      # q.w q.x q.y q.z
PL_POS = 1       # world pos x y z | thrust filter state f
PL_LINVEL = 2    # v_w x y z       | episode_length (int32 bits)
PL_ANGVEL = 3    # omega_w x y z   | packed ints: gate_id | acc_gates<<8 | level<<20 | type<<26 | fresh<<31
PL_TORQUE = 4    # torque filter state x y z | spare
PL_ANGACC = 5    # alpha_w x y z   | spare
PL_FIFO = 6      # action-lag FIFO (a_{t-1})
NUM_HOT_PLANES = 7
# cold planes (read every step, written on reset / startup)
PL_DRAG2 = 7     # quadratic drag x y z(*z_drag) | mass
PL_DRAG1 = 8     # linear drag x y z(*z_drag)    | exp(-dt/thrust_delay)
PL_KP = 9        # rate_gain_p x y z | thr_est_error
PL_KD = 10       # rate_gain_d x y z | spare
PL_ETAU = 11     # exp(-dt/torque_delay) x y z | spare
# command-noise planes (only touched when add_cmd_noise)
PL_NOISE0 = 12   # delta_cur x y z | delta_next x
PL_NOISE1 = 13   # delta_next y z | noise_pos_hi | noise_level
NUM_PLANES = 14
# optional episode-sum planes (reward logging, extras["log"])
PL_EPSUM0 = 14   # episode sums of reward terms 0..3
PL_EPSUM1 = 15   # episode sums of reward terms 4..5 | spare | spare
NUM_PLANES_WITH_STATS = 16

# packed-int field positions in PL_ANGVEL.w
PK_GATE_BITS, PK_GATE_SHIFT = 8, 0
PK_ACC_BITS, PK_ACC_SHIFT = 12, 8
PK_LEVEL_BITS, PK_LEVEL_SHIFT = 6, 20
PK_TYPE_BITS, PK_TYPE_SHIFT = 5, 26
PK_FRESH_SHIFT = 31    # 1 = env was reset at the previous step (action latches are zero)

# --- BPTT tape: 7 float4 planes per env-step ------------------------------------------------
TAPE_PLANES = 7

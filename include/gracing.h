/* gracing.h -- C ABI of libgracing.so, the B200 (sm_100a) racing hot path.
 *
 * The reference (yufengsjtu/GeneralizableRacing) is pure Python/torch; there is no
 * FFI in it.  Each entry point below replaces a group of reference *Python* calls on
 * the hot path (file:line relative to the reference root; QD = extensions/
 * diff.lab_tasks/diff/lab_tasks/tasks/quadcopter_diff, L = extensions/diff.lab/diff/lab,
 * S = standalone).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *  - all pointers are DEVICE pointers owned by the caller (torch-allocated); the
 *    library never allocates, frees or keeps them; no global state.
 *  - every call is asynchronous and stream-ordered on `stream` (pass
 *    torch.cuda.current_stream().cuda_stream); no host synchronisation, no host reads:
 *    every call is CUDA-graph capturable.
 *  - return value: 0 = ok, negative = GrStatus argument error (nothing launched),
 *    positive = cudaError_t of the launch.  Never throws, never exits.
 *  - fp32 everywhere; masks are uint8; ids/counters are int32.
 */
#ifndef GRACING_H_
#define GRACING_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GR_ABI_VERSION 4

/* ---- layout constants (mirrored by generalizableracing_b200/layout.py) ---- */
#define GR_OBS_DIM 16
#define GR_NUM_ACTIONS 4
#define GR_NUM_REWARD_TERMS 6
#define GR_RND_STRIDE 52
#define GR_SRND_STRIDE 16
#define GR_NUM_HOT_PLANES 7
#define GR_NUM_PLANES 14
#define GR_NUM_PLANES_WITH_STATS 16
#define GR_TAPE_PLANES 7
#define GR_TILE_PLANES 16
#define GR_MAX_GATES 32

typedef enum GrStatus {
  GR_OK = 0,
  GR_ERR_NULL = -1,       /* a required pointer is NULL */
  GR_ERR_SIZE = -2,       /* num_envs <= 0, bad stride, bad table shape */
  GR_ERR_ALIGN = -3,      /* a vectorised buffer is not 16-byte aligned */
  GR_ERR_CONFIG = -4,     /* inconsistent GrConfig (e.g. action_lag != 1) */
  GR_ERR_SMEM = -5        /* track slice does not fit in shared memory */
} GrStatus;

/* Task constants: QD/racing_ctbr_env.py, QD/mdp/dynamics/dynamics.yaml,
 * L/controllers/controller_diff_cfg.py (see generalizableracing_b200/config.py for
 * the line of each field). */
typedef struct GrConfig {
  float dt;                 /* sim.dt * decimation = 0.03 */
  int32_t max_episode_length;
  float gravity;            /* 9.81 */
  float grad_decay;         /* 0.92 */
  float inertia[3];         /* diag J */
  float action_scale0;      /* m*g*ratio/2 (thrust scale == offset) */
  float body_rate_bound;    /* 6 */
  float thrust_lo, thrust_hi;        /* gross_thrust_bound (-4.65, 86.83) */
  float update_threshold;   /* 0.35 */
  /* drag model */
  float drag1, drag1_rand, drag2, drag2_rand, z_drag, z_drag_rand;
  int32_t random_drag;
  float thr_err_reset_std, thr_err_init_std;
  /* reset sampler */
  float default_pos[3];
  float reset_pos, reset_roll_pitch, reset_yaw, reset_vel;
  /* startup DR + nominal controller constants */
  float kp[3], kd[3], thrust_delay, torque_delay[3];
  float pid_scale_lo, pid_scale_span, delay_scale_lo, delay_scale_span;  /* span = hi - lo formed in double, as torch does */
  float mass;
  int32_t max_init_level;
  /* terminations */
  int32_t term_oob, term_bad_pose;
  float oob_lo, oob_hi;
  /* rewards: progress, bodyrate, action_rate, perception, success_cross, bad_pose */
  float w_reward[GR_NUM_REWARD_TERMS];
  /* command noise + curricula */
  int32_t add_cmd_noise;
  float cmd_noise_pos, cmd_noise_yaw;
  int32_t level_up_gates, level_down_gates;
  int32_t noise_curriculum, noise_up_gates, noise_down_gates;
  float noise_up, noise_down;
  /* BPTT losses: target, vel, fall */
  float w_loss[3];
  /* observation noise */
  float obs_vel_noise, obs_euler_noise;
} GrConfig;

/* Gate table (QD/mdp/commands.py:188,272; L/terrains/terrain_importer.py:47-55), packed
 * for shared-memory staging: rows[(type*levels + level)*(gates+1) + k] is a float4:
 *   k == 0 : env origin xyz of tile (level,type) | next_gate_id as int32 bits
 *   k >= 1 : position of gate k-1 relative to the tile origin | unused            */
typedef struct GrTrack {
  const float* rows;        /* [types*levels*(gates+1)*4] floats, 16-byte aligned */
  int32_t types, levels, gates;
} GrTrack;

/* Env state: array of 32-env tiles; a tile is GR_TILE_PLANES planes x 32 lanes x float4 (8 KB contiguous), i.e.
 * plane p of env i is the float4 at planes[((i/32)*GR_TILE_PLANES + p)*32 + i%32].  One warp owns one tile, so its
 * loads are a single contiguous block and warps finish loading one after the other (plane ids: layout.py). */
typedef struct GrState {
  float* planes;            /* 16-byte aligned */
  int64_t plane_stride;     /* env capacity of the buffer = 32 * number of tiles, >= num_envs rounded up to 32 */
  int32_t num_envs;
  int32_t num_planes;       /* GR_NUM_PLANES (episode sums off) or GR_NUM_PLANES_WITH_STATS (on); tiles always hold 16 */
  int32_t env_id_offset;    /* global id of env 0 of this shard (Philox key; multi-GPU) */
  int32_t max_types_per_block; /* host-computed bound of distinct terrain types inside one 256-env span */
  int32_t block_threads;    /* threads per block of the env kernels: 0 => default (64); multiple of 32, <= 256 */
  int32_t launch_flags;     /* GR_LAUNCH_* */
  const int32_t* chunk_types;  /* [ceil(N/64)][2] (lowest, highest) terrain type of each 64-env chunk;
                                  written by gr_env_startup, read by every kernel that stages the track */
} GrState;

#define GR_LAUNCH_PDL 1   /* gr_step_fwd: programmatic dependent launch -- the kernel's prologue (Philox, gate-table staging)
                            overlaps the tail of the previous kernel in the stream; it waits before touching any buffer */

#define GR_LAUNCH_PREFETCH 2   /* with GR_LAUNCH_PDL: fetch the read-mostly planes before the grid dependency (see racing_step.cu);
                                 clear it for the first step after the HOST rewrote planes 9..15 of the state */
#define GR_LAUNCH_PREFETCH_L2 4 /* with GR_LAUNCH_PDL, instead of GR_LAUNCH_PREFETCH: before the grid dependency the read-mostly planes are
                                 only pulled into L2 (prefetch.global.L2, no registers, no staleness protocol); every plane is then
                                 loaded after the wait, the read-mostly ones as L2 hits.  Safe after host edits of any plane. */

#define GR_LAUNCH_COOP_RESET 16 /* gr_step_fwd / gr_rollout_fwd, in-kernel Philox: the seven Philox calls of a resetting env are generated by seven lanes
                                 of its warp instead of inside the env's own divergent reset tail (same counters, same bits) */
#define GR_LAUNCH_EARLY_STORE 8 /* gr_step_fwd: write the state planes back before the observation section instead of at the end of the kernel */

/* Random source: dense tensor (parity mode) or in-kernel Philox4x32-10 (throughput mode). */
typedef struct GrRandom {
  const float* rnd;         /* [num_envs, GR_RND_STRIDE] or NULL => Philox */
  uint64_t seed;
  uint32_t step;            /* Philox counter word 1: the caller increments it every call */
} GrRandom;

/* One env.step(): replaces ManagerBasedDiffRLEnv.step (L/envs/manager_based_diff_rl_env.py:160-267)
 * = DiffActionManager.process_action (L/managers/action_manager.py:31-52), DiffActions.process_actions
 * (QD/mdp/diff_action.py:156-206), CTBRController.compute (L/controllers/controller_diff.py:120-138),
 * DroneDynamics.step/align (QD/mdp/dynamics/droneDynamics.py:119-135,156-181), terminations
 * (QD/mdp/termination.py:15-33), rewards (QD/mdp/rewards.py:154-253), _reset_idx (:362-410) with
 * reset_root_state_racing (QD/mdp/events.py:139-177) and curricula (QD/mdp/curriculums.py:25-54),
 * RacingCommand (QD/mdp/commands.py:208-350), losses (QD/mdp/losses.py:72-117) and observations
 * (QD/mdp/observation.py:22-104).  Optional outputs may be NULL. */
typedef struct GrStepIO {
  const float* action;      /* [N,4] */
  float* obs;               /* [N,16] policy observation (noisy)                    required */
  float* critic_obs;        /* [N,16] privileged observation                         optional */
  float* aux_obs;           /* [N]    cross_obs                                       optional */
  float* reward;            /* [N]                                                   required */
  uint8_t* terminated;      /* [N]                                                   required */
  uint8_t* time_out;        /* [N]                                                   required */
  int64_t* dones;           /* [N] (terminated|time_out) as int64, RslRlVecEnvWrapper optional */
  float* reward_terms;      /* [N,6] unweighted-by-dt step rewards (RewardManager._step_reward) optional */
  uint8_t* gate_passed;     /* [N] gate switched in this step (diagnostic)           optional */
  float* loss;              /* [N] BPTT loss (extras["losses"])                      optional */
  float* loss_terms;        /* [N,3] weighted loss terms                             optional */
  float* tape;              /* [tiles][GR_TAPE_PLANES][32] float4 of THIS step       optional (BPTT) */
  int64_t tape_stride;      /* env capacity of one tape step = 32 * tiles */
  uint64_t* phase_times;    /* [num_warps][5] %globaltimer stamps; only written by a -DGR_PHASE_TIMING build (tools/)  optional */
  float* log_accum;         /* [GR_LOG_SHARDS][GR_LOG_SLOTS] float atomics (sum the shards), see GR_LOG_*  optional */
  float* aligned_states;    /* [N,13] extras["aligned_states"] = extras["nominal_states"] in value (closure A.1): local position 3 |
                               quaternion 4 | world linear velocity 3 | world angular velocity 3 of the step BEFORE any reset
                               (L/envs/manager_based_diff_rl_env.py:205-212)                                  optional */
  float* acc;               /* [N,3] extras["acc"]: world linear acceleration of the step (droneDynamics.py:126) optional */
  uint8_t* dones_u8;        /* [N] terminated | time_out as one byte (RolloutStorage keeps dones as bytes)   optional */
  float* pre_reset_pos;     /* [N,3] world position after this step's physics, BEFORE any reset: the state the reward terms see
                               (input of collision_penalty_custom, QD/mdp/rewards.py:226-242 -> gr_uav_collision_ray)  optional */
  float* pre_reset_quat;    /* [N,4] attitude (w,x,y,z) at the same instant                                           optional */
} GrStepIO;

#define GR_LOG_NUM_RESET 0          /* number of envs reset in this step                          */
#define GR_LOG_SUM_GATES 1          /* sum of accumulate_gates of reset envs                      */
#define GR_LOG_SUM_EPSUM 2          /* +k: sum over reset envs of episode sum of reward term k (6) */
#define GR_LOG_NUM_TIMEOUT 8
#define GR_LOG_NUM_TERMINATED 9
/* RacingCommand metrics logged by CommandTerm.reset (QD/mdp/commands.py:258-260): sums over the reset envs of the value the
 * previous step's command update assigned (0 for an env whose last command update was a full reset()) */
#define GR_LOG_SUM_ACTION_RATE 10   /* command_rate_penalty: kept per env as a 16-bit float, rel. error <= 2^-12 */
#define GR_LOG_SUM_LIN_SPD 11       /* |v_w| */
#define GR_LOG_SUM_ANG_SPD 12       /* |omega_b| */
#define GR_LOG_SUM_LOSS 13          /* +k: sum over reset envs of the LossManager episode sum of loss term k (3), differentiable physics
                                       with episode sums on (L/managers/loss_manager.py:71-78); 0 otherwise, as in the reference */
#define GR_LOG_SLOTS 16
#define GR_LOG_SHARDS 256          /* accumulator rows (64 B each), picked by warp id: spreads the RED traffic over L2 */

int gr_abi_version(void);

/* Construction-time domain randomisation: randomize_rate_controller_gain_and_thrust_delay
 * (QD/mdp/events.py:105-137), DiffActions.__init__ thr_est_error (QD/mdp/diff_action.py:86),
 * DroneDynamics.__init__ drag tables (droneDynamics.py:23-34), terrain type/level assignment.
 * srnd: [N, GR_SRND_STRIDE] or NULL => Philox.  terrain_types: [N] int32 (non-decreasing). */
int gr_env_startup(const GrConfig* cfg, const GrTrack* track, const GrState* st,
                   const int32_t* terrain_types, int32_t* chunk_types_out, const float* srnd, uint64_t seed,
                   void* stream);

/* ManagerBasedRLEnv.reset(): _reset_idx(all envs) then observations.  reset_mask NULL => all. */
int gr_env_reset(const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng,
                 const uint8_t* reset_mask, float* obs, float* critic_obs, float* aux_obs, void* stream);

/* observation_manager.compute() alone (RslRlVecEnvWrapper.get_observations). */
int gr_env_observe(const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng,
                   float* obs, float* critic_obs, float* aux_obs, void* stream);

int gr_step_fwd(const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng,
                const GrStepIO* io, void* stream);

/* Reverse sweep of the BPTT window over tape steps [t_begin, t_end) (analytic backward of
 * DroneDynamics.step/align + CTBRController.compute + the tanh action map; replaces
 * torch.autograd through QD/mdp/dynamics/droneDynamics.py:119-181, L/controllers/
 * controller_diff.py:120-138, QD/mdp/diff_action.py:160-176 as driven by S/diff_rl/algorithms/bptt.py:38-44).
 *   tape        [T][tiles][GR_TAPE_PLANES][32] float4 written by gr_step_fwd (tape_stride = 32*tiles)
 *   grad_loss   [T][N] upstream gradient of extras["losses"], or NULL => uniform `grad_scale`
 *   adjoint     [5][adj_stride] float4: carried adjoints (zero at the end of the window); in/out
 *   grad_action [T][N][4]: row t-1 receives dL/da_{t-1} produced by step t (1-step action lag);
 *               rows are OVERWRITTEN for t-1 in [t_begin-1, t_end-2]; row T-1 is not touched. */
typedef struct GrBwdIO {
  const float* tape;
  int64_t tape_stride;
  int32_t t_begin, t_end;
  const float* grad_loss;
  float grad_scale;
  float* adjoint;
  int64_t adj_stride;
  float* grad_action;
} GrBwdIO;
int gr_step_bwd(const GrConfig* cfg, const GrState* st, const GrBwdIO* io, void* stream);

/* ---- T consecutive env.step() calls in ONE launch for actions known in advance ("open-loop" window) -----------------------
 * What the loop `for t in range(T): env.step(actions[t])` does -- the rollout loop of AlgoRunner.learn
 * (standalone/diff_rl/algorithms/runner.py:110-126) / a play-back of recorded actions -- with the env state held in registers
 * over the window: the state planes are read once and written once per WINDOW, a step costs its action (16 B) plus whatever it
 * is asked to record.  Every step is gr_step_fwd's racing_step_body (bit-identical results, tests/test_rollout_window.py); step t
 * uses the random stream (seed, env, rng->step + t) or, in dense mode, rng->rnd + t * N * GR_RND_STRIDE.  With loss / tape set the
 * window is differentiable and gr_step_bwd sweeps it exactly as it sweeps T gr_step_fwd launches. */
typedef struct GrRolloutIO {
  const float* actions;              /* [T,N,4] */
  float* obs_out;                    /* [N,16] policy observations after the last step */
  float* critic_obs_out;             /* [N,16] optional */
  float* aux_out;                    /* [N] optional */
  float* obs_seq;                    /* [T,N,16] optional: policy observations after every step */
  float* reward;                     /* [T,N] optional */
  uint8_t* dones;                    /* [T,N] optional (terminated | time_out) */
  uint8_t* terminated;               /* [T,N] optional */
  uint8_t* time_out;                 /* [T,N] optional */
  float* loss;                       /* [T,N]   differentiable physics: extras["losses"] of every step */
  float* loss_terms;                 /* [T,N,3] optional */
  float* tape;                       /* [T][tiles][GR_TAPE_PLANES][32] float4, as gr_step_fwd writes it */
  int64_t tape_stride;               /* env capacity of one tape step = 32 * tiles */
  float* log_accum;                  /* optional */
  int32_t T;
} GrRolloutIO;
int gr_rollout_fwd(const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng, const GrRolloutIO* io, void* stream);

/* Dense random tensor exactly as the in-kernel Philox path would draw it (parity chain). */
int gr_fill_rand(float* rnd, int32_t num_envs, int32_t env_id_offset, uint64_t seed, uint32_t step, void* stream);
/* Self-test hook: y[i] = the device square root the gate predicate `|gate - pos| < update_threshold` (QD/mdp/commands.py:308-312) is
 * evaluated with.  It must be correctly rounded (IEEE) whatever -prec-sqrt / -prec-div the library is built with; tests/test_misc.py compares
 * it bit for bit with numpy. */
int gr_selftest_sqrt_rn(const float* x, float* y, int64_t n, void* stream);
int gr_fill_startup_rand(float* srnd, int32_t num_envs, int32_t env_id_offset, uint64_t seed, void* stream);

/* ---- rollout storage (S/rsl_rl/ext/storage/rollout_storage.py) ---- */

/* add_transitions (:71-88) fused with the time-out bootstrap of PPO.process_env_step
 * (S/rsl_rl/ext/algorithms/ppo.py:85-97): one launch copies one transition into slot `step`. */
typedef struct GrTransition {
  const float* obs; const float* critic_obs; const float* actions; const float* rewards;
  const void* dones; int32_t dones_is_int64;         /* uint8 or int64 [N] */
  const float* values; const float* log_prob; const float* mu; const float* sigma;
  const uint8_t* time_outs; float gamma;             /* NULL => no bootstrap */
} GrTransition;
typedef struct GrStorage {
  float* obs; float* critic_obs; float* actions; float* rewards; uint8_t* dones;
  float* values; float* log_prob; float* mu; float* sigma; float* returns; float* advantages;
  int32_t T, N, obs_dim, critic_dim, act_dim;
} GrStorage;
int gr_storage_add(const GrStorage* s, const GrTransition* tr, int32_t step, void* stream);

/* compute_returns (:113-127): GAE + advantage normalisation (unbiased std + 1e-8).
 * scratch: >= gr_gae_scratch_bytes(N) bytes.  If moments != NULL the un-normalised
 * (count, mean, M2) of this shard are written there as 3 doubles and, when
 * `normalize` == 0, normalisation is left to gr_advantage_normalize (multi-GPU: all-reduce
 * the moments in between). */
int64_t gr_gae_scratch_bytes(int32_t N);
int gr_compute_returns(const GrStorage* s, const float* last_values, float gamma, float lam,
                       void* scratch, double* moments, int32_t normalize, void* stream);
int gr_advantage_normalize(const GrStorage* s, const double* moments, void* stream);

/* mini_batch_generator gather (:179-187): one launch gathers all nine fields for one mini-batch.
 * indices: [B] int64 into the flattened [T*N] transitions. */
typedef struct GrMiniBatch {
  float* obs; float* critic_obs; float* actions; float* values; float* advantages; float* returns;
  float* log_prob; float* mu; float* sigma;
} GrMiniBatch;
int gr_storage_gather(const GrStorage* s, const int64_t* indices, int32_t B, const GrMiniBatch* out, void* stream);

/* ---- fused PPO collection (policy MLP on the tensor cores + env.step + add_transitions in one launch) --------------
 * Replaces the rollout loop of OnPolicyRunner.learn (S/rsl_rl/ext/runners/on_policy_runner.py:141-175) for the racing
 * task's state-only ActorCritic (QD/agents/rsl_rl_ppo_cfg.py:22-27: 16 -> 128 -> 128 -> 4 / 1, leaky relu, scalar std):
 * T = storage->T steps of PPO.act (S/rsl_rl/ext/algorithms/ppo.py:71-83), env.step, PPO.process_env_step (:85-97) and
 * RolloutStorage.add_transitions (S/rsl_rl/ext/storage/rollout_storage.py:71-88), then last_values = V(last critic obs)
 * (ppo.py:99-100).  Env semantics are those of gr_step_fwd bit for bit (Philox draws of step rng->step + t); the MLPs
 * run with fp16 operands and fp32 accumulation on tcgen05 (measured error vs the fp32 torch modules: DESIGN.md). */
#define GR_PHILOX_CALL_ACTION 16    /* Philox call index of the 4 action-noise normals (calls 0..12 belong to env.step) */
typedef struct GrMlp {               /* torch.nn.Linear parameters, row-major [out][in] fp32, device pointers */
  const float* w1; const float* b1;  /* [hidden, in_dim], [hidden] */
  const float* w2; const float* b2;  /* [hidden2, hidden], [hidden2] */
  const float* w3; const float* b3;  /* [out_dim, hidden2], [out_dim] */
  int32_t in_dim, hidden, hidden2, out_dim;   /* 16; first / second hidden width: (128,128) or (256,128); 1..4 */
} GrMlp;
typedef struct GrPolicy {
  const void* packed;                /* gr_policy_packed_bytes() bytes written by gr_policy_pack (actor, then critic) */
  const float* sigma;                /* device [4]: ActorCritic.std (or exp(log_std)); 16-byte aligned */
  float negative_slope;              /* leaky-relu slope in [0,1] (nn.LeakyReLU default 0.01; 0 = relu) */
} GrPolicy;
typedef struct GrCollectIO {
  const float* obs0;                 /* [N,16] policy observation the rollout starts from (last step's output) */
  const float* critic_obs0;          /* [N,16] */
  float* obs_out;                    /* [N,16] observations after the last step (must not alias obs0) */
  float* critic_obs_out;             /* [N,16] */
  float* aux_out;                    /* [N] optional */
  float* last_values;                /* [N] V(critic_obs_out) */
  float* episode_acc;                /* [N,2] running (reward sum, length) of each env's current episode; in/out, optional */
  float* log_accum;                  /* [GR_LOG_SHARDS][GR_LOG_SLOTS], optional (see GR_LOG_*) */
  float* episode_log;                /* [GR_LOG_SHARDS][4] float atomics, optional: per finished episode += (reward sum, length, 1, 0)
                                        -- the runner's rewbuffer / lenbuffer book keeping (on_policy_runner.py:160-173) */
  float gamma;                       /* time-out bootstrap r += gamma * V(s_t) * time_out */
  int32_t groups_per_cta;            /* 128-env tiles per thread block: 1, 2, 4 or 0 = pick (fewest that fit one wave) */
  int32_t coop_reset_columns;        /* warp-cooperative Philox draws for resetting envs: staging columns per warp (112 B each); 0 = as many as
                                        the shared memory left by weights / activation tiles / gate table holds (<= 4), -1 = off (each
                                        resetting lane draws for itself), k > 0 = at most k.  Same counters, same bits either way. */
} GrCollectIO;
/* packed size of `nets` (1 = actor only, 2 = actor + critic) MLPs of widths 16 -> hidden -> hidden2; < 0: unsupported widths */
int64_t gr_policy_packed_bytes(int32_t hidden, int32_t hidden2, int32_t nets);
/* critic may be NULL (actor only); both nets must have the same widths */
int gr_policy_pack(const GrMlp* actor, const GrMlp* critic, void* packed, void* stream);
int gr_ppo_collect(const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng, const GrPolicy* policy,
                   const GrStorage* storage, const GrCollectIO* io, void* stream);

/* ---- fused BPTT window (policy MLP on the tensor cores + differentiable env.step with tape, one launch per window) -------
 * Replaces the forward half of AlgoRunner.learn's window (S/diff_rl/algorithms/runner.py:110-126): T x [BaseModel.act =
 * actor MLP + rsample (S/diff_rl/algorithms/model.py:63-99), env.step with losses (QD/mdp/losses.py:72-117)].  The tape,
 * the per-step losses and everything env.step does are those of gr_step_fwd; the reverse sweep stays gr_step_bwd.  The
 * kernel records what the policy's backward needs -- the observation and the noise each action was computed from -- so
 * that ONE batched torch forward/backward of the actor over [T*N] rows (action = actor(obs) + sigma * eps, cotangent =
 * gr_step_bwd's grad_action) replaces T small autograd graphs.  Actor widths (128,128) or (256,128)
 * (QD/agents/diff_rl_naive_cfg.py:26-32); policy inference with fp16 operands / fp32 accumulation as in gr_ppo_collect. */
typedef struct GrBpttCollectIO {
  const float* obs0;                 /* [N,16] observation the window starts from */
  float* obs_out;                    /* [N,16] observations after the last step (must not alias obs0) */
  float* critic_obs_out;             /* [N,16] */
  float* aux_out;                    /* [N] optional */
  float* obs_seq;                    /* [T,N,16] observation behind action t */
  float* eps_seq;                    /* [T,N,4]  standard-normal draw behind action t (Philox call GR_PHILOX_CALL_ACTION) */
  float* actions;                    /* [T,N,4]  the actions applied (mu_fp16 + sigma * eps), optional */
  float* loss;                       /* [T,N]    extras["losses"] of every step */
  float* loss_terms;                 /* [T,N,3]  optional */
  float* reward;                     /* [T,N]    optional */
  uint8_t* dones;                    /* [T,N]    optional */
  float* tape;                       /* [T][tiles][GR_TAPE_PLANES][32] float4, as gr_step_fwd writes it */
  int64_t tape_stride;               /* env capacity of one tape step = 32 * tiles */
  float* log_accum;                  /* optional */
  int32_t T;
  int32_t groups_per_cta;            /* 1, 2 (4 for width 128) or 0 = pick */
} GrBpttCollectIO;
int gr_bptt_collect(const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng, const GrPolicy* policy,
                    int32_t hidden, int32_t hidden2, const GrBpttCollectIO* io, void* stream);

/* ---- actor backward (weight gradients of the BPTT actor on the tensor cores) ------------------------------------------
 * d(loss)/d(parameters) of `actions = actor(obs) (+ sigma * eps)` over `rows` (observation, cotangent) pairs, cotangent =
 * d(loss)/d(action): replaces loss.backward() through BaseModel.actor (S/diff_rl/algorithms/bptt.py:38-44 over the graphs of
 * S/diff_rl/algorithms/model.py:63-99) as FusedBpttCollector batches it.  The activations are recomputed from `obs` with the
 * packed fp16 weights (as gr_bptt_collect evaluated them); gradients are ACCUMULATED into `out` (zero it first).
 * `scale`: device scalar; the cotangent is multiplied by it before the fp16 conversion and the result divided by it
 * (pick ~1024 / max|grad_actions|). */
typedef struct GrMlpGrad {           /* shapes as GrMlp, fp32; w3 [out_dim, hidden2]; w1 / w2 16-byte aligned */
  float* w1; float* b1; float* w2; float* b2; float* w3; float* b3;
  int32_t out_dim;
  int32_t scale_is_maxabs;           /* 1: `*scale` holds max|grad_actions| (gr_ppo_loss_grad's sums[8|9]); the kernel uses 1024 / it */
} GrMlpGrad;
int gr_actor_backward(const GrPolicy* policy, int32_t hidden, int32_t hidden2, const float* obs /* [rows,16] */,
                      const float* grad_actions /* [rows,4] */, const float* scale, int64_t rows, const GrMlpGrad* out, void* stream);
/* one launch for up to two nets of the same widths over the same number of rows (a PPO step's actor and critic) */
/* indices (optional, [rows] int64): row r reads obs[indices[r]] -- the mini-batch gather of rollout_storage.py:179-187 done on load */
typedef struct GrBackwardJob { GrPolicy policy; const float* obs; const float* grad_actions; const float* scale; GrMlpGrad out; const int64_t* indices;
                               int32_t obs_stride;   /* floats between consecutive observation rows: 0 = 16 (dense [rows,16]); GR_RECORD_FLOATS when `obs`
                                                        points into transition records (+16 floats for the critic's observation) */
} GrBackwardJob;
int gr_actor_backward_jobs(const GrBackwardJob* jobs, int32_t n_jobs, int32_t hidden, int32_t hidden2, int64_t rows, void* stream);

/* ---- PPO update on the kernels (forward, loss gradients; the weight gradients come from gr_actor_backward) -----------------
 * One mini-batch step of PPO.update (S/rsl_rl/ext/algorithms/ppo.py:118-171) without autograd:
 *   gr_policy_forward : mu = actor(obs), v = critic(critic_obs) for `rows` rows (packed fp16 nets, tensor cores)
 *   gr_ppo_loss_grad  : clipped surrogate + clipped value loss + entropy bonus of the batch -> d(loss)/d(mu) [rows,4],
 *                       d(loss)/d(v) [rows,4] (column 0), sums[0..7] += (surrogate, value loss, KL, d/d(std) x 4, rows) and
 *                       sums[8], sums[9] = max(.., max|d/d(mu)|), max(.., max|d/d(v)|)   (sums: 16 floats, zero them first)
 * followed by two gr_actor_backward launches (actor with d/d(mu), critic with d/d(v)). */
int gr_policy_forward(const GrPolicy* policy /* packed actor + critic, widths (128,128) */, const float* obs, const float* critic_obs,
                      float* mu /* [rows,4] */, float* value /* [rows] */, int64_t rows, void* stream);
/* the same with the mini-batch gather on load: row r evaluates obs[indices[r]] / critic_obs[indices[r]] (indices [rows] int64 into the flattened
 * [T*N] transitions of the rollout storage, as gr_storage_gather takes them); mu / value are dense [rows] */
int gr_policy_forward_gather(const GrPolicy* policy, const float* obs, const float* critic_obs, const int64_t* indices, float* mu, float* value,
                             int64_t rows, void* stream);
typedef struct GrPpoBatch {
  const float* mu; const float* value;                 /* current policy outputs [rows,4], [rows] */
  const float* sigma;                                  /* device [4]: current action std */
  const float* actions; const float* old_log_prob; const float* advantages; const float* returns; const float* old_values;
  const float* old_mu; const float* old_sigma;         /* [rows,4] each */
  float clip_param, value_loss_coef, entropy_coef;
  int32_t use_clipped_value_loss;
  const int64_t* indices;                              /* optional [rows] int64: the STORED columns (actions .. old_sigma) are read at row indices[r]
                                                          (the mini-batch gather done on load); mu / value and the gradients stay dense */
  const float* records;                                /* optional (gr_policy_forward_loss, gr_ppo_fused_step): transition records (gr_storage_pack_records); when set,
                                                          obs / critic_obs and the stored columns of row r all come from record indices[r] */
} GrPpoBatch;
int gr_ppo_loss_grad(const GrPpoBatch* batch, int64_t rows, float* grad_mu /* [rows,4] */, float* grad_value /* [rows,4] */,
                     float* sums /* [16] accumulated */, void* stream);
/* gr_policy_forward_gather + gr_ppo_loss_grad as ONE launch: the thread that holds a row's mean and value evaluates the row's loss on the stored
 * columns it fetched with the row (batch->indices gathers on load).  batch->mu / batch->value: optional OUTPUTS here (dense [rows,4] / [rows]; NULL:
 * the policy outputs are never written); grad_mu / grad_value / sums exactly as gr_ppo_loss_grad writes them (bit-identical gradients). */
int gr_policy_forward_loss(const GrPolicy* policy, const float* obs, const float* critic_obs, const GrPpoBatch* batch, int64_t rows,
                           float* grad_mu /* [rows,4] */, float* grad_value /* [rows,4] */, float* sums /* [16] accumulated */, void* stream);
/* The three launches above (gr_policy_forward_gather, gr_ppo_loss_grad, gr_actor_backward_jobs) as ONE: per 128-row tile and net the kernel
 * recomputes the activations, takes the head (layer 3) on them, evaluates the row's loss gradient in registers and accumulates the weight
 * gradients -- mu / value / d(loss)/d(mu) / d(loss)/d(v) never exist in memory.  batch.mu / batch.value are ignored; batch.indices (optional)
 * gathers obs / critic_obs and the stored columns on load.  actor_grad / critic_grad accumulate (zero them first, out_dim 4 / 1,
 * scale_is_maxabs ignored); sums[0..7] accumulate as gr_ppo_loss_grad's (sums[8..9] are not written).  The cotangent rows go to fp16 as
 * cotangent_scale x the un-normalised per-row gradient (0 => 1/64; the accumulators are divided by rows * cotangent_scale at the flush). */
/* Transition records: the columns one PPO mini-batch row needs, side by side, GR_RECORD_FLOATS floats per transition of the [T*N] storage:
 * [0,16) policy obs | [16,32) critic obs | [32,36) action | [36,40) old mean | [40,44) old std | 44 old log-prob | 45 advantage | 46 return |
 * 47 old value.  Packed once per PPO iteration (after gr_compute_returns); with GrPpoBatch.records / GrBackwardJob.obs_stride set the update
 * kernels read ONE scattered 192-byte record per sampled row instead of nine scattered columns.  Needs obs_dim = critic_dim = 16, act_dim = 4. */
#define GR_RECORD_FLOATS 48
int gr_storage_pack_records(const GrStorage* s, float* records /* [T*N][GR_RECORD_FLOATS], 16-byte aligned */, void* stream);
/* The same with the iteration's mini-batch permutation applied while packing: record r < num holds transition perm[r] (perm [num <= T*N] int64, the
 * torch.randperm of rollout_storage.py:165 -- ONE permutation per update, reused by every epoch), so mini-batch i of every epoch is the
 * contiguous record range [i*mb, (i+1)*mb): pass records + i*mb*GR_RECORD_FLOATS with indices = NULL to the update kernels, which then
 * stream their rows instead of gathering them (same rows in the same order: bit-identical results). */
int gr_storage_pack_records_permuted(const GrStorage* s, const int64_t* perm, int64_t num, float* records, void* stream);
typedef struct GrPpoStep {
  GrPolicy policy;                     /* packed actor + critic, widths (128,128) */
  const float* obs; const float* critic_obs;
  GrPpoBatch batch;
  GrMlpGrad actor_grad, critic_grad;
  float* sums;                         /* device [16] */
  float cotangent_scale;
} GrPpoStep;
int gr_ppo_fused_step(const GrPpoStep* step, int64_t rows, void* stream);

/* Gradient clipping + Adam + the KL-adaptive learning rate of one optimiser step (nn.utils.clip_grad_norm_, torch.optim.Adam.step
 * and ppo.py:124-141), two launches over a flat view of the parameters: replaces ~40 element-wise torch launches of the
 * captured step.  grad / exp_avg / exp_avg_sq are flat buffers indexed alike (tensor k occupies [seg_offsets[k], seg_offsets[k] +
 * seg_sizes[k])); state: device [16] = {lr, step, clip_coef, step_size, 1/sqrt(bias2), sum value loss, sum surrogate, -, sumsq, counter..}
 * (initialise lr and step, zero the rest).  kl_stats: device {.., .., KL sum, .., .., .., .., rows} = gr_ppo_loss_grad's sums, or NULL. */
typedef struct GrAdamStep {
  const int64_t* param_ptrs;         /* device [n_seg]: address of each parameter tensor (fp32, contiguous) */
  const int32_t* seg_offsets;        /* device [n_seg] */
  const int32_t* seg_sizes;          /* device [n_seg] */
  int32_t n_seg, n_flat;
  float* grad; float* exp_avg; float* exp_avg_sq;
  float* state;
  const float* kl_stats;
  float grad_scale;                  /* gradients are multiplied by it first (1 / world size after a sum all-reduce) */
  float beta1, beta2, eps, max_grad_norm, desired_kl, lr_min, lr_max;
} GrAdamStep;
int gr_adam_clip_step(const GrAdamStep* a, void* stream);

/* ---- the policy-gradient all-reduce as one kernel over NVLink peer memory (csrc/peer_reduce.cu) ---------------------------------------
 * BASELINE config C5's exchange step (env-sharded training: one sum of the flat gradient buffer per optimiser step).  Every rank's buffer
 * lives in symmetric memory (mapped into every rank of the node); out[i] = sum over ranks 0..world-1, in that order on every rank (the
 * replicas get the same bits), of peer_bufs[r][i].  Two flag barriers bracket the reads (all gradients complete / all ranks have read), so
 * when the kernel ends this rank's buffer may be overwritten.  Call it on every rank of the group, once per step, in stream order after the
 * kernels that wrote the gradients; it can be captured in a CUDA graph.  Waits are bounded: after max_spins polls *error = 1 and the
 * kernel returns (results undefined) instead of hanging. */
#define GR_PEER_MAX_WORLD 16
typedef struct GrPeerReduce {
  const void* peer_bufs;             /* device [world] uint64: address, in THIS process, of rank r's gradient buffer [n] fp32 (16-byte aligned) */
  const void* peer_flags;            /* device [world] uint64: address of rank r's flag pad, uint32 [2][GR_PEER_MAX_WORLD], zero-initialised once */
  int32_t world, rank;
  int32_t n;                         /* floats, multiple of 4 */
  int32_t reserved;
  int64_t max_spins;                 /* polls (32 ns apart) before a wait gives up */
  uint32_t* epoch;                   /* device, local: launch counter (zero-initialised once, same on every rank) */
  uint32_t* counter;                 /* device, local: block counter (zero-initialised once) */
  int32_t* error;                    /* device, local: set to 1 when a wait gave up */
} GrPeerReduce;
int gr_peer_allreduce(const GrPeerReduce* a, float* out /* device, local [n], 16-byte aligned */, void* stream);

/* ---- env.step() with HOST buffers (the e2e boundary) -------------------------------------------------------------
 * Same call as gr_step_fwd for a caller whose actions / observations live in host memory: replaces
 * `env.step(torch.as_tensor(a).to(device))` + `.cpu()` of obs / reward / dones around ManagerBasedDiffRLEnv.step
 * (L/envs/manager_based_diff_rl_env.py:160-267) as a host-side consumer of RslRlVecEnvWrapper.step would write it.
 * The pipe owns `depth` device staging slots, two copy streams and its events (the ONLY objects this library ever
 * allocates; freed by gr_host_pipe_destroy).  gr_host_pipe_step enqueues H2D(actions) -> step kernel -> D2H(results)
 * and returns at once with a ticket; the host buffers of ticket t are valid after gr_host_pipe_wait(t) and must stay
 * untouched until then.  Up to `depth` steps are in flight, so the copies of step t overlap the kernel of step t+1.
 * Host buffers should be page-locked (cudaHostAlloc / torch pin_memory), otherwise the copies serialise. */
#define GR_HOST_PIPE_MAX_DEPTH 4
typedef struct GrHostPipe GrHostPipe;
typedef struct GrHostStep {
  const float* action;      /* host [N,4]                                    required */
  float* obs;               /* host [N,16]                                   required */
  float* reward;            /* host [N]                                      required */
  int64_t* dones;           /* host [N]                                      optional */
  float* critic_obs;        /* host [N,16]                                   optional */
  uint8_t* time_out;        /* host [N]                                      optional */
  uint8_t* dones_u8;        /* host [N] terminated | time_out as one byte    optional (1 B per env over PCIe instead of 8) */
  int32_t outputs_contiguous; /* 1: obs | reward | dones are ONE host allocation, back to back (reward == obs + 16 N floats, dones == reward + N
                                 floats, N even): they travel as one device->host copy.  (Adjacent addresses alone do not qualify: a copy must
                                 not span separately pinned allocations.) */
} GrHostStep;
/* `action` must stay untouched until gr_host_pipe_wait(ticket) of that step returned; one calling thread per pipe;
 * gr_host_pipe_destroy synchronises the three streams before it frees the staging buffers. */
int gr_host_pipe_create(int32_t num_envs, int32_t depth, void* compute_stream, GrHostPipe** out);
int gr_host_pipe_destroy(GrHostPipe* pipe);
int gr_host_pipe_step(GrHostPipe* pipe, const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng,
                      const GrHostStep* host, float* log_accum /* device, optional */, int64_t* ticket_out);
int gr_host_pipe_wait(GrHostPipe* pipe, int64_t ticket);
/* the copies of one host step alone (actions in; obs, reward and `dones_bytes` bytes of dones per env out), `steps` times, two streams,
 * pinned buffers, no kernel: seconds_out = host time of the loop.  The platform's ceiling for gr_host_pipe_step (bench.py, e2e). */
int gr_host_copy_probe(int32_t num_envs, int32_t steps, int32_t dones_bytes, double* seconds_out);
/* the same with `packed` != 0: obs | reward | dones as ONE device->host copy per step -- what gr_host_pipe_step issues when the caller's three
 * host buffers are contiguous in that order (reward == obs + 16 N floats, dones == reward + N floats; N even): ~7 % less time per step */
int gr_host_copy_probe2(int32_t num_envs, int32_t steps, int32_t dones_bytes, int32_t packed, double* seconds_out);

/* ---- reach-target tasks: the other command modes and tasks sharing the dynamics ----------------
 * QD/reach_target_lv_env.py + QD/reach_target_ctbr_env.py: the same ManagerBasedDiffRLEnv.step ordering as gr_step_fwd
 * with DiffActions driving LVController / PSController (L/controllers/controller_diff.py:172-443) or CTBRController
 * (:37-170), UniformWorldPoseCommand (QD/mdp/commands.py:32-134), the reach-target rewards (QD/mdp/rewards.py:30-101 +
 * Isaac Lab action_rate_l2 / body_lin_acc_l2 / is_terminated), observations (QD/reach_target_lv_env.py:83-104), the
 * reach-target losses (QD/mdp/losses.py:32-67), Isaac Lab reset_root_state_uniform and time_out.  Closure substitutions:
 * oracle/reach_oracle.py R.1-R.5 (contacts := z bounds).  State: 32-env tiles of GR_REACH_PLANES float4 planes. */
#define GR_REACH_OBS_DIM 17
#define GR_REACH_NUM_REWARD_TERMS 10
#define GR_REACH_NUM_LOSS_TERMS 4
#define GR_REACH_RND_STRIDE 24
#define GR_REACH_PLANES 13
#define GR_REACH_TAPE_PLANES 13
#define GR_CTRL_CTBR 0
#define GR_CTRL_LV 1
#define GR_CTRL_PS 2

typedef struct GrReachConfig {
  int32_t controller;       /* GR_CTRL_* = DiffActionCfg.command_type */
  int32_t sim2real_test;    /* QD/mdp/diff_action.py:168-171: raw (a_zb, body-rate) inputs, no tanh, no gradient */
  int32_t last_action_modified;   /* observation term: 0 = mdp.last_action, 1 = modified_last_action (QD/mdp/observation.py:55-63) */
  int32_t random_drag;
  float dt;
  int32_t max_episode_length;
  float gravity, grad_decay, mass;
  float inertia[3];
  float action_scale[4], action_offset[4];    /* QD/mdp/diff_action.py:257-275 */
  float thrust_lo, thrust_hi, body_rate_bound;
  float kp[3], kd[3], thrust_delay, torque_delay[3];                       /* CTBR */
  float speed_gain[3], pose_gain[3], rate_gain[3], pos_gain[3], max_feedback_accel;   /* LV / PS */
  float drag1, drag1_rand, drag2, drag2_rand, z_drag, z_drag_rand;
  float thr_err_reset_std;
  float default_pos[3], reset_lo[6], reset_hi[6];      /* reset_root_state_uniform: x y z roll pitch yaw */
  float cmd_lo[3], cmd_hi[3], resample_time;           /* UniformWorldPoseCommand */
  int32_t term_oob;
  float oob_lo, oob_hi;
  float w_reward[GR_REACH_NUM_REWARD_TERMS];  /* move_towards, orientation, move_in_dir, action_rate, reach_target, smooth_ang_vel,
                                                 smooth_lin_acc, smooth_ang_acc, early_termination, hover_state */
  float move_in_dir_thr, reach_thr, hover_thr, hover_ratio;
  float w_loss[GR_REACH_NUM_LOSS_TERMS];      /* target_diff, orientation_diff, move_in_dir_diff, smooth_vel_diff */
  float loss_dir_thr, loss_smooth_ratio;
} GrReachConfig;

typedef struct GrReachState {
  float* planes;            /* [tiles][GR_REACH_PLANES][32] float4, 16-byte aligned; zero-filled before the first gr_reach_reset */
  int64_t plane_stride;     /* env capacity = 32 * tiles */
  int32_t num_envs;
  int32_t env_id_offset;    /* global id of env 0 of this shard (Philox key) */
} GrReachState;

typedef struct GrReachStepIO {
  const float* action;      /* [N,4] */
  float* obs;               /* [N,17]                                                required */
  float* reward;            /* [N]                                                   required */
  uint8_t* terminated;      /* [N]                                                   required */
  uint8_t* time_out;        /* [N]                                                   required */
  int64_t* dones;           /* [N]                                                   optional */
  float* reward_terms;      /* [N,10] RewardManager._step_reward                     optional */
  float* loss;              /* [N] extras["losses"]                                  optional */
  float* loss_terms;        /* [N,4] weighted loss terms                             optional */
  float* tape;              /* [tiles][GR_REACH_TAPE_PLANES][32] float4 of THIS step optional (BPTT) */
  int64_t tape_stride;
  float* log_accum;         /* [GR_LOG_SHARDS][GR_LOG_SLOTS], GR_REACH_LOG_*         optional */
} GrReachStepIO;

#define GR_REACH_LOG_NUM_RESET 0
#define GR_REACH_LOG_SUM_POS_ERR 1     /* sum over reset envs of Metrics/desired_pos_b/position_error */
#define GR_REACH_LOG_SUM_EPSUM 2       /* +k: episode sum of reward term k (10) */
#define GR_REACH_LOG_NUM_TIMEOUT 12
#define GR_REACH_LOG_NUM_TERMINATED 13

/* ManagerBasedRLEnv.reset() / _reset_idx(mask) then observations.  reset_mask NULL => all envs. */
int gr_reach_reset(const GrReachConfig* cfg, const GrReachState* st, const GrRandom* rng, const uint8_t* reset_mask, float* obs, void* stream);
/* observation_manager.compute() alone. */
int gr_reach_observe(const GrReachConfig* cfg, const GrReachState* st, float* obs, void* stream);
int gr_reach_step_fwd(const GrReachConfig* cfg, const GrReachState* st, const GrRandom* rng, const GrReachStepIO* io, void* stream);

/* T consecutive gr_reach_step_fwd calls in ONE launch for actions known in advance (see gr_rollout_fwd): the state stays in
 * registers over the window, results bit-identical to the T single steps (tests/test_reach_rollout.py).  One state word differs on
 * purpose: plane RPL_ANGACC .w = 2 marks an env whose read-mostly planes were rewritten by a reset in an earlier step of the window,
 * so that the next single step re-reads them after its grid dependency; that step clears the mark. */
typedef struct GrReachRolloutIO {
  const float* actions;     /* [T,N,4] */
  float* obs_out;           /* [N,17] observations after the last step */
  float* obs_seq;           /* [T,N,17] optional: observations after every step */
  float* reward;            /* [T,N] optional */
  uint8_t* dones;           /* [T,N] optional */
  uint8_t* terminated;      /* [T,N] optional */
  uint8_t* time_out;        /* [T,N] optional */
  float* loss;              /* [T,N]   differentiable physics */
  float* loss_terms;        /* [T,N,4] optional (16-byte aligned) */
  float* tape;              /* [T][tiles][GR_REACH_TAPE_PLANES][32] float4 */
  int64_t tape_stride;      /* env capacity of one tape step = 32 * tiles */
  float* log_accum;         /* optional */
  int32_t T;
} GrReachRolloutIO;
int gr_reach_rollout_fwd(const GrReachConfig* cfg, const GrReachState* st, const GrRandom* rng, const GrReachRolloutIO* io, void* stream);
/* Reverse sweep over tape steps [t_begin, t_end): analytic backward of DroneDynamics.step/align, the controller (full 4x4
 * action Jacobian of the LV / PS outer loop, taken in forward mode while the step runs) and the action map; same GrBwdIO
 * contract as gr_step_bwd. */
int gr_reach_step_bwd(const GrReachConfig* cfg, const GrReachState* st, const GrBwdIO* io, void* stream);
/* Dense random tensor [N, GR_REACH_RND_STRIDE] exactly as the in-kernel Philox path draws it. */
int gr_reach_fill_rand(float* rnd, int32_t num_envs, int32_t env_id_offset, uint64_t seed, uint32_t step, void* stream);

/* ---- recurrent mini-batches: RolloutStorage.reccurent_mini_batch_generator
 * (S/rsl_rl/ext/storage/rollout_storage.py:194-254) = rsl_rl.utils.split_and_pad_trajectories (third party rsl-rl-lib 2.x;
 * call sites :197-199) + the hidden-state gather at trajectory starts (:226-237); gr_traj_unpad = rsl_rl.utils.unpad_trajectories.
 * Trajectories are the pieces of every env's [T] column between dones (the last step always closes one), ordered env-major. */
/* dones [T,N] uint8 -> offsets [N+1] (trajectories of envs < n), and per trajectory (capacity T*N each): env, first step, length. */
int gr_traj_index(const uint8_t* dones, int32_t T, int32_t N, int32_t* offsets, int32_t* traj_env, int32_t* traj_start, int32_t* traj_len, void* stream);
/* src [T,N,D] -> padded [T,count,D] (zero padded) and masks [T,count] (optional) for trajectories [first, first+count). */
int gr_traj_pad(const float* src, int32_t T, int32_t N, int32_t D, const int32_t* traj_env, const int32_t* traj_start, const int32_t* traj_len,
                int32_t first, int32_t count, float* padded, uint8_t* masks, void* stream);
/* padded [T,J,D] + masks [T,J] -> out [T,B,D] (B*T valid rows); scratch: 2*J+1 int32. */
int gr_traj_unpad(const float* padded, const uint8_t* masks, int32_t T, int32_t J, int32_t D, int32_t B, int32_t* scratch, float* out, void* stream);
/* saved hidden states [T,L,N,H] -> out [L,count,H]: the state each trajectory started from. */
int gr_traj_hidden(const float* saved, int32_t T, int32_t L, int32_t N, int32_t H, const int32_t* traj_env, const int32_t* traj_start,
                   int32_t first, int32_t count, float* out, void* stream);

/* ---- UAV-vs-terrain-mesh collision count: replaces the reference's only GPU kernel, the Warp kernel
 * check_uav_collision_ray_kernel (L/utils/mesh_tools.py:128-233) and its launcher get_uav_collision_num_ray (:237-295), consumed by the
 * STAGE-0 reward term collision_penalty_custom (QD/mdp/rewards.py:226-242).  The terrain is a triangle mesh (the inputs of wp.Mesh:
 * points [V,3], indices [F,3]); gr_mesh_build_bvh turns it on the HOST into the BVH + face records the device kernels traverse. */
typedef struct GrMesh {
  const float* nodes;       /* device, [num_nodes][8]: (lo.xyz, int32 first face of a leaf | left child), (hi.xyz, int32 face count of a leaf | 0);
                               16-byte aligned; the children of an internal node are adjacent (left, left + 1); node 0 is the root */
  const float* tris;        /* device, [num_faces][12] in BVH order: (v0.xyz, 0), (v1 - v0, 0), (v2 - v0, 0); 16-byte aligned */
  int32_t num_nodes;
  int32_t num_faces;
} GrMesh;
/* upper bound of the nodes gr_mesh_build_bvh writes for `num_faces` faces */
int64_t gr_mesh_bvh_max_nodes(int32_t num_faces);
/* HOST function (no device work): points [V,3] fp32, indices [F,3] int32 -> nodes_out [max_nodes][8], tris_out [F][12], face_ids_out [F]
 * (optional: original index of the face stored at each BVH position), *num_nodes_out.  Median split of the face centroids along the
 * widest axis, <= 4 faces per leaf, boxes padded by a relative 1e-5. */
int gr_mesh_build_bvh(const float* points, const int32_t* indices, int32_t num_points, int32_t num_faces, float* nodes_out, int64_t max_nodes,
                      float* tris_out, int32_t* face_ids_out, int32_t* num_nodes_out);
/* get_uav_collision_num_ray: uav_position [N,3], uav_quat_wxyz [N,4] (Isaac order; the reference re-orders to Warp's xyzw itself,
 * mesh_tools.py:261), lattices [M,3] or NULL with num_lattices = 0 (centre point only) -> collision_num [N] int32 (zeroed here, as the
 * launcher's torch.zeros).  Per lattice point six axis rays (+x -x +y -y +z -z) of length max_dist; the point counts when the closest hit
 * of one of them is a back face. */
int gr_uav_collision_ray(const GrMesh* mesh, const float* uav_position, const float* uav_quat_wxyz, int32_t num_uav, const float* lattices,
                         int32_t num_lattices, float max_dist, float arm_length, float height, int32_t* collision_num, void* stream);
/* wp.mesh_query_ray over arrays: origins / dirs [R,3] -> t_out [R] (max_t where nothing was hit), sign_out [R] (+1 front face, -1 back
 * face, 0 no hit). */
int gr_mesh_query_rays(const GrMesh* mesh, const float* origins, const float* dirs, int64_t num_rays, float max_t, float* t_out, float* sign_out,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRACING_H_ */

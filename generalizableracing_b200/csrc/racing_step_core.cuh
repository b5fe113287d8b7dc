// racing_step_core.cuh -- the body of one env.step() for ONE env held in registers (EnvRegs): sections 1-11 of
// ManagerBasedDiffRLEnv.step (L/envs/manager_based_diff_rl_env.py:160-267).  Shared by the single-step kernel
// (racing_step.cu: state round-trips HBM, observations stream to global memory) and the fused PPO collection kernel
// (ppo_collect.cu: state stays in registers over the rollout, observations feed the in-kernel policy MLP).  Where
// the observations go is the only difference, hence the ObsSink policy.
#pragma once
#include "gr_common.cuh"

namespace gr {

struct EnvRegs {
  Q4 q; V3 w; float f; V3 v; int eplen; V3 om; uint32_t pk; V3 tau; V3 aacc; float4 fifo;
  V3 k2; float m; V3 k1; float ef; V3 kp; float thr; V3 kd; V3 etau;
  V3 dcur, dnext; float noise_hi, noise_level;
  float aux;      // last cross_obs value (RewardManager._step_reward survives resets)
  bool noise_dirty_prev;   // the noise planes were rewritten by the previous step (gate switch / reset)
  float es4, es5; // episode sums of reward terms 4, 5 (ride in the spare words of PL_TORQUE / PL_ANGACC; terms 0..3: PL_EPSUM0)
  float arate;    // command metric "action_rate" assigned by the previous step's command update (commands.py:258), a 16-bit float in HBM
  bool metrics_zero;   // the last command update of this env was a full reset(): CommandTerm.reset zeroed its metrics
};

// PL_LINVEL.w, as bits: episode_length [0:12) | cross flag [12] | noise-planes-rewritten flag [13] | "action_rate" command metric
// [14:30) as a 16-bit float (5-bit exponent biased at 2^-21, 11-bit mantissa, round to nearest: rel. error <= 2^-12 on [2^-21, 2^11)) |
// [30] always 0, so the word never looks like a NaN / Inf to host-side float comparisons of the planes | [31] "metrics are zero" (the
// env's last command update was a full reset(): CommandTerm.reset zeroed them)
constexpr uint32_t kEplenMask = 0xFFFu, kAuxBit = 1u << 12, kNoiseDirtyBit = 1u << 13, kMetricsZeroBit = 1u << 31;
constexpr uint32_t kArateBias = 106u << 11;
__device__ __forceinline__ uint32_t arate_field(float arate) {
  const uint32_t r = (__float_as_uint(arate) + 0x800u) >> 12;               // exponent | 11 mantissa bits, rounded
  return r <= kArateBias ? 0u : (r - kArateBias > 0xFFFFu ? 0xFFFFu : r - kArateBias);
}
__device__ __forceinline__ float arate_value(uint32_t field) { return field ? __uint_as_float((field + kArateBias) << 12) : 0.0f; }
// the value a later load gives back (kernels that keep an env in registers over several steps carry this, like the single steps)
__device__ __forceinline__ float arate_rounded(float arate) { return arate_value(arate_field(arate)); }
__device__ __forceinline__ uint32_t eplen_word(int eplen, bool aux, bool noise_dirty, float arate, bool metrics_zero) {
  return ((uint32_t)eplen & kEplenMask) | (aux ? kAuxBit : 0u) | (noise_dirty ? kNoiseDirtyBit : 0u) | (arate_field(arate) << 14) |
         (metrics_zero ? kMetricsZeroBit : 0u);
}

// EnvRegs.om / .aacc hold the BODY-frame angular velocity / acceleration (PL_ANGVEL / PL_ANGACC): the reference
// round-trips them through the world frame every step (rot(q', w_b') stored, rotinv(q', .) read back: identity up to
// rounding); only w_b is ever consumed.  EnvRegs.fifo holds tanh(a_{t-1}): the lagged action is only used through tanh.
__device__ __forceinline__ void load_hot(EnvRegs& e, const float4* __restrict__ tile) {
  const float4 a0 = ld_plane(tile, PL_QUAT), a1 = ld_plane(tile, PL_POS), a2 = ld_plane(tile, PL_LINVEL), a3 = ld_plane(tile, PL_ANGVEL),
               a4 = ld_plane(tile, PL_TORQUE), a5 = ld_plane(tile, PL_ANGACC), a6 = ld_plane(tile, PL_FIFO);
  const uint32_t ew = __float_as_uint(a2.w);
  e.q = quat(a0); e.w = xyz(a1); e.f = a1.w; e.v = xyz(a2); e.eplen = (int)(ew & kEplenMask);
  e.aux = (ew & kAuxBit) ? 1.0f : 0.0f; e.noise_dirty_prev = (ew & kNoiseDirtyBit) != 0u;
  e.metrics_zero = (ew & kMetricsZeroBit) != 0u; e.arate = arate_value((ew >> 14) & 0xFFFFu);
  e.om = xyz(a3); e.pk = __float_as_uint(a3.w); e.tau = xyz(a4); e.es4 = a4.w; e.aacc = xyz(a5); e.es5 = a5.w; e.fifo = a6;
}
// kVolatile: re-read through L2 (ld.global.cv) after the grid dependency when the prefetched copy may be stale
template <bool kVolatile>
__device__ __forceinline__ float4 ld_cold(const float4* __restrict__ tile, int plane) {
  return kVolatile ? __ldcv(tile + plane * kTile) : __ldg(tile + plane * kTile);
}
template <bool kVolatile>
__device__ __forceinline__ void load_cold(EnvRegs& e, const float4* __restrict__ tile) {
  const float4 c0 = ld_cold<kVolatile>(tile, PL_DRAG2), c1 = ld_cold<kVolatile>(tile, PL_DRAG1), c2 = ld_cold<kVolatile>(tile, PL_KP),
               c3 = ld_cold<kVolatile>(tile, PL_KD), c4 = ld_cold<kVolatile>(tile, PL_ETAU);
  e.k2 = xyz(c0); e.m = c0.w; e.k1 = xyz(c1); e.ef = c1.w; e.kp = xyz(c2); e.thr = c2.w; e.kd = xyz(c3); e.etau = xyz(c4);
}
template <bool kNoise, bool kVolatile>
__device__ __forceinline__ void load_noise(EnvRegs& e, const float4* __restrict__ tile) {
  if (kNoise) {
    const float4 n0 = ld_cold<kVolatile>(tile, PL_NOISE0), n1 = ld_cold<kVolatile>(tile, PL_NOISE1);
    e.dcur = xyz(n0); e.dnext = v3(n0.w, n1.x, n1.y); e.noise_hi = n1.z; e.noise_level = n1.w;
  } else {
    e.dcur = v3(0.f, 0.f, 0.f); e.dnext = v3(0.f, 0.f, 0.f); e.noise_hi = 0.f; e.noise_level = 1.f;
  }
}
template <bool kNoise>
__device__ __forceinline__ void load_env(EnvRegs& e, const float4* __restrict__ tile) {
  load_hot(e, tile);
  load_cold<false>(e, tile);
  load_noise<kNoise, false>(e, tile);
}

__device__ __forceinline__ float4 tanh4(float4 a) { return make_float4(tanhf(a.x), tanhf(a.y), tanhf(a.z), tanhf(a.w)); }

// F.cosine_similarity(a, b, dim=-1, eps=1e-8): sum((a/max(|a|,eps)) * (b/max(|b|,eps)))
__device__ __forceinline__ float cosine_similarity(V3 a, V3 b) {
  const float na = fmaxf(norm(a), 1e-8f), nb = fmaxf(norm(b), 1e-8f);
  const V3 x = a / na, y = b / nb;
  return x.x * y.x + x.y * y.y + x.z * y.z;
}

// noise offsets of QD/mdp/commands.py:287-289: lo + r*(hi - lo), lo = -hi
__device__ __forceinline__ V3 gate_noise(float hi, float u0, float u1, float u2) {
  const float lo = -hi, span = hi - lo;
  return v3(lo + u0 * span, lo + u1 * span, lo + u2 * span);
}

// policy / critic / auxiliary observations (QD/racing_ctbr_env.py:139-174, QD/mdp/observation.py:22-104)
template <bool kNoise>
__device__ __forceinline__ void write_observations(const GrConfig& cfg, const EnvRegs& e, V3 origin, V3 gate_rel, V3 next_rel,
                                                   float4 th_lag, float4 n01, float4 n23, float aux, int i,
                                                   float* __restrict__ obs, float* __restrict__ critic, float* __restrict__ aux_out) {
  const V3 vb = quat_rotate_inverse(e.q, e.v);
  const V3 g_gt = gate_rel + origin, gn_gt = next_rel + origin;
  // modified_last_action (:55-63): ctbr of the lagged raw action, thrust / mass
  float4 ctbr = make_float4(th_lag.x * cfg.action_scale0 + cfg.action_scale0, th_lag.y * cfg.body_rate_bound + 0.0f,
                            th_lag.z * cfg.body_rate_bound + 0.0f, th_lag.w * cfg.body_rate_bound + 0.0f);
  ctbr.x = ctbr.x / e.m;
  // policy: noisy lin vel (:52), noisy attitude row (:27-32), command w.r.t. the noisy gates (commands.py:208-219)
  const V3 vn = v3(vb.x * (1.0f + n01.x * cfg.obs_vel_noise), vb.y * (1.0f + n01.y * cfg.obs_vel_noise),
                   vb.z * (1.0f + n01.z * cfg.obs_vel_noise));
  const Q4 qn = quat_from_euler_xyz(n01.w * cfg.obs_euler_noise, n23.x * cfg.obs_euler_noise, n23.y * cfg.obs_euler_noise);
  const V3 rn = rotmat_row2(quat_mul(e.q, qn));
  V3 g_pol = g_gt, gn_pol = gn_gt;
  if (kNoise) { g_pol = g_gt + e.dcur; gn_pol = gn_gt + e.dnext; }
  const V3 c0 = quat_rotate_inverse(e.q, g_pol - e.w), c1 = quat_rotate_inverse(e.q, gn_pol - g_pol);
  float4* o = reinterpret_cast<float4*>(obs) + (int64_t)i * 4;
  __stcs(o + 0, make_float4(vn.x, vn.y, vn.z, rn.x));
  __stcs(o + 1, make_float4(rn.y, rn.z, c0.x, c0.y));
  __stcs(o + 2, make_float4(c0.z, c1.x, c1.y, c1.z));
  __stcs(o + 3, ctbr);
  if (critic) {
    const V3 r = rotmat_row2(e.q);
    const V3 d0 = quat_rotate_inverse(e.q, g_gt - e.w), d1 = quat_rotate_inverse(e.q, gn_gt - g_gt);
    float4* c = reinterpret_cast<float4*>(critic) + (int64_t)i * 4;
    __stcs(c + 0, make_float4(vb.x, vb.y, vb.z, r.x));
    __stcs(c + 1, make_float4(r.y, r.z, d0.x, d0.y));
    __stcs(c + 2, make_float4(d0.z, d1.x, d1.y, d1.z));
    __stcs(c + 3, ctbr);
  }
  if (aux_out) aux_out[i] = aux;
}

// _reset_idx for one env (L/envs/manager_based_diff_rl_env.py:362-410): curricula, root-state sampler,
// controller/dynamics/command reset.  Mutates e; returns the new origin.
template <bool kNoise, bool kPhilox, class DrawsT>
__device__ __forceinline__ V3 reset_env(const GrConfig& cfg, const TrackSmem& tr, EnvRegs& e, const DrawsT& rs, float thr_normal) {
  const int type = (int)pk_type(e.pk);
  int level = (int)pk_level(e.pk);
  const int acc = (int)pk_acc(e.pk);
  const float4 u_pose0 = rs.get4(2), u_pose1 = rs.get4(3), u_vel = rs.get4(4), u_d0 = rs.get4(5), u_d1 = rs.get4(6);
  // -- terrain curriculum (QD/mdp/curriculums.py:25-38 + TerrainImporter.update_env_origins)
  {
    const int lv = level + (acc >= cfg.level_up_gates ? 1 : 0) - (acc < cfg.level_down_gates ? 1 : 0);
    int rand_lv = (int)floorf(u_d1.w * (float)tr.levels);          // slot 27
    if (rand_lv > tr.levels - 1) rand_lv = tr.levels - 1;
    level = (lv >= tr.levels) ? rand_lv : (lv < 0 ? 0 : lv);
  }
  // -- command-noise curriculum (QD/mdp/curriculums.py:40-54, QD/mdp/commands.py:385-402)
  if (kNoise && cfg.noise_curriculum) {
    const float up = acc >= cfg.noise_up_gates ? cfg.noise_up : 1.0f;
    const float down = acc < cfg.noise_down_gates ? cfg.noise_down : 1.0f;
    e.noise_level *= up; e.noise_level *= down;
    e.noise_hi *= up; e.noise_hi *= down;
  }
  const float4 orow = tr.origin_row(type, level);
  const V3 origin = xyz(orow);
  const int start_gate = __float_as_int(orow.w);
  // -- reset_root_state_racing (QD/mdp/events.py:139-177): slots 8..13 pose, 14..19 velocity
  const float sp = cfg.reset_pos - (-cfg.reset_pos), srp = cfg.reset_roll_pitch - (-cfg.reset_roll_pitch),
              sy = cfg.reset_yaw - (-cfg.reset_yaw), sv = cfg.reset_vel - (-cfg.reset_vel);
  const V3 dpos = v3(u_pose0.x * sp + (-cfg.reset_pos), u_pose0.y * sp + (-cfg.reset_pos), u_pose0.z * sp + (-cfg.reset_pos));
  const float roll = u_pose0.w * srp + (-cfg.reset_roll_pitch), pitch = u_pose1.x * srp + (-cfg.reset_roll_pitch);
  const float dyaw = u_pose1.y * sy + (-cfg.reset_yaw);
  const V3 pos = v3(cfg.default_pos[0], cfg.default_pos[1], cfg.default_pos[2]) + origin + dpos;
  const V3 towards = (tr.gate(type, level, start_gate) + origin) - pos;
  const float yaw = wrap_to_pi_atan2(atan2f(towards.y, towards.x)) + dyaw;
  e.q = quat_mul(Q4{1.f, 0.f, 0.f, 0.f}, quat_from_euler_xyz_fast(roll, pitch, yaw));
  e.w = pos;
  e.v = v3(u_pose1.z * sv + (-cfg.reset_vel), u_pose1.w * sv + (-cfg.reset_vel), u_vel.x * sv + (-cfg.reset_vel));
  // DroneDynamics.reset_state (droneDynamics.py:116): ang_vel_b = rotinv(q, ang_vel_w)
  e.om = quat_rotate_inverse(e.q, v3(u_vel.y * sv + (-cfg.reset_vel), u_vel.z * sv + (-cfg.reset_vel), u_vel.w * sv + (-cfg.reset_vel)));
  e.aacc = v3(0.f, 0.f, 0.f);                      // closure A.1
  // -- CTBRController.reset_idx (L/controllers/controller_diff.py:146-160)
  e.f = 0.f; e.tau = v3(0.f, 0.f, 0.f);
  // -- DroneDynamics.reset_idx (QD/mdp/dynamics/droneDynamics.py:50-57): slots 20 z, 21..23 drag2, 24..26 drag1
  if (cfg.random_drag) {
    const float z = 1.0f * cfg.z_drag + u_d0.x * cfg.z_drag_rand;
    const float b2 = cfg.drag2 * e.m, b1 = cfg.drag1 * e.m;
    e.k2 = v3(b2 + u_d0.y * cfg.drag2_rand, b2 + u_d0.z * cfg.drag2_rand, (b2 + u_d0.w * cfg.drag2_rand) * z);
    e.k1 = v3(b1 + u_d1.x * cfg.drag1_rand, b1 + u_d1.y * cfg.drag1_rand, (b1 + u_d1.z * cfg.drag1_rand) * z);
  }
  // -- DiffActions.reset_idx (QD/mdp/diff_action.py:233)
  e.thr = 1.0f + thr_normal * cfg.thr_err_reset_std;
  // -- RacingCommand._resample_command (QD/mdp/commands.py:262-306)
  if (kNoise) {
    const float4 u_g0 = rs.get4(7), u_g1 = rs.get4(8);          // slots 28..30 current gate xyz, 31..33 next gate xyz
    e.dcur = gate_noise(e.noise_hi, u_g0.x, u_g0.y, u_g0.z);
    e.dnext = gate_noise(e.noise_hi, u_g0.w, u_g1.x, u_g1.y);
  }
  e.pk = pk_make((uint32_t)start_gate, 0u, (uint32_t)level, (uint32_t)type, 1u);
  e.eplen = 0;
  e.es4 = 0.f; e.es5 = 0.f;
  return origin;
}

template <bool kNoise>
__device__ __forceinline__ void store_env(const EnvRegs& e, float4* __restrict__ tile, bool cold_dirty, bool noise_dirty) {
  // the noise-dirty flag = "noise planes rewritten in this step": lets the next step trust its pre-dependency prefetch of them
  st_plane(tile, PL_QUAT, pack(e.q));
  st_plane(tile, PL_POS, pack(e.w, e.f));
  st_plane(tile, PL_LINVEL, pack(e.v, __uint_as_float(eplen_word(e.eplen, e.aux != 0.0f, kNoise && noise_dirty, e.arate, e.metrics_zero))));
  st_plane(tile, PL_ANGVEL, pack(e.om, __uint_as_float(e.pk)));
  st_plane(tile, PL_TORQUE, pack(e.tau, e.es4));
  st_plane(tile, PL_ANGACC, pack(e.aacc, e.es5));
  st_plane(tile, PL_FIFO, e.fifo);
  if (cold_dirty) {
    tile[PL_DRAG2 * kTile] = pack(e.k2, e.m);
    tile[PL_DRAG1 * kTile] = pack(e.k1, e.ef);
    tile[PL_KP * kTile] = pack(e.kp, e.thr);
  }
  if (kNoise && noise_dirty) {
    tile[PL_NOISE0 * kTile] = make_float4(e.dcur.x, e.dcur.y, e.dcur.z, e.dnext.x);
    tile[PL_NOISE1 * kTile] = make_float4(e.dnext.y, e.dnext.z, e.noise_hi, e.noise_level);
  }
}

// =============================================================================================
// the step kernel
//
// Latency notes (ncu, 65,536 envs = one wave of ~14 warps/SM): the kernel is bound by the per-warp dependent
// instruction chain, not by issue slots or DRAM, so the hot path uses MUFU-based fast math (no slow-path
// branches / calls), hoists the shared quaternion terms (RotQ), and computes every view of the post-step
// state once: the reward section and the observation section share them unless the env was reset or
// switched gate in this step (rare, recomputed in a divergent tail).
// =============================================================================================
#ifdef GR_PHASE_TIMING
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define GR_STAMP(k) do { if (io.phase_times && (threadIdx.x & 31) == 0) io.phase_times[(size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 5 + (k)] = gtime(); } while (0)
#else
#define GR_STAMP(k) do { } while (0)
#endif

__device__ __forceinline__ float4 fm_tanh4(float4 a) { return make_float4(fm_tanh(a.x), fm_tanh(a.y), fm_tanh(a.z), fm_tanh(a.w)); }

// what the body hands back besides the updated EnvRegs
struct StepOut {
  float reward; float terms[GR_NUM_REWARD_TERMS];     // terms: weighted step rewards / dt (RewardManager._step_reward)
  bool terminated, time_out, reset, passed, noise_dirty;
  bool stored;       // the sink already wrote the state back (GlobalObsSink with an early store)
};

// observation sink of the single-step kernel.  A thread owns one 64-byte row; written directly, every store instruction of a
// warp would touch 32 half-filled sectors.  The warp parks its rows in shared memory (quarter index rotated by row/2: both
// the 16-byte row writes and the transposed reads are bank-conflict free) and streams them out as four contiguous 512-byte
// stores.  `live` = the lanes of this warp that own an env (contiguous from lane 0); `stage` = 128 float4 per warp.
struct GlobalObsSink {
  const GrStepIO& io;
  unsigned live;
  float4* stage;
  float4* tile = nullptr;            // set => the state is written back as soon as it is final (before the observation section), so the
                                     // stores of a warp start draining while it still computes its observations
  // called by the body when the post-step state (e, sums) is final; returns true if it stored the state
  template <bool kNoise, bool kDiff, bool kStats>
  __device__ __forceinline__ bool state_final(EnvRegs& e, float4& eps0, const float4& lsum, const float (&terms)[GR_NUM_REWARD_TERMS], float dt, bool reset,
                                              bool noise_dirty) const;
  __device__ __forceinline__ void rows(float4* __restrict__ out, int i, float4 o0, float4 o1, float4 o2, float4 o3) const {
#ifdef GR_CPU_EMUL
    float4* o = out + (int64_t)i * 4;
    o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3;
#else
    const int lane = threadIdx.x & 31, rot = lane >> 1;
    float4* s = stage + lane * 4;
    s[(0 + rot) & 3] = o0; s[(1 + rot) & 3] = o1; s[(2 + rot) & 3] = o2; s[(3 + rot) & 3] = o3;
    __syncwarp(live);
    float4* dst = out + (int64_t)(i - lane) * 4;
    if (live == 0xffffffffu) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = k * 32 + lane, r = j >> 2, c = j & 3;
        __stcs(dst + j, stage[r * 4 + ((c + (r >> 1)) & 3)]);
      }
    } else {                                                // the ragged last warp: only the live lanes are here
      const int nlive = __popc(live);
      for (int j = lane; j < nlive * 4; j += nlive) {
        const int r = j >> 2, c = j & 3;
        __stcs(dst + j, stage[r * 4 + ((c + (r >> 1)) & 3)]);
      }
    }
    __syncwarp(live);                                       // the buffer is reused by the next group of rows
#endif
  }
  __device__ __forceinline__ void policy(int i, float4 o0, float4 o1, float4 o2, float4 o3) const { rows(reinterpret_cast<float4*>(io.obs), i, o0, o1, o2, o3); }
  __device__ __forceinline__ constexpr bool wants_policy() const { return true; }
  __device__ __forceinline__ bool wants_critic() const { return io.critic_obs != nullptr; }
  __device__ __forceinline__ void critic(int i, float4 c0, float4 c1, float4 c2, float4 c3) const { rows(reinterpret_cast<float4*>(io.critic_obs), i, c0, c1, c2, c3); }
  __device__ __forceinline__ void aux(int i, float v) const { if (io.aux_obs) io.aux_obs[i] = v; }
};

// episode sums of the reward terms: 0..3 in `eps0` (PL_EPSUM0), 4..5 in the env's hot planes (e.es4 / e.es5)
__device__ __forceinline__ void add_episode_sums(float4& eps0, EnvRegs& e, const float (&terms)[GR_NUM_REWARD_TERMS], float dt) {
  eps0.x += terms[0] * dt; eps0.y += terms[1] * dt; eps0.z += terms[2] * dt; eps0.w += terms[3] * dt;
  e.es4 += terms[4] * dt; e.es5 += terms[5] * dt;
}

template <bool kNoise, bool kDiff, bool kStats>
__device__ __forceinline__ bool GlobalObsSink::state_final(EnvRegs& e, float4& eps0, const float4& lsum, const float (&terms)[GR_NUM_REWARD_TERMS], float dt,
                                                           bool reset, bool noise_dirty) const {
  if (!tile) return false;
  if (kStats && !reset) add_episode_sums(eps0, e, terms, dt);
  store_env<kNoise>(e, tile, reset, noise_dirty);
  if (kStats) { st_plane(tile, PL_EPSUM0, eps0); if (kDiff) st_plane(tile, PL_LOSSSUM, lsum); }
  return true;
}

// Sections 1-11.  `e` = state before the step in, state after the step (and after a reset) out, e.fifo included;
// eps0 / e.es4 / e.es5 = episode sums of the reward terms (kStats): logged + zeroed here on reset, the non-reset accumulation
// is left to the caller (add_episode_sums with out.terms), who knows when its copy of the sums has arrived; lsum = LossManager
// episode sums (kDiff && kStats, PL_LOSSSUM).  Returns false for a lane past the last env.
template <bool kNoise, bool kDiff, bool kPhilox, bool kStats, class ObsSink, class DrawsT>
__device__ __forceinline__ bool racing_step_body(const GrConfig& cfg, const TrackSmem& tr, EnvRegs& e, const float4 a_t, const float4 n01,
                                                 const float4 n23, const DrawsT& draws_in, float4& eps0, float4& lsum, const GrStepIO& io,
                                                 const int i, const bool active, const ObsSink& sink, StepOut& out) {
  const int type = (int)pk_type(e.pk);
  int level = (int)pk_level(e.pk);
  int gate_id = (int)pk_gate(e.pk);
  const uint32_t fresh = pk_fresh(e.pk);
#ifdef GR_PHASE_TIMING
  if (__float_as_uint(e.q.w + e.w.x + e.v.x + e.om.x + e.tau.x + e.aacc.x + e.fifo.x + e.k2.x + e.k1.x + e.kp.x + e.kd.x + e.etau.x + a_t.x + eps0.x + lsum.x + e.dcur.x + e.noise_hi) != 0x7fc12345u) GR_STAMP(2);
#endif
  V3 origin = xyz(tr.origin_row(type, level));
  V3 gate_rel = tr.gate(type, level, gate_id);
  const float dt = cfg.dt;
  // RacingCommand metrics of the previous command update (commands.py:258-260), logged if this step resets the env: functions of
  // the state entering the step
  const float lin_spd_in = dot(e.v, e.v), ang_spd_in = dot(e.om, e.om);
  const V3 J = v3(cfg.inertia[0], cfg.inertia[1], cfg.inertia[2]);
  const V3 Jinv = v3(fm_rcp(J.x), fm_rcp(J.y), fm_rcp(J.z));
  const float inv_m = fm_rcp(e.m);
  const float s0 = cfg.action_scale0, sb = cfg.body_rate_bound;

  // ---- 1. process_action (L/managers/action_manager.py:44-45; QD/mdp/diff_action.py:160-176) ----
  // FIFO (lag 1): the applied action is a_{t-1}; prev_action is a_{t-1} unless the latches were zeroed by a reset
  const float4 th_lag = e.fifo;                 // tanh(a_{t-1}), stored by the previous step
  const float4 th_a = fm_tanh4(a_t);
  const float4 th_prev = fresh ? make_float4(0.f, 0.f, 0.f, 0.f) : th_lag;
  // get_state_from_sim (QD/mdp/diff_action.py:126-154)
  const RotQ R0(e.q);
  const V3 p = e.w - origin;
  const V3 om_b = e.om;
  const V3 v_b = R0.rotinv(e.v);
  const V3 aacc_b = e.aacc;
  const float cmd0 = (th_lag.x * s0 + s0) * e.thr;
  const V3 cmd_rate = v3(th_lag.y * sb, th_lag.z * sb, th_lag.w * sb);

  // ---- CTBRController.compute (L/controllers/controller_diff.py:120-138) ----
  const float thrust_des = fminf(fmaxf(cmd0, cfg.thrust_lo), cfg.thrust_hi);
  const float f_new = (1.0f - e.ef) * thrust_des + e.ef * e.f;
  const V3 rate_c = v3(fminf(fmaxf(cmd_rate.x, -sb), sb), fminf(fmaxf(cmd_rate.y, -sb), sb), fminf(fmaxf(cmd_rate.z, -sb), sb));
  const V3 gyro = cross(om_b, J * om_b);
  const V3 torque_des = J * (e.kp * (rate_c - om_b)) + gyro - e.kd * aacc_b;
  const V3 one_m_etau = v3(1.0f - e.etau.x, 1.0f - e.etau.y, 1.0f - e.etau.z);
  const V3 tau_new = one_m_etau * torque_des + e.etau * e.tau;

  // ---- DroneDynamics.step (QD/mdp/dynamics/droneDynamics.py:119-135) ----
  const V3 F_b = v3(0.f, 0.f, f_new) - e.k2 * v_b * vabs(v_b) - e.k1 * v_b;
  const V3 acc = v3(0.f, 0.f, -cfg.gravity) + R0.rot(F_b) * inv_m;
  const V3 alpha = Jinv * (tau_new - gyro);
  const V3 p1 = p + e.v * dt + (0.5f * dt * dt) * acc;
  const float hdt = 0.5f * dt;
  // q + 0.5*dt*qmul(q, (0, om_b)): Hamilton product with a pure quaternion
  const Q4 qt = Q4{e.q.w - hdt * (e.q.x * om_b.x + e.q.y * om_b.y + e.q.z * om_b.z),
                   e.q.x + hdt * (e.q.w * om_b.x + e.q.y * om_b.z - e.q.z * om_b.y),
                   e.q.y + hdt * (e.q.w * om_b.y + e.q.z * om_b.x - e.q.x * om_b.z),
                   e.q.z + hdt * (e.q.w * om_b.z + e.q.x * om_b.y - e.q.y * om_b.x)};
  const float inv_qn = fm_rsqrt(qt.w * qt.w + qt.x * qt.x + qt.y * qt.y + qt.z * qt.z);
  const Q4 q1 = Q4{qt.w * inv_qn, qt.x * inv_qn, qt.y * inv_qn, qt.z * inv_qn};
  const V3 v1 = e.v + acc * dt;
  const V3 omb1 = om_b + alpha * dt;
  const RotQ R1(q1);

  // ---- BPTT tape planes 0..5 of this step (consumer: racing_bwd.cu; SURVEY.md A.6/A.7, derived form) ----
  if (kDiff && io.tape && active) {
    float4* __restrict__ T = reinterpret_cast<float4*>(io.tape);
    const float m0 = (cmd0 >= cfg.thrust_lo && cmd0 <= cfg.thrust_hi) ? 1.0f : 0.0f;
    const float A0 = m0 * e.thr * s0 * (1.0f - th_lag.x * th_lag.x) * (1.0f - e.ef);
    const V3 mk = v3((cmd_rate.x >= -sb && cmd_rate.x <= sb) ? sb : 0.0f, (cmd_rate.y >= -sb && cmd_rate.y <= sb) ? sb : 0.0f,
                     (cmd_rate.z >= -sb && cmd_rate.z <= sb) ? sb : 0.0f);
    const V3 A = mk * v3(1.0f - th_lag.y * th_lag.y, 1.0f - th_lag.z * th_lag.z, 1.0f - th_lag.w * th_lag.w) * one_m_etau * J * e.kp;
    const V3 D = -(2.0f * (e.k2 * vabs(v_b)) + e.k1);
    __stcs(T + tidx(0, i), pack(e.q));
    __stcs(T + tidx(1, i), pack(om_b, A0));
    __stcs(T + tidx(2, i), pack(F_b, A.x));
    __stcs(T + tidx(3, i), pack(D, A.y));
    __stcs(T + tidx(4, i), pack(v1, A.z));
    __stcs(T + tidx(5, i), pack(omb1, __uint_as_float(fresh)));
  }

  // ---- 2. "physics": closure A.1, truth := nominal; world pose, last angular acceleration ----
  e.w = p1 + origin; e.q = q1; e.v = v1; e.om = omb1; e.aacc = alpha;
  e.f = f_new; e.tau = tau_new;
  // ---- 3. align (droneDynamics.py:156-181): value = sim-derived state; aligned local position for the loss ----
  const V3 p_al = e.w - origin;
  const V3 v_al = v1;
  if (kDiff && active && io.aligned_states) {       // extras["aligned_states"] / ["nominal_states"] (:205-212), before any reset
    float* __restrict__ al = io.aligned_states + (int64_t)i * 13;
    const V3 om_w = R1.rot(omb1);
    al[0] = p1.x; al[1] = p1.y; al[2] = p1.z; al[3] = q1.w; al[4] = q1.x; al[5] = q1.y; al[6] = q1.z;
    al[7] = v1.x; al[8] = v1.y; al[9] = v1.z; al[10] = om_w.x; al[11] = om_w.y; al[12] = om_w.z;
    if (io.acc) { io.acc[(int64_t)i * 3] = acc.x; io.acc[(int64_t)i * 3 + 1] = acc.y; io.acc[(int64_t)i * 3 + 2] = acc.z; }
  }

  if (io.pre_reset_pos && active) {                  // pose the reward terms see (mesh collision term: a separate launch on this snapshot)
    float* __restrict__ pp = io.pre_reset_pos + (int64_t)i * 3;
    pp[0] = e.w.x; pp[1] = e.w.y; pp[2] = e.w.z;
    if (io.pre_reset_quat) __stcs(reinterpret_cast<float4*>(io.pre_reset_quat) + i, pack(q1));
  }

  // ---- 4./5. counters + terminations (QD/mdp/termination.py:15-33; Isaac Lab mdp.time_out) ----
  e.eplen += 1;
  const bool time_out = e.eplen >= cfg.max_episode_length;
  bool terminated = false;
  if (cfg.term_oob) terminated = terminated || (e.w.z < cfg.oob_lo) || (e.w.z > cfg.oob_hi);
  // bad_pose: |wrap_to_pi(roll)| > pi/2  <=>  cos_roll < 0 ; the pitch clause can never fire (|asin| <= pi/2).
  // Equivalent to the literal atan2/asin/%2pi/wrap_to_pi chain of termination.py:29-33 except inside the few-ulp
  // band |cos_roll| ~ 1e-7 where the literal form is itself a rounding lottery (tests/test_bad_pose.py).
  const bool bad = (1.0f - 2.0f * (q1.x * q1.x + q1.y * q1.y)) < 0.0f;
  if (cfg.term_bad_pose) terminated = terminated || bad;

  // ---- 6. rewards (QD/mdp/rewards.py:154-253; RewardManager.compute: sum_i term_i * w_i * dt) ----
  const V3 vec = (gate_rel + origin) - e.w;                       // gate_pose_gt_w - root_pos_w
  const float d2 = vec.x * vec.x + vec.y * vec.y + vec.z * vec.z;
  const float dist = sqrt_rn(d2);
  const bool pass_pre = dist < cfg.update_threshold;
  const V3 vb1 = R1.rotinv(v1);
  const V3 cg0 = R1.rotinv(vec);                                  // command_gt[:, :3]
  float (&terms)[GR_NUM_REWARD_TERMS] = out.terms;
  float reward = 0.0f, arate_now;
  {
    const float inv_nc = fm_rsqrt(fmaxf(dot(cg0, cg0), 1e-16f));
    const float inv_nv = fm_rsqrt(fmaxf(dot(vb1, vb1), 1e-16f));
    terms[0] = dot(vb1, cg0) * inv_nv * inv_nc;                                                   // :154-161
    terms[1] = sb * fm_sqrt(th_a.y * th_a.y + th_a.z * th_a.z + th_a.w * th_a.w);                 // :188-194
    {
      const float d0 = (th_a.x - th_prev.x) * s0, d1 = (th_a.y - th_prev.y) * sb, d2r = (th_a.z - th_prev.z) * sb,
                  d3 = (th_a.w - th_prev.w) * sb;                                                  // :196-206
      terms[2] = d0 * d0 + d1 * d1 + d2r * d2r + d3 * d3;
    }
    arate_now = terms[2];                                                                         // commands.py:258 (same function, step 8)
    terms[3] = cg0.x * inv_nc;                                                                    // :171-179
    terms[4] = pass_pre ? fm_rcp(d2 + 1.0f) : 0.0f;                                               // :215-224
    terms[5] = bad ? 1.0f : 0.0f;                                                                  // :244-253
#pragma unroll
    for (int k = 0; k < GR_NUM_REWARD_TERMS; ++k) {
      const float tw = cfg.w_reward[k] != 0.0f ? terms[k] * cfg.w_reward[k] : 0.0f;
      const float value = tw * dt;
      reward += value;
      terms[k] = tw;                                              // RewardManager._step_reward = value / dt; episode sums below
    }
  }
  e.aux = terms[4] > 0.0f ? 1.0f : 0.0f;                 // cross_obs (QD/mdp/observation.py:97-104)

  // ---- 7. reset (L/envs/manager_based_diff_rl_env.py:232-247,362-410) ----
  const bool reset = terminated || time_out;
  const DrawsT draws = draws_in.staged(reset && active, kNoise);       // (warp-cooperative reset draws, when the caller gave a staging area)
  bool noise_dirty = false;
  bool passed = pass_pre;                                 // 8. on an env that did not reset this is the same test
  if (reset) {
    if (kStats) add_episode_sums(eps0, e, terms, dt);          // RewardManager.compute adds this step's values before the reset logs them
    // episode log (extras["log"], manager_based_diff_rl_env.py:380-407): fire-and-forget RED ops on one of GR_LOG_SHARDS
    // accumulator rows picked by warp id (a single row serialises on one L2 line: 3x the kernel time at a 5 % reset rate)
    if (io.log_accum && active) {
      float* acc_row = io.log_accum + (size_t)((i >> 5) & (GR_LOG_SHARDS - 1)) * GR_LOG_SLOTS;
      atomicAdd(acc_row + GR_LOG_NUM_RESET, 1.0f);
      atomicAdd(acc_row + GR_LOG_SUM_GATES, (float)pk_acc(e.pk));
      if (time_out) atomicAdd(acc_row + GR_LOG_NUM_TIMEOUT, 1.0f);
      if (terminated) atomicAdd(acc_row + GR_LOG_NUM_TERMINATED, 1.0f);
      if (kStats) {
#pragma unroll
        for (int k = 0; k < 4; ++k) atomicAdd(acc_row + GR_LOG_SUM_EPSUM + k, (&eps0.x)[k]);
        atomicAdd(acc_row + GR_LOG_SUM_EPSUM + 4, e.es4);
        atomicAdd(acc_row + GR_LOG_SUM_EPSUM + 5, e.es5);
        if (!e.metrics_zero) {
          atomicAdd(acc_row + GR_LOG_SUM_ACTION_RATE, e.arate);
          atomicAdd(acc_row + GR_LOG_SUM_LIN_SPD, sqrtf(lin_spd_in));
          atomicAdd(acc_row + GR_LOG_SUM_ANG_SPD, sqrtf(ang_spd_in));
        }
        if (kDiff) { atomicAdd(acc_row + GR_LOG_SUM_LOSS, lsum.x); atomicAdd(acc_row + GR_LOG_SUM_LOSS + 1, lsum.y); atomicAdd(acc_row + GR_LOG_SUM_LOSS + 2, lsum.z); }
      }
    }
    if (kStats) { eps0 = make_float4(0.f, 0.f, 0.f, 0.f); lsum = eps0; }       // (e.es4 / e.es5: reset_env)
    origin = reset_env<kNoise, kPhilox>(cfg, tr, e, draws, draws.thr_normal(n23.z));
    level = (int)pk_level(e.pk);
    gate_id = (int)pk_gate(e.pk);
    gate_rel = tr.gate(type, level, gate_id);
    noise_dirty = true;
    // ---- 8. on the fresh state (QD/mdp/commands.py:247-260,308-312)
    const V3 dv = (gate_rel + origin) - e.w;
    passed = sqrt_rn(dv.x * dv.x + dv.y * dv.y + dv.z * dv.z) < cfg.update_threshold;
  } else {
    e.pk &= 0x7FFFFFFFu;      // latches hold a_t again
  }
  // _update_metrics of this step's command update (step 8): ActionManager.reset zeroed the latches of an env reset in this step
  e.arate = reset ? 0.0f : arate_rounded(arate_now);
  e.metrics_zero = false;

  // ---- 8. command update (QD/mdp/commands.py:247-260 then :308-350) ----
  if (passed) {
    const uint32_t acc_g = pk_acc(e.pk) + 1u;
    gate_id = (gate_id + 1) % tr.gates;
    e.pk = pk_make((uint32_t)gate_id, acc_g, (uint32_t)level, (uint32_t)type, pk_fresh(e.pk));
    gate_rel = tr.gate(type, level, gate_id);
    if (kNoise) {
      const float4 u0 = draws.get4(10), u1 = draws.get4(11);     // slots 40..42 gate xyz, 43..45 next gate xyz
      e.dcur = gate_noise(e.noise_hi, u0.x, u0.y, u0.z);
      e.dnext = gate_noise(e.noise_hi, u0.w, u1.x, u1.y);
      noise_dirty = true;
    }
  }
  const V3 next_rel = tr.gate(type, level, (gate_id + 1) % tr.gates);
  const V3 g_gt = gate_rel + origin, gn_gt = next_rel + origin;

  // ---- 9. BPTT losses (QD/mdp/losses.py:72-80,95-101,111-117) + tape plane 6 ----
  if (kDiff && active) {
    const V3 desired = g_gt - origin;
    const V3 dvec = desired - p_al;
    const float ld = norm(dvec);
    const float l_target = ld * cfg.w_loss[0];
    const float l_vel = ((v_al.x * v_al.x + v_al.y * v_al.y + v_al.z * v_al.z) / 3.0f) * cfg.w_loss[1];
    const float z = p_al.z;
    const float den = 1.0f + 1.0f * z + 10.0f * (z * z);
    const float l_fall = (1.0f / den) * cfg.w_loss[2];
    if (kStats) { lsum.x += l_target * dt; lsum.y += l_vel * dt; lsum.z += l_fall * dt; }      // LossManager.compute (:103), after the reset
    if (io.loss) io.loss[i] = ((0.0f + l_target) + l_vel) + l_fall;
    if (io.loss_terms) { io.loss_terms[i * 3 + 0] = l_target; io.loss_terms[i * 3 + 1] = l_vel; io.loss_terms[i * 3 + 2] = l_fall; }
    if (io.tape) {
      // plane 6: d loss / d aligned position (target + falling terms); d loss / d velocity is rebuilt from v1
      const float inv = ld > 0.0f ? cfg.w_loss[0] / ld : 0.0f;
      const float dfall = -cfg.w_loss[2] * (1.0f + 20.0f * z) / (den * den);
      __stcs(reinterpret_cast<float4*>(io.tape) + tidx(6, i),
             make_float4(-dvec.x * inv, -dvec.y * inv, -dvec.z * inv + dfall, 0.0f));
    }
  }

  if (!active) return false;      // (the caller leaves too)
  e.fifo = th_a;
  out.stored = sink.template state_final<kNoise, kDiff, kStats>(e, eps0, lsum, terms, dt, reset, noise_dirty);

  // ---- 11. observations on the post-reset state (QD/racing_ctbr_env.py:139-174, QD/mdp/observation.py:22-104) ----
  if (sink.wants_policy()) {
    V3 vb = vb1, d0 = cg0;                       // common case: same state and gate as the reward section
    RotQ Rq = R1;
    if (reset || passed) {                       // rare: recompute the views on the new state / gate
      Rq = RotQ(e.q);
      vb = Rq.rotinv(e.v);
      d0 = Rq.rotinv(g_gt - e.w);
    }
    const V3 d1 = Rq.rotinv(gn_gt - g_gt);
    V3 c0 = d0, c1 = d1;
    if (kNoise) {
      const V3 g_pol = g_gt + e.dcur, gn_pol = gn_gt + e.dnext;
      c0 = Rq.rotinv(g_pol - e.w);
      c1 = Rq.rotinv(gn_pol - g_pol);
    }
    // modified_last_action (:55-63): ctbr of the lagged raw action, thrust / mass
    const float4 ctbr = make_float4((th_lag.x * s0 + s0) * inv_m, th_lag.y * sb, th_lag.z * sb, th_lag.w * sb);
    // noisy lin vel (:52) and noisy attitude row (:27-32): R(q (x) q_noise)[2,:]
    const float nv = cfg.obs_vel_noise, ne = cfg.obs_euler_noise;
    const V3 vn = v3(vb.x * (1.0f + n01.x * nv), vb.y * (1.0f + n01.y * nv), vb.z * (1.0f + n01.z * nv));
    float sr, cr, sp, cp, sy, cy;
    fm_sincos(0.5f * ne * n01.w, &sr, &cr);
    fm_sincos(0.5f * ne * n23.x, &sp, &cp);
    fm_sincos(0.5f * ne * n23.y, &sy, &cy);
    const Q4 qn = Q4{cy * cr * cp + sy * sr * sp, cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp};
    const Q4 a = e.q;
    const Q4 qq = Q4{a.w * qn.w - a.x * qn.x - a.y * qn.y - a.z * qn.z, a.w * qn.x + a.x * qn.w + a.y * qn.z - a.z * qn.y,
                     a.w * qn.y - a.x * qn.z + a.y * qn.w + a.z * qn.x, a.w * qn.z + a.x * qn.y - a.y * qn.x + a.z * qn.w};
    const float two_s = 2.0f * fm_rcp(qq.w * qq.w + qq.x * qq.x + qq.y * qq.y + qq.z * qq.z);
    const V3 rn = v3(two_s * (qq.x * qq.z - qq.y * qq.w), two_s * (qq.y * qq.z + qq.x * qq.w), 1.0f - two_s * (qq.x * qq.x + qq.y * qq.y));
    sink.policy(i, make_float4(vn.x, vn.y, vn.z, rn.x), make_float4(rn.y, rn.z, c0.x, c0.y), make_float4(c0.z, c1.x, c1.y, c1.z), ctbr);
    if (sink.wants_critic()) {
      const float ts = 2.0f * fm_rcp(a.w * a.w + a.x * a.x + a.y * a.y + a.z * a.z);
      const V3 r = v3(ts * (a.x * a.z - a.y * a.w), ts * (a.y * a.z + a.x * a.w), 1.0f - ts * (a.x * a.x + a.y * a.y));
      sink.critic(i, make_float4(vb.x, vb.y, vb.z, r.x), make_float4(r.y, r.z, d0.x, d0.y), make_float4(d0.z, d1.x, d1.y, d1.z), ctbr);
    }
    sink.aux(i, e.aux);
  }
  out.reward = reward; out.terminated = terminated; out.time_out = time_out; out.reset = reset; out.passed = passed; out.noise_dirty = noise_dirty;
  return true;
}

}  // namespace gr

// reach_step.cu -- one env.step() of the reach-target tasks per launch, one thread per env.
//
// Replaces ManagerBasedDiffRLEnv.step (L/envs/manager_based_diff_rl_env.py:160-267) for QD/reach_target_lv_env.py and
// QD/reach_target_ctbr_env.py: DiffActions.process_actions (QD/mdp/diff_action.py:156-206) with LVController /
// PSController (L/controllers/controller_diff.py:172-443) or CTBRController (:37-170), DroneDynamics.step/align
// (QD/mdp/dynamics/droneDynamics.py:119-181), time_out + z bounds, the reach-target rewards (QD/mdp/rewards.py:30-101),
// _reset_idx with reset_root_state_uniform, UniformWorldPoseCommand (QD/mdp/commands.py:98-134), the reach-target losses
// (QD/mdp/losses.py:32-67) and the 17-wide observation (QD/reach_target_lv_env.py:83-104).  Closure: oracle/reach_oracle.py.
//
// HBM-bound like racing_step.cu: 11 planes read + written (352 B) + 2 read-mostly planes (32 B) + action 16 + obs 68 +
// reward / masks / dones 14 = 482 B per env-step.  The diff variant also carries the controller's 4x4 action Jacobian in
// forward mode (reach_core.cuh D4) and writes the 13-plane tape consumed by reach_bwd.cu:
//   0: q (pre-step)   1: omega_b | cut   2: F_b | gQ.w   3: D | gQ.x   4: v' | gQ.y   5: omega_b' | gQ.z
//   6: gP   7: gV   8: gW   (g* = d loss / d aligned p, q, v, omega_w)   9..12: rows of d (f', tau') / d a_lag
#include "reach_core.cuh"

namespace gr {

constexpr int kReachBlock = 64;

// lanes of this warp that own an env (call right after the `i >= num_envs` exit)
__device__ __forceinline__ unsigned warp_live_mask() {
#ifdef GR_CPU_EMUL
  return 1u;
#else
  return __activemask();
#endif
}

template <bool kPhilox>
struct ReachRand {
  const float* row; Philox ph;
  __device__ __forceinline__ ReachRand(const GrRandom& r, int i, int env_id)
      : row(kPhilox ? nullptr : r.rnd + (int64_t)i * GR_REACH_RND_STRIDE), ph(r.seed, (uint32_t)env_id, r.step) {}
  // 4 consecutive uniform slots [4*call, 4*call + 4)
  __device__ __forceinline__ float4 u4(int call) const {
    if (!kPhilox) return *reinterpret_cast<const float4*>(row + 4 * call);
    const uint4 w = ph((uint32_t)call);
    return make_float4(u01(w.x), u01(w.y), u01(w.z), u01(w.w));
  }
  // slot 13: the one normal draw; Philox mode takes it from the two spare words of call 5
  __device__ __forceinline__ float normal13() const {
    if (!kPhilox) return row[13];
    const uint4 w = ph(5u);
    return box_muller(w.z, w.w).x;
  }
};

struct ReachRegs {
  Q4 q; V3 p; float f; V3 v; int eplen; V3 om; float time_left; V3 tau; V3 aacc; bool fresh, cold_stale; float4 fifo; V3 target; float4 raw;
  float4 eps0, eps1; float eps8, eps9; V3 k2, k1; float thr;
};

// read-mostly planes (drag, thr_est_error): rewritten only by a reset, which the hot planes flag (fresh)
template <bool kVolatile>
__device__ __forceinline__ void reach_load_cold(ReachRegs& e, const float4* __restrict__ P, int i) {
  const float4 c0 = kVolatile ? __ldcv(P + ridx(RPL_DRAG2, i)) : __ldg(P + ridx(RPL_DRAG2, i)), c1 = kVolatile ? __ldcv(P + ridx(RPL_DRAG1, i)) : __ldg(P + ridx(RPL_DRAG1, i));
  e.k2 = xyz(c0); e.k1 = xyz(c1); e.thr = c1.w;
}
__device__ __forceinline__ void reach_load_hot(ReachRegs& e, const float4* __restrict__ P, int i) {
  const float4 a0 = __ldcs(P + ridx(RPL_QUAT, i)), a1 = __ldcs(P + ridx(RPL_POS, i)), a2 = __ldcs(P + ridx(RPL_LINVEL, i)), a3 = __ldcs(P + ridx(RPL_ANGVEL, i)),
               a4 = __ldcs(P + ridx(RPL_TORQUE, i)), a5 = __ldcs(P + ridx(RPL_ANGACC, i)), a6 = __ldcs(P + ridx(RPL_FIFO, i)), a7 = __ldcs(P + ridx(RPL_TARGET, i)),
               a8 = __ldcs(P + ridx(RPL_EPSUM0, i)), a9 = __ldcs(P + ridx(RPL_EPSUM1, i)), a10 = __ldcs(P + ridx(RPL_EPSUM2, i));
  e.q = quat(a0); e.p = xyz(a1); e.f = a1.w; e.v = xyz(a2); e.eplen = __float_as_int(a2.w); e.om = xyz(a3); e.time_left = a3.w;
  // ANGACC.w: 1 = the env was reset by the previous step (ActionManager latches zeroed), 2 = a window kernel rewrote the read-mostly
  // planes in an earlier step of its window (only the prefetch needs to know), 0 = neither
  e.tau = xyz(a4); e.aacc = xyz(a5); e.fresh = a5.w == 1.0f; e.cold_stale = a5.w == 2.0f; e.fifo = a6; e.target = xyz(a7);
  e.raw = make_float4(a7.w, a10.z, a10.w, a4.w);
  e.eps0 = a8; e.eps1 = a9; e.eps8 = a10.x; e.eps9 = a10.y;
}
__device__ __forceinline__ void reach_load(ReachRegs& e, const float4* __restrict__ P, int i) {
  reach_load_cold<false>(e, P, i);
  reach_load_hot(e, P, i);
}
__device__ __forceinline__ void reach_store(const ReachRegs& e, float4* __restrict__ P, int i, bool cold_dirty, bool flag_stale = false) {
  __stcs(P + ridx(RPL_QUAT, i), pack(e.q));
  __stcs(P + ridx(RPL_POS, i), pack(e.p, e.f));
  __stcs(P + ridx(RPL_LINVEL, i), pack(e.v, __int_as_float(e.eplen)));
  __stcs(P + ridx(RPL_ANGVEL, i), pack(e.om, e.time_left));
  __stcs(P + ridx(RPL_TORQUE, i), pack(e.tau, e.raw.w));
  __stcs(P + ridx(RPL_ANGACC, i), pack(e.aacc, e.fresh ? 1.0f : (flag_stale ? 2.0f : 0.0f)));
  __stcs(P + ridx(RPL_FIFO, i), e.fifo);
  __stcs(P + ridx(RPL_TARGET, i), pack(e.target, e.raw.x));
  __stcs(P + ridx(RPL_EPSUM0, i), e.eps0);
  __stcs(P + ridx(RPL_EPSUM1, i), e.eps1);
  __stcs(P + ridx(RPL_EPSUM2, i), make_float4(e.eps8, e.eps9, e.raw.y, e.raw.z));
  if (cold_dirty) {
    P[ridx(RPL_DRAG2, i)] = pack(e.k2, 0.0f);
    P[ridx(RPL_DRAG1, i)] = pack(e.k1, e.thr);
  }
}

// tanh through ex2 + rcp (gr_math.cuh fm_tanh, |err| ~ 1e-7), as in racing_step_core.cuh
__device__ __forceinline__ float4 tanh4p(float4 a) { return make_float4(fm_tanh(a.x), fm_tanh(a.y), fm_tanh(a.z), fm_tanh(a.w)); }

// _reset_idx for one env (L/envs/manager_based_diff_rl_env.py:362-410): reset_root_state_uniform, ActionManager.reset,
// controller / dynamics reset with the drag re-draw (droneDynamics.py:50-58), thr_est_error (diff_action.py:233), command resample.
template <bool kPhilox>
__device__ __forceinline__ void reach_reset_env(const GrReachConfig& c, ReachRegs& e, const ReachRand<kPhilox>& rs) {
  const float4 u0 = rs.u4(0), u1 = rs.u4(1), u2 = rs.u4(2), u3 = rs.u4(3), u4 = rs.u4(4);
  const float us[6] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y};
  float sm[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) sm[k] = us[k] * (c.reset_hi[k] - c.reset_lo[k]) + c.reset_lo[k];       // sample_uniform
  e.p = v3(c.default_pos[0] + sm[0], c.default_pos[1] + sm[1], c.default_pos[2] + sm[2]);
  e.q = quat_from_euler_xyz(sm[3], sm[4], sm[5]);                     // quat_mul((1,0,0,0), delta) == delta
  e.v = v3(0.f, 0.f, 0.f); e.om = e.v; e.aacc = e.v; e.tau = e.v; e.f = 0.0f;
  if (c.random_drag) {
    const float z = c.z_drag + u1.z * c.z_drag_rand;
    const float b2 = c.drag2 * c.mass, b1 = c.drag1 * c.mass;
    e.k2 = v3(b2 + u1.w * c.drag2_rand, b2 + u2.x * c.drag2_rand, (b2 + u2.y * c.drag2_rand) * z);
    e.k1 = v3(b1 + u2.z * c.drag1_rand, b1 + u2.w * c.drag1_rand, (b1 + u3.x * c.drag1_rand) * z);
  }
  e.thr = 1.0f + rs.normal13() * c.thr_err_reset_std;
  e.target = v3(e.p.x + (u3.z * (c.cmd_hi[0] - c.cmd_lo[0]) + c.cmd_lo[0]), e.p.y + (u3.w * (c.cmd_hi[1] - c.cmd_lo[1]) + c.cmd_lo[1]),
                e.p.z + (u4.x * (c.cmd_hi[2] - c.cmd_lo[2]) + c.cmd_lo[2]));
  e.time_left = c.resample_time;
  e.eplen = 0;
  e.fresh = true;
  e.eps0 = make_float4(0.f, 0.f, 0.f, 0.f); e.eps1 = e.eps0; e.eps8 = 0.f; e.eps9 = 0.f;
}

// observation_manager.compute (QD/reach_target_lv_env.py:83-104): 17 floats, row-major
__device__ __forceinline__ void reach_write_obs(const GrReachConfig& c, const ReachRegs& e, float4 last_action, float* __restrict__ obs, int i, unsigned live) {
  const V3 vb = quat_rotate_inverse(e.q, e.v);
  const V3 d = quat_rotate_inverse(e.q, e.target - e.p);
  float4 la = last_action;
  if (c.last_action_modified) {                                           // QD/mdp/observation.py:55-63
    const float4 th = tanh4p(e.raw);
    la = make_float4((th.x * c.action_scale[0] + c.action_offset[0]) / c.mass, th.y * c.action_scale[1] + c.action_offset[1],
                     th.z * c.action_scale[2] + c.action_offset[2], th.w * c.action_scale[3] + c.action_offset[3]);
  }
#ifdef GR_CPU_EMUL
  float* o = obs + (int64_t)i * GR_REACH_OBS_DIM;
#else
  // rows are 68 B: a warp parks its 32 rows in shared memory (stride 17: conflict-free) and streams them out as one
  // contiguous 2176-byte block; `live` = the lanes of this warp that own an env (contiguous from lane 0)
  __shared__ float stage[kReachBlock / 32][32 * GR_REACH_OBS_DIM];
  const int lane = threadIdx.x & 31;
  float* o = stage[threadIdx.x >> 5] + lane * GR_REACH_OBS_DIM;
#endif
  o[0] = vb.x; o[1] = vb.y; o[2] = vb.z; o[3] = e.om.x; o[4] = e.om.y; o[5] = e.om.z;
  o[6] = la.x; o[7] = la.y; o[8] = la.z; o[9] = la.w;
  o[10] = e.q.w; o[11] = e.q.x; o[12] = e.q.y; o[13] = e.q.z; o[14] = d.x; o[15] = d.y; o[16] = d.z;
#ifndef GR_CPU_EMUL
  __syncwarp(live);
  const int nlive = __popc(live);
  const float* src = stage[threadIdx.x >> 5];
  float* dst = obs + (int64_t)(i - lane) * GR_REACH_OBS_DIM;
  for (int k = lane; k < nlive * GR_REACH_OBS_DIM; k += nlive) __stcs(dst + k, src[k]);
  __syncwarp(live);                                       // the staging rows are reused by the next call (window kernel)
#endif
}

// step-invariant constants (formed before the grid dependency in the single-step kernel, once per window in the rollout kernel)
struct ReachConsts { float dt; V3 J, Jinv; float inv_m, ef; V3 etau; };
__device__ __forceinline__ ReachConsts reach_consts(const GrReachConfig& cfg) {
  ReachConsts k;
  k.dt = cfg.dt;
  k.J = v3(cfg.inertia[0], cfg.inertia[1], cfg.inertia[2]);
  k.Jinv = v3(1.0f / k.J.x, 1.0f / k.J.y, 1.0f / k.J.z);
  k.inv_m = 1.0f / cfg.mass;
  k.ef = expf(-k.dt / cfg.thrust_delay);
  k.etau = v3(0.f, 0.f, 0.f);
  if (cfg.controller == GR_CTRL_CTBR) k.etau = v3(expf(-k.dt / cfg.torque_delay[0]), expf(-k.dt / cfg.torque_delay[1]), expf(-k.dt / cfg.torque_delay[2]));
  return k;
}
// destinations of ONE step (nullptr: skip) and what the body hands back
struct ReachStepPtrs { float* loss; float* loss_terms; float4* tape; float* log_accum; float* obs; float* obs2; };
struct ReachStepOut { float reward; bool terminated, time_out, reset; float terms[GR_REACH_NUM_REWARD_TERMS]; };

// Sections 1-11 of ManagerBasedDiffRLEnv.step for ONE env held in registers: shared by the single-step kernel and the
// multi-step window kernel (gr_reach_rollout_fwd).  `e` = state before the step in, after the step (and a reset) out.
template <bool kDiff, bool kPhilox>
__device__ __forceinline__ void reach_step_body(const GrReachConfig& cfg, const ReachConsts& kc, ReachRegs& e, const float4 a_t,
                                                const ReachRand<kPhilox>& rs, const ReachStepPtrs& io, const int i, const unsigned live,
                                                ReachStepOut& out) {
  const float dt = kc.dt, inv_m = kc.inv_m, ef = kc.ef;
  const V3 J = kc.J, Jinv = kc.Jinv, etau = kc.etau;
  // ---- 1. process_action (L/managers/action_manager.py:44-45; QD/mdp/diff_action.py:156-176) ----
  const bool fresh0 = e.fresh;
  const float4 a_lag = e.fifo;                                   // lag 1: the applied action is a_{t-1}
  const float4 prev = e.fresh ? make_float4(0.f, 0.f, 0.f, 0.f) : a_lag;      // ActionManager.prev_action after the latch
  e.raw = a_lag;
  const float4 th = tanh4p(a_lag);
  float cmd[4];
  float dcmd[4];                                                 // d cmd_k / d a_lag_k
  if (cfg.sim2real_test) {
    cmd[0] = a_lag.x * cfg.mass; cmd[1] = a_lag.y; cmd[2] = a_lag.z; cmd[3] = a_lag.w;
    dcmd[0] = dcmd[1] = dcmd[2] = dcmd[3] = 0.0f;                // actions.clone().detach()
  } else {
    cmd[0] = (th.x * cfg.action_scale[0] + cfg.action_offset[0]) * e.thr;
    cmd[1] = th.y * cfg.action_scale[1] + cfg.action_offset[1];
    cmd[2] = th.z * cfg.action_scale[2] + cfg.action_offset[2];
    cmd[3] = th.w * cfg.action_scale[3] + cfg.action_offset[3];
    dcmd[0] = (1.0f - th.x * th.x) * cfg.action_scale[0] * e.thr;
    dcmd[1] = (1.0f - th.y * th.y) * cfg.action_scale[1];
    dcmd[2] = (1.0f - th.z * th.z) * cfg.action_scale[2];
    dcmd[3] = (1.0f - th.w * th.w) * cfg.action_scale[3];
  }
  // get_state_from_sim (QD/mdp/diff_action.py:126-154)
  const Q4 q0 = e.q; const V3 p0 = e.p, v0 = e.v, om_b = e.om, target0 = e.target;
  const V3 v_b = quat_rotate_inverse(q0, v0);
  const V3 gyro = cross(om_b, J * om_b);

  // ---- controller ----
  float f_new; V3 tau_new;
  float Jac[4][4];                                               // d (f', tau') / d a_lag
  if (cfg.controller == GR_CTRL_CTBR) {                          // L/controllers/controller_diff.py:120-138
    const float sb = cfg.body_rate_bound;
    const float thrust_des = fminf(fmaxf(cmd[0], cfg.thrust_lo), cfg.thrust_hi);
    f_new = (1.0f - ef) * thrust_des + ef * e.f;
    const V3 rate_c = v3(fminf(fmaxf(cmd[1], -sb), sb), fminf(fmaxf(cmd[2], -sb), sb), fminf(fmaxf(cmd[3], -sb), sb));
    const V3 kp = v3(cfg.kp[0], cfg.kp[1], cfg.kp[2]), kd = v3(cfg.kd[0], cfg.kd[1], cfg.kd[2]);
    const V3 torque_des = J * (kp * (rate_c - om_b)) + gyro - kd * e.aacc;
    const V3 om1 = v3(1.0f - etau.x, 1.0f - etau.y, 1.0f - etau.z);
    tau_new = om1 * torque_des + etau * e.tau;
    if (kDiff) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) Jac[r][k] = 0.0f;
      Jac[0][0] = (cmd[0] >= cfg.thrust_lo && cmd[0] <= cfg.thrust_hi) ? dcmd[0] * (1.0f - ef) : 0.0f;
      Jac[1][1] = (cmd[1] >= -sb && cmd[1] <= sb) ? dcmd[1] * om1.x * J.x * kp.x : 0.0f;
      Jac[2][2] = (cmd[2] >= -sb && cmd[2] <= sb) ? dcmd[2] * om1.y * J.y * kp.y : 0.0f;
      Jac[3][3] = (cmd[3] >= -sb && cmd[3] <= sb) ? dcmd[3] * om1.z * J.z * kp.z : 0.0f;
    }
  } else {                                                       // LVController / PSController (:242-291, :378-430)
    if (kDiff) {
      D4 c4[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { c4[k] = D4(cmd[k]); c4[k].d[k] = dcmd[k]; }
      D4 thr_d; D4 tau_d[3];
      outer_loop<D4>(cfg, p0, q0, v0, om_b, c4, thr_d, tau_d);
      f_new = (1.0f - ef) * thr_d.v + ef * e.f;
      tau_new = v3(tau_d[0].v, tau_d[1].v, tau_d[2].v);
#pragma unroll
      for (int k = 0; k < 4; ++k) { Jac[0][k] = (1.0f - ef) * thr_d.d[k]; Jac[1][k] = tau_d[0].d[k]; Jac[2][k] = tau_d[1].d[k]; Jac[3][k] = tau_d[2].d[k]; }
    } else {
      float thr_f; float tau_f[3];
      outer_loop<float>(cfg, p0, q0, v0, om_b, cmd, thr_f, tau_f);
      f_new = (1.0f - ef) * thr_f + ef * e.f;
      tau_new = v3(tau_f[0], tau_f[1], tau_f[2]);
    }
  }

  // ---- DroneDynamics.step (QD/mdp/dynamics/droneDynamics.py:119-135) ----
  const V3 F_b = v3(0.f, 0.f, f_new) - e.k2 * v_b * vabs(v_b) - e.k1 * v_b;
  const V3 acc = v3(0.f, 0.f, -cfg.gravity) + quat_rotate(q0, F_b) * inv_m;
  const V3 alpha = Jinv * (tau_new - gyro);
  const V3 p1 = p0 + v0 * dt + (0.5f * dt * dt) * acc;
  const float hdt = 0.5f * dt;
  const Q4 qt = Q4{q0.w - hdt * (q0.x * om_b.x + q0.y * om_b.y + q0.z * om_b.z), q0.x + hdt * (q0.w * om_b.x + q0.y * om_b.z - q0.z * om_b.y),
                   q0.y + hdt * (q0.w * om_b.y + q0.z * om_b.x - q0.x * om_b.z), q0.z + hdt * (q0.w * om_b.z + q0.x * om_b.y - q0.y * om_b.x)};
  const float qn = sqrtf(qt.w * qt.w + qt.x * qt.x + qt.y * qt.y + qt.z * qt.z);
  const Q4 q1 = Q4{qt.w / qn, qt.x / qn, qt.y / qn, qt.z / qn};
  const V3 v1 = v0 + acc * dt;
  const V3 omb1 = om_b + alpha * dt;
  const V3 omw1 = quat_rotate(q1, omb1);
  const V3 Dg = -(2.0f * (e.k2 * vabs(v_b)) + e.k1);             // d F_b / d v_b (diagonal), with the drag of THIS step

  float4* __restrict__ T = io.tape;
  const bool tape = kDiff && io.tape != nullptr;

  // ---- 2./3. closure physics (oracle/reach_oracle.py R.1) + align: the carried state is the nominal one ----
  e.p = p1; e.q = q1; e.v = v1; e.om = omb1; e.aacc = alpha; e.f = f_new; e.tau = tau_new;

  // ---- 4./5. counters + terminations ----
  e.eplen += 1;
  const bool time_out = e.eplen >= cfg.max_episode_length;
  const bool terminated = cfg.term_oob && ((p1.z < cfg.oob_lo) || (p1.z > cfg.oob_hi));

  // ---- 6. rewards (QD/mdp/rewards.py:30-101); the body-frame command is the one of the previous command update ----
  const V3 des_b = quat_rotate_inverse(q0, target0 - p0);
  const float dist = norm(des_b);
  float terms[GR_REACH_NUM_REWARD_TERMS];
  {
    const V3 vb1 = quat_rotate_inverse(q1, v1);
    terms[0] = 1.0f / (1.0f + dist);
    const float qe = sqrtf((q1.w - 1.0f) * (q1.w - 1.0f) + q1.x * q1.x + q1.y * q1.y + q1.z * q1.z);
    terms[1] = 1.0f / (1.0f + qe);
    const V3 nv = vb1 / fmaxf(norm(vb1), 1e-12f), nd = des_b / fmaxf(dist, 1e-12f);
    terms[2] = dist < cfg.move_in_dir_thr ? 1.0f : dot(nv, nd);
    const float d0 = a_t.x - prev.x, d1 = a_t.y - prev.y, d2 = a_t.z - prev.z, d3 = a_t.w - prev.w;
    terms[3] = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    terms[4] = dist < cfg.reach_thr ? 1.0f : 0.0f;
    terms[5] = omb1.x * omb1.x + omb1.y * omb1.y + (omb1.z * 5.0f) * (omb1.z * 5.0f);
    terms[6] = norm(acc);
    terms[7] = norm(quat_rotate(q1, alpha));
    terms[8] = terminated ? 1.0f : 0.0f;
    const float nw = norm(omw1);
    terms[9] = (dist < cfg.hover_thr ? 1.0f : 0.0f) / (1.0f + sqrtf(dot(v1, v1) + dot(omw1, omw1)) + cfg.hover_ratio * nw);
  }
  float reward = 0.0f;
#pragma unroll
  for (int k = 0; k < GR_REACH_NUM_REWARD_TERMS; ++k) {
    const float tw = cfg.w_reward[k] != 0.0f ? terms[k] * cfg.w_reward[k] : 0.0f;
    const float value = tw * dt;
    reward += value;
    terms[k] = tw;
    if (k < 4) (&e.eps0.x)[k] += value; else if (k < 8) (&e.eps1.x)[k - 4] += value; else if (k == 8) e.eps8 += value; else e.eps9 += value;
  }

  // ---- 7. reset ----
  const bool reset = terminated || time_out;
  if (reset) {
    if (io.log_accum) {
      float* row = io.log_accum + (size_t)((i >> 5) & (GR_LOG_SHARDS - 1)) * GR_LOG_SLOTS;
      atomicAdd(row + GR_REACH_LOG_NUM_RESET, 1.0f);
      atomicAdd(row + GR_REACH_LOG_SUM_POS_ERR, norm(target0 - p0));           // the metric of the previous command update
      if (time_out) atomicAdd(row + GR_REACH_LOG_NUM_TIMEOUT, 1.0f);
      if (terminated) atomicAdd(row + GR_REACH_LOG_NUM_TERMINATED, 1.0f);
#pragma unroll
      for (int k = 0; k < GR_REACH_NUM_REWARD_TERMS; ++k)
        atomicAdd(row + GR_REACH_LOG_SUM_EPSUM + k, k < 4 ? (&e.eps0.x)[k] : (k < 8 ? (&e.eps1.x)[k - 4] : (k == 8 ? e.eps8 : e.eps9)));
    }
    reach_reset_env<kPhilox>(cfg, e, rs);
  } else {
    e.fresh = false;
  }

  // ---- 8. CommandTerm.compute: timer, resample, body-frame command (QD/mdp/commands.py:98-134) ----
  e.time_left -= dt;
  if (e.time_left <= 0.0f) {
    const float4 u4 = rs.u4(4), u5 = rs.u4(5);
    e.time_left = cfg.resample_time;
    e.target = v3(e.p.x + (u4.z * (cfg.cmd_hi[0] - cfg.cmd_lo[0]) + cfg.cmd_lo[0]), e.p.y + (u4.w * (cfg.cmd_hi[1] - cfg.cmd_lo[1]) + cfg.cmd_lo[1]),
                  e.p.z + (u5.x * (cfg.cmd_hi[2] - cfg.cmd_lo[2]) + cfg.cmd_lo[2]));
  }

  // ---- 9. losses on the aligned (pre-reset) state against the current target (QD/mdp/losses.py:32-67) + tape ----
  if (kDiff) {
    const V3 dd = e.target - p1;
    const float nd = norm(dd);
    const float qe = sqrtf((q1.w - 1.0f) * (q1.w - 1.0f) + q1.x * q1.x + q1.y * q1.y + q1.z * q1.z);
    const float nv = norm(v1), nw = norm(omw1);
    const float cs = cos_sim(v1, dd);
    const float far = nd > cfg.loss_dir_thr ? 1.0f : 0.0f;
    const float l0 = cfg.w_loss[0] != 0.0f ? nd * cfg.w_loss[0] : 0.0f;
    const float l1 = cfg.w_loss[1] != 0.0f ? qe * cfg.w_loss[1] : 0.0f;
    const float l2 = cfg.w_loss[2] != 0.0f ? ((1.0f - cs) * far) * cfg.w_loss[2] : 0.0f;
    const float l3 = cfg.w_loss[3] != 0.0f ? (nv + cfg.loss_smooth_ratio * nw) * cfg.w_loss[3] : 0.0f;
    if (io.loss) io.loss[i] = (((0.0f + l0) + l1) + l2) + l3;
    if (io.loss_terms) *reinterpret_cast<float4*>(io.loss_terms + (int64_t)i * 4) = make_float4(l0, l1, l2, l3);
    if (tape) {
      // d loss / d aligned (p, q, v, omega_w)
      V3 gP = nd > 0.0f ? dd * (-cfg.w_loss[0] / nd) : v3(0.f, 0.f, 0.f);
      Q4 gQ = Q4{0.f, 0.f, 0.f, 0.f};
      if (qe > 0.0f) { const float s = cfg.w_loss[1] / qe; gQ = Q4{(q1.w - 1.0f) * s, q1.x * s, q1.y * s, q1.z * s}; }
      V3 gV = nv > 0.0f ? v1 * (cfg.w_loss[3] / nv) : v3(0.f, 0.f, 0.f);
      const V3 gW = nw > 0.0f ? omw1 * (cfg.w_loss[3] * cfg.loss_smooth_ratio / nw) : v3(0.f, 0.f, 0.f);
      if (cfg.w_loss[2] != 0.0f && far != 0.0f) {
        const float nx = fmaxf(nv, 1e-8f), ny = fmaxf(nd, 1e-8f);
        const V3 dcdx = dd / (nx * ny) - v1 * (cs / (nx * nx));
        const V3 dcdy = v1 / (nx * ny) - dd * (cs / (ny * ny));
        gV = gV - cfg.w_loss[2] * dcdx;
        gP = gP + cfg.w_loss[2] * dcdy;                                   // d dd / d p = -1
      }
      __stcs(T + rtidx(0, i), pack(q0));
      __stcs(T + rtidx(1, i), pack(om_b, fresh0 ? 1.0f : 0.0f));
      __stcs(T + rtidx(2, i), pack(F_b, gQ.w));
      __stcs(T + rtidx(3, i), pack(Dg, gQ.x));
      __stcs(T + rtidx(4, i), pack(v1, gQ.y));
      __stcs(T + rtidx(5, i), pack(omb1, gQ.z));
      __stcs(T + rtidx(6, i), pack(gP, 0.0f));
      __stcs(T + rtidx(7, i), pack(gV, 0.0f));
      __stcs(T + rtidx(8, i), pack(gW, 0.0f));
#pragma unroll
      for (int r = 0; r < 4; ++r) __stcs(T + rtidx(9 + r, i), make_float4(Jac[r][0], Jac[r][1], Jac[r][2], Jac[r][3]));
    }
  }

  // ---- 11. observations on the post-reset state ----
  if (io.obs) reach_write_obs(cfg, e, reset ? make_float4(0.f, 0.f, 0.f, 0.f) : a_t, io.obs, i, live);      // (warp-uniform)
  if (io.obs2) reach_write_obs(cfg, e, reset ? make_float4(0.f, 0.f, 0.f, 0.f) : a_t, io.obs2, i, live);
  e.fifo = a_t;
  out.reward = reward; out.terminated = terminated; out.time_out = time_out; out.reset = reset;
#pragma unroll
  for (int k = 0; k < GR_REACH_NUM_REWARD_TERMS; ++k) out.terms[k] = terms[k];
}

template <bool kDiff, bool kPhilox>
__global__ void __launch_bounds__(kReachBlock) reach_step_fwd_kernel(const GrReachConfig cfg, const GrReachState st, const GrRandom rng, const GrReachStepIO io) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.num_envs) return;
  const unsigned live = warp_live_mask();
  float4* __restrict__ P = reinterpret_cast<float4*>(st.planes);
  ReachRegs e;
  // Programmatic dependent launch: everything that does not depend on the previous kernel in the stream -- the read-mostly
  // planes, the filter constants -- is fetched / formed before the grid dependency and overlaps that kernel's tail.
  reach_load_cold<false>(e, P, i);
  const ReachRand<kPhilox> rs(rng, i, st.env_id_offset + i);
  const ReachConsts kc = reach_consts(cfg);
  pdl_wait();
  reach_load_hot(e, P, i);
  const float4 a_t = __ldcs(reinterpret_cast<const float4*>(io.action) + i);
  pdl_launch_dependents();
  if (e.fresh || e.cold_stale) reach_load_cold<true>(e, P, i);   // a reset rewrote this env's read-mostly planes: the prefetched copy may be stale

  const ReachStepPtrs sp{io.loss, io.loss_terms, reinterpret_cast<float4*>(io.tape), io.log_accum, io.obs, nullptr};
  ReachStepOut so;
  reach_step_body<kDiff, kPhilox>(cfg, kc, e, a_t, rs, sp, i, live, so);
  io.reward[i] = so.reward;
  io.terminated[i] = so.terminated ? 1 : 0;
  io.time_out[i] = so.time_out ? 1 : 0;
  if (io.dones) io.dones[i] = so.reset ? 1 : 0;
  if (io.reward_terms) {
#pragma unroll
    for (int k = 0; k < GR_REACH_NUM_REWARD_TERMS; ++k) io.reward_terms[(int64_t)i * GR_REACH_NUM_REWARD_TERMS + k] = so.terms[k];
  }
  reach_store(e, P, i, so.reset, false);
}

// T steps in ONE launch for actions known in advance (gr_reach_rollout_fwd): the same body, the env state in registers over the
// window.  Step t uses the random stream (seed, env, rng.step + t) or, in dense mode, rng.rnd + t * N * GR_REACH_RND_STRIDE.
template <bool kDiff, bool kPhilox>
__global__ void __launch_bounds__(kReachBlock) reach_rollout_fwd_kernel(const GrReachConfig cfg, const GrReachState st, const GrRandom rng, const GrReachRolloutIO rio) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int N = st.num_envs, T = rio.T;
  if (i >= N) return;
  const unsigned live = warp_live_mask();
  float4* __restrict__ P = reinterpret_cast<float4*>(st.planes);
  ReachRegs e;
  reach_load(e, P, i);
  const ReachConsts kc = reach_consts(cfg);
  const int64_t tape_step = (int64_t)(rio.tape_stride / 32) * GR_REACH_TAPE_PLANES * 32;          // float4 per tape step
  bool any_reset = false;
  float4 a_next = __ldcs(reinterpret_cast<const float4*>(rio.actions) + i);
#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    const int64_t tn = (int64_t)t * N + i;
    const float4 a_t = a_next;
    if (t + 1 < T) a_next = __ldcs(reinterpret_cast<const float4*>(rio.actions) + (int64_t)(t + 1) * N + i);     // one step ahead
    GrRandom rt = rng;
    rt.step = rng.step + (uint32_t)t;
    if (!kPhilox) rt.rnd = rng.rnd + (int64_t)t * N * GR_REACH_RND_STRIDE;
    const ReachRand<kPhilox> rs(rt, i, st.env_id_offset + i);
    ReachStepPtrs sp{nullptr, nullptr, nullptr, rio.log_accum, nullptr, nullptr};
    if (kDiff) {
      sp.loss = rio.loss ? rio.loss + (int64_t)t * N : nullptr;
      sp.loss_terms = rio.loss_terms ? rio.loss_terms + (int64_t)t * N * GR_REACH_NUM_LOSS_TERMS : nullptr;
      sp.tape = rio.tape ? reinterpret_cast<float4*>(rio.tape) + (int64_t)t * tape_step : nullptr;
    }
    if (rio.obs_seq) sp.obs = rio.obs_seq + (int64_t)t * N * GR_REACH_OBS_DIM;
    if (t == T - 1) sp.obs2 = rio.obs_out;       // the "observation after the window" buffer
    ReachStepOut so;
    reach_step_body<kDiff, kPhilox>(cfg, kc, e, a_t, rs, sp, i, live, so);
    any_reset |= so.reset;
    if (rio.reward) rio.reward[tn] = so.reward;
    if (rio.dones) rio.dones[tn] = so.reset ? 1 : 0;
    if (rio.terminated) rio.terminated[tn] = so.terminated ? 1 : 0;
    if (rio.time_out) rio.time_out[tn] = so.time_out ? 1 : 0;
  }
  // cold planes rewritten if any step of the window reset this env; unless the LAST step did (fresh), the next single step could not
  // tell that its pre-dependency prefetch of them is stale: flag it
  reach_store(e, P, i, any_reset, any_reset && !e.fresh);
}

// ManagerBasedRLEnv.reset / _reset_idx(mask) + observations; mode 0 = masked, 1 = all, 2 = observe only
template <bool kPhilox>
__global__ void __launch_bounds__(kReachBlock) reach_reset_kernel(const GrReachConfig cfg, const GrReachState st, const GrRandom rng, const uint8_t* __restrict__ mask,
                                                                  const int mode, float* __restrict__ obs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.num_envs) return;
  const unsigned live = warp_live_mask();
  float4* __restrict__ P = reinterpret_cast<float4*>(st.planes);
  ReachRegs e;
  reach_load(e, P, i);
  const bool doit = mode == 1 || (mode == 0 && mask[i] != 0);
  if (doit) {
    if (!cfg.random_drag || (e.k2.x == 0.0f && e.k1.x == 0.0f)) {        // first reset of a zero-filled state: nominal drag (droneDynamics.py:23-34)
      const float b2 = cfg.drag2 * cfg.mass, b1 = cfg.drag1 * cfg.mass;
      e.k2 = v3(b2, b2, b2 * cfg.z_drag); e.k1 = v3(b1, b1, b1 * cfg.z_drag);
    }
    const ReachRand<kPhilox> rs(rng, i, st.env_id_offset + i);
    reach_reset_env<kPhilox>(cfg, e, rs);
  }
  if (obs) reach_write_obs(cfg, e, e.fresh ? make_float4(0.f, 0.f, 0.f, 0.f) : e.fifo, obs, i, live);
  if (doit) reach_store(e, P, i, true);
}

__global__ void reach_fill_rand_kernel(float* __restrict__ rnd, int num_envs, int env_id_offset, uint64_t seed, uint32_t step) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_envs) return;
  const GrRandom r{nullptr, seed, step};
  const ReachRand<true> rs(r, i, env_id_offset + i);
  float4* row = reinterpret_cast<float4*>(rnd + (int64_t)i * GR_REACH_RND_STRIDE);
  for (int c = 0; c < GR_REACH_RND_STRIDE / 4; ++c) row[c] = rs.u4(c);
  rnd[(int64_t)i * GR_REACH_RND_STRIDE + 13] = rs.normal13();
}

}  // namespace gr

#ifndef GR_CPU_EMUL
using namespace gr;

static int reach_check(const GrReachConfig* cfg, const GrReachState* st) {
  if (!cfg || !st || !st->planes) return GR_ERR_NULL;
  if (st->num_envs <= 0 || st->plane_stride < ((st->num_envs + 31) & ~31)) return GR_ERR_SIZE;
  if (reinterpret_cast<uintptr_t>(st->planes) & 15u) return GR_ERR_ALIGN;
  if (cfg->controller < GR_CTRL_CTBR || cfg->controller > GR_CTRL_PS || cfg->dt <= 0.0f || cfg->mass <= 0.0f || cfg->thrust_delay <= 0.0f) return GR_ERR_CONFIG;
  return GR_OK;
}

extern "C" int gr_reach_step_fwd(const GrReachConfig* cfg, const GrReachState* st, const GrRandom* rng, const GrReachStepIO* io, void* stream) {
  if (!rng || !io || !io->action || !io->obs || !io->reward || !io->terminated || !io->time_out) return GR_ERR_NULL;
  if (const int rc = reach_check(cfg, st)) return rc;
  auto mis = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; };
  if (mis(io->action) || (io->tape && mis(io->tape)) || (io->loss_terms && mis(io->loss_terms)) || (rng->rnd && mis(rng->rnd))) return GR_ERR_ALIGN;
  if (io->tape && io->tape_stride < ((st->num_envs + 31) & ~31)) return GR_ERR_SIZE;
  const bool diff = io->loss || io->tape || io->loss_terms, philox = rng->rnd == nullptr;
  const int grid = (st->num_envs + kReachBlock - 1) / kReachBlock;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid);
  lc.blockDim = dim3((unsigned)kReachBlock);
  lc.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // the kernel's prologue overlaps the previous kernel's tail
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  if (diff) return (int)(philox ? cudaLaunchKernelEx(&lc, reach_step_fwd_kernel<true, true>, *cfg, *st, *rng, *io) : cudaLaunchKernelEx(&lc, reach_step_fwd_kernel<true, false>, *cfg, *st, *rng, *io));
  return (int)(philox ? cudaLaunchKernelEx(&lc, reach_step_fwd_kernel<false, true>, *cfg, *st, *rng, *io) : cudaLaunchKernelEx(&lc, reach_step_fwd_kernel<false, false>, *cfg, *st, *rng, *io));
}

extern "C" int gr_reach_rollout_fwd(const GrReachConfig* cfg, const GrReachState* st, const GrRandom* rng, const GrReachRolloutIO* io, void* stream) {
  if (!rng || !io || !io->actions || !io->obs_out) return GR_ERR_NULL;
  if (const int rc = reach_check(cfg, st)) return rc;
  if (io->T < 1) return GR_ERR_SIZE;
  auto mis = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; };
  if (mis(io->actions) || (io->tape && mis(io->tape)) || (io->loss_terms && mis(io->loss_terms)) || (rng->rnd && mis(rng->rnd))) return GR_ERR_ALIGN;
  if (io->tape && (io->tape_stride < ((st->num_envs + 31) & ~31) || (io->tape_stride & 31))) return GR_ERR_SIZE;
  const bool diff = io->loss || io->tape || io->loss_terms, philox = rng->rnd == nullptr;
  const int grid = (st->num_envs + kReachBlock - 1) / kReachBlock;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (diff) {
    if (philox) reach_rollout_fwd_kernel<true, true><<<grid, kReachBlock, 0, s>>>(*cfg, *st, *rng, *io);
    else reach_rollout_fwd_kernel<true, false><<<grid, kReachBlock, 0, s>>>(*cfg, *st, *rng, *io);
  } else {
    if (philox) reach_rollout_fwd_kernel<false, true><<<grid, kReachBlock, 0, s>>>(*cfg, *st, *rng, *io);
    else reach_rollout_fwd_kernel<false, false><<<grid, kReachBlock, 0, s>>>(*cfg, *st, *rng, *io);
  }
  return (int)cudaGetLastError();
}

static int reach_reset_launch(const GrReachConfig* cfg, const GrReachState* st, const GrRandom* rng, const uint8_t* mask, int mode, float* obs, void* stream) {
  if (const int rc = reach_check(cfg, st)) return rc;
  const int grid = (st->num_envs + kReachBlock - 1) / kReachBlock;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const GrRandom none{nullptr, 0, 0};
  const GrRandom& r = rng ? *rng : none;
  if (r.rnd == nullptr) reach_reset_kernel<true><<<grid, kReachBlock, 0, s>>>(*cfg, *st, r, mask, mode, obs);
  else reach_reset_kernel<false><<<grid, kReachBlock, 0, s>>>(*cfg, *st, r, mask, mode, obs);
  return (int)cudaGetLastError();
}

extern "C" int gr_reach_reset(const GrReachConfig* cfg, const GrReachState* st, const GrRandom* rng, const uint8_t* reset_mask, float* obs, void* stream) {
  if (!rng) return GR_ERR_NULL;
  if (rng->rnd && (reinterpret_cast<uintptr_t>(rng->rnd) & 15u)) return GR_ERR_ALIGN;
  return reach_reset_launch(cfg, st, rng, reset_mask, reset_mask ? 0 : 1, obs, stream);
}

extern "C" int gr_reach_observe(const GrReachConfig* cfg, const GrReachState* st, float* obs, void* stream) {
  if (!obs) return GR_ERR_NULL;
  return reach_reset_launch(cfg, st, nullptr, nullptr, 2, obs, stream);
}

extern "C" int gr_reach_fill_rand(float* rnd, int32_t num_envs, int32_t env_id_offset, uint64_t seed, uint32_t step, void* stream) {
  if (!rnd) return GR_ERR_NULL;
  if (num_envs <= 0) return GR_ERR_SIZE;
  if (reinterpret_cast<uintptr_t>(rnd) & 15u) return GR_ERR_ALIGN;
  reach_fill_rand_kernel<<<(num_envs + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(rnd, num_envs, env_id_offset, seed, step);
  return (int)cudaGetLastError();
}
#endif  // GR_CPU_EMUL

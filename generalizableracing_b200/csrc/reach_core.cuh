// reach_core.cuh -- shared device code of the reach-target tasks (reach_step.cu, reach_bwd.cu): state / tape tile indexing,
// the forward-mode scalar used to take the controller's action Jacobian while the step runs, and the LV / PS outer loop
// (L/controllers/controller_diff.py:242-291, :378-430) written once for float and for that scalar.
#pragma once
#include "gr_common.cuh"

namespace gr {

// plane ids: generalizableracing_b200/layout.py RPL_*
enum ReachPlane : int {
  RPL_QUAT = 0,    // q.w q.x q.y q.z
  RPL_POS = 1,     // pos xyz        | thrust filter state
  RPL_LINVEL = 2,  // v_w xyz        | episode_length (int bits)
  RPL_ANGVEL = 3,  // omega_b xyz    | command time_left
  RPL_TORQUE = 4,  // CTBR torque filter state | raw_actions.w
  RPL_ANGACC = 5,  // alpha_b xyz    | fresh (1.0 = reset at the previous step)
  RPL_FIFO = 6,    // action-lag FIFO: raw a_{t-1}
  RPL_TARGET = 7,  // pose_command_w xyz | raw_actions.x
  RPL_EPSUM0 = 8,  // episode sums 0..3
  RPL_EPSUM1 = 9,  // 4..7
  RPL_EPSUM2 = 10, // 8..9 | raw_actions.y | raw_actions.z      (raw_actions = the lagged action the last step applied)
  RPL_DRAG2 = 11,  // quadratic drag xyz (z * z_drag) | spare          -- planes 11, 12 are rewritten only on reset
  RPL_DRAG1 = 12   // linear drag xyz (z * z_drag)    | thr_est_error
};
__device__ __forceinline__ int64_t ridx(int plane, int i) { return ((int64_t)(i >> 5) * GR_REACH_PLANES + plane) * kTile + (i & 31); }
__device__ __forceinline__ int64_t rtidx(int plane, int i) { return ((int64_t)(i >> 5) * GR_REACH_TAPE_PLANES + plane) * kTile + (i & 31); }

// ---- forward-mode scalar: value + 4 tangents (d / d a_lag[0..3]) ------------------------------------------------------
struct D4 {
  float v; float d[4];
  __device__ __forceinline__ D4() {}
  __device__ __forceinline__ D4(float c) : v(c) { d[0] = d[1] = d[2] = d[3] = 0.0f; }
};
__device__ __forceinline__ D4 operator+(D4 a, D4 b) { D4 r; r.v = a.v + b.v; _Pragma("unroll") for (int k = 0; k < 4; ++k) r.d[k] = a.d[k] + b.d[k]; return r; }
__device__ __forceinline__ D4 operator-(D4 a, D4 b) { D4 r; r.v = a.v - b.v; _Pragma("unroll") for (int k = 0; k < 4; ++k) r.d[k] = a.d[k] - b.d[k]; return r; }
__device__ __forceinline__ D4 operator-(D4 a) { D4 r; r.v = -a.v; _Pragma("unroll") for (int k = 0; k < 4; ++k) r.d[k] = -a.d[k]; return r; }
__device__ __forceinline__ D4 operator*(D4 a, D4 b) { D4 r; r.v = a.v * b.v; _Pragma("unroll") for (int k = 0; k < 4; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k]; return r; }
__device__ __forceinline__ D4 operator/(D4 a, D4 b) {
  D4 r; const float ib = 1.0f / b.v; r.v = a.v * ib;
  _Pragma("unroll") for (int k = 0; k < 4; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * ib;
  return r;
}
__device__ __forceinline__ D4 s_sqrt(D4 a) { D4 r; r.v = sqrtf(a.v); const float h = r.v > 0.0f ? 0.5f / r.v : 0.0f; _Pragma("unroll") for (int k = 0; k < 4; ++k) r.d[k] = a.d[k] * h; return r; }
__device__ __forceinline__ D4 s_sin(D4 a) { D4 r; const float s = sinf(a.v), c = cosf(a.v); r.v = s; _Pragma("unroll") for (int k = 0; k < 4; ++k) r.d[k] = a.d[k] * c; return r; }
__device__ __forceinline__ D4 s_cos(D4 a) { D4 r; const float s = sinf(a.v), c = cosf(a.v); r.v = c; _Pragma("unroll") for (int k = 0; k < 4; ++k) r.d[k] = -a.d[k] * s; return r; }
__device__ __forceinline__ D4 s_minc(D4 a, float c) { return a.v < c ? a : D4(c); }          // torch.min(x, const)
__device__ __forceinline__ D4 s_maxc(D4 a, float c) { return a.v > c ? a : D4(c); }          // clamp_min (F.normalize eps)
__device__ __forceinline__ D4 s_clamp(D4 a, float lo, float hi) { return a.v < lo ? D4(lo) : (a.v > hi ? D4(hi) : a); }
__device__ __forceinline__ float s_val(D4 a) { return a.v; }
__device__ __forceinline__ float s_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ float s_sin(float a) { return sinf(a); }
__device__ __forceinline__ float s_cos(float a) { return cosf(a); }
__device__ __forceinline__ float s_minc(float a, float c) { return a < c ? a : c; }
__device__ __forceinline__ float s_maxc(float a, float c) { return a > c ? a : c; }
__device__ __forceinline__ float s_clamp(float a, float lo, float hi) { return a < lo ? lo : (a > hi ? hi : a); }
__device__ __forceinline__ float s_val(float a) { return a; }

template <class S> struct T3 { S x, y, z; };
template <class S> __device__ __forceinline__ S dot3(const T3<S>& a, const T3<S>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class S> __device__ __forceinline__ S dot3(const T3<S>& a, V3 b) { return a.x * S(b.x) + a.y * S(b.y) + a.z * S(b.z); }
template <class S> __device__ __forceinline__ T3<S> cross3(const T3<S>& a, const T3<S>& b) {
  return T3<S>{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// LVController.compute / PSController.compute up to (clamped gross_thrust_des, torque_des); the thrust low-pass is the caller's.
// State inputs are detached in the reference (DiffActions.get_state_from_sim): only `cmd` carries tangents.
template <class S>
__device__ __forceinline__ void outer_loop(const GrReachConfig& c, V3 p, Q4 q, V3 v_w, V3 om_b, const S (&cmd)[4], S& thrust_des, S (&tau)[3]) {
  T3<S> vd{cmd[1], cmd[2], cmd[3]};
  if (c.controller == GR_CTRL_PS) {                                       // controller_diff.py:383-384
    vd.x = S(c.pos_gain[0]) * (cmd[1] - S(p.x)); vd.y = S(c.pos_gain[1]) * (cmd[2] - S(p.y)); vd.z = S(c.pos_gain[2]) * (cmd[3] - S(p.z));
  }
  const T3<S> e{vd.x - S(v_w.x), vd.y - S(v_w.y), vd.z - S(v_w.z)};
  const T3<S> ge{S(c.speed_gain[0]) * e.x, S(c.speed_gain[1]) * e.y, S(c.speed_gain[2]) * e.z};
  const S mag = s_minc(s_sqrt(dot3(ge, ge)), c.max_feedback_accel);      // :247-248
  const S en = s_maxc(s_sqrt(dot3(e, e)), 1e-12f);
  const S sc = mag / en;
  const S m = S(c.mass);
  const T3<S> F{m * (sc * e.x), m * (sc * e.y), m * (sc * e.z + S(c.gravity))};       // des_F = mass * (acc_fb - g), g = (0, 0, -9.81)
  // gross_thrust_des = quat_rotate_inverse(q, des_F).z (literal Isaac Lab form: v (2w^2 - 1) - 2w (u x v) + 2u (u . v))
  const S udF = S(q.x) * F.x + S(q.y) * F.y + S(q.z) * F.z;
  const S thrust = F.z * S(2.0f * q.w * q.w - 1.0f) - S(2.0f * q.w) * (S(q.x) * F.y - S(q.y) * F.x) + S(2.0f * q.z) * udF;
  // matrix_from_quat(q): columns r0 r1 r2
  const float two_s = 2.0f / (q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  const V3 r0 = v3(1.0f - two_s * (q.y * q.y + q.z * q.z), two_s * (q.x * q.y + q.z * q.w), two_s * (q.x * q.z - q.y * q.w));
  const V3 r1 = v3(two_s * (q.x * q.y - q.z * q.w), 1.0f - two_s * (q.x * q.x + q.z * q.z), two_s * (q.y * q.z + q.x * q.w));
  const V3 r2 = v3(two_s * (q.x * q.z + q.y * q.w), two_s * (q.y * q.z - q.x * q.w), 1.0f - two_s * (q.x * q.x + q.y * q.y));
  const T3<S> b1{s_cos(cmd[0]), s_sin(cmd[0]), S(0.0f)};
  const S Fn = s_maxc(s_sqrt(dot3(F, F)), 1e-12f);
  const T3<S> b3{F.x / Fn, F.y / Fn, F.z / Fn};
  T3<S> b2 = cross3(b3, b1);
  const S b2n = s_maxc(s_sqrt(dot3(b2, b2)), 1e-12f);
  b2 = T3<S>{b2.x / b2n, b2.y / b2n, b2.z / b2n};
  const T3<S> c0 = cross3(b2, b3);                                        // R_des = [b2 x b3, b2, b3]
  // m = 0.5 (R_des^T R - R^T R_des): m_ab = 0.5 (c_a . r_b - r_a . c_b); pose_err = (m12, -m02, m01)
  const S m12 = S(0.5f) * (dot3(b2, r2) - dot3(b3, r1));
  const S m02 = S(0.5f) * (dot3(c0, r2) - dot3(b3, r0));
  const S m01 = S(0.5f) * (dot3(c0, r1) - dot3(b2, r0));
  const float bb = c.body_rate_bound;
  const S rd0 = s_clamp(S(c.pose_gain[0]) * m12, -bb, bb), rd1 = s_clamp(S(c.pose_gain[1]) * (-m02), -bb, bb), rd2 = s_clamp(S(c.pose_gain[2]) * m01, -bb, bb);
  thrust_des = s_clamp(thrust, c.thrust_lo, c.thrust_hi);
  const V3 J = v3(c.inertia[0], c.inertia[1], c.inertia[2]);
  const V3 gyro = cross(om_b, J * om_b);
  tau[0] = S(J.x * c.rate_gain[0]) * (rd0 - S(om_b.x)) + S(gyro.x);
  tau[1] = S(J.y * c.rate_gain[1]) * (rd1 - S(om_b.y)) + S(gyro.y);
  tau[2] = S(J.z * c.rate_gain[2]) * (rd2 - S(om_b.z)) + S(gyro.z);
}

// F.cosine_similarity(a, b, dim=-1, eps=1e-8)
__device__ __forceinline__ float cos_sim(V3 a, V3 b) {
  const float na = fmaxf(norm(a), 1e-8f), nb = fmaxf(norm(b), 1e-8f);
  return dot(a / na, b / nb);
}

}  // namespace gr

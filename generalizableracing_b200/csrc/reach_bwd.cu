// reach_bwd.cu -- analytic reverse sweep of the reach-target step over a BPTT window (one thread per env, adjoints in
// registers over the horizon).  Same derivation as racing_bwd.cu for DroneDynamics.step/align; what differs:
//  * the controller is any of CTBR / LV / PS: the tape carries the full 4x4 Jacobian d(f', tau')/d a_lag that reach_step.cu
//    took in forward mode through the tanh action map and the controller (L/controllers/controller_diff.py:120-138, :242-291,
//    :378-430; QD/mdp/diff_action.py:160-176); the LV / PS torque is not low-passed, so no torque adjoint is carried for them;
//  * the loss terms (QD/mdp/losses.py:32-67) touch position, attitude, linear and angular velocity: the tape carries
//    d loss / d aligned (p, q, v, omega_w).
// Tape planes: see reach_step.cu.  Driven like gr_step_bwd by BPTT.update (standalone/diff_rl/algorithms/bptt.py:38-44).
#include "reach_core.cuh"

namespace gr {

struct RQAdj { float w; V3 u; };
__device__ __forceinline__ RQAdj r_rot_q_adj(Q4 q, V3 v, V3 yb) {          // adjoint of quat_rotate(q, v) w.r.t. q (literal Isaac Lab form)
  const V3 u = v3(q.x, q.y, q.z);
  const V3 uxv = cross(u, v);
  const float wb = dot(yb, 4.0f * q.w * v + 2.0f * uxv);
  const V3 ub = 2.0f * q.w * cross(v, yb) + 2.0f * dot(u, v) * yb + 2.0f * dot(u, yb) * v;
  return RQAdj{wb, ub};
}
__device__ __forceinline__ RQAdj r_rotinv_q_adj(Q4 q, V3 v, V3 yb) {       // adjoint of quat_rotate_inverse(q, v) w.r.t. q
  const V3 u = v3(q.x, q.y, q.z);
  const V3 uxv = cross(u, v);
  const float wb = dot(yb, 4.0f * q.w * v - 2.0f * uxv);
  const V3 ub = -2.0f * q.w * cross(v, yb) + 2.0f * dot(u, v) * yb + 2.0f * dot(u, yb) * v;
  return RQAdj{wb, ub};
}

constexpr int kReachBwdBlock = 64;

__global__ void __launch_bounds__(kReachBwdBlock) reach_step_bwd_kernel(const GrReachConfig cfg, const GrReachState st, const GrBwdIO io) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.num_envs) return;
  const float m = cfg.mass;
  const float dt = cfg.dt, decay = cfg.grad_decay;
  const float ef = expf(-dt / cfg.thrust_delay);
  V3 etau = v3(0.f, 0.f, 0.f);
  if (cfg.controller == GR_CTRL_CTBR) etau = v3(expf(-dt / cfg.torque_delay[0]), expf(-dt / cfg.torque_delay[1]), expf(-dt / cfg.torque_delay[2]));
  const V3 J = v3(cfg.inertia[0], cfg.inertia[1], cfg.inertia[2]);
  const V3 Jinv = v3(1.0f / J.x, 1.0f / J.y, 1.0f / J.z);

  float4* __restrict__ A = reinterpret_cast<float4*>(io.adjoint);
  const int64_t AS = io.adj_stride;
  const float4 a0 = A[0 * AS + i], a1 = A[1 * AS + i], a2 = A[2 * AS + i], a3 = A[3 * AS + i], a4 = A[4 * AS + i];
  V3 lP = xyz(a0); float lF = a0.w;
  Q4 lQ = quat(a1);
  V3 lV = xyz(a2), lVb = xyz(a3), lWb = xyz(a4);
  V3 lTau = v3(a2.w, a3.w, a4.w);

  const float4* __restrict__ T = reinterpret_cast<const float4*>(io.tape);
  const int64_t TS = io.tape_stride;
  const int N = st.num_envs;

  float4 c[GR_REACH_TAPE_PLANES];
  if (io.t_end > io.t_begin) {
#pragma unroll
    for (int k = 0; k < GR_REACH_TAPE_PLANES; ++k) c[k] = __ldcs(T + (int64_t)(io.t_end - 1) * GR_REACH_TAPE_PLANES * TS + rtidx(k, i));
  }
  for (int t = io.t_end - 1; t >= io.t_begin; --t) {
    float4 nx[GR_REACH_TAPE_PLANES];
    if (t - 1 >= io.t_begin) {                                         // one-step prefetch of the next (earlier) tape step
#pragma unroll
      for (int k = 0; k < GR_REACH_TAPE_PLANES; ++k) nx[k] = __ldcs(T + (int64_t)(t - 1) * GR_REACH_TAPE_PLANES * TS + rtidx(k, i));
    }
    const float g = io.grad_loss ? __ldg(io.grad_loss + (int64_t)t * N + i) : io.grad_scale;
    const Q4 q = quat(c[0]);
    const V3 om_b = xyz(c[1]); const bool cut = c[1].w != 0.0f;
    const V3 F_b = xyz(c[2]), D = xyz(c[3]), v1 = xyz(c[4]), omb1 = xyz(c[5]);
    const Q4 gQ = Q4{c[2].w, c[3].w, c[4].w, c[5].w};
    const V3 gP = xyz(c[6]), gV = xyz(c[7]), gW = xyz(c[8]);

    // loss gradient of this step w.r.t. the aligned state, captured before the reset detach
    lP = lP + g * gP;
    lQ = Q4{lQ.w + g * gQ.w, lQ.x + g * gQ.x, lQ.y + g * gQ.y, lQ.z + g * gQ.z};
    lV = lV + g * gV;
    const V3 lWw = g * gW;                                            // aligned omega_w: read only by the loss, never carried

    // recompute q' and omega_w' (droneDynamics.py:129-134)
    const Q4 dq = quat_mul(q, Q4{0.f, om_b.x, om_b.y, om_b.z});
    const Q4 qt = Q4{q.w + 0.5f * dq.w * dt, q.x + 0.5f * dq.x * dt, q.y + 0.5f * dq.y * dt, q.z + 0.5f * dq.z * dt};
    const float qn = sqrtf(qt.w * qt.w + qt.x * qt.x + qt.y * qt.y + qt.z * qt.z);
    const Q4 q1 = Q4{qt.w / qn, qt.x / qn, qt.y / qn, qt.z / qn};
    const V3 omw1 = quat_rotate(q1, omb1);

    // align (droneDynamics.py:174-179): d/d nominal = decay * d/d aligned
    const V3 p1b = decay * lP;
    float q1b_w = decay * lQ.w; V3 q1b_u = decay * v3(lQ.x, lQ.y, lQ.z);
    V3 v1b = decay * lV;
    {   // aligned v_b = rotinv(q', v')
      const V3 yb = decay * lVb;
      v1b = v1b + quat_rotate(q1, yb);
      const RQAdj a = r_rotinv_q_adj(q1, v1, yb);
      q1b_w += a.w; q1b_u = q1b_u + a.u;
    }
    V3 omw1b = decay * lWw;
    {   // aligned omega_b = rotinv(q', omega_w')
      const V3 yb = decay * lWb;
      omw1b = omw1b + quat_rotate(q1, yb);
      const RQAdj a = r_rotinv_q_adj(q1, omw1, yb);
      q1b_w += a.w; q1b_u = q1b_u + a.u;
    }
    // omega_w' = rot(q', omega_b')
    const V3 omb1b = quat_rotate_inverse(q1, omw1b);
    {
      const RQAdj a = r_rot_q_adj(q1, omb1, omw1b);
      q1b_w += a.w; q1b_u = q1b_u + a.u;
    }
    V3 ombb = omb1b;
    const V3 alphab = dt * omb1b;
    const V3 vb_w = v1b + dt * p1b;
    const V3 accb = dt * v1b + (0.5f * dt * dt) * p1b;
    const V3 pb = p1b;
    // q' = qt / |qt|
    const float qdot = q1.w * q1b_w + q1.x * q1b_u.x + q1.y * q1b_u.y + q1.z * q1b_u.z;
    const float tb_w = (q1b_w - q1.w * qdot) / qn;
    const V3 tb_u = v3((q1b_u.x - q1.x * qdot) / qn, (q1b_u.y - q1.y * qdot) / qn, (q1b_u.z - q1.z * qdot) / qn);
    // qt = q + 0.5 dt * qmul(q, (0, omega_b))
    float qb_w = tb_w; V3 qb_u = tb_u;
    {
      const float ow = 0.5f * dt * tb_w; const V3 o = (0.5f * dt) * tb_u;
      qb_w += o.x * om_b.x + o.y * om_b.y + o.z * om_b.z;
      qb_u.x += -ow * om_b.x - o.y * om_b.z + o.z * om_b.y;
      qb_u.y += -ow * om_b.y + o.x * om_b.z - o.z * om_b.x;
      qb_u.z += -ow * om_b.z - o.x * om_b.y + o.y * om_b.x;
      ombb.x += -ow * q.x + o.x * q.w + o.y * q.z - o.z * q.y;
      ombb.y += -ow * q.y - o.x * q.z + o.y * q.w + o.z * q.x;
      ombb.z += -ow * q.z + o.x * q.y - o.y * q.x + o.z * q.w;
    }
    // a = g + rot(q, F_b)/m
    const V3 yb = accb / m;
    const V3 Fb = quat_rotate_inverse(q, yb);
    {
      const RQAdj a = r_rot_q_adj(q, F_b, yb);
      qb_w += a.w; qb_u = qb_u + a.u;
    }
    // alpha = Jinv*tau' - Jinv*(omega_b x J omega_b)
    const V3 taub = lTau + Jinv * alphab;
    {
      const V3 z = -(Jinv * alphab);
      ombb = ombb + cross(J * om_b, z) + J * cross(z, om_b);
    }
    // F_b = f' e_z - k2 v_b |v_b| - k1 v_b
    const float fb = lF + Fb.z;
    const V3 vbb = D * Fb;
    // controller + action map through the taped Jacobian; 1-step lag -> a_{t-1}
    if (t >= 1) {
      reinterpret_cast<float4*>(io.grad_action)[(int64_t)(t - 1) * N + i] =
          make_float4(c[9].x * fb + c[10].x * taub.x + c[11].x * taub.y + c[12].x * taub.z, c[9].y * fb + c[10].y * taub.x + c[11].y * taub.y + c[12].y * taub.z,
                      c[9].z * fb + c[10].z * taub.x + c[11].z * taub.y + c[12].z * taub.z, c[9].w * fb + c[10].w * taub.x + c[11].w * taub.y + c[12].w * taub.z);
    }
    if (cut) {
      lP = v3(0.f, 0.f, 0.f); lQ = Q4{0.f, 0.f, 0.f, 0.f}; lV = lP; lVb = lP; lWb = lP; lF = 0.f; lTau = lP;
    } else {
      lP = pb; lQ = Q4{qb_w, qb_u.x, qb_u.y, qb_u.z}; lV = vb_w; lVb = vbb; lWb = ombb;
      lF = ef * fb; lTau = etau * taub;
    }
#pragma unroll
    for (int k = 0; k < GR_REACH_TAPE_PLANES; ++k) c[k] = nx[k];
  }
  A[0 * AS + i] = pack(lP, lF);
  A[1 * AS + i] = pack(lQ);
  A[2 * AS + i] = pack(lV, lTau.x);
  A[3 * AS + i] = pack(lVb, lTau.y);
  A[4 * AS + i] = pack(lWb, lTau.z);
}

}  // namespace gr

#ifndef GR_CPU_EMUL
using namespace gr;

extern "C" int gr_reach_step_bwd(const GrReachConfig* cfg, const GrReachState* st, const GrBwdIO* io, void* stream) {
  if (!cfg || !st || !io || !io->tape || !io->adjoint || !io->grad_action) return GR_ERR_NULL;
  if (st->num_envs <= 0 || io->tape_stride < ((st->num_envs + 31) & ~31) || io->adj_stride < st->num_envs) return GR_ERR_SIZE;
  if (io->t_begin < 0 || io->t_end < io->t_begin) return GR_ERR_SIZE;
  auto mis = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; };
  if (mis(io->tape) || mis(io->adjoint) || mis(io->grad_action)) return GR_ERR_ALIGN;
  if (cfg->controller < GR_CTRL_CTBR || cfg->controller > GR_CTRL_PS || cfg->thrust_delay <= 0.0f) return GR_ERR_CONFIG;
  const int grid = (st->num_envs + kReachBwdBlock - 1) / kReachBwdBlock;
  reach_step_bwd_kernel<<<grid, kReachBwdBlock, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*cfg, *st, *io);
  return (int)cudaGetLastError();
}
#endif  // GR_CPU_EMUL

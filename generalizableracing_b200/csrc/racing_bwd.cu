// racing_bwd.cu -- analytic reverse sweep of the racing step over a BPTT window (one thread per env,
// adjoints in registers across the whole horizon, tape streamed from HBM with a one-step prefetch).
//
// Replaces torch.autograd through DroneDynamics.step/align (QD/mdp/dynamics/droneDynamics.py:119-181),
// CTBRController.compute (L/controllers/controller_diff.py:120-138) and the tanh action map with 1-step lag
// (QD/mdp/diff_action.py:160-176), driven by BPTT.update (standalone/diff_rl/algorithms/bptt.py:38-44).
// Derivation: SURVEY.md Appendix A.6; the quaternion Jacobians differentiate the *literal* Isaac Lab
// formulas (gr_math.cuh), not rotation matrices.
//
// Tape per env-step, written by racing_step.cu (7 float4 planes):
//   0: q (pre-step)          1: omega_b | A0      2: F_b | A1      3: D | A2
//   4: v' | A3               5: omega_b' | cut    6: dloss/dP | -
//   A0 = dF'/da_lag[0] = clampmask*thr_err*scale0*(1-tanh^2)*(1-e_f);  A_i = dtau'_i/da_lag[i]
//   D  = dF_b/dv_b (diagonal) = -(2 k2 |v_b| + k1);  cut = env state at the start of the step was fresh from a reset
#include "gr_common.cuh"

namespace gr {

struct QAdj { float w; V3 u; };

// adjoint of y = quat_rotate(q, v) w.r.t. q
__device__ __forceinline__ QAdj rot_q_adj(Q4 q, V3 v, V3 yb) {
  const V3 u = v3(q.x, q.y, q.z);
  const V3 uxv = cross(u, v);
  const float wb = dot(yb, 4.0f * q.w * v + 2.0f * uxv);
  const V3 ub = 2.0f * q.w * cross(v, yb) + 2.0f * dot(u, v) * yb + 2.0f * dot(u, yb) * v;
  return QAdj{wb, ub};
}
// adjoint of y = quat_rotate_inverse(q, v) w.r.t. q
__device__ __forceinline__ QAdj rotinv_q_adj(Q4 q, V3 v, V3 yb) {
  const V3 u = v3(q.x, q.y, q.z);
  const V3 uxv = cross(u, v);
  const float wb = dot(yb, 4.0f * q.w * v - 2.0f * uxv);
  const V3 ub = -2.0f * q.w * cross(v, yb) + 2.0f * dot(u, v) * yb + 2.0f * dot(u, yb) * v;
  return QAdj{wb, ub};
}

constexpr int kBwdBlock = 64;
constexpr int kBwdStages = 6;

__device__ __forceinline__ void cp_async16(float4* smem_dst, const float4* gmem_src) {
#ifdef GR_CPU_EMUL
  *smem_dst = *gmem_src;
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#ifndef GR_CPU_EMUL
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int kN>
__device__ __forceinline__ void cp_async_wait() {
#ifndef GR_CPU_EMUL
  asm volatile("cp.async.wait_group %0;" ::"n"(kN) : "memory");
#endif
}

// ---------------------------------------------------------------------------------------------------------------------------
// One reverse step, written as two halves that only meet in three places, so that the two-lane kernel below can give each half
// its own lane (the one-lane kernel calls them back to back):
//   translational half (T): carries the adjoints of p, v, v_b and the thrust filter; produces the adjoint of the quaternion through
//                           the two places where a translational quantity is rotated (qa1 at q', qa2 at q) and d loss / d thrust (fb)
//   rotational half (R):    carries the adjoints of q, omega_b and the torque filter; consumes qa1, qa2
// Both recompute q' from the tape (a few dozen flops) instead of exchanging it.
// ---------------------------------------------------------------------------------------------------------------------------
struct BwdConst { float m, ef, dt, decay, wv; V3 etau, J, Jinv; };
struct Q1 { Q4 q1; float qn; };

__device__ __forceinline__ Q1 recompute_q1(Q4 q, V3 om_b, float dt) {          // droneDynamics.py:129-131
  const Q4 dq = quat_mul(q, Q4{0.f, om_b.x, om_b.y, om_b.z});
  const Q4 qt = Q4{q.w + 0.5f * dq.w * dt, q.x + 0.5f * dq.x * dt, q.y + 0.5f * dq.y * dt, q.z + 0.5f * dq.z * dt};
  const float qn = sqrtf(qt.w * qt.w + qt.x * qt.x + qt.y * qt.y + qt.z * qt.z);
  return Q1{Q4{qt.w / qn, qt.x / qn, qt.y / qn, qt.z / qn}, qn};
}

struct AdjT { V3 lP, lV, lVb; float lF; };
struct AdjR { Q4 lQ; V3 lWb, lTau; };

// T half.  In: tape (q, q', F_b, D, v', dloss/dP), loss weight g.  Out: qa1, qa2, fb; the carried adjoints are advanced to step t-1.
__device__ __forceinline__ void bwd_translational(const BwdConst& k, AdjT& a, Q4 q, Q4 q1, V3 F_b, V3 D, V3 v1, V3 dLdP, float g, bool cut,
                                                  QAdj& qa1, QAdj& qa2, float& fb) {
  // loss gradient of this step (QD/mdp/losses.py:72-80,95-101,111-117), captured before the reset detach
  a.lP = a.lP + g * dLdP;
  a.lV = a.lV + (g * k.wv) * v1;
  // align (droneDynamics.py:174-179): d/d nominal = decay * d/d aligned; aligned v_b = rotinv(q', v')
  const V3 p1b = k.decay * a.lP;
  V3 v1b = k.decay * a.lV;
  const V3 yb = k.decay * a.lVb;
  v1b = v1b + quat_rotate(q1, yb);
  qa1 = rotinv_q_adj(q1, v1, yb);
  // v' = v + a dt ; p' = p + v dt + 0.5 a dt^2
  const V3 vb_w = v1b + k.dt * p1b;
  const V3 accb = k.dt * v1b + (0.5f * k.dt * k.dt) * p1b;
  // a = g + rot(q, F_b)/m
  const V3 y2 = accb / k.m;
  const V3 Fb = quat_rotate_inverse(q, y2);
  qa2 = rot_q_adj(q, F_b, y2);
  // F_b = f' e_z - k2 v_b |v_b| - k1 v_b
  fb = a.lF + Fb.z;
  const V3 vbb = D * Fb;
  if (cut) { a.lP = v3(0.f, 0.f, 0.f); a.lV = a.lP; a.lVb = a.lP; a.lF = 0.f; }       // the state at the start of step t came from a reset
  else { a.lP = p1b; a.lV = vb_w; a.lVb = vbb; a.lF = k.ef * fb; }
}

// R half.  In: tape (q, omega_b, q', |q~|, omega_b'), qa1 / qa2 from the T half.  Out: d loss / d torque (taub); adjoints advanced.
__device__ __forceinline__ void bwd_rotational(const BwdConst& k, AdjR& a, Q4 q, V3 om_b, Q4 q1, float qn, V3 omb1, bool cut, const QAdj& qa1,
                                               const QAdj& qa2, V3& taub) {
  const V3 omw1 = quat_rotate(q1, omb1);                 // omega_w' (droneDynamics.py:134)
  float q1b_w = k.decay * a.lQ.w; V3 q1b_u = k.decay * v3(a.lQ.x, a.lQ.y, a.lQ.z);
  q1b_w += qa1.w; q1b_u = q1b_u + qa1.u;
  V3 omw1b;
  {   // aligned omega_b = rotinv(q', omega_w')
    const V3 yb = k.decay * a.lWb;
    omw1b = quat_rotate(q1, yb);
    const QAdj x = rotinv_q_adj(q1, omw1, yb);
    q1b_w += x.w; q1b_u = q1b_u + x.u;
  }
  // omega_w' = rot(q', omega_b')
  const V3 omb1b = quat_rotate_inverse(q1, omw1b);
  {
    const QAdj x = rot_q_adj(q1, omb1, omw1b);
    q1b_w += x.w; q1b_u = q1b_u + x.u;
  }
  // omega_b' = omega_b + alpha dt
  V3 ombb = omb1b;
  const V3 alphab = k.dt * omb1b;
  // q' = qt / |qt|
  const float qdot = q1.w * q1b_w + q1.x * q1b_u.x + q1.y * q1b_u.y + q1.z * q1b_u.z;
  const float tb_w = (q1b_w - q1.w * qdot) / qn;
  const V3 tb_u = v3((q1b_u.x - q1.x * qdot) / qn, (q1b_u.y - q1.y * qdot) / qn, (q1b_u.z - q1.z * qdot) / qn);
  // qt = q + 0.5 dt * qmul(q, (0, omega_b))
  float qb_w = tb_w; V3 qb_u = tb_u;
  {
    const float ow = 0.5f * k.dt * tb_w; const V3 o = (0.5f * k.dt) * tb_u;
    qb_w += o.x * om_b.x + o.y * om_b.y + o.z * om_b.z;
    qb_u.x += -ow * om_b.x - o.y * om_b.z + o.z * om_b.y;
    qb_u.y += -ow * om_b.y + o.x * om_b.z - o.z * om_b.x;
    qb_u.z += -ow * om_b.z - o.x * om_b.y + o.y * om_b.x;
    ombb.x += -ow * q.x + o.x * q.w + o.y * q.z - o.z * q.y;
    ombb.y += -ow * q.y - o.x * q.z + o.y * q.w + o.z * q.x;
    ombb.z += -ow * q.z + o.x * q.y - o.y * q.x + o.z * q.w;
  }
  qb_w += qa2.w; qb_u = qb_u + qa2.u;                     // a = g + rot(q, F_b)/m
  // alpha = Jinv*tau' - Jinv*(omega_b x J omega_b)
  taub = a.lTau + k.Jinv * alphab;
  {
    const V3 z = -(k.Jinv * alphab);
    ombb = ombb + cross(k.J * om_b, z) + k.J * cross(z, om_b);
  }
  if (cut) { a.lQ = Q4{0.f, 0.f, 0.f, 0.f}; a.lWb = v3(0.f, 0.f, 0.f); a.lTau = a.lWb; }
  else { a.lQ = Q4{qb_w, qb_u.x, qb_u.y, qb_u.z}; a.lWb = ombb; a.lTau = k.etau * taub; }
}

__device__ __forceinline__ BwdConst bwd_constants(const GrConfig& cfg, const float4* __restrict__ P, int i) {
  BwdConst k;
  k.m = __ldg(&P[pidx(PL_DRAG2, i)]).w;
  k.ef = __ldg(&P[pidx(PL_DRAG1, i)]).w;
  k.etau = xyz(__ldg(&P[pidx(PL_ETAU, i)]));
  k.J = v3(cfg.inertia[0], cfg.inertia[1], cfg.inertia[2]);
  k.Jinv = v3(1.0f / k.J.x, 1.0f / k.J.y, 1.0f / k.J.z);
  k.dt = cfg.dt; k.decay = cfg.grad_decay;
  k.wv = cfg.w_loss[1] * (2.0f / 3.0f);
  return k;
}

// ---- one lane per env.  Measured dead ends at C3 (16,384 envs x 32 steps, 22.7 us per sweep under graph replay): giving the two
// halves to two lanes of ONE warp serialises them (47 us); giving them to two partner WARPS (translational warp one step ahead, hand-over
// through shared memory and a 64-thread named barrier per step) takes 22.5 us -- the rotational half alone is as long a dependent chain
// as the whole step, the translational work was already filling its latency gaps (profiles/r2_bptt_sweep_variants.md).
__global__ void __launch_bounds__(kBwdBlock) racing_step_bwd_kernel(const GrConfig cfg, const GrState st, const GrBwdIO io) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= st.num_envs) return;
  const float4* __restrict__ P = reinterpret_cast<const float4*>(st.planes);
  const BwdConst k = bwd_constants(cfg, P, i);

  float4* __restrict__ A = reinterpret_cast<float4*>(io.adjoint);
  const int64_t AS = io.adj_stride;
  float4 a0 = A[0 * AS + i], a1 = A[1 * AS + i], a2 = A[2 * AS + i], a3 = A[3 * AS + i], a4 = A[4 * AS + i];
  AdjT aT{xyz(a0), xyz(a2), xyz(a3), a0.w};
  AdjR aR{quat(a1), xyz(a4), v3(a2.w, a3.w, a4.w)};

  const float4* __restrict__ T = reinterpret_cast<const float4*>(io.tape);
  const int64_t TS = io.tape_stride;
  const int N = st.num_envs;

  // Tape pipeline: a serial dependence over t, so the tape of the next kBwdStages steps is kept in flight with cp.async (LDGSTS)
  // into a per-thread shared-memory ring; every thread reads back only what it copied itself, so commit/wait groups are the only
  // synchronisation.
  GR_DYN_SMEM(float4, ring);
  float4* my = ring + threadIdx.x;                                   // slot(stage, plane) = my[(stage*7 + plane) * blockDim.x]
  const int n_steps = io.t_end - io.t_begin;
#pragma unroll 1
  for (int s = 0; s < kBwdStages; ++s) {                              // prologue: steps t_end-1 .. t_end-kBwdStages
    const int t = io.t_end - 1 - s;
    if (s < n_steps) {
#pragma unroll
      for (int p = 0; p < GR_TAPE_PLANES; ++p) cp_async16(my + (s * GR_TAPE_PLANES + p) * blockDim.x, T + (int64_t)t * GR_TAPE_PLANES * TS + tidx(p, i));
    }
    cp_async_commit();
  }
  int stage = 0;
  for (int t = io.t_end - 1; t >= io.t_begin; --t) {
    cp_async_wait<kBwdStages - 1>();                                  // the oldest group (step t) has landed
    float4 c[GR_TAPE_PLANES];
#pragma unroll
    for (int p = 0; p < GR_TAPE_PLANES; ++p) c[p] = my[(stage * GR_TAPE_PLANES + p) * blockDim.x];
    const float g = io.grad_loss ? __ldg(io.grad_loss + (int64_t)t * N + i) : io.grad_scale;
    {                                                                  // refill this stage with step t - kBwdStages
      const int tn = t - kBwdStages;
      if (tn >= io.t_begin) {
#pragma unroll
        for (int p = 0; p < GR_TAPE_PLANES; ++p) cp_async16(my + (stage * GR_TAPE_PLANES + p) * blockDim.x, T + (int64_t)tn * GR_TAPE_PLANES * TS + tidx(p, i));
      }
      cp_async_commit();
    }
    stage = stage + 1 == kBwdStages ? 0 : stage + 1;
    const Q4 q = quat(c[0]);
    const V3 om_b = xyz(c[1]); const float A0 = c[1].w;
    const V3 Again = v3(c[2].w, c[3].w, c[4].w);
    const bool cut = __float_as_uint(c[5].w) != 0u;
    const Q1 r = recompute_q1(q, om_b, k.dt);
    QAdj qa1, qa2; float fb; V3 taub;
    bwd_translational(k, aT, q, r.q1, xyz(c[2]), xyz(c[3]), xyz(c[4]), xyz(c[6]), g, cut, qa1, qa2, fb);
    bwd_rotational(k, aR, q, om_b, r.q1, r.qn, xyz(c[5]), cut, qa1, qa2, taub);
    // controller filters + action map (controller_diff.py:128-135; diff_action.py:174-176); 1-step lag -> a_{t-1}
    if (t >= 1) {
      reinterpret_cast<float4*>(io.grad_action)[(int64_t)(t - 1) * N + i] =
          make_float4(A0 * fb, Again.x * taub.x, Again.y * taub.y, Again.z * taub.z);
    }
  }
  A[0 * AS + i] = pack(aT.lP, aT.lF);
  A[1 * AS + i] = pack(aR.lQ);
  A[2 * AS + i] = pack(aT.lV, aR.lTau.x);
  A[3 * AS + i] = pack(aT.lVb, aR.lTau.y);
  A[4 * AS + i] = pack(aR.lWb, aR.lTau.z);
}

}  // namespace gr

#ifndef GR_CPU_EMUL
using namespace gr;

extern "C" int gr_step_bwd(const GrConfig* cfg, const GrState* st, const GrBwdIO* io, void* stream) {
  if (!cfg || !st || !io || !st->planes || !io->tape || !io->adjoint || !io->grad_action) return GR_ERR_NULL;
  if (st->num_envs <= 0 || st->plane_stride < ((st->num_envs + 31) & ~31) || io->tape_stride < ((st->num_envs + 31) & ~31) || io->adj_stride < st->num_envs) return GR_ERR_SIZE;
  if (io->t_begin < 0 || io->t_end < io->t_begin) return GR_ERR_SIZE;
  auto mis = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; };
  if (mis(st->planes) || mis(io->tape) || mis(io->adjoint) || mis(io->grad_action)) return GR_ERR_ALIGN;
  const size_t smem = (size_t)kBwdStages * GR_TAPE_PLANES * kBwdBlock * sizeof(float4);      // 43 KB ring per block (64 envs)
  const int grid = (st->num_envs + kBwdBlock - 1) / kBwdBlock;
  racing_step_bwd_kernel<<<grid, kBwdBlock, smem, reinterpret_cast<cudaStream_t>(stream)>>>(*cfg, *st, *io);
  return (int)cudaGetLastError();
}
#endif  // GR_CPU_EMUL

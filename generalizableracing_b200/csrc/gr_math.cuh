// Device math for the racing hot path: 3-vectors, (w,x,y,z) quaternions with the *literal*
// Isaac Lab formulas (their Jacobians w.r.t. a non-unit quaternion differ from rotation-matrix
// Jacobians -- SURVEY.md Appendix B), torch-style remainder / wrap_to_pi, Philox4x32-10.
#pragma once
#ifndef GR_CPU_EMUL
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace gr {

struct V3 { float x, y, z; };
struct Q4 { float w, x, y, z; };

__device__ __forceinline__ V3 v3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return V3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
__device__ __forceinline__ V3 operator-(V3 a) { return V3{-a.x, -a.y, -a.z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
  return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ float norm(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ V3 vabs(V3 a) { return V3{fabsf(a.x), fabsf(a.y), fabsf(a.z)}; }

__device__ __forceinline__ V3 xyz(float4 f) { return V3{f.x, f.y, f.z}; }
__device__ __forceinline__ Q4 quat(float4 f) { return Q4{f.x, f.y, f.z, f.w}; }
__device__ __forceinline__ float4 pack(V3 a, float w) { return make_float4(a.x, a.y, a.z, w); }
__device__ __forceinline__ float4 pack(Q4 q) { return make_float4(q.w, q.x, q.y, q.z); }

// quat_rotate(q, v) = v(2w^2-1) + 2w(u x v) + 2u(u.v)         (Isaac Lab math.quat_rotate)
__device__ __forceinline__ V3 quat_rotate(Q4 q, V3 v) {
  const V3 u = v3(q.x, q.y, q.z);
  const float s = 2.0f * q.w * q.w - 1.0f;
  const V3 c = cross(u, v);
  const float d = dot(u, v);
  return V3{v.x * s + c.x * q.w * 2.0f + u.x * d * 2.0f,
            v.y * s + c.y * q.w * 2.0f + u.y * d * 2.0f,
            v.z * s + c.z * q.w * 2.0f + u.z * d * 2.0f};
}
// quat_rotate_inverse(q, v) = v(2w^2-1) - 2w(u x v) + 2u(u.v)
__device__ __forceinline__ V3 quat_rotate_inverse(Q4 q, V3 v) {
  const V3 u = v3(q.x, q.y, q.z);
  const float s = 2.0f * q.w * q.w - 1.0f;
  const V3 c = cross(u, v);
  const float d = dot(u, v);
  return V3{v.x * s - c.x * q.w * 2.0f + u.x * d * 2.0f,
            v.y * s - c.y * q.w * 2.0f + u.y * d * 2.0f,
            v.z * s - c.z * q.w * 2.0f + u.z * d * 2.0f};
}
// Isaac Lab quat_mul: the 8-multiplication form, same association.
__device__ __forceinline__ Q4 quat_mul(Q4 a, Q4 b) {
  const float ww = (a.z + a.x) * (b.x + b.y);
  const float yy = (a.w - a.y) * (b.w + b.z);
  const float zz = (a.w + a.y) * (b.w - b.z);
  const float xx = ww + yy + zz;
  const float qq = 0.5f * (xx + (a.z - a.x) * (b.x - b.y));
  return Q4{qq - ww + (a.z - a.y) * (b.y - b.z),
            qq - xx + (a.x + a.w) * (b.x + b.w),
            qq - yy + (a.w - a.x) * (b.y + b.z),
            qq - zz + (a.z + a.y) * (b.w - b.x)};
}
__device__ __forceinline__ Q4 quat_from_euler_xyz(float roll, float pitch, float yaw) {
  float sy, cy, sr, cr, sp, cp;
  sincosf(yaw * 0.5f, &sy, &cy);
  sincosf(roll * 0.5f, &sr, &cr);
  sincosf(pitch * 0.5f, &sp, &cp);
  return Q4{cy * cr * cp + sy * sr * sp, cy * sr * cp - sy * cr * sp,
            cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp};
}
// same with MUFU sin/cos; |angle/2| stays below ~2 rad on the reset path (abs error < 5e-7)
__device__ __forceinline__ Q4 quat_from_euler_xyz_fast(float roll, float pitch, float yaw);
// third row of matrix_from_quat (QD/mdp/observation.py:31-32)
__device__ __forceinline__ V3 rotmat_row2(Q4 q) {
  const float two_s = 2.0f / (q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  return V3{two_s * (q.x * q.z - q.y * q.w), two_s * (q.y * q.z + q.x * q.w),
            1.0f - two_s * (q.x * q.x + q.y * q.y)};
}

#define GR_PI_F 3.14159265358979323846f
#define GR_2PI_F 6.28318530717958647692f
#define GR_HALF_PI_F 1.57079632679489661923f

// torch.remainder(a, b) for b > 0
__device__ __forceinline__ float remainder_pos(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.0f && m < 0.0f) m += b;
  return m;
}
__device__ __forceinline__ float wrap_to_pi(float a) {
  const float w = remainder_pos(a + GR_PI_F, GR_2PI_F);
  return (w == 0.0f && a > 0.0f) ? GR_PI_F : w - GR_PI_F;
}
// wrap_to_pi restricted to a in [-pi, pi] (the range of atan2f): bit-identical to the general form, no fmod
__device__ __forceinline__ float wrap_to_pi_atan2(float a) {
  float w = a + GR_PI_F;
  if (w >= GR_2PI_F) w -= GR_2PI_F;
  return (w == 0.0f && a > 0.0f) ? GR_PI_F : w - GR_PI_F;
}
// bad_pose (QD/mdp/termination.py:24-33): euler_xyz_from_quat -> % 2pi -> wrap_to_pi -> |.| > pi/2
__device__ __forceinline__ bool bad_pose(Q4 q) {
  const float sin_roll = 2.0f * (q.w * q.x + q.y * q.z);
  const float cos_roll = 1.0f - 2.0f * (q.x * q.x + q.y * q.y);
  const float roll = wrap_to_pi(remainder_pos(atan2f(sin_roll, cos_roll), GR_2PI_F));
  const float sin_pitch = 2.0f * (q.w * q.y - q.z * q.x);
  const float pitch0 = fabsf(sin_pitch) >= 1.0f ? copysignf(GR_HALF_PI_F, sin_pitch) : asinf(sin_pitch);
  const float pitch = wrap_to_pi(remainder_pos(pitch0, GR_2PI_F));
  return (fabsf(roll) > GR_HALF_PI_F) || (fabsf(pitch) > GR_HALF_PI_F);
}

// ---- fast math (hot path only; the rare reset path keeps libm accuracy) ---------------------------
// Absolute errors ~1e-7 on O(1) results, far inside the 1e-5 relative tolerance of the parity tests.
#ifdef GR_CPU_EMUL
__device__ __forceinline__ float fm_rcp(float x) { return 1.0f / x; }
__device__ __forceinline__ float fm_rsqrt(float x) { return 1.0f / sqrtf(x); }
__device__ __forceinline__ float fm_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ float fm_exp2(float x) { return exp2f(x); }
__device__ __forceinline__ float fm_log2(float x) { return log2f(x); }
__device__ __forceinline__ void fm_sincos(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ float sqrt_rn(float x) { return sqrtf(x); }
#else
__device__ __forceinline__ float fm_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fm_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fm_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fm_exp2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fm_log2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void fm_sincos(float x, float* s, float* c) { *s = __sinf(x); *c = __cosf(x); }
__device__ __forceinline__ float sqrt_rn(float x) { return __fsqrt_rn(x); }      // correctly rounded: gate predicate
#endif
// tanh(x) = 1 - 2/(1 + e^{2x}); saturates correctly for |x| large (e^{2x} -> inf / 0)
__device__ __forceinline__ float fm_tanh(float x) { return 1.0f - 2.0f * fm_rcp(1.0f + fm_exp2(x * 2.8853900817779268f)); }

// Rotation by a quaternion with the shared terms hoisted (same literal formula, re-associated):
//   rot(q,v) = s v + w2 (u x v) + u2 (u.v),  rotinv: minus on the cross term;  s = 2w^2-1, w2 = 2w, u2 = 2u
struct RotQ {
  V3 u, u2; float s, w2;
  __device__ __forceinline__ explicit RotQ(Q4 q) : u(v3(q.x, q.y, q.z)), u2(v3(2.0f * q.x, 2.0f * q.y, 2.0f * q.z)), s(2.0f * q.w * q.w - 1.0f), w2(2.0f * q.w) {}
  __device__ __forceinline__ V3 rot(V3 v) const {
    const V3 c = cross(u, v); const float d = dot(u, v);
    return V3{v.x * s + c.x * w2 + u2.x * d, v.y * s + c.y * w2 + u2.y * d, v.z * s + c.z * w2 + u2.z * d};
  }
  __device__ __forceinline__ V3 rotinv(V3 v) const {
    const V3 c = cross(u, v); const float d = dot(u, v);
    return V3{v.x * s - c.x * w2 + u2.x * d, v.y * s - c.y * w2 + u2.y * d, v.z * s - c.z * w2 + u2.z * d};
  }
};

__device__ __forceinline__ Q4 quat_from_euler_xyz_fast(float roll, float pitch, float yaw) {
  float sy, cy, sr, cr, sp, cp;
  fm_sincos(yaw * 0.5f, &sy, &cy);
  fm_sincos(roll * 0.5f, &sr, &cr);
  fm_sincos(pitch * 0.5f, &sp, &cp);
  return Q4{cy * cr * cp + sy * sr * sp, cy * sr * cp - sy * cr * sp, cy * cr * sp + sy * sr * cp, sy * cr * cp - cy * sr * sp};
}

// ---- Philox4x32-10 (Salmon et al. 2011) ------------------------------------------------
struct Philox {
  uint32_t k0, k1;      // key = seed
  uint32_t c0, c1;      // counter words 0,1 = global env id, step
  __device__ __forceinline__ Philox(uint64_t seed, uint32_t env, uint32_t step)
      : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c0(env), c1(step) {}
  // call index -> 4 x uint32
  __device__ __forceinline__ uint4 operator()(uint32_t call) const {
    uint32_t x0 = c0, x1 = c1, x2 = call, x3 = 0u, a = k0, b = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
      const uint32_t y0 = hi1 ^ x1 ^ a, y1 = lo1, y2 = hi0 ^ x3 ^ b, y3 = lo0;
      x0 = y0; x1 = y1; x2 = y2; x3 = y3;
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return make_uint4(x0, x1, x2, x3);
  }
};
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-8f; }   // [0,1), 24 bits
// Box-Muller on two raw words: (n0, n1)
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = (float)((a >> 8) + 1u) * 5.9604644775390625e-8f;    // (0,1]
  const float u2 = (float)(b >> 8) * 5.9604644775390625e-8f;
  // r = sqrt(-2 ln u1) = sqrt(-2 ln2 * log2 u1); angle in (-pi, pi] so the MUFU sin/cos stay in their accurate range
  const float r = fm_sqrt(-1.3862943611198906f * fm_log2(u1));
  float s, c;
  fm_sincos((u2 - 0.5f) * GR_2PI_F, &s, &c);
  return make_float2(r * c, r * s);
}

// Box-Muller pair from ONE 32-bit word: u1 = (low16 + 1) / 65536 in (0,1], u2 = high16 / 65536 in [0,1)
__device__ __forceinline__ float2 box_muller16(uint32_t x) {
  const float u1 = (float)((x & 0xFFFFu) + 1u) * 1.52587890625e-5f;
  const float u2 = (float)(x >> 16) * 1.52587890625e-5f;
  const float r = fm_sqrt(-1.3862943611198906f * fm_log2(u1));
  float s, c;
  fm_sincos((u2 - 0.5f) * GR_2PI_F, &s, &c);
  return make_float2(r * c, r * s);
}

}  // namespace gr

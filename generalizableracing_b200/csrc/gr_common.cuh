// Shared device-side definitions: plane ids, packed-int fields, state load/store, random source,
// gate-table staging.  Mirrors generalizableracing_b200/layout.py and include/gracing.h.
#pragma once
#include "gr_math.cuh"
#include "../../include/gracing.h"

#ifdef GR_CPU_EMUL
#define GR_DYN_SMEM(T, name) static T name[1 << 14]
#else
#define GR_DYN_SMEM(T, name) extern __shared__ T name[]
#endif

namespace gr {

enum Plane : int {
  PL_QUAT = 0, PL_POS = 1, PL_LINVEL = 2, PL_ANGVEL = 3, PL_TORQUE = 4, PL_ANGACC = 5, PL_FIFO = 6,
  PL_DRAG2 = 7, PL_DRAG1 = 8, PL_KP = 9, PL_KD = 10, PL_ETAU = 11, PL_NOISE0 = 12, PL_NOISE1 = 13,
  PL_EPSUM0 = 14, PL_EPSUM1 = 15
};

// packed ints in PL_ANGVEL.w : gate_id[0:8) | acc_gates[8:20) | level[20:26) | type[26:31) | fresh[31]
__device__ __forceinline__ uint32_t pk_gate(uint32_t p) { return p & 0xFFu; }
__device__ __forceinline__ uint32_t pk_acc(uint32_t p) { return (p >> 8) & 0xFFFu; }
__device__ __forceinline__ uint32_t pk_level(uint32_t p) { return (p >> 20) & 0x3Fu; }
__device__ __forceinline__ uint32_t pk_type(uint32_t p) { return (p >> 26) & 0x1Fu; }
__device__ __forceinline__ uint32_t pk_fresh(uint32_t p) { return p >> 31; }
__device__ __forceinline__ uint32_t pk_make(uint32_t gate, uint32_t acc, uint32_t level, uint32_t type, uint32_t fresh) {
  return (gate & 0xFFu) | ((acc > 0xFFFu ? 0xFFFu : acc) << 8) | ((level & 0x3Fu) << 20) | ((type & 0x1Fu) << 26) | (fresh << 31);
}

// 128-bit plane access.  Hot planes are read once and written once per step: streaming hints keep
// them from displacing the (tiny, reused) gate table in L1.
__device__ __forceinline__ float4 ld_plane(const float4* __restrict__ base, int64_t stride, int plane, int i) {
  return __ldcs(base + (int64_t)plane * stride + i);
}
__device__ __forceinline__ float4 ld_plane_ro(const float4* __restrict__ base, int64_t stride, int plane, int i) {
  return __ldg(base + (int64_t)plane * stride + i);
}
__device__ __forceinline__ void st_plane(float4* __restrict__ base, int64_t stride, int plane, int i, float4 v) {
  __stcs(base + (int64_t)plane * stride + i, v);
}

// Random source.  get4(call) returns slots [4*call, 4*call+4): calls 0,1 are standard normals,
// calls >= 2 uniforms in [0,1).  Dense mode reads the caller's tensor, Philox mode generates.
template <bool kPhilox>
struct RandSrc;

template <>
struct RandSrc<false> {
  const float4* row;
  __device__ __forceinline__ RandSrc(const GrRandom& r, int i, int /*env_id*/)
      : row(reinterpret_cast<const float4*>(r.rnd) + (int64_t)i * (GR_RND_STRIDE / 4)) {}
  __device__ __forceinline__ float4 get4(int call) const { return __ldg(row + call); }
};

template <>
struct RandSrc<true> {
  Philox ph;
  __device__ __forceinline__ RandSrc(const GrRandom& r, int /*i*/, int env_id) : ph(r.seed, (uint32_t)env_id, r.step) {}
  __device__ __forceinline__ float4 get4(int call) const {
    const uint4 x = ph((uint32_t)call);
    if (call < 2) {
      const float2 a = box_muller(x.x, x.y), b = box_muller(x.z, x.w);
      return make_float4(a.x, a.y, b.x, b.y);
    }
    return make_float4(u01(x.x), u01(x.y), u01(x.z), u01(x.w));
  }
};

// full-warp sum (all 32 lanes participate)
__device__ __forceinline__ float warp_sum(float v) {
#ifndef GR_CPU_EMUL
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}

// Gate table slice staged in shared memory: rows of types [type_lo, type_lo + ntypes).
struct TrackSmem {
  const float4* rows;   // shared
  int type_lo, levels, gates;
  __device__ __forceinline__ int base(int type, int level) const { return ((type - type_lo) * levels + level) * (gates + 1); }
  __device__ __forceinline__ float4 origin_row(int type, int level) const { return rows[base(type, level)]; }
  __device__ __forceinline__ V3 gate(int type, int level, int g) const { return xyz(rows[base(type, level) + 1 + g]); }
};

// All threads of the block call this.  chunk_types: int2 (lo, hi) per 64-env chunk.
__device__ __forceinline__ TrackSmem stage_track(const GrTrack& tr, const int2* __restrict__ chunk_types, int num_envs,
                                                 float4* smem) {
  const int first = blockIdx.x * blockDim.x;
  int last = first + blockDim.x - 1;
  if (last > num_envs - 1) last = num_envs - 1;
  const int tlo = __ldg(&chunk_types[first >> 6]).x;
  const int thi = __ldg(&chunk_types[last >> 6]).y;
  const int per_type = tr.levels * (tr.gates + 1);
  const int n = (thi - tlo + 1) * per_type;
  const float4* src = reinterpret_cast<const float4*>(tr.rows) + (int64_t)tlo * per_type;
  for (int k = threadIdx.x; k < n; k += blockDim.x) smem[k] = __ldg(src + k);
  __syncthreads();
  return TrackSmem{smem, tlo, tr.levels, tr.gates};
}

}  // namespace gr

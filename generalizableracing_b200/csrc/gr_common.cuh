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

// State layout: array of 32-env TILES, each tile = 16 planes x 32 lanes x float4 (8 KB, contiguous).  A warp owns one
// tile: its 16 plane loads are consecutive 512 B rows of the same 8 KB block, so DRAM serves them together and warps
// complete their loads one after the other (load -> compute -> store pipelines ACROSS warps).  With plane-major arrays
// the memory system streams plane by plane and every warp gets its last plane only at the end of the transfer.
// Written planes come first (0..7, plus 8 in differentiable mode), read-mostly planes after.
enum Plane : int {
  PL_QUAT = 0, PL_POS = 1, PL_LINVEL = 2, PL_ANGVEL = 3, PL_TORQUE = 4, PL_ANGACC = 5, PL_FIFO = 6,
  PL_EPSUM0 = 7,      // episode sums of reward terms 0..3 (terms 4, 5: PL_TORQUE.w / PL_ANGACC.w)            -- with episode sums on
  PL_LOSSSUM = 8,     // LossManager episode sums of the 3 loss terms | spare                              -- sums on + differentiable
  PL_DRAG2 = 9, PL_DRAG1 = 10, PL_KP = 11, PL_KD = 12, PL_ETAU = 13, PL_NOISE0 = 14, PL_NOISE1 = 15
};
constexpr int kTile = 32, kTilePlanes = 16;
__device__ __forceinline__ int64_t pidx(int plane, int i) { return ((int64_t)(i >> 5) * kTilePlanes + plane) * kTile + (i & 31); }
// BPTT tape of one step: tiles of 7 planes x 32 lanes x float4
__device__ __forceinline__ int64_t tidx(int plane, int i) { return ((int64_t)(i >> 5) * GR_TAPE_PLANES + plane) * kTile + (i & 31); }

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
// `tile` = address of lane (i & 31) in plane 0 of env i's tile: plane p is at the constant offset p * 32.
__device__ __forceinline__ float4* tile_ptr(float4* base, int i) { return base + (int64_t)(i >> 5) * (kTilePlanes * kTile) + (i & 31); }
__device__ __forceinline__ float4 ld_plane(const float4* __restrict__ tile, int plane) { return __ldcs(tile + plane * kTile); }
__device__ __forceinline__ float4 ld_plane_ro(const float4* __restrict__ tile, int plane) { return __ldg(tile + plane * kTile); }
__device__ __forceinline__ void st_plane(float4* __restrict__ tile, int plane, float4 v) { __stcs(tile + plane * kTile, v); }
// pull the line holding *p into L2 (no register, no L1 allocation)
__device__ __forceinline__ void prefetch_l2_line(const void* p) {
#ifndef GR_CPU_EMUL
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}

// Random source.  normals8() returns slots 0..7 (standard normals), get4(call >= 2) returns the uniform slots
// [4*call, 4*call+4).  Dense mode reads the caller's tensor; Philox mode generates: the eight normals come from ONE
// Philox call (each 32-bit word feeds a Box-Muller pair with two 16-bit uniforms: 4.7 sigma tails, ample for the
// 3 % / 0.05 rad observation noise), uniforms use 24 bits of each word.
template <bool kPhilox>
struct RandSrc;

template <>
struct RandSrc<false> {
  const float4* row;
  __device__ __forceinline__ RandSrc(const GrRandom& r, int i, int /*env_id*/)
      : row(reinterpret_cast<const float4*>(r.rnd) + (int64_t)i * (GR_RND_STRIDE / 4)) {}
  __device__ __forceinline__ float4 get4(int call) const { return __ldg(row + call); }
  __device__ __forceinline__ void normals8(float4& a, float4& b) const { a = __ldg(row); b = __ldg(row + 1); }
  __device__ __forceinline__ float normal6() const { return __ldg(row + 1).z; }        // slot 6 alone
  __device__ __forceinline__ uint32_t ph_env() const { return 0u; }
};

template <>
struct RandSrc<true> {
  Philox ph;
  __device__ __forceinline__ RandSrc(const GrRandom& r, int /*i*/, int env_id) : ph(r.seed, (uint32_t)env_id, r.step) {}
  __device__ __forceinline__ float4 get4(int call) const {
    const uint4 x = ph((uint32_t)call);
    return make_float4(u01(x.x), u01(x.y), u01(x.z), u01(x.w));
  }
  __device__ __forceinline__ void normals8(float4& a, float4& b) const {
    const uint4 x = ph(0u);
    const float2 p0 = box_muller16(x.x), p1 = box_muller16(x.y), p2 = box_muller16(x.z), p3 = box_muller16(x.w);
    a = make_float4(p0.x, p0.y, p1.x, p1.y);
    b = make_float4(p2.x, p2.y, p3.x, p3.y);
  }
  __device__ __forceinline__ float normal6() const { return box_muller16(ph(0u).w).x; }        // slot 6 alone: the same bits normals8 gives
  __device__ __forceinline__ uint32_t ph_env() const { return ph.c0; }
};

// Programmatic dependent launch (sm_90+): wait = block until the previous kernel in the stream has completed and its
// writes are visible; launch_dependents = allow the next PDL-launched kernel to start its prologue.  Both are no-ops
// when the kernel was not launched with the programmatic-stream-serialization attribute.
__device__ __forceinline__ void pdl_wait() {
#ifndef GR_CPU_EMUL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_launch_dependents() {
#ifndef GR_CPU_EMUL
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

// full-warp sum (all 32 lanes participate)
__device__ __forceinline__ float warp_sum(float v) {
#ifndef GR_CPU_EMUL
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}

// Draws of the rare paths (reset: calls 2..8, gate pass: calls 10..11).  A resetting env needs seven Philox calls (~110 instructions
// each); left to the env's own lane they run inside the divergent reset tail with one or two lanes active -- ~770 issue slots of the warp
// for every warp that holds a resetting env (77 % of the warps at a 4.5 % reset rate).  With `spec` set (GR_LAUNCH_COOP_RESET) the WARP
// generates them instead: for each resetting lane, lanes 0..6 compute one call each and park the result in shared memory
// (stage_reset_draws), the reset tail reads them back.  Same counters, same bits.  Dense mode and the standalone reset kernel read /
// generate directly.
__device__ __forceinline__ int spec_slot(int call) { return call < 10 ? call - 2 : call - 3; }     // 2..8 -> 0..6, 10..11 -> 7..8
constexpr int kSpecCalls = 9;

template <bool kPhilox>
struct Draws {
  const RandSrc<kPhilox>& rs;
  const float4* spec;       // shared: [kSpecCalls][blockDim.x] or nullptr
  bool have_normals = true; // false: the caller skipped normals8() (a window step that records no observation); slot 6 is then drawn on demand
  __device__ __forceinline__ float4 get4(int call) const {
    if (kPhilox && spec && call < 10) return spec[spec_slot(call) * blockDim.x + threadIdx.x];      // reset draws staged by stage_reset_draws()
    return rs.get4(call);
  }
  // slot 6 (thr_est_error normal, consumed by a reset only)
  __device__ __forceinline__ float thr_normal(float from_normals8) const { return have_normals ? from_normals8 : rs.normal6(); }
  // called by the step body with the warp converged once `reset` is known: stages the reset draws (if a staging area was given) and returns
  // the source the reset tail reads from
  __device__ __forceinline__ Draws staged(bool reset, bool with_noise) const;
};

// Warp-cooperative generation of the reset draws (Philox mode): call with all 32 lanes converged; `reset` = this lane's env resets in
// this step, `env_id` = its global id.  spec layout: [kSpecCalls][blockDim.x] float4, column = the thread the draws belong to.
__device__ __forceinline__ void stage_reset_draws(const RandSrc<true>& rs, float4* spec, bool reset, uint32_t env_id, bool with_noise) {
#ifndef GR_CPU_EMUL
  unsigned m = __ballot_sync(0xffffffffu, reset);
  if (m == 0u) return;
  const int lane = threadIdx.x & 31;
  const int calls = with_noise ? 7 : 5;                        // calls 2..6: pose, velocity, drag, level; 7..8: gate noise
  while (m) {
    const int r = __ffs(m) - 1;
    m &= m - 1;
    Philox ph = rs.ph;
    ph.c0 = __shfl_sync(0xffffffffu, env_id, r);
    if (lane < calls) {
      const uint4 x = ph((uint32_t)(2 + lane));
      spec[lane * blockDim.x + (threadIdx.x - lane + r)] = make_float4(u01(x.x), u01(x.y), u01(x.z), u01(x.w));
    }
  }
  __syncwarp();
#endif
}
__device__ __forceinline__ void stage_reset_draws(const RandSrc<false>&, float4*, bool, uint32_t, bool) {}
template <bool kPhilox>
__device__ __forceinline__ Draws<kPhilox> Draws<kPhilox>::staged(bool reset, bool with_noise) const {
  if (kPhilox && spec) stage_reset_draws(rs, const_cast<float4*>(spec), reset, rs.ph_env(), with_noise);
  return *this;
}

// The same hand-over for kernels whose shared memory is nearly full (the fused collection kernels: weights + activation tiles): each WARP
// owns a compact staging area of [7 calls][K columns] float4 (112 * K bytes), enough for the first K resetting lanes of a step; a lane
// beyond them (P(>= 3 resets in a warp) = 15 % at a 4.2 % reset rate, 1.2 % for >= 5) draws for itself as before.  Same counters, same bits.
struct CompactDraws {
  const RandSrc<true>& rs;
  float4* spec;             // shared: this warp's [7][K] columns, or nullptr (every lane draws for itself)
  int K;
  int col = -1;             // this lane's column after staged(), -1: not staged
  bool have_normals = true;
  __device__ __forceinline__ float4 get4(int call) const {
    if (col >= 0 && call < 10) return spec[spec_slot(call) * K + col];
    return rs.get4(call);
  }
  __device__ __forceinline__ float thr_normal(float from_normals8) const { return have_normals ? from_normals8 : rs.normal6(); }
  __device__ __forceinline__ CompactDraws staged(bool reset, bool with_noise) const {
    CompactDraws d = *this;
#ifndef GR_CPU_EMUL
    if (!spec) return d;
    unsigned m = __ballot_sync(0xffffffffu, reset);
    if (m == 0u) return d;
    const int lane = threadIdx.x & 31;
    const int rank = __popc(m & ((1u << lane) - 1u));               // this lane's place among the resetting lanes
    if (reset && rank < K) d.col = rank;
    const int calls = with_noise ? 7 : 5;
    __syncwarp();                                                   // (the columns may still be read by a straggler of the previous step)
    for (int s = 0; s < K && m; ++s) {
      const int r = __ffs(m) - 1;
      m &= m - 1;
      Philox ph = rs.ph;
      ph.c0 = __shfl_sync(0xffffffffu, rs.ph.c0, r);
      if (lane < calls) {
        const uint4 x = ph((uint32_t)(2 + lane));
        spec[lane * K + s] = make_float4(u01(x.x), u01(x.y), u01(x.z), u01(x.w));
      }
    }
    __syncwarp();
#endif
    return d;
  }
};

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

// mlp_tc.cuh -- a 16 -> H1 -> H2 -> (<= 4) leaky-relu MLP evaluated for 128 envs at a time on the sm_100a tensor cores,
// as the fused collection kernels (ppo_collect.cu, bptt_collect.cu) use it: one thread per env = one row of the UMMA tile
// (M = 128), operands in shared memory (fp16, K-major, un-swizzled [K/8 chunks][rows][8 halfs]), fp32 accumulators in
// tensor memory, thread-local epilogues (the accumulator row of env i is TMEM lane i = thread i).  See ppo_collect.cu for
// the design notes; tools/umma_probe.cu pins the descriptor conventions on hardware.
#pragma once
#include <cuda_fp16.h>
#include "../../include/gracing.h"
#include "umma.cuh"

namespace gr {

using namespace umma;

constexpr int kTileEnvs = 128;                    // M of one UMMA tile = threads per group
constexpr int kObsDim = 16, kK1 = 32, kOutPad = 16;
constexpr int kChunkA = kTileEnvs * 16;           // byte stride between K chunks of an A operand (128 rows x 16 B)

// packed parameters of one net (bytes); every block is 128-byte aligned
template <int H1, int H2>
struct NetLayout {
  static_assert(H1 % 16 == 0 && H2 % 16 == 0 && H1 <= 256 && H2 <= 256, "layer widths: multiples of 16 up to 256 (one MMA in N)");
  static constexpr int kH1 = H1, kH2 = H2;
  static constexpr int kW1Off = 0, kW1Bytes = H1 * kK1 * 2;                    // [4][H1][8] halfs: 16 inputs | b1 hi | b1 lo | 0...
  static constexpr int kW2Off = kW1Off + kW1Bytes, kW2Bytes = H2 * H1 * 2;     // [H1/8][H2][8]
  static constexpr int kW3Off = kW2Off + kW2Bytes, kW3Bytes = kOutPad * H2 * 2;   // [H2/8][16][8], rows >= out_dim are zero
  static constexpr int kB2Off = kW3Off + kW3Bytes, kB2Bytes = H2 * 2;          // fp16 [H2]
  static constexpr int kB3Off = kB2Off + kB2Bytes, kB3Bytes = 128;             // fp32 [16] (+ pad)
  static constexpr int kNetBytes = kB3Off + kB3Bytes;
  static constexpr int kCols = H1 > H2 ? H1 : H2;                              // TMEM columns of one group's accumulator
  static constexpr int kHBytes = kTileEnvs * kCols * 2;                        // activation tile of one group
  static_assert(kNetBytes % 128 == 0, "packed net must keep 128-byte alignment");
};

// ---------------------------------------------------------------------------------------------
// parameter packing: torch Linear weights [out][in] fp32 -> fp16, UMMA operand order (one net per blockIdx.y)
// ---------------------------------------------------------------------------------------------
template <class NL>
__global__ void policy_pack_kernel(const GrMlp net0, const GrMlp net1, uint8_t* __restrict__ packed) {
  constexpr int H1 = NL::kH1, H2 = NL::kH2;
  const GrMlp& m = blockIdx.y == 0 ? net0 : net1;
  uint8_t* out = packed + (size_t)blockIdx.y * NL::kNetBytes;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  __half* w1 = reinterpret_cast<__half*>(out + NL::kW1Off);
  __half* w2 = reinterpret_cast<__half*>(out + NL::kW2Off);
  __half* w3 = reinterpret_cast<__half*>(out + NL::kW3Off);
  if (idx < H1 * kK1) {
    const int n = idx / kK1, k = idx % kK1;
    float v = 0.0f;
    if (k < kObsDim) v = m.w1[n * kObsDim + k];
    else if (k == kObsDim) v = __half2float(__float2half_rn(m.b1[n]));
    else if (k == kObsDim + 1) v = m.b1[n] - __half2float(__float2half_rn(m.b1[n]));
    w1[(k >> 3) * (H1 * 8) + n * 8 + (k & 7)] = __float2half_rn(v);
  }
  if (idx < H2 * H1) {
    const int n = idx / H1, k = idx % H1;
    w2[(k >> 3) * (H2 * 8) + n * 8 + (k & 7)] = __float2half_rn(m.w2[n * H1 + k]);
  }
  if (idx < kOutPad * H2) {
    const int n = idx / H2, k = idx % H2;
    w3[(k >> 3) * (kOutPad * 8) + n * 8 + (k & 7)] = __float2half_rn(n < m.out_dim ? m.w3[n * H2 + k] : 0.0f);
  }
  if (idx < H2) reinterpret_cast<__half*>(out + NL::kB2Off)[idx] = __float2half_rn(m.b2[idx]);
  if (idx < NL::kB3Bytes / 4) reinterpret_cast<float*>(out + NL::kB3Off)[idx] = idx < m.out_dim ? m.b3[idx] : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 bits_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
__device__ __forceinline__ uint4 pack8(float4 a, float4 b) {
  return make_uint4(h2_bits(__floats2half2_rn(a.x, a.y)), h2_bits(__floats2half2_rn(a.z, a.w)), h2_bits(__floats2half2_rn(b.x, b.y)),
                    h2_bits(__floats2half2_rn(b.z, b.w)));
}

// first-layer operand row of one env: chunks 0,1 = the 16 observations, chunk 2 = (1, 1, 0...) against the bias rows, chunk 3 = 0
__device__ __forceinline__ void write_x_row(uint8_t* hrow, uint4 c0, uint4 c1) {
  *reinterpret_cast<uint4*>(hrow) = c0;
  *reinterpret_cast<uint4*>(hrow + kChunkA) = c1;
  *reinterpret_cast<uint4*>(hrow + 2 * kChunkA) = make_uint4(0x3C003C00u, 0u, 0u, 0u);      // half2(1, 1)
  *reinterpret_cast<uint4*>(hrow + 3 * kChunkA) = make_uint4(0u, 0u, 0u, 0u);
}

// 8 accumulator columns (fp32 bits) -> (+ bias) -> leaky relu -> 8 halfs
template <bool kBias>
__device__ __forceinline__ uint4 activate8(const uint32_t* r, const uint4* __restrict__ bias, int chunk, __half2 slope) {
  __half2 h[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) h[q] = __floats2half2_rn(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
  if (kBias) {
    const uint4 b = bias[chunk];
    h[0] = __hadd2(h[0], bits_h2(b.x)); h[1] = __hadd2(h[1], bits_h2(b.y)); h[2] = __hadd2(h[2], bits_h2(b.z)); h[3] = __hadd2(h[3], bits_h2(b.w));
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) h[q] = __hmax2(h[q], __hmul2(h[q], slope));
  return make_uint4(h2_bits(h[0]), h2_bits(h[1]), h2_bits(h[2]), h2_bits(h[3]));
}

// hidden-layer epilogue of one env row: D[row][0..NCOLS) (TMEM) -> (+ bias) -> leaky relu -> fp16 -> the row of the next A
// operand.  The TMEM loads are double-buffered (16 columns each): the next load is in flight while this one is processed.
template <bool kBias, int NCOLS>
__device__ __forceinline__ void hidden_epilogue(uint32_t taddr, uint8_t* hrow, const uint4* __restrict__ bias, __half2 slope) {
  constexpr int kBlocks = NCOLS / 16;
  uint32_t ra[16], rb[16];
  tmem_ld_x16(taddr, ra);
#pragma unroll 1
  for (int c = 0; c < kBlocks; c += 2) {          // c = index of the 16-column block held by ra
    tmem_ld_wait();
    tmem_ld_x16(taddr + (c + 1) * 16, rb);
    *reinterpret_cast<uint4*>(hrow + (2 * c) * kChunkA) = activate8<kBias>(ra, bias, 2 * c, slope);
    *reinterpret_cast<uint4*>(hrow + (2 * c + 1) * kChunkA) = activate8<kBias>(ra + 8, bias, 2 * c + 1, slope);
    tmem_ld_wait();
    if (c + 2 < kBlocks) tmem_ld_x16(taddr + (c + 2) * 16, ra);
    *reinterpret_cast<uint4*>(hrow + (2 * c + 2) * kChunkA) = activate8<kBias>(rb, bias, 2 * c + 2, slope);
    *reinterpret_cast<uint4*>(hrow + (2 * c + 3) * kChunkA) = activate8<kBias>(rb + 8, bias, 2 * c + 3, slope);
  }
}

// one WARP, converged (one elected lane issues, umma.cuh): D[tmem] = A[smem: KSTEPS x 16 K-columns] . B[smem]^T, then arrive on `bar` when
// done.  The descriptors of successive K steps differ only in the start-address field (bytes >> 4), so they are formed by integer adds.
// kWarp = false: ONE thread issues (the issuing lane of a diverged warp: ~16 instructions of R2UR / ELECT / BRA.U.ANY hand-over per UTCHMMA,
// but no extra live registers -- what ppo_collect_kernel<G = 4> at its 128-register cap needs: the converged form spills 152 B there and
// measured 14.5-14.9 instead of 13.4 us per step).
template <int KSTEPS, int LBO_B, bool kWarp>
__device__ __forceinline__ void issue_layer(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint64_t* bar) {
  tc_fence_after_sync();
#pragma unroll
  for (int kk = 0; kk < KSTEPS; ++kk) {
    if (kWarp) mma_f16_ss_warp(d_tmem, a_desc + (uint64_t)(kk * ((2 * kChunkA) >> 4)), b_desc + (uint64_t)(kk * ((2 * LBO_B) >> 4)), idesc, kk > 0);
    else mma_f16_ss(d_tmem, a_desc + (uint64_t)(kk * ((2 * kChunkA) >> 4)), b_desc + (uint64_t)(kk * ((2 * LBO_B) >> 4)), idesc, kk > 0);
  }
  if (kWarp) tc_commit_warp(bar);
  else tc_commit(bar);
}

// what one group needs to run a net
struct GroupCtx {
  uint8_t* hbuf;        // this group's activation tile (shared)
  uint8_t* hrow;        // hbuf + row * 16
  uint32_t hbuf_addr;   // shared-space address of hbuf
  uint32_t d_tmem;      // accumulator columns of this group
  uint32_t taddr;       // d_tmem + (lane quarter << 16): what this warp may tcgen05.ld
  uint64_t* bar;
  uint32_t phase;
  int bar_id;           // named barrier of the group (bar_id + 8: its release barrier)
  bool issuer;          // the lane that polls the mbarrier
  bool issuer_warp;     // its warp: issues the group's MMAs
  __half2 slope;
};

// A layer = [every thread of the group has written its operand row] -> group barrier -> one thread issues the MMAs ->
// ... independent work ... -> stage_wait -> the accumulator is readable.
enum Layer : int { kL1 = 0, kL2 = 1, kL3 = 2 };
template <class NL, bool kWarp = true>
__device__ __forceinline__ void stage_issue(const GroupCtx& g, uint32_t net_addr, const int layer) {
  constexpr int H1 = NL::kH1, H2 = NL::kH2;
  fence_proxy_async_smem();                 // this thread's st.shared operand rows -> async proxy
  tc_fence_before_sync();                   // this thread's tcgen05.ld of the columns about to be overwritten
  bar_sync(g.bar_id, kTileEnvs);
  if (kWarp ? g.issuer_warp : g.issuer) {   // kWarp: warp-uniform -- all 32 lanes run the issue code, one elected lane issues
    if (kWarp) __syncwarp();
    const uint64_t a_desc = make_smem_desc(g.hbuf_addr, kChunkA, 128);
    if (layer == kL1) issue_layer<kK1 / 16, H1 * 16, kWarp>(g.d_tmem, a_desc, make_smem_desc(net_addr + NL::kW1Off, H1 * 16, 128), make_idesc_f16(kTileEnvs, H1), g.bar);
    else if (layer == kL2) issue_layer<H1 / 16, H2 * 16, kWarp>(g.d_tmem, a_desc, make_smem_desc(net_addr + NL::kW2Off, H2 * 16, 128), make_idesc_f16(kTileEnvs, H2), g.bar);
    else issue_layer<H2 / 16, kOutPad * 16, kWarp>(g.d_tmem, a_desc, make_smem_desc(net_addr + NL::kW3Off, kOutPad * 16, 128), make_idesc_f16(kTileEnvs, kOutPad), g.bar);
  }
}
// Only the issuing thread polls the mbarrier; everybody else blocks in hardware on the group's second named barrier.
__device__ __forceinline__ void stage_wait(GroupCtx& g) {
  if (g.issuer) {
    mbar_wait(g.bar, g.phase);
    g.phase ^= 1u;
  }
  __syncwarp();
  bar_sync(g.bar_id + 8, kTileEnvs);
  tc_fence_after_sync();
}
// first 4 outputs of layer 3 (+ fp32 bias)
template <class NL>
__device__ __forceinline__ float4 read_head(const GroupCtx& g, const uint8_t* net_smem) {
  uint32_t r[4];
  tmem_ld_x4(g.taddr, r);
  tmem_ld_wait();
  const float4 b3 = *reinterpret_cast<const float4*>(net_smem + NL::kB3Off);
  return make_float4(__uint_as_float(r[0]) + b3.x, __uint_as_float(r[1]) + b3.y, __uint_as_float(r[2]) + b3.z, __uint_as_float(r[3]) + b3.w);
}
// the two hidden epilogues
template <class NL>
__device__ __forceinline__ void epilogue1(const GroupCtx& g) { hidden_epilogue<false, NL::kH1>(g.taddr, g.hrow, nullptr, g.slope); }
template <class NL>
__device__ __forceinline__ void epilogue2(const GroupCtx& g, const uint8_t* net_smem) {
  hidden_epilogue<true, NL::kH2>(g.taddr, g.hrow, reinterpret_cast<const uint4*>(net_smem + NL::kB2Off), g.slope);
}
// a whole net with nothing overlapped
template <class NL, bool kWarp = true>
__device__ __forceinline__ float4 run_net(GroupCtx& g, const uint8_t* net_smem, uint32_t net_addr) {
  stage_issue<NL, kWarp>(g, net_addr, kL1); stage_wait(g);
  epilogue1<NL>(g);
  stage_issue<NL, kWarp>(g, net_addr, kL2); stage_wait(g);
  epilogue2<NL>(g, net_smem);
  stage_issue<NL, kWarp>(g, net_addr, kL3); stage_wait(g);
  return read_head<NL>(g, net_smem);
}

// per-thread view of its group: `cols_per_group` TMEM columns and `h_bytes` of activation tile per group
__device__ __forceinline__ GroupCtx make_group_ctx(uint8_t* h_smem, int h_bytes, uint64_t* bars, uint32_t tmem_base, int cols_per_group, int grp, int row,
                                                   float negative_slope) {
  GroupCtx g;
  g.hbuf = h_smem + grp * h_bytes;
  g.hrow = g.hbuf + row * 16;
  g.hbuf_addr = smem_u32(g.hbuf);
  g.d_tmem = tmem_base + (uint32_t)(grp * cols_per_group);
  g.taddr = g.d_tmem + ((uint32_t)((row >> 5) * 32) << 16);
  g.bar = &bars[grp];
  g.phase = 0u;
  g.bar_id = 1 + grp;
  g.issuer = row == 32 * (grp & 3);            // lane 0 of a different warp per group: the issuers sit on different SM sub-partitions
  g.issuer_warp = (row >> 5) == (grp & 3);
  g.slope = __float2half2_rn(negative_slope);
  return g;
}

}  // namespace gr

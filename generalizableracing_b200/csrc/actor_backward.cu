// actor_backward.cu -- d(loss)/d(parameters) of the BPTT actor MLP (16 -> H1 -> H2 -> 4, leaky relu) over R = T*N rows, on the
// sm_100a tensor cores, in one persistent kernel.
//
// Reference path replaced: loss.backward() through BaseModel.actor (S/diff_rl/algorithms/bptt.py:38-44 over the T graphs
// built by model.py:63-99), as FusedBpttCollector batches it: rows = the (observation, cotangent) pairs of a whole window,
// cotangent = gr_step_bwd's d(loss)/d(action) (= d/d(mean): action = mean + sigma * eps).  Measured: the batched fp32 torch
// version of this (cuBLAS SGEMM over 524,288 rows) is 4.9 ms of a 5.5 ms fused BPTT iteration.
//
// Per 128-row tile (one thread per row, as in mlp_tc.cuh): recompute H1, H2 (forward layers 1-2), then
//     dW3^T += H2^T . dA          dH2 = dA . W3          dH2' = dH2 * lrelu'(H2)
//     dW2   += dH2'^T . H1        db2 += dH2'^T . 1      dH1 = dH2' . W2      dH1' = dH1 * lrelu'(H1)
//     dW1|b1 += dH1'^T . [X | 1]
// Every GEMM is a tcgen05.mma on operands that ALREADY sit in shared memory in the per-row layout the forward uses
// ([chunk of 8 columns][row][8 halfs]): read K-major when the contraction runs over columns (dgrad), MN-major when it runs
// over the rows (weight gradients) -- same bytes, two descriptors (tools/umma_probe.cu checks both on hardware); the packed
// forward weights serve the dgrad GEMMs as MN-major B operands, so no transposed copy exists anywhere.  The weight-gradient
// accumulators stay in tensor memory across all tiles of the CTA (fp32) and are flushed once with atomics.
// fp16 operands need a loss scale: the cotangent is multiplied by *scale (device scalar, caller picks ~1024 / max|G|) on
// load and the accumulators are divided by it at the flush.
#include <cstdlib>

#include "mlp_tc.cuh"
#include "ppo_loss.cuh"

namespace gr {

// leaky-relu derivative applied to 8 accumulator columns, gated by the sign of the stored forward activation (same chunk,
// same thread), written back in place as the next operand
__device__ __forceinline__ uint4 dact8(const uint32_t* r, uint4 fwd, __half2 slope) {
  const __half2 zero = __float2half2_rn(0.0f), one_minus = __hsub2(__float2half2_rn(1.0f), slope);
  const uint32_t f[4] = {fwd.x, fwd.y, fwd.z, fwd.w};
  uint32_t o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const __half2 g = __floats2half2_rn(__uint_as_float(r[2 * q]), __uint_as_float(r[2 * q + 1]));
    const __half2 gate = __hfma2(__hgt2(bits_h2(f[q]), zero), one_minus, slope);        // 1 where the unit was active, slope elsewhere
    o[q] = h2_bits(__hmul2(g, gate));
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}
// D[row][0..128) * lrelu'(stored activation) -> fp16, in place over the activation chunks [chunk0, chunk0 + 16) of the row
__device__ __forceinline__ void dact_epilogue(uint32_t taddr, uint8_t* row_base, int chunk0, __half2 slope) {
  uint32_t ra[16], rb[16];
  tmem_ld_x16(taddr, ra);
#pragma unroll 1
  for (int c = 0; c < 8; c += 2) {
    tmem_ld_wait();
    tmem_ld_x16(taddr + (c + 1) * 16, rb);
    uint4* p0 = reinterpret_cast<uint4*>(row_base + (chunk0 + 2 * c) * kChunkA);
    uint4* p1 = reinterpret_cast<uint4*>(row_base + (chunk0 + 2 * c + 1) * kChunkA);
    *p0 = dact8(ra, *p0, slope);
    *p1 = dact8(ra + 8, *p1, slope);
    tmem_ld_wait();
    if (c + 2 < 8) tmem_ld_x16(taddr + (c + 2) * 16, ra);
    uint4* p2 = reinterpret_cast<uint4*>(row_base + (chunk0 + 2 * c + 2) * kChunkA);
    uint4* p3 = reinterpret_cast<uint4*>(row_base + (chunk0 + 2 * c + 3) * kChunkA);
    *p2 = dact8(rb, *p2, slope);
    *p3 = dact8(rb + 8, *p3, slope);
  }
}

// fire-and-forget vector reduction: 4 consecutive floats in one L2 operation (sm_90+), 16-byte aligned
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr uint32_t kMnA = 1u << 15, kMnB = 1u << 16;          // instruction-descriptor bits: A / B operand is MN-major

// descriptors over the per-row layout [chunk][128 rows][8 halfs]
__device__ __forceinline__ uint64_t desc_rows_k(uint32_t addr) { return make_smem_desc(addr, kChunkA, 128); }     // rows = M, columns = K
__device__ __forceinline__ uint64_t desc_rows_mn(uint32_t addr) { return make_smem_desc(addr, 128, kChunkA); }    // columns = M|N, rows = K

// blockIdx.y selects the job: the actor and the critic of a PPO step share one launch (same widths, same row count).
//
// Warp-specialised: kGroups producer groups of 128 threads (one thread per row of a 128-row tile: operand rows in, epilogues out) and
// ONE issuing warp whose lane 0 issues every tcgen05.mma of the CTA in program order.  A stage of a group = [its 128 threads have
// written the operands: `full[g]`, 128 arrivals] -> the issuer queues the stage's MMAs and commits them to `done[g]` -> the group reads
// the accumulator and writes the next operands.  With two groups the issuer alternates between them, so the tensor pipe works on one
// tile while the other group's threads run their epilogue (one group alone leaves the pipe idle during every epilogue: 23 % tensor-
// pipe activity, 14 % issue-slot use).  The weight-gradient accumulators are shared by the groups -- a single thread issues all MMAs,
// so the accumulations are ordered -- and each group owns 128 scratch columns.  TMEM: 16 -> 128 -> 128: 2 x 128 scratch + 128 (dW2) +
// 32 (dW1|db1) + 16 (dW3) + 16 (db2) = 448 columns, two groups; 16 -> 256 -> 128: 128 + 256 + 64 + 32 = 480 columns, one group.
//
// kPpo (gr_ppo_fused_step): the same kernel with the PPO mini-batch step's forward head and loss inside.  One more stage per tile -- layer 3
// (N = 16) on the recomputed H2 -- gives the row's mean (job 0: actor) or value (job 1: critic); the row's thread evaluates the loss of
// ppo_loss.cuh on it and on the stored columns it fetched at the start of the tile, and writes the cotangent row itself.  The separate
// gr_policy_forward / gr_ppo_loss_grad launches and their mu / value / gradient round trips through HBM disappear.  The cotangent goes to
// fp16 with a STATIC scale (fz.cot_scale x the un-normalised per-row gradient, 1 / rows applied to the fp32 accumulators at the flush):
// the batch maximum the stand-alone path scales by would need a grid-wide reduction between the loss and the weight gradients.
struct PpoFusedArgs { GrPpoBatch b; float* sums; float cot_scale; };

template <class NL, int kGroups, bool kPpo>
__global__ void __launch_bounds__(kGroups * kTileEnvs + 32, 1) actor_backward_kernel(const GrBackwardJob job0, const GrBackwardJob job1, const int64_t R,
                                                                                    const PpoFusedArgs fz) {
  const GrBackwardJob& job = blockIdx.y == 0 ? job0 : job1;
  const GrPolicy pol = job.policy;
  const float* __restrict__ X = job.obs;
  const float* __restrict__ G = job.grad_actions;
  const float* __restrict__ scale_ptr = job.scale;
  const int64_t* __restrict__ idx = job.indices;
  const int64_t x_stride = job.obs_stride > 0 ? job.obs_stride : kObsDim;      // floats per observation row (transition records: 48)
  const GrMlpGrad out = job.out;
  constexpr int H1 = NL::kH1, H2 = NL::kH2, kHalves = H1 / 128;
  static_assert(H2 == 128 && (H1 == 128 || H1 == 256), "built for 16 -> 128|256 -> 128 -> 4");
  // tensor memory: one scratch accumulator per group | dW2 [H2 x H1] | dW1 (+ db1) [H1 x 32] as `kHalves` blocks | dW3^T [H2 x 16] | db2 [H2 x 16]
  constexpr uint32_t kColD = 0, kColW2 = 128 * kGroups, kColW1 = kColW2 + H1, kColW3 = kColW1 + 32 * kHalves, kColB2 = kColW3 + 16, kColsUsed = kColB2 + 16;
  static_assert(kColsUsed <= 512, "tensor memory budget");
  constexpr int kStages = 2 * kHalves + 3 + (kPpo ? 1 : 0); // forward L1 (per half), forward L2, [forward L3], dW3 | dH2, dW2 | db2 | dH1 (per half), dW1
  constexpr int kGroupBytes = (4 + H1 / 8 + H2 / 8 + 2) * kChunkA;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* w_smem = smem;                                   // forward-packed net
  uint8_t* grp_smem = w_smem + NL::kNetBytes;               // per group: xs [4][128][8] | h1s [H1/8][128][8] | h2s [H2/8][128][8] | das [2][128][8]
  uint8_t* ones = grp_smem + kGroups * kGroupBytes;         // [2][128][8]   column 0 = 1 (shared, read-only)
  uint64_t* full = reinterpret_cast<uint64_t*>(ones + 2 * kChunkA);
  uint64_t* done = full + kGroups;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + kGroups);

  const int tid = threadIdx.x;
  const bool producer = tid < kGroups * kTileEnvs;
  const int grp = producer ? tid / kTileEnvs : 0, row = tid % kTileEnvs;
  {
    const uint4* src = reinterpret_cast<const uint4*>(pol.packed);
    uint4* dst = reinterpret_cast<uint4*>(w_smem);
    for (int k = tid; k < NL::kNetBytes / 16; k += kGroups * kTileEnvs + 32) dst[k] = __ldg(src + k);
  }
  uint8_t* xs = grp_smem + grp * kGroupBytes;
  uint8_t* h1s = xs + 4 * kChunkA;
  uint8_t* h2s = h1s + (H1 / 8) * kChunkA;
  uint8_t* das = h2s + (H2 / 8) * kChunkA;
  if (tid < kTileEnvs) {
    *reinterpret_cast<uint4*>(ones + row * 16) = make_uint4(0x00003C00u, 0u, 0u, 0u);            // half(1) in column 0
    *reinterpret_cast<uint4*>(ones + kChunkA + row * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (producer) *reinterpret_cast<uint4*>(das + kChunkA + row * 16) = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    for (int g = 0; g < kGroups; ++g) { mbar_init(&full[g], kTileEnvs); mbar_init(&done[g], 1); }
  }
  __syncwarp();
  if (tid < 32) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();

  const uint32_t tm = *tmem_slot;
  const int64_t tiles = (R + kTileEnvs - 1) / kTileEnvs;
  const int64_t stride = (int64_t)gridDim.x * kGroups;
  const float inv_rows = 1.0f / (float)R;
  const float scale = kPpo ? fz.cot_scale : (out.scale_is_maxabs ? 1024.0f / fmaxf(__ldg(scale_ptr), 1e-30f) : __ldg(scale_ptr));
  const bool critic_job = blockIdx.y != 0;

  if (producer) {
    const uint32_t lane_sel = (uint32_t)((row >> 5) * 32) << 16;
    const uint32_t d_addr = tm + kColD + 128u * grp + lane_sel;
    const __half2 slope = __float2half2_rn(pol.negative_slope);
    float4 gsum = make_float4(0.f, 0.f, 0.f, 0.f);          // db3 = sum of the (unscaled) cotangent rows
    float lsum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // kPpo: this thread's share of gr_ppo_loss_grad's sums
    uint32_t ph = 0u;
    // one stage hand-over: operands written -> arrive; wait for the stage's MMAs
#define GR_HANDOVER() do { fence_proxy_async_smem(); tc_fence_before_sync(); mbar_arrive(&full[grp]); mbar_wait(&done[grp], ph); ph ^= 1u; tc_fence_after_sync(); } while (0)
    // The rows of a tile: observation row, and either the cotangent row (plain backward) or the stored columns of the row (kPpo).  With the
    // mini-batch gather on load every row is a scattered 64-byte read (plus five scattered scalars / 16-byte rows under kPpo) behind a
    // scattered index read, and one tile at a time per group leaves nothing else to hide two DRAM latencies behind (ncu: long-scoreboard
    // 11 of 16 stall cycles per issue).  So the loads run ahead of the tile they belong to: the INDEX of tile t + 2 and the ROWS of tile
    // t + 1 are requested while tile t is processed -- when the rows are requested their address has been known for a whole tile.  The
    // requests are unconditional (rows past the end are clamped to the last row and zeroed at their use): a divergent or skipped request
    // would make the compiler wait for the outstanding loads where the paths meet.
    struct TileRows {
      float4 o0, o1, o2, o3, gr4, st_a, st_omu, st_osg;
      float st_adv, st_logp, st_ret, st_ov;
    };
    const bool gather = idx != nullptr;
    auto index_of = [&](const int64_t tile) -> int64_t {
      int64_t r = tile * kTileEnvs + row;
      r = r < R ? r : R - 1;
      return gather ? __ldg(idx + r) : r;
    };
    auto fetch = [&](const int64_t tile, const int64_t q, TileRows& t) {
      int64_t r = tile * kTileEnvs + row;
      r = r < R ? r : R - 1;
      const float4* xr = reinterpret_cast<const float4*>(X + q * x_stride);
      t.o0 = __ldcs(xr); t.o1 = __ldcs(xr + 1); t.o2 = __ldcs(xr + 2); t.o3 = __ldcs(xr + 3);
      if (!kPpo) {
        t.gr4 = __ldcs(reinterpret_cast<const float4*>(G) + r);
      } else if (fz.b.records) {                   // transition records: the stored columns sit behind the observations of the same record
        const float4* p = reinterpret_cast<const float4*>(fz.b.records) + q * (GR_RECORD_FLOATS / 4);
        const float4 sc = __ldcs(p + 11);
        if (!critic_job) { t.st_a = __ldcs(p + 8); t.st_omu = __ldcs(p + 9); t.st_osg = __ldcs(p + 10); t.st_logp = sc.x; t.st_adv = sc.y; }
        else { t.st_ret = sc.z; t.st_ov = sc.w; }
      } else if (!critic_job) {
        t.st_a = __ldg(reinterpret_cast<const float4*>(fz.b.actions) + q);
        t.st_omu = __ldg(reinterpret_cast<const float4*>(fz.b.old_mu) + q);
        t.st_osg = __ldg(reinterpret_cast<const float4*>(fz.b.old_sigma) + q);
        t.st_adv = __ldg(fz.b.advantages + q);
        t.st_logp = __ldg(fz.b.old_log_prob + q);
      } else {
        t.st_ret = __ldg(fz.b.returns + q);
        t.st_ov = fz.b.use_clipped_value_loss ? __ldg(fz.b.old_values + q) : 0.0f;
      }
    };
    TileRows cur = {}, nxt = {};
    const int64_t first_tile = (int64_t)blockIdx.x * kGroups + grp;
    int64_t q_next = 0, q_next2 = 0;
    if (first_tile < tiles) {
      fetch(first_tile, index_of(first_tile), cur);
      q_next = index_of(first_tile + stride);
    }
#pragma unroll 1
    for (int64_t tile = first_tile; tile < tiles; tile += stride) {
      const int64_t r = tile * kTileEnvs + row;
      q_next2 = index_of(tile + 2 * stride);
      fetch(tile + stride, q_next, nxt);
      const bool live = r < R;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 o0 = live ? cur.o0 : z4, o1 = live ? cur.o1 : z4, o2 = live ? cur.o2 : z4, o3 = live ? cur.o3 : z4, gr4 = live ? cur.gr4 : z4;
      const float4 st_a = cur.st_a, st_omu = cur.st_omu, st_osg = cur.st_osg;
      const float st_adv = cur.st_adv, st_logp = cur.st_logp, st_ret = cur.st_ret, st_ov = cur.st_ov;
      // (the previous tile's last stage -- dW1 -- read xs and h1s: its hand-over waited for it)
      write_x_row(xs + row * 16, pack8(o0, o1), pack8(o2, o3));
      if (!kPpo) {
        gsum.x += gr4.x; gsum.y += gr4.y; gsum.z += gr4.z; gsum.w += gr4.w;
        *reinterpret_cast<uint4*>(das + row * 16) =
            make_uint4(h2_bits(__floats2half2_rn(gr4.x * scale, gr4.y * scale)), h2_bits(__floats2half2_rn(gr4.z * scale, gr4.w * scale)), 0u, 0u);
      }
      // ---- forward layer 1 (128 units at a time) and layer 2: recompute the activations
#pragma unroll 1
      for (int h = 0; h < kHalves; ++h) {
        GR_HANDOVER();
        hidden_epilogue<false, 128>(d_addr, h1s + row * 16 + h * 16 * kChunkA, nullptr, slope);
      }
      GR_HANDOVER();
      hidden_epilogue<true, H2>(d_addr, h2s + row * 16, reinterpret_cast<const uint4*>(w_smem + NL::kB2Off), slope);
      if (kPpo) {
        // ---- forward layer 3 -> this row's mean / value -> its loss gradient = the cotangent row (ppo.py:118-171, ppo_loss.cuh)
        GR_HANDOVER();
        uint32_t hr[4];
        tmem_ld_x4(d_addr, hr);
        tmem_ld_wait();
        const float4 b3 = *reinterpret_cast<const float4*>(w_smem + NL::kB3Off);
        float4 cot = make_float4(0.f, 0.f, 0.f, 0.f);          // un-normalised d(loss)/d(head); 1 / rows goes to the accumulators at the flush
        if (r < R) {
          if (!critic_job) {
            const float4 mu = make_float4(__uint_as_float(hr[0]) + b3.x, __uint_as_float(hr[1]) + b3.y, __uint_as_float(hr[2]) + b3.z, __uint_as_float(hr[3]) + b3.w);
            const float4 sg = __ldg(reinterpret_cast<const float4*>(fz.b.sigma));
            const PpoActorRow ar = ppo_actor_row(mu, sg, st_a, st_omu, st_osg, st_adv, st_logp, fz.b.clip_param);
            const float g_logp = ar.g_logp * inv_rows;
            float gs[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float d = ar.d[k], inv_s = ar.inv_s[k];
              (&cot.x)[k] = ar.g_logp * d * inv_s * inv_s;
              gs[k] = g_logp * (d * d * inv_s * inv_s * inv_s - inv_s) - fz.b.entropy_coef * inv_rows * inv_s;
            }
            lsum[0] += ar.surrogate; lsum[2] += ar.kl; lsum[3] += gs[0]; lsum[4] += gs[1]; lsum[5] += gs[2]; lsum[6] += gs[3]; lsum[7] += 1.0f;
          } else {
            const PpoCriticRow cr = ppo_critic_row(__uint_as_float(hr[0]) + b3.x, st_ret, st_ov, fz.b.use_clipped_value_loss != 0, fz.b.clip_param);
            cot.x = cr.g_v * fz.b.value_loss_coef;
            lsum[1] += cr.vloss;
          }
        }
        gsum.x += cot.x * inv_rows; gsum.y += cot.y * inv_rows; gsum.z += cot.z * inv_rows; gsum.w += cot.w * inv_rows;
        const float hi = 60000.0f;                             // (fp16 saturation guard; |cotangent * scale| this large does not occur in practice)
        *reinterpret_cast<uint4*>(das + row * 16) =
            make_uint4(h2_bits(__floats2half2_rn(fminf(fmaxf(cot.x * scale, -hi), hi), fminf(fmaxf(cot.y * scale, -hi), hi))),
                       h2_bits(__floats2half2_rn(fminf(fmaxf(cot.z * scale, -hi), hi), fminf(fmaxf(cot.w * scale, -hi), hi))), 0u, 0u);
      }
      // ---- dW3^T += H2^T . dA ;  dH2 = dA . W3
      GR_HANDOVER();
      dact_epilogue(d_addr, h2s + row * 16, 0, slope);                    // h2s now holds dH2'
      // ---- dW2 += dH2'^T . H1 ; db2 += dH2'^T . 1 ; dH1 = dH2' . W2, 128 units at a time
#pragma unroll 1
      for (int h = 0; h < kHalves; ++h) {
        GR_HANDOVER();
        // (h == 0 with two halves: dW2 also read ALL of h1s -- it completed with this stage, so overwriting is safe)
        dact_epilogue(d_addr, h1s + row * 16, h * 16, slope);             // h1s chunks [16h, 16h+16) now hold dH1'
      }
      // ---- dW1 | db1 += dH1'^T . [X | 1 1 0..]: nothing to read back, but xs / h1s stay in use until it completes
      GR_HANDOVER();
      cur = nxt;
      q_next = q_next2;
    }
#undef GR_HANDOVER
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      gsum.x += __shfl_xor_sync(0xffffffffu, gsum.x, o); gsum.y += __shfl_xor_sync(0xffffffffu, gsum.y, o);
      gsum.z += __shfl_xor_sync(0xffffffffu, gsum.z, o); gsum.w += __shfl_xor_sync(0xffffffffu, gsum.w, o);
    }
    if ((row & 31) == 0) {
      atomicAdd(out.b3 + 0, gsum.x);
      if (out.out_dim > 1) atomicAdd(out.b3 + 1, gsum.y);
      if (out.out_dim > 2) atomicAdd(out.b3 + 2, gsum.z);
      if (out.out_dim > 3) atomicAdd(out.b3 + 3, gsum.w);
    }
    if (kPpo) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = lsum[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((row & 31) == 0 && v != 0.0f) atomicAdd(fz.sums + k, v);
      }
    }
  } else {
    // ---- the issuing warp (all 32 lanes run this code converged, one elected lane issues): every MMA of the CTA, in program order
    const uint32_t w_addr = smem_u32(w_smem), ones_a = smem_u32(ones), grp0_a = smem_u32(grp_smem);
    // The issuer serves whichever group has its next stage ready (non-blocking test of `full[g]`), each group at its own stage: a group
    // that waits for its rows from HBM at the start of a tile does not hold the other one back (strict alternation did: ncu showed both
    // groups' threads in their mbarrier waits half of the time with the tensor pipe 15-30 % busy).
    int64_t left[kGroups];
    uint32_t ph[kGroups];
    int stage[kGroups];
    int active = 0;
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
      const int64_t first_tile = (int64_t)blockIdx.x * kGroups + g;
      left[g] = first_tile < tiles ? (tiles - first_tile + stride - 1) / stride : 0;
      ph[g] = 0u;
      stage[g] = 0;
      active += left[g] > 0 ? 1 : 0;
    }
    bool acc_w3 = false, acc_w2 = false, acc_w1 = false;     // the shared accumulators have been written at least once
    while (active > 0) {
      bool served = false;
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        if (left[g] <= 0 || !mbar_test(&full[g], ph[g])) continue;
        served = true;
        ph[g] ^= 1u;
        tc_fence_after_sync();
        const int s = stage[g];
        const uint32_t xs_a = grp0_a + g * kGroupBytes, h1_a = xs_a + 4 * kChunkA, h2_a = h1_a + (H1 / 8) * kChunkA, da_a = h2_a + (H2 / 8) * kChunkA;
        const uint32_t d_tm = tm + kColD + 128u * g;
        const int s_bwd = (kPpo && s > kHalves + 1) ? s - 1 : s;      // stage index in the plain backward's numbering
        if (kPpo && s == kHalves + 1) {                      // forward layer 3 (N = 16): the head of the net
#pragma unroll
          for (int kk = 0; kk < H2 / 16; ++kk)
            mma_f16_ss_warp(d_tm, desc_rows_k(h2_a + kk * 2 * kChunkA), make_smem_desc(w_addr + NL::kW3Off + kk * 2 * (kOutPad * 16), kOutPad * 16, 128),
                       make_idesc_f16(128, kOutPad), kk > 0);
        } else if (s < kHalves) {                            // forward layer 1, units [128 s, 128 s + 128)
#pragma unroll
          for (int kk = 0; kk < kK1 / 16; ++kk)
            mma_f16_ss_warp(d_tm, desc_rows_k(xs_a + kk * 2 * kChunkA), make_smem_desc(w_addr + NL::kW1Off + s * 128 * 16 + kk * 2 * (H1 * 16), H1 * 16, 128),
                       make_idesc_f16(128, 128), kk > 0);
        } else if (s == kHalves) {                           // forward layer 2
#pragma unroll
          for (int kk = 0; kk < H1 / 16; ++kk)
            mma_f16_ss_warp(d_tm, desc_rows_k(h1_a + kk * 2 * kChunkA), make_smem_desc(w_addr + NL::kW2Off + kk * 2 * (H2 * 16), H2 * 16, 128),
                       make_idesc_f16(128, H2), kk > 0);
        } else if (s_bwd == kHalves + 1) {                   // dW3^T += H2^T . dA (rows are K: both operands MN-major) ; dH2 = dA . W3 (W3 as MN-major B)
#pragma unroll
          for (int kk = 0; kk < kTileEnvs / 16; ++kk)
            mma_f16_ss_warp(tm + kColW3, desc_rows_mn(h2_a + kk * 256), desc_rows_mn(da_a + kk * 256), make_idesc_f16(H2, 16) | kMnA | kMnB, acc_w3 || kk > 0);
          acc_w3 = true;
          mma_f16_ss_warp(d_tm, desc_rows_k(da_a), make_smem_desc(w_addr + NL::kW3Off, 128, kOutPad * 16), make_idesc_f16(128, H2) | kMnB, false);
        } else if (s_bwd < 2 * kHalves + 2) {                // dW2 | db2 (first half only) ; dH1 units [128 h, 128 h + 128) = dH2' . W2 (W2 as MN-major B)
          const int h = s_bwd - (kHalves + 2);
          if (h == 0) {
#pragma unroll
            for (int kk = 0; kk < kTileEnvs / 16; ++kk) {
              mma_f16_ss_warp(tm + kColW2, desc_rows_mn(h2_a + kk * 256), desc_rows_mn(h1_a + kk * 256), make_idesc_f16(H2, H1) | kMnA | kMnB, acc_w2 || kk > 0);
              mma_f16_ss_warp(tm + kColB2, desc_rows_mn(h2_a + kk * 256), desc_rows_mn(ones_a + kk * 256), make_idesc_f16(H2, 16) | kMnA | kMnB, acc_w2 || kk > 0);
            }
            acc_w2 = true;
          }
#pragma unroll
          for (int kk = 0; kk < H2 / 16; ++kk)
            mma_f16_ss_warp(d_tm, desc_rows_k(h2_a + kk * 2 * kChunkA), make_smem_desc(w_addr + NL::kW2Off + h * 16 * (H2 * 16) + kk * 256, 128, H2 * 16),
                       make_idesc_f16(128, 128) | kMnB, kk > 0);
        } else {                                             // dW1 | db1 += dH1'^T . [X | 1 1 0..]
#pragma unroll
          for (int h = 0; h < kHalves; ++h)
#pragma unroll
            for (int kk = 0; kk < kTileEnvs / 16; ++kk)
              mma_f16_ss_warp(tm + kColW1 + 32 * h, desc_rows_mn(h1_a + h * 16 * kChunkA + kk * 256), desc_rows_mn(xs_a + kk * 256), make_idesc_f16(128, 32) | kMnA | kMnB,
                         acc_w1 || kk > 0);
          acc_w1 = true;
        }
        tc_commit_warp(&done[g]);
        if (++stage[g] == kStages) {
          stage[g] = 0;
          if (--left[g] == 0) --active;
        }
      }
      __syncwarp();
      if (!served) __nanosleep(20);
    }
  }
  // every group waited for its own last stage; after this barrier every MMA of the CTA has completed
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();

  // ---- flush (group 0): accumulator row m = TMEM lane m = this thread; everything divided by the loss scale
  if (tid < kTileEnvs && (int64_t)blockIdx.x * kGroups < tiles) {
    const uint32_t lane_sel = (uint32_t)((row >> 5) * 32) << 16;
    const float inv = kPpo ? inv_rows / scale : 1.0f / scale;
    const int j = row;                                       // unit of layer 2 (dW2, db2, dW3) / unit within a 128-block of layer 1 (dW1)
#pragma unroll 1
    for (int c0 = 0; c0 < H1; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tm + kColW2 + lane_sel + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 16; k += 4)
        red_add_v4(out.w2 + (int64_t)j * H1 + c0 + k, __uint_as_float(v[k]) * inv, __uint_as_float(v[k + 1]) * inv, __uint_as_float(v[k + 2]) * inv,
                   __uint_as_float(v[k + 3]) * inv);
    }
#pragma unroll 1
    for (int h = 0; h < kHalves; ++h) {
      uint32_t v[16], w[16];
      tmem_ld_x16(tm + kColW1 + 32 * h + lane_sel, v);
      tmem_ld_x16(tm + kColW1 + 32 * h + lane_sel + 16, w);
      tmem_ld_wait();
#pragma unroll
      for (int k = 0; k < 16; k += 4)
        red_add_v4(out.w1 + (int64_t)(128 * h + j) * kObsDim + k, __uint_as_float(v[k]) * inv, __uint_as_float(v[k + 1]) * inv, __uint_as_float(v[k + 2]) * inv,
                   __uint_as_float(v[k + 3]) * inv);
      atomicAdd(out.b1 + 128 * h + j, __uint_as_float(w[0]) * inv);
    }
    {
      uint32_t v[16], w[16];
      tmem_ld_x16(tm + kColW3 + lane_sel, v);
      tmem_ld_x16(tm + kColB2 + lane_sel, w);
      tmem_ld_wait();
#pragma unroll
      for (int a = 0; a < GR_NUM_ACTIONS; ++a) if (a < out.out_dim) atomicAdd(out.w3 + (int64_t)a * H2 + j, __uint_as_float(v[a]) * inv);
      atomicAdd(out.b2 + j, __uint_as_float(w[0]) * inv);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tm, 512);
}

}  // namespace gr

using namespace gr;

template <class NL, int kGroups, bool kPpo = false>
static int launch_actor_backward(const GrBackwardJob* jobs, int n_jobs, int64_t R, cudaStream_t s, const PpoFusedArgs* fz = nullptr) {
  const size_t bytes = (size_t)NL::kNetBytes + (size_t)kGroups * (4 + NL::kH1 / 8 + NL::kH2 / 8 + 2) * kChunkA + 2 * kChunkA + 128;
  if (bytes > 227 * 1024) return GR_ERR_SMEM;
  auto kernel = actor_backward_kernel<NL, kGroups, kPpo>;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = (R + kTileEnvs - 1) / kTileEnvs;
  const int64_t units = (tiles + kGroups - 1) / kGroups;      // a CTA works on kGroups tiles at a time
  const int per_job = sms / n_jobs;
  const int grid = (int)(units < per_job ? units : per_job);
  const PpoFusedArgs none = {};
  kernel<<<dim3(grid, n_jobs), kGroups * kTileEnvs + 32, bytes, s>>>(jobs[0], jobs[n_jobs - 1], R, fz ? *fz : none);
  return (int)cudaGetLastError();
}

static int check_job(const GrBackwardJob* j) {
  if (!j->policy.packed || !j->obs || !j->grad_actions || !j->scale) return GR_ERR_NULL;
  const GrMlpGrad* out = &j->out;
  if (!out->w1 || !out->b1 || !out->w2 || !out->b2 || !out->w3 || !out->b3) return GR_ERR_NULL;
  if (out->out_dim < 1 || out->out_dim > 4) return GR_ERR_SIZE;
  if (j->policy.negative_slope < 0.0f || j->policy.negative_slope > 1.0f) return GR_ERR_CONFIG;
  if (j->obs_stride < 0 || (j->obs_stride & 3) || (j->obs_stride > 0 && j->obs_stride < kObsDim)) return GR_ERR_SIZE;
  if ((reinterpret_cast<uintptr_t>(j->policy.packed) | reinterpret_cast<uintptr_t>(j->obs) | reinterpret_cast<uintptr_t>(j->grad_actions) |
       reinterpret_cast<uintptr_t>(out->w1) | reinterpret_cast<uintptr_t>(out->w2)) & 15u)
    return GR_ERR_ALIGN;
  return GR_OK;
}

extern "C" int gr_actor_backward_jobs(const GrBackwardJob* jobs, int32_t n_jobs, int32_t hidden, int32_t hidden2, int64_t rows, void* stream) {
  if (!jobs) return GR_ERR_NULL;
  if (n_jobs < 1 || n_jobs > 2 || rows <= 0) return GR_ERR_SIZE;
  if (!((hidden == 128 || hidden == 256) && hidden2 == 128)) return GR_ERR_SIZE;
  for (int k = 0; k < n_jobs; ++k) { const int rc = check_job(&jobs[k]); if (rc != GR_OK) return rc; }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // two tiles in flight per CTA where tensor memory has room for two scratch accumulators (16 -> 128 -> 128); one otherwise.
  // (a job with very few tiles keeps one group per CTA so that the tiles spread over more SMs)
  if (hidden == 256) return launch_actor_backward<NetLayout<256, 128>, 1>(jobs, n_jobs, rows, s);
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = (rows + kTileEnvs - 1) / kTileEnvs;
  bool two = tiles > sms / n_jobs;
  if (const char* e = getenv("GRACING_ACTOR_BACKWARD_GROUPS")) two = e[0] == '2';      // (A/B tests: the two variants must agree to accumulation order)
  return two ? launch_actor_backward<NetLayout<128, 128>, 2>(jobs, n_jobs, rows, s) : launch_actor_backward<NetLayout<128, 128>, 1>(jobs, n_jobs, rows, s);
}

extern "C" int gr_actor_backward(const GrPolicy* policy, int32_t hidden, int32_t hidden2, const float* obs, const float* grad_actions,
                                 const float* scale, int64_t rows, const GrMlpGrad* out, void* stream) {
  if (!policy || !out) return GR_ERR_NULL;
  GrBackwardJob job;
  job.policy = *policy; job.obs = obs; job.grad_actions = grad_actions; job.scale = scale; job.out = *out; job.indices = nullptr; job.obs_stride = 0;
  return gr_actor_backward_jobs(&job, 1, hidden, hidden2, rows, stream);
}

extern "C" int gr_ppo_fused_step(const GrPpoStep* st, int64_t rows, void* stream) {
  if (!st) return GR_ERR_NULL;
  const GrPpoBatch& b = st->batch;
  if (!st->policy.packed || !st->sums || !b.sigma) return GR_ERR_NULL;
  const bool rec = b.records != nullptr;
  if (!rec && (!st->obs || !st->critic_obs || !b.actions || !b.old_log_prob || !b.advantages || !b.returns || !b.old_mu || !b.old_sigma)) return GR_ERR_NULL;
  if (!rec && b.use_clipped_value_loss && !b.old_values) return GR_ERR_NULL;
  if (rec && (reinterpret_cast<uintptr_t>(b.records) & 15u)) return GR_ERR_ALIGN;
  if (rows <= 0) return GR_ERR_SIZE;
  if (st->cotangent_scale < 0.0f) return GR_ERR_CONFIG;
  GrBackwardJob jobs[2];
  const int net_bytes = NetLayout<128, 128>::kNetBytes;
  for (int k = 0; k < 2; ++k) {
    jobs[k].policy = st->policy;
    if (k == 1) jobs[k].policy.packed = static_cast<const uint8_t*>(st->policy.packed) + net_bytes;
    jobs[k].obs = rec ? b.records + 16 * k : (k == 0 ? st->obs : st->critic_obs);
    jobs[k].grad_actions = jobs[k].obs;            // (unused by the fused kernel; non-null for check_job)
    jobs[k].scale = st->sums;                      // (unused)
    jobs[k].out = k == 0 ? st->actor_grad : st->critic_grad;
    jobs[k].indices = b.indices;
    jobs[k].obs_stride = rec ? GR_RECORD_FLOATS : 0;
    const int rc = check_job(&jobs[k]);
    if (rc != GR_OK) return rc;
  }
  if (reinterpret_cast<uintptr_t>(b.sigma) & 15u) return GR_ERR_ALIGN;
  if (!rec && ((reinterpret_cast<uintptr_t>(b.actions) | reinterpret_cast<uintptr_t>(b.old_mu) | reinterpret_cast<uintptr_t>(b.old_sigma)) & 15u)) return GR_ERR_ALIGN;
  if (jobs[0].out.out_dim != 4 || jobs[1].out.out_dim != 1) return GR_ERR_SIZE;
  PpoFusedArgs fz;
  fz.b = b;
  fz.sums = st->sums;
  fz.cot_scale = st->cotangent_scale > 0.0f ? st->cotangent_scale : 1.0f / 64.0f;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = (rows + kTileEnvs - 1) / kTileEnvs;
  return tiles > sms / 2 ? launch_actor_backward<NetLayout<128, 128>, 2, true>(jobs, 2, rows, s, &fz) : launch_actor_backward<NetLayout<128, 128>, 1, true>(jobs, 2, rows, s, &fz);
}

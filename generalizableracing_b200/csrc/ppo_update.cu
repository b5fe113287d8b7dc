// ppo_update.cu -- one PPO mini-batch step without autograd: policy forward on the tensor cores and the loss gradients.
//
// Reference path replaced (S = standalone): the body of PPO.update's mini-batch loop (S/rsl_rl/ext/algorithms/ppo.py:118-171):
//   policy.act / get_actions_log_prob / evaluate            -> policy_forward_kernel  (mu = actor(obs), v = critic(critic_obs))
//   ratio, clipped surrogate, clipped value loss, entropy,  -> ppo_loss_grad_kernel   (d loss / d mu, d loss / d v, d loss / d std,
//   KL for the adaptive learning rate, loss.backward() up      loss and KL sums)
//   to the network outputs
// The weight gradients then come from gr_actor_backward (actor with d/d mu, critic with d/d v); clipping and Adam stay torch.
#include <cstdlib>
#include "mlp_tc.cuh"
#include "ppo_loss.cuh"

namespace gr {

using NLp = NetLayout<128, 128>;
constexpr int kFwdGroups = 2;                // two 128-row tiles in flight per CTA (one covers the other's MMA latency)

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// what the forward kernel needs to evaluate the loss of a row itself (kLoss): the stored columns, the outputs, the sums
struct FwdLossArgs { GrPpoBatch b; float* grad_mu; float* grad_value; float* sums; };

// mu [rows,4], value [rows]: persistent over 128-row tiles, both nets resident in shared memory.  The rows of a tile (gathered on load:
// scattered 64-byte reads behind a scattered index read) are requested one tile ahead, their indices two tiles ahead, unconditionally
// (rows past the end are clamped and never stored) -- see actor_backward.cu.
// kLoss: the thread that holds a row's mean and value also evaluates the row's loss (ppo_loss.cuh) on the stored columns it fetched with
// the row, and writes d(loss)/d(mu), d(loss)/d(v) and its share of the sums: gr_ppo_loss_grad's launch and the mu / value round trip go.
template <bool kLoss, bool kInter>
__global__ void __launch_bounds__(kFwdGroups * kTileEnvs, 1) policy_forward_kernel(const GrPolicy pol, const float* __restrict__ obs,
                                                                                  const float* __restrict__ critic_obs, const int64_t* __restrict__ idx,
                                                                                  float* __restrict__ mu, float* __restrict__ value, const int64_t R,
                                                                                  const FwdLossArgs la) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* w_smem = smem;
  uint8_t* h_smem = smem + 2 * NLp::kNetBytes;                     // activation tiles: [actor | critic][group]
  constexpr int kCtx = kInter ? 2 * kFwdGroups : kFwdGroups;      // (activation tile, accumulator, mbarrier) sets of the CTA
  uint64_t* bars = reinterpret_cast<uint64_t*>(h_smem + kCtx * NLp::kHBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + kCtx);
  const int tid = threadIdx.x, grp = tid / kTileEnvs, row = tid % kTileEnvs;
  {
    const uint4* src = reinterpret_cast<const uint4*>(pol.packed);
    uint4* dst = reinterpret_cast<uint4*>(w_smem);
    for (int k = tid; k < 2 * NLp::kNetBytes / 16; k += kFwdGroups * kTileEnvs) dst[k] = __ldg(src + k);
  }
  if (tid < kCtx) mbar_init(&bars[tid], 1);
  __syncwarp();
  if (tid < 32) tmem_alloc(tmem_slot, kCtx * NLp::kCols);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  // The two nets of a tile are independent chains (actor on the policy observation, critic on the critic's): each group runs them
  // INTERLEAVED on two activation tiles / accumulators / mbarriers of its own, so that the epilogue of one net's layer runs while the
  // other net's MMAs are in flight -- four of the six MMA waits of a tile disappear behind epilogue work (they were 26 % of the stall
  // samples, profiles/r2_ncu_ppo_update_kernels_summary.md).  The 128 threads of a group pass the same named barriers in program order,
  // so both contexts share the group's two barrier ids; the issuing thread queues actor and critic MMAs in that order.
  GroupCtx g = make_group_ctx(h_smem, NLp::kHBytes, bars, *tmem_slot, NLp::kCols, grp, row, pol.negative_slope);
  GroupCtx gc = make_group_ctx(h_smem, NLp::kHBytes, bars, *tmem_slot, NLp::kCols, kInter ? kFwdGroups + grp : grp, row, pol.negative_slope);
  gc.bar_id = g.bar_id;
  gc.issuer = g.issuer;
  gc.issuer_warp = g.issuer_warp;
  const uint32_t w_addr = smem_u32(w_smem);
  const uint8_t* wc_smem = w_smem + NLp::kNetBytes;
  const uint32_t wc_addr = w_addr + NLp::kNetBytes;
  const int64_t tiles = (R + kTileEnvs - 1) / kTileEnvs;
  const int64_t stride = (int64_t)gridDim.x * kFwdGroups;
  const bool gather = idx != nullptr;
  struct Rows { float4 o0, o1, o2, o3, c0, c1, c2, c3, a, omu, osg; float adv, logp, ret, ov; };
  auto index_of = [&](const int64_t tile) -> int64_t {
    int64_t r = tile * kTileEnvs + row;
    r = r < R ? r : R - 1;
    return gather ? __ldg(idx + r) : r;
  };
  const float4* __restrict__ rec = kLoss ? reinterpret_cast<const float4*>(la.b.records) : nullptr;
  auto fetch = [&](const int64_t q, Rows& t) {
    if (kLoss && rec) {                            // one 192-byte transition record (gr_storage_pack_records): 12 contiguous 16-byte loads
      const float4* p = rec + q * (GR_RECORD_FLOATS / 4);
      t.o0 = __ldcs(p); t.o1 = __ldcs(p + 1); t.o2 = __ldcs(p + 2); t.o3 = __ldcs(p + 3);
      t.c0 = __ldcs(p + 4); t.c1 = __ldcs(p + 5); t.c2 = __ldcs(p + 6); t.c3 = __ldcs(p + 7);
      t.a = __ldcs(p + 8); t.omu = __ldcs(p + 9); t.osg = __ldcs(p + 10);
      const float4 sc = __ldcs(p + 11);
      t.logp = sc.x; t.adv = sc.y; t.ret = sc.z; t.ov = sc.w;
      return;
    }
    const float4* xo = reinterpret_cast<const float4*>(obs) + q * 4;
    const float4* xc = reinterpret_cast<const float4*>(critic_obs) + q * 4;
    t.o0 = __ldcs(xo); t.o1 = __ldcs(xo + 1); t.o2 = __ldcs(xo + 2); t.o3 = __ldcs(xo + 3);
    t.c0 = __ldcs(xc); t.c1 = __ldcs(xc + 1); t.c2 = __ldcs(xc + 2); t.c3 = __ldcs(xc + 3);
    if (kLoss) {
      t.a = __ldg(reinterpret_cast<const float4*>(la.b.actions) + q);
      t.omu = __ldg(reinterpret_cast<const float4*>(la.b.old_mu) + q);
      t.osg = __ldg(reinterpret_cast<const float4*>(la.b.old_sigma) + q);
      t.adv = __ldg(la.b.advantages + q);
      t.logp = __ldg(la.b.old_log_prob + q);
      t.ret = __ldg(la.b.returns + q);
      t.ov = la.b.use_clipped_value_loss ? __ldg(la.b.old_values + q) : 0.0f;
    }
  };
  Rows cur = {}, nxt = {};
  const int64_t first = (int64_t)blockIdx.x * kFwdGroups + grp;
  int64_t q_next = 0, q_next2 = 0;
  // (every thread of the CTA runs the same number of rounds -- the named barriers are per group, the tile loop is not -- so the requests
  //  are issued whether or not this group's tile exists)
  fetch(index_of(first), cur);
  q_next = index_of(first + stride);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float max_gm = 0.0f, max_gv = 0.0f;
  const float inv_rows = 1.0f / (float)R;
#pragma unroll 1
  for (int64_t base = (int64_t)blockIdx.x * kFwdGroups; base < tiles; base += stride) {
    const int64_t tile = base + grp;
    const int64_t r = tile * kTileEnvs + row;
    const bool live = tile < tiles && r < R;
    q_next2 = index_of(tile + 2 * stride);
    fetch(q_next, nxt);
    float4 m, v;
    if (kInter) {
      write_x_row(g.hrow, pack8(cur.o0, cur.o1), pack8(cur.o2, cur.o3));
      stage_issue<NLp>(g, w_addr, kL1);
      write_x_row(gc.hrow, pack8(cur.c0, cur.c1), pack8(cur.c2, cur.c3));
      stage_issue<NLp>(gc, wc_addr, kL1);
      stage_wait(g);  epilogue1<NLp>(g);            stage_issue<NLp>(g, w_addr, kL2);
      stage_wait(gc); epilogue1<NLp>(gc);           stage_issue<NLp>(gc, wc_addr, kL2);
      stage_wait(g);  epilogue2<NLp>(g, w_smem);    stage_issue<NLp>(g, w_addr, kL3);
      stage_wait(gc); epilogue2<NLp>(gc, wc_smem);  stage_issue<NLp>(gc, wc_addr, kL3);
      stage_wait(g);
      m = read_head<NLp>(g, w_smem);
      stage_wait(gc);
      v = read_head<NLp>(gc, wc_smem);
    } else {
      write_x_row(g.hrow, pack8(cur.o0, cur.o1), pack8(cur.o2, cur.o3));
      m = run_net<NLp>(g, w_smem, w_addr);
      tc_fence_before_sync();
      write_x_row(g.hrow, pack8(cur.c0, cur.c1), pack8(cur.c2, cur.c3));
      v = run_net<NLp>(g, wc_smem, wc_addr);
    }
    tc_fence_before_sync();
    if (live) {
      if (mu) __stcs(reinterpret_cast<float4*>(mu) + r, m);
      if (value) value[r] = v.x;
      if (kLoss) {       // exactly ppo_loss_grad_kernel's arithmetic on this row
        const float4 sg = __ldg(reinterpret_cast<const float4*>(la.b.sigma));
        const PpoActorRow ar = ppo_actor_row(m, sg, cur.a, cur.omu, cur.osg, cur.adv, cur.logp, la.b.clip_param);
        const float g_logp = ar.g_logp * inv_rows;
        const PpoCriticRow cr = ppo_critic_row(v.x, cur.ret, cur.ov, la.b.use_clipped_value_loss != 0, la.b.clip_param);
        float g_v = cr.g_v;
        g_v *= la.b.value_loss_coef * inv_rows;
        float4 gm;
        float gs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float d = ar.d[k], inv_s = ar.inv_s[k];
          (&gm.x)[k] = g_logp * d * inv_s * inv_s;
          gs[k] = g_logp * (d * d * inv_s * inv_s * inv_s - inv_s) - la.b.entropy_coef * inv_rows * inv_s;
        }
        __stcs(reinterpret_cast<float4*>(la.grad_mu) + r, gm);
        __stcs(reinterpret_cast<float4*>(la.grad_value) + r, make_float4(g_v, 0.f, 0.f, 0.f));
        max_gm = fmaxf(max_gm, fmaxf(fmaxf(fabsf(gm.x), fabsf(gm.y)), fmaxf(fabsf(gm.z), fabsf(gm.w))));
        max_gv = fmaxf(max_gv, fabsf(g_v));
        acc[0] += ar.surrogate; acc[1] += cr.vloss; acc[2] += ar.kl; acc[3] += gs[0]; acc[4] += gs[1]; acc[5] += gs[2]; acc[6] += gs[3]; acc[7] += 1.0f;
      }
    }
    cur = nxt;
    q_next = q_next2;
  }
  if (kLoss) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float sacc = warp_sum_f(acc[k]);
      if (lane == 0 && sacc != 0.0f) atomicAdd(la.sums + k, sacc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { max_gm = fmaxf(max_gm, __shfl_xor_sync(0xffffffffu, max_gm, o)); max_gv = fmaxf(max_gv, __shfl_xor_sync(0xffffffffu, max_gv, o)); }
    if (lane == 0) {       // non-negative floats order like their bit patterns
      atomicMax(reinterpret_cast<unsigned int*>(la.sums) + 8, __float_as_uint(max_gm));
      atomicMax(reinterpret_cast<unsigned int*>(la.sums) + 9, __float_as_uint(max_gv));
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (tid < 32) tmem_dealloc(*tmem_slot, kCtx * NLp::kCols);
}

// ---------------------------------------------------------------------------------------------
// loss gradients of one mini-batch (ppo.py:118-171), one thread per row
//   sums[0] += sum surrogate, [1] += sum value loss, [2] += sum KL, [3..6] += d loss / d std, [7] += rows
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ppo_loss_grad_kernel(const GrPpoBatch b, const int64_t R, float* __restrict__ grad_mu, float* __restrict__ grad_value,
                                                           float* __restrict__ sums) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float max_gm = 0.0f, max_gv = 0.0f;
  if (r < R) {
    const float inv_rows = 1.0f / (float)R;
    const int64_t q = b.indices ? __ldg(b.indices + r) : r;    // row of the stored columns (gather on load)
    const float4 mu = __ldg(reinterpret_cast<const float4*>(b.mu) + r), a = __ldg(reinterpret_cast<const float4*>(b.actions) + q);
    const float4 sg = __ldg(reinterpret_cast<const float4*>(b.sigma));
    const float4 omu = __ldg(reinterpret_cast<const float4*>(b.old_mu) + q), osg = __ldg(reinterpret_cast<const float4*>(b.old_sigma) + q);
    const PpoActorRow ar = ppo_actor_row(mu, sg, a, omu, osg, b.advantages[q], b.old_log_prob[q], b.clip_param);
    const float kl = ar.kl, surrogate = ar.surrogate;
    const float g_logp = ar.g_logp * inv_rows;
    const PpoCriticRow cr = ppo_critic_row(b.value[r], b.returns[q], b.use_clipped_value_loss ? b.old_values[q] : 0.0f, b.use_clipped_value_loss != 0, b.clip_param);
    const float vloss = cr.vloss;
    float g_v = cr.g_v;
    g_v *= b.value_loss_coef * inv_rows;
    float4 gm;
    float gs[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float d = ar.d[k], inv_s = ar.inv_s[k];
      (&gm.x)[k] = g_logp * d * inv_s * inv_s;
      // d logp / d std = d^2 / std^3 - 1 / std ; entropy = sum(0.5 + 0.5 log 2pi + log std) enters with -entropy_coef * mean
      gs[k] = g_logp * (d * d * inv_s * inv_s * inv_s - inv_s) - b.entropy_coef * inv_rows * inv_s;
    }
    __stcs(reinterpret_cast<float4*>(grad_mu) + r, gm);
    __stcs(reinterpret_cast<float4*>(grad_value) + r, make_float4(g_v, 0.f, 0.f, 0.f));
    max_gm = fmaxf(fmaxf(fabsf(gm.x), fabsf(gm.y)), fmaxf(fabsf(gm.z), fabsf(gm.w)));
    max_gv = fabsf(g_v);
    acc[0] = surrogate; acc[1] = vloss; acc[2] = kl; acc[3] = gs[0]; acc[4] = gs[1]; acc[5] = gs[2]; acc[6] = gs[3]; acc[7] = 1.0f;
  }
  __shared__ float part[8][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float s = warp_sum_f(acc[k]);
    if (lane == 0) part[wid][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += part[w][threadIdx.x];
    atomicAdd(sums + threadIdx.x, s);
  }
  // loss scales of the two backward launches: non-negative floats order like their bit patterns
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { max_gm = fmaxf(max_gm, __shfl_xor_sync(0xffffffffu, max_gm, o)); max_gv = fmaxf(max_gv, __shfl_xor_sync(0xffffffffu, max_gv, o)); }
  if (lane == 0) {
    atomicMax(reinterpret_cast<unsigned int*>(sums) + 8, __float_as_uint(max_gm));
    atomicMax(reinterpret_cast<unsigned int*>(sums) + 9, __float_as_uint(max_gv));
  }
}

// ---------------------------------------------------------------------------------------------
// clip_grad_norm_ + Adam + KL-adaptive learning rate in two launches over the flat parameter view
// ---------------------------------------------------------------------------------------------
enum AdamState : int { kLr = 0, kStep = 1, kClip = 2, kStepSize = 3, kInvSqrtBc2 = 4, kSumVLoss = 5, kSumSurr = 6, kSumSq = 8, kCounter = 9 };

// tensor k owns flat[offs[k], offs[k] + sizes[k]); everything else in the flat buffer (alignment padding, the loss kernel's sums) is skipped
__device__ __forceinline__ int segment_of(int i, const int* offs, const int* sizes, int n_seg) {
  int k = 0;
  while (k + 1 < n_seg && i >= offs[k + 1]) ++k;
  return (i >= offs[k] && i - offs[k] < sizes[k]) ? k : -1;
}

__global__ void __launch_bounds__(256) adam_norm_kernel(const GrAdamStep a) {
  __shared__ int offs[32], sizes[32];
  if (threadIdx.x < a.n_seg) { offs[threadIdx.x] = a.seg_offsets[threadIdx.x]; sizes[threadIdx.x] = a.seg_sizes[threadIdx.x]; }
  __syncthreads();
  float s = 0.0f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_flat; i += gridDim.x * blockDim.x) {
    if (segment_of(i, offs, sizes, a.n_seg) < 0) continue;
    const float g = a.grad[i] * a.grad_scale;
    s += g * g;
  }
  s = warp_sum_f(s);
  __shared__ float part[8];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    atomicAdd(a.state + kSumSq, t);
    __threadfence();
    last = atomicAdd(reinterpret_cast<unsigned int*>(a.state) + kCounter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {                       // the last block to finish sees every partial sum: finalise the step's scalars
    __threadfence();
    const float norm = sqrtf(*reinterpret_cast<volatile float*>(a.state + kSumSq));
    float lr = a.state[kLr];
    if (a.kl_stats) {                                   // ppo.py:124-141
      const float kl_mean = a.kl_stats[2] / a.kl_stats[7];
      if (kl_mean > a.desired_kl * 2.0f) lr = fmaxf(a.lr_min, lr / 1.5f);
      else if (kl_mean < a.desired_kl / 2.0f && kl_mean > 0.0f) lr = fminf(a.lr_max, lr * 1.5f);
      a.state[kSumSurr] += a.kl_stats[0] / a.kl_stats[7];
      a.state[kSumVLoss] += a.kl_stats[1] / a.kl_stats[7];
    }
    const float step = a.state[kStep] + 1.0f;
    a.state[kLr] = lr;
    a.state[kStep] = step;
    a.state[kClip] = fminf(1.0f, a.max_grad_norm / (norm + 1e-6f));            // clip_grad_norm_: clamp(max_norm / (total_norm + 1e-6), max=1)
    a.state[kStepSize] = lr / (1.0f - powf(a.beta1, step));
    a.state[kInvSqrtBc2] = rsqrtf(1.0f - powf(a.beta2, step));
    a.state[kSumSq] = 0.0f;
    reinterpret_cast<unsigned int*>(a.state)[kCounter] = 0u;
  }
}

__global__ void __launch_bounds__(256) adam_apply_kernel(const GrAdamStep a) {
  __shared__ int offs[32], sizes[32];
  __shared__ float* ptrs[32];
  if (threadIdx.x < a.n_seg) { offs[threadIdx.x] = a.seg_offsets[threadIdx.x]; sizes[threadIdx.x] = a.seg_sizes[threadIdx.x]; ptrs[threadIdx.x] = reinterpret_cast<float*>(a.param_ptrs[threadIdx.x]); }
  __syncthreads();
  const float clip = a.state[kClip] * a.grad_scale, step_size = a.state[kStepSize], inv_sqrt_bc2 = a.state[kInvSqrtBc2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_flat; i += gridDim.x * blockDim.x) {
    const int k = segment_of(i, offs, sizes, a.n_seg);
    if (k < 0) continue;
    const int j = i - offs[k];
    const float g = a.grad[i] * clip;
    const float m = a.beta1 * a.exp_avg[i] + (1.0f - a.beta1) * g;
    const float v = a.beta2 * a.exp_avg_sq[i] + (1.0f - a.beta2) * g * g;
    a.exp_avg[i] = m;
    a.exp_avg_sq[i] = v;
    ptrs[k][j] -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + a.eps);
  }
}

}  // namespace gr

using namespace gr;

template <bool kLoss>
static int launch_policy_forward(const GrPolicy* policy, const float* obs, const float* critic_obs, const int64_t* indices, float* mu, float* value, int64_t rows,
                                 const FwdLossArgs& la, cudaStream_t s) {
  // GRACING_FWD_INTERLEAVE=1: the two nets of a tile as interleaved chains (two activation tiles / accumulators per group); measured, see DESIGN 4d
  static const bool inter = [] { const char* e = getenv("GRACING_FWD_INTERLEAVE"); return e && e[0] == '1'; }();
  const size_t bytes = 2 * (size_t)NLp::kNetBytes + (inter ? 2 : 1) * (size_t)kFwdGroups * NLp::kHBytes + 128;      // weights | activation tiles
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t pairs = ((rows + kTileEnvs - 1) / kTileEnvs + kFwdGroups - 1) / kFwdGroups;
  const int grid = (int)(pairs < sms ? pairs : sms);
  if (inter) {
    auto kernel = policy_forward_kernel<kLoss, true>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    kernel<<<grid, kFwdGroups * kTileEnvs, bytes, s>>>(*policy, obs, critic_obs, indices, mu, value, rows, la);
  } else {
    auto kernel = policy_forward_kernel<kLoss, false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    kernel<<<grid, kFwdGroups * kTileEnvs, bytes, s>>>(*policy, obs, critic_obs, indices, mu, value, rows, la);
  }
  return (int)cudaGetLastError();
}

extern "C" int gr_policy_forward_gather(const GrPolicy* policy, const float* obs, const float* critic_obs, const int64_t* indices, float* mu, float* value,
                                        int64_t rows, void* stream) {
  if (!policy || !policy->packed || !obs || !critic_obs || !mu || !value) return GR_ERR_NULL;
  if (rows <= 0) return GR_ERR_SIZE;
  if (policy->negative_slope < 0.0f || policy->negative_slope > 1.0f) return GR_ERR_CONFIG;
  if ((reinterpret_cast<uintptr_t>(policy->packed) | reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(critic_obs) | reinterpret_cast<uintptr_t>(mu)) & 15u)
    return GR_ERR_ALIGN;
  const FwdLossArgs none = {};
  return launch_policy_forward<false>(policy, obs, critic_obs, indices, mu, value, rows, none, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gr_policy_forward_loss(const GrPolicy* policy, const float* obs, const float* critic_obs, const GrPpoBatch* b, int64_t rows, float* grad_mu,
                                      float* grad_value, float* sums, void* stream) {
  if (!policy || !policy->packed || !b || !grad_mu || !grad_value || !sums || !b->sigma) return GR_ERR_NULL;
  if (rows <= 0) return GR_ERR_SIZE;
  if (policy->negative_slope < 0.0f || policy->negative_slope > 1.0f) return GR_ERR_CONFIG;
  uintptr_t align = reinterpret_cast<uintptr_t>(policy->packed) | reinterpret_cast<uintptr_t>(b->sigma) | reinterpret_cast<uintptr_t>(grad_mu) |
                    reinterpret_cast<uintptr_t>(grad_value) | (b->mu ? reinterpret_cast<uintptr_t>(b->mu) : 0);
  if (b->records) {                                // everything a row needs comes from its transition record
    align |= reinterpret_cast<uintptr_t>(b->records);
  } else {
    if (!obs || !critic_obs || !b->actions || !b->old_log_prob || !b->advantages || !b->returns || !b->old_mu || !b->old_sigma) return GR_ERR_NULL;
    if (b->use_clipped_value_loss && !b->old_values) return GR_ERR_NULL;
    align |= reinterpret_cast<uintptr_t>(obs) | reinterpret_cast<uintptr_t>(critic_obs) | reinterpret_cast<uintptr_t>(b->actions) |
             reinterpret_cast<uintptr_t>(b->old_mu) | reinterpret_cast<uintptr_t>(b->old_sigma);
  }
  if (align & 15u) return GR_ERR_ALIGN;
  FwdLossArgs la;
  la.b = *b; la.grad_mu = grad_mu; la.grad_value = grad_value; la.sums = sums;
  return launch_policy_forward<true>(policy, obs, critic_obs, b->indices, const_cast<float*>(b->mu), const_cast<float*>(b->value), rows, la,
                                     reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gr_policy_forward(const GrPolicy* policy, const float* obs, const float* critic_obs, float* mu, float* value, int64_t rows, void* stream) {
  return gr_policy_forward_gather(policy, obs, critic_obs, nullptr, mu, value, rows, stream);
}

extern "C" int gr_adam_clip_step(const GrAdamStep* a, void* stream) {
  if (!a || !a->param_ptrs || !a->seg_offsets || !a->seg_sizes || !a->grad || !a->exp_avg || !a->exp_avg_sq || !a->state) return GR_ERR_NULL;
  if (a->n_seg < 1 || a->n_seg > 32 || a->n_flat < 1) return GR_ERR_SIZE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = (a->n_flat + 256 * 4 - 1) / (256 * 4);
  adam_norm_kernel<<<blocks < 148 ? blocks : 148, 256, 0, s>>>(*a);
  adam_apply_kernel<<<(a->n_flat + 255) / 256, 256, 0, s>>>(*a);
  return (int)cudaGetLastError();
}

extern "C" int gr_ppo_loss_grad(const GrPpoBatch* b, int64_t rows, float* grad_mu, float* grad_value, float* sums, void* stream) {
  if (!b || !grad_mu || !grad_value || !sums) return GR_ERR_NULL;
  if (!b->mu || !b->value || !b->sigma || !b->actions || !b->old_log_prob || !b->advantages || !b->returns || !b->old_mu || !b->old_sigma) return GR_ERR_NULL;
  if (b->use_clipped_value_loss && !b->old_values) return GR_ERR_NULL;
  if (rows <= 0) return GR_ERR_SIZE;
  if ((reinterpret_cast<uintptr_t>(b->mu) | reinterpret_cast<uintptr_t>(b->actions) | reinterpret_cast<uintptr_t>(b->old_mu) | reinterpret_cast<uintptr_t>(b->old_sigma) |
       reinterpret_cast<uintptr_t>(b->sigma) | reinterpret_cast<uintptr_t>(grad_mu) | reinterpret_cast<uintptr_t>(grad_value)) & 15u)
    return GR_ERR_ALIGN;
  ppo_loss_grad_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*b, rows, grad_mu, grad_value, sums);
  return (int)cudaGetLastError();
}

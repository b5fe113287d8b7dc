// rollout.cu -- rsl_rl PPO rollout storage on sm_100a: fused add_transitions (+ time-out bootstrap),
// GAE + advantage normalisation in two launches, fused mini-batch gather.
//
// Reference path replaced (S = standalone): S/rsl_rl/ext/storage/rollout_storage.py:71-88 (9 copy_ launches
// per step), :113-127 (~8T+6 launches of [N,1] elementwise ops), :152-191 (9 fancy-index gathers per
// mini-batch) and S/rsl_rl/ext/algorithms/ppo.py:85-97.  All of it is HBM-bound copy / scan work: one thread
// per env for the time scan (coalesced across envs at every t), 128-bit vector copies for the rows.
#ifndef GR_CPU_EMUL
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#endif
#include <stdint.h>
#include "../../include/gracing.h"

namespace gr {

// ---------------------------------------------------------------------------------------------
// add_transitions: one launch; a thread copies one float4 of one of the row-shaped fields, the first N
// threads additionally handle the scalar columns (reward with bootstrap, done, value, log-prob).
// ---------------------------------------------------------------------------------------------
// Row widths that are multiples of 4 floats (the racing task: 16 / 16 / 4) move as 128-bit words (V = 4); any other width
// (the 17-wide reach-target observation) moves as scalars (V = 1).
template <int V> struct RowVec { using type = float4; };
template <> struct RowVec<1> { using type = float; };

template <int V>
__device__ __forceinline__ void copy_rows(const float* __restrict__ src, float* __restrict__ dst, int64_t k4) {
  using VT = typename RowVec<V>::type;
  reinterpret_cast<VT*>(dst)[k4] = __ldcs(reinterpret_cast<const VT*>(src) + k4);
}
#define copy_rows copy_rows<V>

template <int V>
__global__ void storage_add_kernel(const GrStorage s, const GrTransition tr, const int step) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t N = s.N;
  const int64_t n_obs4 = N * s.obs_dim / V, n_cri4 = s.critic_obs ? N * s.critic_dim / V : 0, n_act4 = N * s.act_dim / V;
  int64_t k = tid;
  if (k < n_obs4) { copy_rows(tr.obs, s.obs + (int64_t)step * N * s.obs_dim, k); }
  else if ((k -= n_obs4) < n_cri4) { copy_rows(tr.critic_obs, s.critic_obs + (int64_t)step * N * s.critic_dim, k); }
  else if ((k -= n_cri4) < n_act4) { copy_rows(tr.actions, s.actions + (int64_t)step * N * s.act_dim, k); }
  else if ((k -= n_act4) < n_act4) { if (s.mu) copy_rows(tr.mu, s.mu + (int64_t)step * N * s.act_dim, k); }
  else if ((k -= n_act4) < n_act4) { if (s.sigma) copy_rows(tr.sigma, s.sigma + (int64_t)step * N * s.act_dim, k); }
  else if ((k -= n_act4) < N) {
    const int64_t n = k;
    const float v = tr.values ? tr.values[n] : 0.0f;
    float r = tr.rewards[n];
    // ppo.py:89-92: rewards += gamma * squeeze(values * time_outs.unsqueeze(1), 1)
    if (tr.time_outs) r += tr.gamma * (v * (tr.time_outs[n] ? 1.0f : 0.0f));
    s.rewards[(int64_t)step * N + n] = r;
    const bool d = tr.dones_is_int64 ? (reinterpret_cast<const int64_t*>(tr.dones)[n] != 0) : (reinterpret_cast<const uint8_t*>(tr.dones)[n] != 0);
    s.dones[(int64_t)step * N + n] = d ? 1 : 0;
    if (s.values) s.values[(int64_t)step * N + n] = v;
    if (s.log_prob) s.log_prob[(int64_t)step * N + n] = tr.log_prob[n];
  }
}
#undef copy_rows

// ---------------------------------------------------------------------------------------------
// GAE: thread per env scans t = T-1..0 (rollout_storage.py:113-123), writes returns and raw advantages and
// accumulates (count, mean, M2) in fp64 (Chan/Welford merge: warp shuffle -> block -> one partial per block).
// ---------------------------------------------------------------------------------------------
struct Moments { double n, mean, m2; };
__device__ __forceinline__ Moments merge(Moments a, Moments b) {
  if (b.n == 0.0) return a;
  if (a.n == 0.0) return b;
  const double n = a.n + b.n, d = b.mean - a.mean;
  return Moments{n, a.mean + d * (b.n / n), a.m2 + b.m2 + d * d * (a.n * b.n / n)};
}
__device__ __forceinline__ Moments warp_merge(Moments m) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Moments b{__shfl_down_sync(0xffffffffu, m.n, o), __shfl_down_sync(0xffffffffu, m.mean, o), __shfl_down_sync(0xffffffffu, m.m2, o)};
    m = merge(m, b);
  }
  return m;
}
__device__ __forceinline__ Moments block_merge(Moments m, Moments* sh) {
  m = warp_merge(m);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) sh[wid] = m;
  __syncthreads();
  if (wid == 0) {
    m = lane < nw ? sh[lane] : Moments{0.0, 0.0, 0.0};
    m = warp_merge(m);
  }
  return m;   // valid in thread 0
}

constexpr int kGaeBlock = 128;
constexpr int kGaeChunk = 8;

__global__ void __launch_bounds__(kGaeBlock) gae_kernel(const GrStorage s, const float* __restrict__ last_values, const float gamma, const float lam,
                                                        double* __restrict__ partials) {
  __shared__ Moments sh[kGaeBlock / 32];
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t N = s.N;
  double cnt = 0.0, s1 = 0.0, s2 = 0.0;
  if (n < N) {
    float next_value = last_values[n];
    float adv = 0.0f;
    // the recurrence is serial in t but its loads are not: fetch kGaeChunk steps at once (independent loads in flight),
    // then run the scan on registers -- T dependent DRAM latencies become T / kGaeChunk
    for (int t0 = s.T - 1; t0 >= 0; t0 -= kGaeChunk) {
      float r[kGaeChunk], v[kGaeChunk];
      uint8_t d[kGaeChunk];
#pragma unroll
      for (int j = 0; j < kGaeChunk; ++j) {
        const int t = t0 - j;
        if (t >= 0) {
          const int64_t k = (int64_t)t * N + n;
          r[j] = __ldcs(s.rewards + k); v[j] = __ldcs(s.values + k); d[j] = __ldcs(s.dones + k);
        }
      }
#pragma unroll
      for (int j = 0; j < kGaeChunk; ++j) {
        const int t = t0 - j;
        if (t >= 0) {
          const int64_t k = (int64_t)t * N + n;
          const float not_term = 1.0f - (float)d[j];
          const float delta = r[j] + not_term * gamma * next_value - v[j];
          adv = delta + not_term * gamma * lam * adv;
          const float ret = adv + v[j];
          s.returns[k] = ret;
          const float a = ret - v[j];          // rollout_storage.py:126 (returns - values, not `adv`)
          s.advantages[k] = a;
          cnt += 1.0; s1 += (double)a; s2 += (double)a * (double)a;     // fp64 power sums: no division in the scan
          next_value = v[j];
        }
      }
    }
  }
  // per-thread power sums -> (count, mean, M2), then the Chan merge over the block (warp shuffles, one partial per block)
  Moments mom{cnt, cnt > 0.0 ? s1 / cnt : 0.0, cnt > 0.0 ? s2 - s1 * s1 / cnt : 0.0};
  mom = block_merge(mom, sh);
  if (threadIdx.x == 0) { partials[3 * blockIdx.x] = mom.n; partials[3 * blockIdx.x + 1] = mom.mean; partials[3 * blockIdx.x + 2] = mom.m2; }
}

// merge the per-block partials (single block), optionally export them
__global__ void gae_moments_kernel(const double* __restrict__ partials, const int num_partials, double* __restrict__ moments) {
  __shared__ Moments sh[32];
  Moments m{0.0, 0.0, 0.0};
  for (int k = threadIdx.x; k < num_partials; k += blockDim.x) m = merge(m, Moments{partials[3 * k], partials[3 * k + 1], partials[3 * k + 2]});
  m = block_merge(m, sh);
  if (threadIdx.x == 0) { moments[0] = m.n; moments[1] = m.mean; moments[2] = m.m2; }
}

// (adv - mean) / (std_unbiased + 1e-8)   (rollout_storage.py:127)
__global__ void adv_normalize_kernel(float* __restrict__ adv, const int64_t total, const double* __restrict__ moments) {
  const double n = moments[0], mean = moments[1], m2 = moments[2];
  const float mu = (float)mean;
  const float sd = (float)sqrt(m2 / (n - 1.0));
  const float inv_den = sd + 1e-8f;
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t base = i4 * 4;
  if (base + 3 < total && ((reinterpret_cast<uintptr_t>(adv) & 15u) == 0)) {
    float4 a = reinterpret_cast<float4*>(adv)[i4];
    a.x = (a.x - mu) / inv_den; a.y = (a.y - mu) / inv_den; a.z = (a.z - mu) / inv_den; a.w = (a.w - mu) / inv_den;
    reinterpret_cast<float4*>(adv)[i4] = a;
  } else {
    for (int64_t k = base; k < total && k < base + 4; ++k) adv[k] = (adv[k] - mu) / inv_den;
  }
}

#ifndef GR_CPU_EMUL
// compute_returns as ONE cooperative launch (the three kernels above are ~5 us of work spread over three dependent launches):
// phase 1 = gae_kernel's scan and block partial, grid-wide barrier, phase 2 = every block merges the (few hundred) partials itself,
// phase 3 = each thread normalises the T advantages of its own env (just written: L2 hits).  Needs all blocks co-resident; the host
// side falls back to the three launches when they are not (more than ~2,000 blocks of 128 envs).
__global__ void __launch_bounds__(kGaeBlock) gae_fused_kernel(const GrStorage s, const float* __restrict__ last_values, const float gamma, const float lam,
                                                              double* __restrict__ partials, double* __restrict__ moments, const int normalize) {
  __shared__ Moments sh[kGaeBlock / 32];
  __shared__ Moments total;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t N = s.N;
  double cnt = 0.0, s1 = 0.0, s2 = 0.0;
  if (n < N) {
    float next_value = last_values[n];
    float adv = 0.0f;
    for (int t0 = s.T - 1; t0 >= 0; t0 -= kGaeChunk) {
      float r[kGaeChunk], v[kGaeChunk];
      uint8_t d[kGaeChunk];
#pragma unroll
      for (int j = 0; j < kGaeChunk; ++j) {
        const int t = t0 - j;
        if (t >= 0) {
          const int64_t k = (int64_t)t * N + n;
          r[j] = __ldcs(s.rewards + k); v[j] = __ldcs(s.values + k); d[j] = __ldcs(s.dones + k);
        }
      }
#pragma unroll
      for (int j = 0; j < kGaeChunk; ++j) {
        const int t = t0 - j;
        if (t >= 0) {
          const int64_t k = (int64_t)t * N + n;
          const float not_term = 1.0f - (float)d[j];
          const float delta = r[j] + not_term * gamma * next_value - v[j];
          adv = delta + not_term * gamma * lam * adv;
          const float ret = adv + v[j];
          s.returns[k] = ret;
          const float a = ret - v[j];
          s.advantages[k] = a;
          cnt += 1.0; s1 += (double)a; s2 += (double)a * (double)a;
          next_value = v[j];
        }
      }
    }
  }
  Moments mom{cnt, cnt > 0.0 ? s1 / cnt : 0.0, cnt > 0.0 ? s2 - s1 * s1 / cnt : 0.0};
  mom = block_merge(mom, sh);
  if (threadIdx.x == 0) { partials[3 * blockIdx.x] = mom.n; partials[3 * blockIdx.x + 1] = mom.mean; partials[3 * blockIdx.x + 2] = mom.m2; }
  cooperative_groups::this_grid().sync();
  Moments m{0.0, 0.0, 0.0};
  for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) m = merge(m, Moments{__ldcg(partials + 3 * k), __ldcg(partials + 3 * k + 1), __ldcg(partials + 3 * k + 2)});
  __syncthreads();                                   // (sh is reused)
  m = block_merge(m, sh);
  if (threadIdx.x == 0) {
    total = m;
    if (blockIdx.x == 0) { moments[0] = m.n; moments[1] = m.mean; moments[2] = m.m2; }
  }
  __syncthreads();
  if (!normalize || n >= N) return;
  const float mu = (float)total.mean;
  const float inv_den = (float)sqrt(total.m2 / (total.n - 1.0)) + 1e-8f;
  for (int t = 0; t < s.T; ++t) {
    const int64_t k = (int64_t)t * N + n;
    s.advantages[k] = (__ldcg(s.advantages + k) - mu) / inv_den;
  }
}
#endif

// ---------------------------------------------------------------------------------------------
// mini-batch gather: one launch for the nine fields; a thread moves one float4 of a row-shaped field
// or one element of the five scalar fields.
// ---------------------------------------------------------------------------------------------
template <int V>
__global__ void storage_gather_kernel(const GrStorage s, const int64_t* __restrict__ idx, const int B, const GrMiniBatch o) {
  using float4 = typename RowVec<V>::type;          // (the row word of this instantiation)
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int o4 = s.obs_dim / V, c4 = s.critic_obs ? s.critic_dim / V : 0, a4 = s.act_dim / V;
  const int per_row = o4 + c4 + 3 * a4 + 1;
  const int64_t b = tid / per_row;
  if (b >= B) return;
  int k = (int)(tid - b * per_row);
  const int64_t r = __ldg(idx + b);
  if (k < o4) { reinterpret_cast<float4*>(o.obs)[b * o4 + k] = __ldg(reinterpret_cast<const float4*>(s.obs) + r * o4 + k); return; }
  k -= o4;
  if (k < c4) { reinterpret_cast<float4*>(o.critic_obs)[b * c4 + k] = __ldg(reinterpret_cast<const float4*>(s.critic_obs) + r * c4 + k); return; }
  k -= c4;
  if (k < a4) { reinterpret_cast<float4*>(o.actions)[b * a4 + k] = __ldg(reinterpret_cast<const float4*>(s.actions) + r * a4 + k); return; }
  k -= a4;
  if (k < a4) { reinterpret_cast<float4*>(o.mu)[b * a4 + k] = __ldg(reinterpret_cast<const float4*>(s.mu) + r * a4 + k); return; }
  k -= a4;
  if (k < a4) { reinterpret_cast<float4*>(o.sigma)[b * a4 + k] = __ldg(reinterpret_cast<const float4*>(s.sigma) + r * a4 + k); return; }
  o.values[b] = __ldg(s.values + r);
  o.advantages[b] = __ldg(s.advantages + r);
  o.returns[b] = __ldg(s.returns + r);
  o.log_prob[b] = __ldg(s.log_prob + r);
}

// ---------------------------------------------------------------------------------------------
// transition records: the columns one PPO mini-batch row needs, side by side (GR_RECORD_FLOATS = 48 floats = 192 B per transition):
//   [0,16) policy obs | [16,32) critic obs | [32,36) action | [36,40) old mean | [40,44) old std | 44 old log-prob | 45 advantage |
//   46 return | 47 old value
// Packed once per PPO iteration (after compute_returns); the update kernels then read ONE 192-byte record per sampled row
// instead of nine scattered columns.  One thread per (record, 16-byte chunk).
// perm (optional, [rows] int64): record r holds transition perm[r] -- the mini-batch permutation of rollout_storage.py:165 applied ONCE per
// iteration (the reference reuses one permutation for every epoch), so that mini-batch i of every epoch is the CONTIGUOUS record range
// [i * mb, (i + 1) * mb): the update kernels stream their rows instead of gathering them.
// ---------------------------------------------------------------------------------------------
__global__ void storage_pack_records_kernel(const GrStorage s, const int64_t* __restrict__ perm, float4* __restrict__ rec, const int64_t rows) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t d = tid / 12;
  if (d >= rows) return;
  const int c = (int)(tid - d * 12);
  const int64_t r = perm ? __ldg(perm + d) : d;
  float4 v;
  if (c < 4) v = __ldg(reinterpret_cast<const float4*>(s.obs) + r * 4 + c);
  else if (c < 8) v = __ldg(reinterpret_cast<const float4*>(s.critic_obs ? s.critic_obs : s.obs) + r * 4 + (c - 4));
  else if (c == 8) v = __ldg(reinterpret_cast<const float4*>(s.actions) + r);
  else if (c == 9) v = __ldg(reinterpret_cast<const float4*>(s.mu) + r);
  else if (c == 10) v = __ldg(reinterpret_cast<const float4*>(s.sigma) + r);
  else v = make_float4(__ldg(s.log_prob + r), __ldg(s.advantages + r), __ldg(s.returns + r), __ldg(s.values + r));
  rec[d * 12 + c] = v;
}

}  // namespace gr

#ifndef GR_CPU_EMUL   // host API (the CPU emulation harness in tests/emul includes only the device code)
using namespace gr;

static inline bool mis16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

static inline bool rows_vec4(const GrStorage* s) { return !((s->obs_dim & 3) || (s->act_dim & 3) || (s->critic_obs && (s->critic_dim & 3))); }

static int check_storage(const GrStorage* s) {
  if (!s || !s->obs || !s->actions || !s->rewards || !s->dones) return GR_ERR_NULL;
  if (s->T <= 0 || s->N <= 0 || s->obs_dim <= 0 || s->act_dim <= 0) return GR_ERR_SIZE;
  if (mis16(s->obs) || mis16(s->actions) || (s->critic_obs && mis16(s->critic_obs)) || (s->mu && mis16(s->mu)) || (s->sigma && mis16(s->sigma)))
    return GR_ERR_ALIGN;
  return GR_OK;
}

extern "C" int gr_storage_add(const GrStorage* s, const GrTransition* tr, int32_t step, void* stream) {
  int rc = check_storage(s);
  if (rc != GR_OK) return rc;
  if (!tr || !tr->obs || !tr->actions || !tr->rewards || !tr->dones) return GR_ERR_NULL;
  if (s->critic_obs && !tr->critic_obs) return GR_ERR_NULL;
  if ((s->values && !tr->values) || (s->log_prob && !tr->log_prob) || (s->mu && !tr->mu) || (s->sigma && !tr->sigma)) return GR_ERR_NULL;
  if (tr->time_outs && !tr->values) return GR_ERR_NULL;
  if (step < 0 || step >= s->T) return GR_ERR_SIZE;
  if (mis16(tr->obs) || mis16(tr->actions) || (tr->critic_obs && mis16(tr->critic_obs)) || (tr->mu && mis16(tr->mu)) || (tr->sigma && mis16(tr->sigma)))
    return GR_ERR_ALIGN;
  const int64_t N = s->N;
  const int V = rows_vec4(s) ? 4 : 1;
  const int64_t total = N * s->obs_dim / V + (s->critic_obs ? N * s->critic_dim / V : 0) + 3 * (N * s->act_dim / V) + N;
  if (V == 4) storage_add_kernel<4><<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*s, *tr, step);
  else storage_add_kernel<1><<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*s, *tr, step);
  return (int)cudaGetLastError();
}

extern "C" int64_t gr_gae_scratch_bytes(int32_t N) {
  const int64_t blocks = ((int64_t)N + kGaeBlock - 1) / kGaeBlock;
  return (blocks * 3 + 3) * (int64_t)sizeof(double);
}

extern "C" int gr_advantage_normalize(const GrStorage* s, const double* moments, void* stream) {
  if (!s || !s->advantages || !moments) return GR_ERR_NULL;
  const int64_t total = (int64_t)s->T * s->N;
  if (total <= 0) return GR_ERR_SIZE;
  const int64_t n4 = (total + 3) / 4;
  adv_normalize_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(s->advantages, total, moments);
  return (int)cudaGetLastError();
}

extern "C" int gr_compute_returns(const GrStorage* s, const float* last_values, float gamma, float lam, void* scratch, double* moments,
                                  int32_t normalize, void* stream) {
  if (!s || !s->rewards || !s->dones || !s->values || !s->returns || !s->advantages || !last_values || !scratch) return GR_ERR_NULL;
  if (s->T <= 0 || s->N <= 0) return GR_ERR_SIZE;
  if (reinterpret_cast<uintptr_t>(scratch) & 7u) return GR_ERR_ALIGN;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = (s->N + kGaeBlock - 1) / kGaeBlock;
  double* partials = reinterpret_cast<double*>(scratch);
  double* mom = moments ? moments : partials + 3 * (int64_t)blocks;
  // one cooperative launch when every block can be resident at once (always at the sizes of BASELINE's configs), else three launches
  static int resident_blocks = -1;
  if (resident_blocks < 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gae_fused_kernel, kGaeBlock, 0);
    resident_blocks = sms * per_sm;
  }
  if (blocks <= resident_blocks) {
    GrStorage sv = *s;
    int norm = normalize;
    void* args[] = {&sv, (void*)&last_values, &gamma, &lam, &partials, &mom, &norm};
    return (int)cudaLaunchCooperativeKernel((void*)gae_fused_kernel, dim3((unsigned)blocks), dim3(kGaeBlock), args, 0, st);
  }
  gae_kernel<<<blocks, kGaeBlock, 0, st>>>(*s, last_values, gamma, lam, partials);
  gae_moments_kernel<<<1, 256, 0, st>>>(partials, blocks, mom);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  if (normalize) return gr_advantage_normalize(s, mom, stream);
  return GR_OK;
}

static int pack_records_impl(const GrStorage* s, const int64_t* perm, int64_t num, float* records, void* stream) {
  int rc = check_storage(s);
  if (rc != GR_OK) return rc;
  if (!records || !s->values || !s->advantages || !s->returns || !s->log_prob || !s->mu || !s->sigma) return GR_ERR_NULL;
  if (s->obs_dim != 16 || s->act_dim != 4 || (s->critic_obs && s->critic_dim != 16)) return GR_ERR_SIZE;
  if (mis16(records)) return GR_ERR_ALIGN;
  const int64_t rows = perm ? num : (int64_t)s->T * s->N;
  if (rows < 1 || rows > (int64_t)s->T * s->N) return GR_ERR_SIZE;
  storage_pack_records_kernel<<<(unsigned)((rows * 12 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*s, perm, reinterpret_cast<float4*>(records), rows);
  return (int)cudaGetLastError();
}

extern "C" int gr_storage_pack_records(const GrStorage* s, float* records, void* stream) { return pack_records_impl(s, nullptr, 0, records, stream); }

extern "C" int gr_storage_pack_records_permuted(const GrStorage* s, const int64_t* perm, int64_t num, float* records, void* stream) {
  if (!perm) return GR_ERR_NULL;
  return pack_records_impl(s, perm, num, records, stream);
}

extern "C" int gr_storage_gather(const GrStorage* s, const int64_t* indices, int32_t B, const GrMiniBatch* out, void* stream) {
  int rc = check_storage(s);
  if (rc != GR_OK) return rc;
  if (!indices || !out || !out->obs || !out->actions || !out->values || !out->advantages || !out->returns || !out->log_prob || !out->mu || !out->sigma)
    return GR_ERR_NULL;
  if (!s->values || !s->advantages || !s->returns || !s->log_prob || !s->mu || !s->sigma) return GR_ERR_NULL;
  if (s->critic_obs && !out->critic_obs) return GR_ERR_NULL;
  if (B <= 0) return GR_ERR_SIZE;
  if (mis16(out->obs) || mis16(out->actions) || mis16(out->mu) || mis16(out->sigma) || (out->critic_obs && mis16(out->critic_obs))) return GR_ERR_ALIGN;
  const int V = rows_vec4(s) ? 4 : 1;
  const int per_row = s->obs_dim / V + (s->critic_obs ? s->critic_dim / V : 0) + 3 * (s->act_dim / V) + 1;
  const int64_t total = (int64_t)B * per_row;
  if (V == 4) storage_gather_kernel<4><<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*s, indices, B, *out);
  else storage_gather_kernel<1><<<(unsigned)((total + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*s, indices, B, *out);
  return (int)cudaGetLastError();
}
#endif  // GR_CPU_EMUL

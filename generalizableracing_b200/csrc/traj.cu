// traj.cu -- trajectory split / pad / unpad for recurrent mini-batches.
//
// Replaces, for RolloutStorage.reccurent_mini_batch_generator (S/rsl_rl/ext/storage/rollout_storage.py:194-254):
//   * rsl_rl.utils.split_and_pad_trajectories (third party rsl-rl-lib 2.x, not vendored; call sites :197-199): split every env's
//     [T] column at its dones (the last step always ends a trajectory), order the pieces env-major, pad each to T rows
//     -> padded [T, J, D] + masks [T, J];
//   * the boolean-mask gather of the saved hidden states at trajectory starts (:226-237);
//   * rsl_rl.utils.unpad_trajectories (used by the recurrent policy's Memory in batch mode): the inverse scatter.
// The reference does this with nonzero / tolist / torch.split / pad_sequence (a host round trip and J small tensors).
// Here: one thread per env counts its trajectories, a single-block scan turns counts into env-major trajectory offsets,
// one thread per env writes (env, start, length) of its trajectories, and output-driven kernels move the rows
// (coalesced along the row, zeros written in the same pass).  HBM-bound byte work; integer results bit-exact.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gracing.h"

namespace gr {

__global__ void traj_count_kernel(const uint8_t* __restrict__ dones, const int T, const int N, int32_t* __restrict__ counts) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int c = 1;                                            // dones[-1] = 1: the window end closes a trajectory
  for (int t = 0; t < T - 1; ++t) c += dones[(int64_t)t * N + n] != 0;
  counts[n] = c;
}

// out[0] = 0, out[i+1] = in[0] + ... + in[i]; one block of 1024 threads walks the array in chunks (n <= a few 100 k)
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int32_t* __restrict__ in, const int n, int32_t* __restrict__ out) {
  __shared__ int32_t warp_tot[32];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) { carry = 0; out[0] = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    int32_t v = i < n ? in[i] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int32_t u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    if (wid == 0) {
      int32_t w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int32_t u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
      warp_tot[lane] = w;
    }
    __syncthreads();
    const int32_t incl = v + (wid ? warp_tot[wid - 1] : 0) + carry;
    if (i < n) out[i + 1] = incl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
}

__global__ void traj_fill_kernel(const uint8_t* __restrict__ dones, const int T, const int N, const int32_t* __restrict__ offsets,
                                 int32_t* __restrict__ traj_env, int32_t* __restrict__ traj_start, int32_t* __restrict__ traj_len) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int j = offsets[n], start = 0;
  for (int t = 0; t < T; ++t) {
    if (t == T - 1 || dones[(int64_t)t * N + n] != 0) {
      traj_env[j] = n; traj_start[j] = start; traj_len[j] = t + 1 - start;
      ++j; start = t + 1;
    }
  }
}

// padded[t', j - first, :] = t' < len_j ? src[start_j + t', env_j, :] : 0 ; masks[t', j - first] = t' < len_j
__global__ void traj_pad_kernel(const float* __restrict__ src, const int T, const int N, const int D, const int32_t* __restrict__ traj_env,
                                const int32_t* __restrict__ traj_start, const int32_t* __restrict__ traj_len, const int first, const int count,
                                float* __restrict__ padded, uint8_t* __restrict__ masks) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)T * count * D;
  if (tid >= total) return;
  const int d = (int)(tid % D);
  const int64_t r = tid / D;
  const int jl = (int)(r % count), tp = (int)(r / count);
  const int j = first + jl;
  const bool valid = tp < __ldg(traj_len + j);
  padded[tid] = valid ? __ldg(src + ((int64_t)(__ldg(traj_start + j) + tp) * N + __ldg(traj_env + j)) * D + d) : 0.0f;
  if (masks && d == 0) masks[r] = valid ? 1 : 0;
}

// lengths from masks (column sums), one thread per trajectory
__global__ void traj_len_from_masks_kernel(const uint8_t* __restrict__ masks, const int T, const int J, int32_t* __restrict__ len) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= J) return;
  int c = 0;
  for (int t = 0; t < T; ++t) c += masks[(int64_t)t * J + j] != 0;
  len[j] = c;
}

// unpad_trajectories: the valid rows, concatenated trajectory after trajectory, refill [B, T] env-major:
// out[f % T, f / T, :] = padded[t', j, :] with f = cum_j + t'
__global__ void traj_unpad_kernel(const float* __restrict__ padded, const int32_t* __restrict__ len, const int32_t* __restrict__ cum, const int T,
                                  const int J, const int D, const int B, float* __restrict__ out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)T * J * D;
  if (tid >= total) return;
  const int d = (int)(tid % D);
  const int64_t r = tid / D;
  const int j = (int)(r % J), tp = (int)(r / J);
  if (tp >= __ldg(len + j)) return;
  const int f = __ldg(cum + j) + tp;
  const int env = f / T, t = f - env * T;
  if (env < B) out[((int64_t)t * B + env) * D + d] = padded[tid];
}

// hidden states at trajectory starts: out[l, j - first, h] = saved[start_j, l, env_j, h]   (saved: [T, L, N, H])
__global__ void traj_hidden_kernel(const float* __restrict__ saved, const int L, const int N, const int H, const int32_t* __restrict__ traj_env,
                                   const int32_t* __restrict__ traj_start, const int first, const int count, float* __restrict__ out) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)L * count * H;
  if (tid >= total) return;
  const int h = (int)(tid % H);
  const int64_t r = tid / H;
  const int jl = (int)(r % count), l = (int)(r / count);
  const int j = first + jl;
  out[tid] = __ldg(saved + (((int64_t)__ldg(traj_start + j) * L + l) * N + __ldg(traj_env + j)) * H + h);
}

}  // namespace gr

using namespace gr;

static inline unsigned grid_for(int64_t total, int block) { return (unsigned)((total + block - 1) / block); }

extern "C" int gr_traj_index(const uint8_t* dones, int32_t T, int32_t N, int32_t* offsets, int32_t* traj_env, int32_t* traj_start, int32_t* traj_len,
                             void* stream) {
  if (!dones || !offsets || !traj_env || !traj_start || !traj_len) return GR_ERR_NULL;
  if (T <= 0 || N <= 0) return GR_ERR_SIZE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // counts are staged in traj_len[0..N) (rewritten by the fill kernel only after the scan has consumed them: N <= number of trajectories)
  traj_count_kernel<<<grid_for(N, 128), 128, 0, s>>>(dones, T, N, traj_len);
  exclusive_scan_kernel<<<1, 1024, 0, s>>>(traj_len, N, offsets);
  traj_fill_kernel<<<grid_for(N, 128), 128, 0, s>>>(dones, T, N, offsets, traj_env, traj_start, traj_len);
  return (int)cudaGetLastError();
}

extern "C" int gr_traj_pad(const float* src, int32_t T, int32_t N, int32_t D, const int32_t* traj_env, const int32_t* traj_start, const int32_t* traj_len,
                           int32_t first, int32_t count, float* padded, uint8_t* masks, void* stream) {
  if (!src || !traj_env || !traj_start || !traj_len || !padded) return GR_ERR_NULL;
  if (T <= 0 || N <= 0 || D <= 0 || first < 0 || count <= 0) return GR_ERR_SIZE;
  traj_pad_kernel<<<grid_for((int64_t)T * count * D, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, T, N, D, traj_env, traj_start, traj_len, first,
                                                                                                             count, padded, masks);
  return (int)cudaGetLastError();
}

extern "C" int gr_traj_unpad(const float* padded, const uint8_t* masks, int32_t T, int32_t J, int32_t D, int32_t B, int32_t* scratch, float* out, void* stream) {
  if (!padded || !masks || !scratch || !out) return GR_ERR_NULL;
  if (T <= 0 || J <= 0 || D <= 0 || B <= 0) return GR_ERR_SIZE;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  int32_t* len = scratch;               // [J]
  int32_t* cum = scratch + J;           // [J + 1]
  traj_len_from_masks_kernel<<<grid_for(J, 128), 128, 0, s>>>(masks, T, J, len);
  exclusive_scan_kernel<<<1, 1024, 0, s>>>(len, J, cum);
  traj_unpad_kernel<<<grid_for((int64_t)T * J * D, 256), 256, 0, s>>>(padded, len, cum, T, J, D, B, out);
  return (int)cudaGetLastError();
}

extern "C" int gr_traj_hidden(const float* saved, int32_t T, int32_t L, int32_t N, int32_t H, const int32_t* traj_env, const int32_t* traj_start, int32_t first,
                              int32_t count, float* out, void* stream) {
  if (!saved || !traj_env || !traj_start || !out) return GR_ERR_NULL;
  if (T <= 0 || L <= 0 || N <= 0 || H <= 0 || first < 0 || count <= 0) return GR_ERR_SIZE;
  traj_hidden_kernel<<<grid_for((int64_t)L * count * H, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(saved, L, N, H, traj_env, traj_start, first, count, out);
  return (int)cudaGetLastError();
}

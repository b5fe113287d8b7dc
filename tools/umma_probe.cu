// umma_probe.cu -- hardware check of the tcgen05 conventions ppo_collect.cu relies on (run once on a B200):
//   * shared-memory matrix descriptor for a K-major, non-swizzled operand stored as [K/8 chunks][rows][8 halfs]
//     (LBO = byte stride between K chunks, SBO = byte stride between 8-row groups = 128),
//   * instruction descriptor for kind::f16 (fp16 x fp16 -> fp32), M = 128, N = 128 / 16,
//   * accumulator row i -> TMEM lane i, column j -> TMEM column j; tcgen05.ld.32x32b readback by the warp owning the lane quarter,
//   * commit -> mbarrier -> wait round trip, two accumulating K steps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu ; run: tools/umma_probe
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../generalizableracing_b200/csrc/umma.cuh"

using namespace gr::umma;

template <int N>
__global__ void __launch_bounds__(128) probe_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ D,
                                                    const int K, const int swap_offsets) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sA = smem;                       // [K/8][128][8] halfs
  uint8_t* sB = smem + (size_t)K * 128 * 2; // [K/8][N][8] halfs
  // operands arrive row-major [rows][K]; each thread copies its row(s) into the chunked layout
  for (int c = 0; c < K / 8; ++c) {
    *reinterpret_cast<uint4*>(sA + (size_t)c * 128 * 16 + tid * 16) = *reinterpret_cast<const uint4*>(A + (size_t)tid * K + c * 8);
    if (tid < N) *reinterpret_cast<uint4*>(sB + (size_t)c * N * 16 + tid * 16) = *reinterpret_cast<const uint4*>(B + (size_t)tid * K + c * 8);
  }
  if (tid == 0) mbar_init(&bar, 1);
  if (warp == 0) tmem_alloc(&tmem_base_slot, 128);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t lboA = 128 * 16, lboB = N * 16, sbo = 128;
    for (int kk = 0; kk < K / 16; ++kk) {
      const uint64_t da = swap_offsets ? make_smem_desc(smem_u32(sA) + kk * 2 * lboA, sbo, lboA) : make_smem_desc(smem_u32(sA) + kk * 2 * lboA, lboA, sbo);
      const uint64_t db = swap_offsets ? make_smem_desc(smem_u32(sB) + kk * 2 * lboB, sbo, lboB) : make_smem_desc(smem_u32(sB) + kk * 2 * lboB, lboB, sbo);
      mma_f16_ss(tmem, da, db, idesc, kk > 0);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld_x16(taddr + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

// MN-major operands: the SAME per-row layout ([chunk of 8 columns][row][8 halfs], a thread writes its own row) read
// "transposed": D[m][n] = sum over rows r of P[r][m] * Q[r][n]  (weight-gradient GEMMs: the row dimension is K).
//   A = P as MN-major: 8 m contiguous (16 B), K = rows at 16 B stride inside a core matrix; LBO = 128 B (next 8 rows),
//   SBO = R*16 B (next 8 columns); same for B = Q.
template <int M, int N>
__global__ void __launch_bounds__(128) probe_mn_kernel(const __half* __restrict__ P, const __half* __restrict__ Q, float* __restrict__ D, const int R) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sP = smem;                              // [M/8][R][8]
  uint8_t* sQ = smem + (size_t)M * R * 2;          // [N/8][R][8]
  for (int r = tid; r < R; r += 128) {
    for (int c = 0; c < M / 8; ++c) *reinterpret_cast<uint4*>(sP + (size_t)c * R * 16 + r * 16) = *reinterpret_cast<const uint4*>(P + (size_t)r * M + c * 8);
    for (int c = 0; c < N / 8; ++c) *reinterpret_cast<uint4*>(sQ + (size_t)c * R * 16 + r * 16) = *reinterpret_cast<const uint4*>(Q + (size_t)r * N + c * 8);
  }
  if (tid == 0) mbar_init(&bar, 1);
  if (warp == 0) tmem_alloc(&tmem_base_slot, 256);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_base_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(M, N) | (1u << 15) | (1u << 16);          // A and B MN-major
    for (int kk = 0; kk < R / 16; ++kk) {           // 16 rows (K) per instruction = two 8-row core-matrix groups: +256 B
      const uint64_t da = make_smem_desc(smem_u32(sP) + kk * 256, 128, R * 16);
      const uint64_t db = make_smem_desc(smem_u32(sQ) + kk * 256, 128, R * 16);
      mma_f16_ss(tmem, da, db, idesc, kk > 0);
    }
    tc_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after_sync();
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t r[16];
    tmem_ld_x16(taddr + c0, r);
    tmem_ld_wait();
    if (tid < M) for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

template <int M, int N>
static double run_mn(int R) {
  std::vector<__half> hP((size_t)R * M), hQ((size_t)R * N);
  std::vector<float> ref((size_t)M * N), out((size_t)M * N);
  srand(11 + R + N);
  for (auto* v : {&hP, &hQ}) for (auto& x : *v) x = __float2half((float)(rand() % 2001 - 1000) / 1000.0f);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int r = 0; r < R; ++r) s += (double)__half2float(hP[(size_t)r * M + m]) * __half2float(hQ[(size_t)r * N + n]); ref[(size_t)m * N + n] = (float)s; }
  __half *dP, *dQ; float* dD;
  cudaMalloc(&dP, hP.size() * 2); cudaMalloc(&dQ, hQ.size() * 2); cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dP, hP.data(), hP.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dQ, hQ.data(), hQ.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, out.size() * 4);
  const size_t smem = (size_t)R * (M + N) * 2;
  cudaFuncSetAttribute(probe_mn_kernel<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_mn_kernel<M, N><<<1, 128, smem>>>(dP, dQ, dD, R);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  CUDA error (MN-major probe): %s\n", cudaGetErrorString(e)); exit(2); }
  cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  double err = 0;
  for (size_t i = 0; i < out.size(); ++i) err = fmax(err, fabs((double)out[i] - ref[i]));
  cudaFree(dP); cudaFree(dQ); cudaFree(dD);
  return err;
}

template <int N>
static double run(int K, int swap_offsets) {
  std::vector<__half> hA(128 * K), hB(N * K);
  std::vector<float> fA(128 * K), fB(N * K), ref(128 * N), out(128 * N);
  srand(7 + K + N);
  for (auto* v : {&hA, &hB}) for (auto& x : *v) x = __float2half((float)(rand() % 2001 - 1000) / 1000.0f);
  for (int i = 0; i < 128 * K; ++i) fA[i] = __half2float(hA[i]);
  for (int i = 0; i < N * K; ++i) fB[i] = __half2float(hB[i]);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k]; ref[m * N + n] = (float)s; }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, out.size() * 4);
  const size_t smem = (size_t)K * 128 * 2 + (size_t)K * N * 2;
  cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<N><<<1, 128, smem>>>(dA, dB, dD, K, swap_offsets);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("  CUDA error: %s\n", cudaGetErrorString(e)); exit(2); }
  cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  double err = 0;
  for (size_t i = 0; i < out.size(); ++i) err = fmax(err, fabs((double)out[i] - ref[i]));
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return err;
}

int main() {
  int bad = 0;
  for (int swap_offsets = 0; swap_offsets < 1; ++swap_offsets) {      // (the swapped reading of LBO/SBO faults: measured once, kept out)
    const double e1 = run<128>(32, swap_offsets), e2 = run<128>(128, swap_offsets), e3 = run<16>(128, swap_offsets);
    printf("%s: max|err| N=128,K=32: %.3g   N=128,K=128: %.3g   N=16,K=128: %.3g\n",
           swap_offsets ? "desc(start, SBO-first)" : "desc(start, LBO=chunk stride, SBO=128)", e1, e2, e3);
    if (!swap_offsets && (e1 > 1e-3 || e2 > 1e-3 || e3 > 1e-3)) bad = 1;
  }
  {
    const double m1 = run_mn<128, 16>(128), m2 = run_mn<128, 256>(128), m3 = run_mn<128, 32>(128);
    printf("MN-major A and B (K = rows): max|err| M=128,N=16: %.3g   M=128,N=256: %.3g   M=128,N=32: %.3g\n", m1, m2, m3);
    if (m1 > 1e-3 || m2 > 1e-3 || m3 > 1e-3) bad = 1;
  }
  printf(bad ? "PROBE FAILED for the convention ppo_collect.cu uses\n" : "PROBE OK\n");
  return bad;
}

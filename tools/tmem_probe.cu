// tmem_probe.cu -- how fast can the warps of one SM read tensor memory back (tcgen05.ld 32x32b)?  The epilogues of every tcgen05 kernel in
// this library (ppo_collect / bptt_collect / policy_forward / actor_backward) read fp32 accumulators of 128 rows x 128..256 columns per
// layer; this probe gives the denominator: bytes per cycle per SM with 1 / 4 / 8 / 16 warps, loads of 16 or 32 columns, 1..4 loads in
// flight.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tmem_probe tools/tmem_probe.cu && tools/tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void ld<16>(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// X columns per load, D loads in flight before the wait
template <int X, int D>
__global__ void __launch_bounds__(512, 1) probe(int iters, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5;
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[D][X];
#pragma unroll
    for (int d = 0; d < D; ++d) ld<X>(base + (uint32_t)(((it * D + d) * X) & 511 & ~(X - 1)), r[d]);
    ld_wait();
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
      for (int k = 0; k < X; ++k) acc ^= r[d][k];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int X, int D>
static void run(int warps, long long* d_cycles, uint32_t* d_sink) {
  const int iters = 2000;
  probe<X, D><<<1, warps * 32>>>(iters, d_cycles, d_sink);
  probe<X, D><<<1, warps * 32>>>(iters, d_cycles, d_sink);
  long long c = 0;
  cudaError_t e = cudaMemcpy(&c, d_cycles, sizeof(c), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return; }
  const double bytes = (double)warps * iters * D * X * 32 * 4;
  printf("{\"cols_per_load\": %d, \"loads_in_flight\": %d, \"warps\": %d, \"cycles_per_round\": %.1f, \"bytes_per_cycle_sm\": %.1f, \"bytes_per_cycle_warp\": %.1f}\n", X, D, warps,
         (double)c / iters, bytes / (double)c, bytes / (double)c / warps);
}

int main() {
  long long* d_cycles; uint32_t* d_sink;
  cudaMalloc(&d_cycles, 1024);
  cudaMalloc(&d_sink, 4096);
  const int ws[] = {1, 2, 4, 8, 16};
  for (int w : ws) { run<16, 1>(w, d_cycles, d_sink); run<16, 2>(w, d_cycles, d_sink); run<16, 4>(w, d_cycles, d_sink); run<32, 1>(w, d_cycles, d_sink); run<32, 2>(w, d_cycles, d_sink); }
  return 0;
}

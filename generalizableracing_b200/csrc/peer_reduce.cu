// peer_reduce.cu -- the policy-gradient all-reduce of an env-sharded PPO / BPTT step as ONE kernel over NVLink peer memory.
//
// Reference path replaced: none in the reference (it trains on one GPU); BASELINE.json config C5 shards the envs over 8 GPUs and sums the
// policy gradients once per optimiser step -- `torch.distributed.all_reduce(flat)` over NCCL in algorithms/ppo.py.  The buffer is small
// (38,040 floats = 152 KB: 13 parameter gradients + 16 loss / KL sums) and the step around it is a captured graph of a few 10-200 us
// kernels, so the collective is pure latency: NCCL takes 16.4 us at 2 GPUs and 33.7 us at 8 (profiles/r2_bench_8gpu_late.json), 20 times
// per PPO iteration.  Here every rank READS the other ranks' buffers directly (NVSwitch: every peer at full bandwidth, 1 MB per rank per
// step) between two flag barriers:
//   A  "my gradients are complete"  -- each rank stores the launch's epoch into slot [A][rank] of EVERY rank's flag pad (st.release.sys)
//                                      and waits until its own pad shows the epoch in all slots (ld.acquire.sys).  The gradients were
//                                      written by earlier kernels of the same stream, the release orders them before the flag.
//   sum                              -- out[i] = sum over ranks r = 0..R-1 (FIXED order: every rank computes the same bits, the replicas
//                                      stay identical) of peer[r][i], 16 bytes per load
//   B  "I have read everybody"      -- the last block to finish stores the epoch into [B][rank] of every pad and waits for all ranks: when
//                                      the kernel ends no peer is still reading this rank's buffer, so the stream may overwrite it.
// The buffers live in symmetric memory (torch.distributed._symmetric_memory: cuMem allocations mapped into every rank; plumbing only --
// `buffer_ptrs` gives the peer addresses).  Waits are bounded (GrPeerReduce.max_spins): a lost peer sets *error instead of hanging the GPU.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gracing.h"

namespace gr {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// lanes 0..world-1 of one warp: wait until slot r of the local pad shows `epoch`
__device__ __forceinline__ bool wait_slots(const uint32_t* slots, int world, uint32_t epoch, int64_t max_spins, int lane) {
  bool ok = true;
  if (lane < world) {
    int64_t spins = 0;
    while (ld_acquire_sys(slots + lane) != epoch) {
      if (++spins > max_spins) { ok = false; break; }
      __nanosleep(32);
    }
  }
  return __all_sync(0xffffffffu, ok);
}

__global__ void __launch_bounds__(256) peer_allreduce_kernel(const GrPeerReduce a, float* __restrict__ out) {
  const int world = a.world, rank = a.rank;
  const uint64_t* bufs = reinterpret_cast<const uint64_t*>(a.peer_bufs);
  const uint64_t* pads = reinterpret_cast<const uint64_t*>(a.peer_flags);
  uint32_t* my_pad = reinterpret_cast<uint32_t*>(pads[rank]);             // [2][GR_PEER_MAX_WORLD]
  const uint32_t epoch = *a.epoch + 1u;                                   // (written back by the last block, after everybody has read it)
  __shared__ int s_last;
  // ---- A: announce (block 0), then every block waits on the LOCAL pad
  if (blockIdx.x == 0 && threadIdx.x < world) st_release_sys(reinterpret_cast<uint32_t*>(pads[threadIdx.x]) + rank, epoch);
  if (threadIdx.x < 32) {
    const bool ok = wait_slots(my_pad, world, epoch, a.max_spins, threadIdx.x);
    if (!ok && threadIdx.x == 0) *a.error = 1;
  }
  __syncthreads();
  // ---- sum, rank order 0..R-1
  const int n4 = a.n >> 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const float4 v = __ldcv(reinterpret_cast<const float4*>(bufs[r]) + i);           // (never cached: the peers rewrite their buffers every step)
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = s;
  }
  // ---- B: the last block of this rank tells everybody, and waits for everybody
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(a.counter, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x < world) st_release_sys(reinterpret_cast<uint32_t*>(pads[threadIdx.x]) + GR_PEER_MAX_WORLD + rank, epoch);
  if (threadIdx.x < 32) {
    const bool ok = wait_slots(my_pad + GR_PEER_MAX_WORLD, world, epoch, a.max_spins, threadIdx.x);
    if (threadIdx.x == 0) {
      if (!ok) *a.error = 1;
      *a.counter = 0u;
      *a.epoch = epoch;
    }
  }
}

}  // namespace gr

using namespace gr;

extern "C" int gr_peer_allreduce(const GrPeerReduce* a, float* out, void* stream) {
  if (!a || !out || !a->peer_bufs || !a->peer_flags || !a->epoch || !a->counter || !a->error) return GR_ERR_NULL;
  if (a->world < 1 || a->world > GR_PEER_MAX_WORLD || a->rank < 0 || a->rank >= a->world || a->n < 4 || (a->n & 3) || a->max_spins < 1) return GR_ERR_SIZE;
  if (reinterpret_cast<uintptr_t>(out) & 15u) return GR_ERR_ALIGN;
  const int n4 = a->n >> 2;
  int blocks = (n4 + 255) / 256;
  blocks = blocks > 64 ? 64 : blocks;               // all blocks must be resident at once (they wait on each other's peers): far below 148 SMs
  peer_allreduce_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*a, out);
  return (int)cudaGetLastError();
}

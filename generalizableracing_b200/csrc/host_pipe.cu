// host_pipe.cu -- env.step() for callers whose buffers live in HOST memory (the e2e boundary of bench.py).
//
// The reference's runners hand the env device tensors (S/rsl_rl/ext/runners/on_policy_runner.py:141-157); a host-side
// consumer (a CPU policy server, a logger, another process) pays a host->device copy of the actions and a device->host
// copy of obs / reward / dones around every ManagerBasedDiffRLEnv.step (L/envs/manager_based_diff_rl_env.py:160-267).
// This pipe keeps those three stages on three streams with `depth` staging slots, so the copies of step t overlap the
// kernel of step t+1 and PCIe runs full duplex:
//
//   h2d stream    : actions(t+1) ------------>|
//   compute stream:      step kernel(t) |---->| step kernel(t+1)
//   d2h stream    :                     obs/reward/dones(t) ----------->
//
// The env state advances in compute-stream order, exactly as with gr_step_fwd.
//
// Wire format: `dones` (int64 for the caller, RslRlVecEnvWrapper's `.long()`) crosses PCIe as the two uint8 masks the kernel writes
// anyway (terminated, time_out: 2 B per env instead of 8) into pinned staging owned by the pipe; gr_host_pipe_wait widens them into
// the caller's int64 / bool buffers on the host.  Contract for the caller: the `action` buffer of a step must stay untouched until
// that step's ticket has been waited for (the host->device copy is asynchronous), and every call must come from one thread.
#include <cuda_runtime.h>
#include <new>
#include "../../include/gracing.h"

struct GrHostPipe {
  int32_t num_envs, depth;
  cudaStream_t compute, h2d, d2h;
  int64_t issued;                    // tickets handed out so far
  struct Slot {
    float* action; float* obs; float* critic; float* reward; uint8_t* terminated; uint8_t* time_out;
    uint8_t* h_masks;                  // pinned host staging: [terminated N][time_out N]
    int64_t* host_dones; uint8_t* host_time_out;      // the caller's buffers of the step in flight (filled by finish())
    cudaEvent_t h2d_done, kernel_done, d2h_done;
    bool busy, widened;
  } slot[GR_HOST_PIPE_MAX_DEPTH];
  void* arena;
  void* host_arena;
};

// after the slot's copies have landed: dones = terminated | time_out as int64, time_outs as bytes, into the caller's buffers
static void finish(GrHostPipe* p, GrHostPipe::Slot& sl) {
  if (sl.widened) return;
  const size_t N = (size_t)p->num_envs;
  const uint8_t* __restrict__ term = sl.h_masks;
  const uint8_t* __restrict__ to = sl.h_masks + N;
  if (sl.host_dones) { int64_t* __restrict__ d = sl.host_dones; for (size_t i = 0; i < N; ++i) d[i] = (int64_t)((term[i] | to[i]) != 0); }
  if (sl.host_time_out) { uint8_t* __restrict__ t = sl.host_time_out; for (size_t i = 0; i < N; ++i) t[i] = to[i]; }
  sl.widened = true;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int gr_host_pipe_create(int32_t num_envs, int32_t depth, void* compute_stream, GrHostPipe** out) {
  if (!out) return GR_ERR_NULL;
  *out = nullptr;
  if (num_envs <= 0 || depth < 1 || depth > GR_HOST_PIPE_MAX_DEPTH) return GR_ERR_SIZE;
  GrHostPipe* p = new (std::nothrow) GrHostPipe();
  if (!p) return (int)cudaErrorMemoryAllocation;
  p->num_envs = num_envs; p->depth = depth; p->issued = 0; p->arena = nullptr; p->host_arena = nullptr;
  p->compute = reinterpret_cast<cudaStream_t>(compute_stream);
  const size_t N = (size_t)num_envs;
  const size_t per_slot = align256(N * 16) + 2 * align256(N * 64) + align256(N * 4) + align256(2 * N);
  cudaError_t e = cudaMalloc(&p->arena, per_slot * depth);
  if (e != cudaSuccess) { delete p; return (int)e; }
  e = cudaHostAlloc(&p->host_arena, align256(2 * N) * depth, cudaHostAllocDefault);
  if (e != cudaSuccess) { cudaFree(p->arena); delete p; return (int)e; }
  e = cudaStreamCreateWithFlags(&p->h2d, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->d2h, cudaStreamNonBlocking);
  if (e != cudaSuccess) { cudaFree(p->arena); cudaFreeHost(p->host_arena); delete p; return (int)e; }
  char* base = static_cast<char*>(p->arena);
  for (int s = 0; s < depth; ++s) {
    GrHostPipe::Slot& sl = p->slot[s];
    char* q = base + per_slot * s;
    sl.action = reinterpret_cast<float*>(q); q += align256(N * 16);
    sl.obs = reinterpret_cast<float*>(q); q += align256(N * 64);
    sl.critic = reinterpret_cast<float*>(q); q += align256(N * 64);
    sl.reward = reinterpret_cast<float*>(q); q += align256(N * 4);
    sl.terminated = reinterpret_cast<uint8_t*>(q);          // the two masks are contiguous: ONE device->host copy
    sl.time_out = sl.terminated + N;
    sl.h_masks = static_cast<uint8_t*>(p->host_arena) + align256(2 * N) * s;
    sl.host_dones = nullptr; sl.host_time_out = nullptr;
    sl.busy = false; sl.widened = true;
    cudaEventCreateWithFlags(&sl.h2d_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&sl.kernel_done, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&sl.d2h_done, cudaEventDisableTiming);
  }
  *out = p;
  return GR_OK;
}

extern "C" int gr_host_pipe_destroy(GrHostPipe* p) {
  if (!p) return GR_ERR_NULL;
  cudaStreamSynchronize(p->h2d);
  cudaStreamSynchronize(p->compute);                 // a step kernel may still read / write the slots' device buffers
  cudaStreamSynchronize(p->d2h);
  for (int s = 0; s < p->depth; ++s) {
    if (p->slot[s].busy) finish(p, p->slot[s]);
    cudaEventDestroy(p->slot[s].h2d_done);
    cudaEventDestroy(p->slot[s].kernel_done);
    cudaEventDestroy(p->slot[s].d2h_done);
  }
  cudaStreamDestroy(p->h2d);
  cudaStreamDestroy(p->d2h);
  cudaFree(p->arena);
  cudaFreeHost(p->host_arena);
  delete p;
  return GR_OK;
}

extern "C" int gr_host_pipe_step(GrHostPipe* p, const GrConfig* cfg, const GrTrack* track, const GrState* st, const GrRandom* rng,
                                 const GrHostStep* host, float* log_accum, int64_t* ticket_out) {
  if (!p || !host || !host->action || !host->obs || !host->reward || !st) return GR_ERR_NULL;
  if (st->num_envs != p->num_envs) return GR_ERR_SIZE;
  const size_t N = (size_t)p->num_envs;
  GrHostPipe::Slot& sl = p->slot[p->issued % p->depth];
  cudaError_t e;
  if (sl.busy) {           // the slot's previous device->host copies must have landed before its buffers are rewritten
    e = cudaEventSynchronize(sl.d2h_done);
    if (e != cudaSuccess) return (int)e;
    finish(p, sl);         // (a caller that never waited for that ticket still gets its dones)
  }
  // stage 1: actions host -> device.  The slot's previous kernel (its reader) finished before d2h_done, waited above.
  e = cudaMemcpyAsync(sl.action, host->action, N * 16, cudaMemcpyHostToDevice, p->h2d);
  if (e != cudaSuccess) return (int)e;
  cudaEventRecord(sl.h2d_done, p->h2d);
  // stage 2: the step kernel, in compute-stream order behind the previous step
  cudaStreamWaitEvent(p->compute, sl.h2d_done, 0);
  GrStepIO io = {};
  io.action = sl.action; io.obs = sl.obs; io.critic_obs = host->critic_obs ? sl.critic : nullptr;
  io.reward = sl.reward; io.terminated = sl.terminated; io.time_out = sl.time_out; io.dones = nullptr;
  io.log_accum = log_accum;
  const int rc = gr_step_fwd(cfg, track, st, rng, &io, p->compute);
  if (rc != GR_OK) return rc;
  cudaEventRecord(sl.kernel_done, p->compute);
  // stage 3: results device -> host
  cudaStreamWaitEvent(p->d2h, sl.kernel_done, 0);
  e = cudaMemcpyAsync(host->obs, sl.obs, N * 64, cudaMemcpyDeviceToHost, p->d2h);
  if (e == cudaSuccess) e = cudaMemcpyAsync(host->reward, sl.reward, N * 4, cudaMemcpyDeviceToHost, p->d2h);
  if (e == cudaSuccess && (host->dones || host->time_out)) e = cudaMemcpyAsync(sl.h_masks, sl.terminated, 2 * N, cudaMemcpyDeviceToHost, p->d2h);
  if (e == cudaSuccess && host->critic_obs) e = cudaMemcpyAsync(host->critic_obs, sl.critic, N * 64, cudaMemcpyDeviceToHost, p->d2h);
  if (e != cudaSuccess) return (int)e;
  cudaEventRecord(sl.d2h_done, p->d2h);
  sl.host_dones = host->dones; sl.host_time_out = reinterpret_cast<uint8_t*>(host->time_out);
  sl.widened = !(host->dones || host->time_out);
  sl.busy = true;
  if (ticket_out) *ticket_out = p->issued;
  p->issued += 1;
  return GR_OK;
}

extern "C" int gr_host_pipe_wait(GrHostPipe* p, int64_t ticket) {
  if (!p) return GR_ERR_NULL;
  if (ticket < 0 || ticket >= p->issued) return GR_ERR_SIZE;
  if (ticket + p->depth < p->issued) return GR_OK;       // slot already recycled: its copies were waited for then
  GrHostPipe::Slot& sl = p->slot[ticket % p->depth];
  const cudaError_t e = cudaEventSynchronize(sl.d2h_done);
  if (e != cudaSuccess) return (int)e;
  finish(p, sl);
  return GR_OK;
}

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
// `dones` reaches the caller either as int64 (RslRlVecEnvWrapper's `.long()`; 8 B per env over PCIe) or as one byte per env
// (GrHostStep.dones_u8, written by the kernel itself).  Measured on the B200 box: widening the byte masks to int64 on the HOST instead
// (a 512 KB store stream per 65,536-env step from the calling thread) costs more than letting the DMA engine carry the 8 bytes
// (108.6 vs 100.8 us per step), so the int64 form is produced on the device.  Contract for the caller: the `action` buffer of a step
// must stay untouched until that step's ticket has been waited for (the host->device copy is asynchronous), one calling thread per pipe.
#include <cuda_runtime.h>
#include <new>
#include "../../include/gracing.h"

struct GrHostPipe {
  int32_t num_envs, depth;
  cudaStream_t compute, h2d, d2h;
  int64_t issued;                    // tickets handed out so far
  struct Slot {
    float* action; float* obs; float* critic; float* reward; uint8_t* terminated; uint8_t* time_out; uint8_t* dones_u8; int64_t* dones;
    cudaEvent_t h2d_done, kernel_done, d2h_done;
    bool busy;
  } slot[GR_HOST_PIPE_MAX_DEPTH];
  void* arena;
};

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int gr_host_pipe_create(int32_t num_envs, int32_t depth, void* compute_stream, GrHostPipe** out) {
  if (!out) return GR_ERR_NULL;
  *out = nullptr;
  if (num_envs <= 0 || depth < 1 || depth > GR_HOST_PIPE_MAX_DEPTH) return GR_ERR_SIZE;
  GrHostPipe* p = new (std::nothrow) GrHostPipe();
  if (!p) return (int)cudaErrorMemoryAllocation;
  p->num_envs = num_envs; p->depth = depth; p->issued = 0; p->arena = nullptr;
  p->compute = reinterpret_cast<cudaStream_t>(compute_stream);
  const size_t N = (size_t)num_envs;
  // obs | reward | int64 dones sit back to back (no padding in between: 68 N is a multiple of 8 for even N), so that a caller whose three
  // host buffers are laid out the same way gets ONE device->host copy per step (measured with tools/pcie_probe.cu on the B200 box:
  // 108.9 -> 101.1 us per 65,536-env step for the same bytes; the copy alone, without the concurrent action upload, takes 90.2 us)
  const size_t per_slot = align256(N * 16) + align256(N * 64 + N * 4 + N * 8 + 8) + align256(N * 64) + 3 * align256(N);
  cudaError_t e = cudaMalloc(&p->arena, per_slot * depth);
  if (e != cudaSuccess) { delete p; return (int)e; }
  e = cudaStreamCreateWithFlags(&p->h2d, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->d2h, cudaStreamNonBlocking);
  if (e != cudaSuccess) { cudaFree(p->arena); delete p; return (int)e; }
  char* base = static_cast<char*>(p->arena);
  for (int s = 0; s < depth; ++s) {
    GrHostPipe::Slot& sl = p->slot[s];
    char* q = base + per_slot * s;
    sl.action = reinterpret_cast<float*>(q); q += align256(N * 16);
    sl.obs = reinterpret_cast<float*>(q);
    sl.reward = reinterpret_cast<float*>(q + N * 64);
    sl.dones = reinterpret_cast<int64_t*>(q + ((N * 68 + 7) & ~(size_t)7));
    q += align256(N * 64 + N * 4 + N * 8 + 8);
    sl.critic = reinterpret_cast<float*>(q); q += align256(N * 64);
    sl.terminated = reinterpret_cast<uint8_t*>(q); q += align256(N);
    sl.time_out = reinterpret_cast<uint8_t*>(q); q += align256(N);
    sl.dones_u8 = reinterpret_cast<uint8_t*>(q);
    sl.busy = false;
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
    cudaEventDestroy(p->slot[s].h2d_done);
    cudaEventDestroy(p->slot[s].kernel_done);
    cudaEventDestroy(p->slot[s].d2h_done);
  }
  cudaStreamDestroy(p->h2d);
  cudaStreamDestroy(p->d2h);
  cudaFree(p->arena);
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
  }
  // stage 1: actions host -> device.  The slot's previous kernel (its reader) finished before d2h_done, waited above.
  e = cudaMemcpyAsync(sl.action, host->action, N * 16, cudaMemcpyHostToDevice, p->h2d);
  if (e != cudaSuccess) return (int)e;
  cudaEventRecord(sl.h2d_done, p->h2d);
  // stage 2: the step kernel, in compute-stream order behind the previous step
  cudaStreamWaitEvent(p->compute, sl.h2d_done, 0);
  GrStepIO io = {};
  io.action = sl.action; io.obs = sl.obs; io.critic_obs = host->critic_obs ? sl.critic : nullptr;
  io.reward = sl.reward; io.terminated = sl.terminated; io.time_out = sl.time_out;
  io.dones = host->dones ? sl.dones : nullptr; io.dones_u8 = host->dones_u8 ? sl.dones_u8 : nullptr;
  io.log_accum = log_accum;
  const int rc = gr_step_fwd(cfg, track, st, rng, &io, p->compute);
  if (rc != GR_OK) return rc;
  cudaEventRecord(sl.kernel_done, p->compute);
  // stage 3: results device -> host
  cudaStreamWaitEvent(p->d2h, sl.kernel_done, 0);
  const char* h_obs = reinterpret_cast<const char*>(host->obs);
  // the caller's buffers mirror the slot AND are one allocation (its promise: a copy may not span separately pinned allocations): one copy
  const bool packed = host->outputs_contiguous != 0 && (N & 1) == 0 && host->dones && reinterpret_cast<const char*>(host->reward) == h_obs + N * 64 &&
                      reinterpret_cast<const char*>(host->dones) == h_obs + N * 68;
  if (packed) {
    e = cudaMemcpyAsync(host->obs, sl.obs, N * 76, cudaMemcpyDeviceToHost, p->d2h);
  } else {
    e = cudaMemcpyAsync(host->obs, sl.obs, N * 64, cudaMemcpyDeviceToHost, p->d2h);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host->reward, sl.reward, N * 4, cudaMemcpyDeviceToHost, p->d2h);
    if (e == cudaSuccess && host->dones) e = cudaMemcpyAsync(host->dones, sl.dones, N * 8, cudaMemcpyDeviceToHost, p->d2h);
  }
  if (e == cudaSuccess && host->dones_u8) e = cudaMemcpyAsync(host->dones_u8, sl.dones_u8, N, cudaMemcpyDeviceToHost, p->d2h);
  if (e == cudaSuccess && host->critic_obs) e = cudaMemcpyAsync(host->critic_obs, sl.critic, N * 64, cudaMemcpyDeviceToHost, p->d2h);
  if (e == cudaSuccess && host->time_out) e = cudaMemcpyAsync(host->time_out, sl.time_out, N, cudaMemcpyDeviceToHost, p->d2h);
  if (e != cudaSuccess) return (int)e;
  cudaEventRecord(sl.d2h_done, p->d2h);
  sl.busy = true;
  if (ticket_out) *ticket_out = p->issued;
  p->issued += 1;
  return GR_OK;
}

extern "C" int gr_host_pipe_wait(GrHostPipe* p, int64_t ticket) {
  if (!p) return GR_ERR_NULL;
  if (ticket < 0 || ticket >= p->issued) return GR_ERR_SIZE;
  if (ticket + p->depth < p->issued) return GR_OK;       // slot already recycled: its copies were waited for then
  return (int)cudaEventSynchronize(p->slot[ticket % p->depth].d2h_done);
}

// What the platform gives the pipe's copies alone: `steps` iterations of the same transfers -- actions host->device on one stream,
// obs / reward / dones device->host on another -- from pinned buffers, no kernel, no dependencies.  Returns the seconds the loop took
// (host clock between two device synchronisations).  Several ranks call it concurrently to find the aggregate limit of a node.
#include <chrono>
extern "C" int gr_host_copy_probe2(int32_t num_envs, int32_t steps, int32_t dones_bytes, int32_t packed, double* seconds_out);
extern "C" int gr_host_copy_probe(int32_t num_envs, int32_t steps, int32_t dones_bytes, double* seconds_out) {
  return gr_host_copy_probe2(num_envs, steps, dones_bytes, 0, seconds_out);
}
// packed != 0: obs | reward | dones as ONE device->host copy per step (what gr_host_pipe_step does for a caller whose buffers are contiguous)
extern "C" int gr_host_copy_probe2(int32_t num_envs, int32_t steps, int32_t dones_bytes, int32_t packed, double* seconds_out) {
  if (!seconds_out) return GR_ERR_NULL;
  if (num_envs <= 0 || steps <= 0 || dones_bytes < 0 || dones_bytes > 8) return GR_ERR_SIZE;
  const size_t N = (size_t)num_envs, in_b = N * 16, out_b[3] = {N * 64, N * 4, N * (size_t)dones_bytes};
  void *h_in = nullptr, *d_in = nullptr, *h_out = nullptr, *d_out = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;
  const size_t out_total = align256(out_b[0]) + align256(out_b[1]) + align256(out_b[2] + 1);
  cudaError_t e = cudaHostAlloc(&h_in, in_b, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaHostAlloc(&h_out, out_total, cudaHostAllocDefault);
  if (e == cudaSuccess) e = cudaMalloc(&d_in, in_b);
  if (e == cudaSuccess) e = cudaMalloc(&d_out, out_total);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking);
  if (e == cudaSuccess) {
    auto loop = [&](int n) {
      for (int t = 0; t < n; ++t) {
        cudaMemcpyAsync(d_in, h_in, in_b, cudaMemcpyHostToDevice, s_in);
        if (packed) {
          cudaMemcpyAsync(h_out, d_out, out_b[0] + out_b[1] + out_b[2], cudaMemcpyDeviceToHost, s_out);
          continue;
        }
        size_t off = 0;
        for (int k = 0; k < 3; ++k) {
          if (out_b[k]) cudaMemcpyAsync(static_cast<char*>(h_out) + off, static_cast<char*>(d_out) + off, out_b[k], cudaMemcpyDeviceToHost, s_out);
          off += align256(out_b[k]);
        }
      }
      cudaStreamSynchronize(s_in);
      return cudaStreamSynchronize(s_out);
    };
    e = loop(3);
    const auto t0 = std::chrono::steady_clock::now();
    if (e == cudaSuccess) e = loop(steps);
    *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  if (s_in) cudaStreamDestroy(s_in);
  if (s_out) cudaStreamDestroy(s_out);
  cudaFree(d_in); cudaFree(d_out); cudaFreeHost(h_in); cudaFreeHost(h_out);
  return (int)e;
}

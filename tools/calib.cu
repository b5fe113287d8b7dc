// Calibration (tool, not product): a kernel with the SAME memory traffic as racing_step_fwd (16 float4 plane loads +
// action in; 9 plane stores + obs/critic rows + scalars out) and ~no compute.  Its time is the floor of this
// access pattern / launch shape on a single wave of 65,536 threads.
#include <cuda_runtime.h>
#include <stdint.h>
extern "C" __global__ void calib_kernel(float4* __restrict__ P, int64_t S, int N, const float4* __restrict__ act, float4* __restrict__ obs,
                                        float4* __restrict__ critic, float* __restrict__ rew, uint8_t* __restrict__ term, uint8_t* __restrict__ to,
                                        int64_t* __restrict__ dones, int nload, int nstore) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  float4 acc = __ldcs(act + i);
  float4 v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) if (k < nload) v[k] = __ldcs(P + (int64_t)k * S + i);
#pragma unroll
  for (int k = 0; k < 16; ++k) if (k < nload) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
  // write back hot planes 0..6 + stats 14,15
  const int wr[9] = {0, 1, 2, 3, 4, 5, 6, 14, 15};
#pragma unroll
  for (int k = 0; k < 9; ++k) if (k < nstore) __stcs(P + (int64_t)wr[k] * S + i, make_float4(v[wr[k] % 16].x + 1e-9f * acc.x, v[wr[k] % 16].y, v[wr[k] % 16].z, v[wr[k] % 16].w));
#pragma unroll
  for (int k = 0; k < 4; ++k) { __stcs(obs + (int64_t)i * 4 + k, acc); __stcs(critic + (int64_t)i * 4 + k, acc); }
  rew[i] = acc.x; term[i] = 0; to[i] = 0; dones[i] = 0;
}
extern "C" int calib_launch(void* P, int64_t S, int N, const void* act, void* obs, void* critic, void* rew, void* term, void* to, void* dones,
                            int block, int nload, int nstore, void* stream) {
  calib_kernel<<<(N + block - 1) / block, block, 0, (cudaStream_t)stream>>>((float4*)P, S, N, (const float4*)act, (float4*)obs, (float4*)critic,
                                                                            (float*)rew, (uint8_t*)term, (uint8_t*)to, (int64_t*)dones, nload, nstore);
  return (int)cudaGetLastError();
}

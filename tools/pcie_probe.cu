// tools/pcie_probe.cu -- what limits the host-buffer step (gr_host_pipe_*): per-step time of the step's transfers alone under different
// copy patterns.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/pcie_probe.cu -o tools/pcie_probe
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void copy_to_host_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main() {
  const size_t N = 65536, in_b = N * 16, o_obs = N * 64, o_rew = N * 4, o_done = N * 8, out_b = o_obs + o_rew + o_done;
  void *h_in, *h_out, *d_in, *d_out;
  cudaHostAlloc(&h_in, in_b, cudaHostAllocDefault);
  cudaHostAlloc(&h_out, out_b, cudaHostAllocMapped);
  cudaMalloc(&d_in, in_b);
  cudaMalloc(&d_out, out_b);
  void* h_out_dev = nullptr;
  cudaHostGetDevicePointer(&h_out_dev, h_out, 0);
  cudaStream_t s_in, s_out;
  cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking);
  const int steps = 2000;
  auto run = [&](const char* name, int mode) {
    for (int rep = 0; rep < 2; ++rep) {
      const int n = rep == 0 ? 50 : steps;
      cudaDeviceSynchronize();
      const auto t0 = std::chrono::steady_clock::now();
      for (int t = 0; t < n; ++t) {
        if (mode != 2 && mode != 4) cudaMemcpyAsync(d_in, h_in, in_b, cudaMemcpyHostToDevice, s_in);
        char *h = static_cast<char*>(h_out), *d = static_cast<char*>(d_out);
        if (mode == 0) {          // three copies (obs, reward, dones)
          cudaMemcpyAsync(h, d, o_obs, cudaMemcpyDeviceToHost, s_out);
          cudaMemcpyAsync(h + o_obs, d + o_obs, o_rew, cudaMemcpyDeviceToHost, s_out);
          cudaMemcpyAsync(h + o_obs + o_rew, d + o_obs + o_rew, o_done, cudaMemcpyDeviceToHost, s_out);
        } else if (mode == 1 || mode == 2) {   // one packed copy
          cudaMemcpyAsync(h, d, out_b, cudaMemcpyDeviceToHost, s_out);
        } else {                  // 3, 4: a kernel stores into mapped host memory
          copy_to_host_kernel<<<148 * 4, 256, 0, s_out>>>(static_cast<const float4*>(d_out), static_cast<float4*>(h_out_dev), out_b / 16);
        }
      }
      cudaStreamSynchronize(s_in);
      cudaStreamSynchronize(s_out);
      const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (rep == 1) printf("%-58s %8.2f us/step  D2H %6.2f GB/s\n", name, s / n * 1e6, out_b / (s / n) / 1e9);
    }
  };
  run("3 D2H copies + H2D (the pipe's pattern)", 0);
  run("1 packed D2H copy + H2D", 1);
  run("1 packed D2H copy alone", 2);
  run("kernel stores to mapped host memory + H2D copy", 3);
  run("kernel stores to mapped host memory alone", 4);
  return 0;
}

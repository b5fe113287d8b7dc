// TEST INFRASTRUCTURE -- CPU emulation shim for the device code in generalizableracing_b200/csrc.
// Compiles the *same* kernel sources with g++ (GR_CPU_EMUL) and runs them one thread per block,
// sequentially, so kernel logic can be debugged against the oracle in a container without a GPU.
// It is never loaded by the product path (which fails loudly without libgracing.so + a GPU).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __restrict__

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct uint4 { unsigned x, y, z, w; };
struct dim3e { unsigned x = 1, y = 1, z = 1; };
static thread_local dim3e threadIdx_, blockIdx_, blockDim_, gridDim_;
#define threadIdx threadIdx_
#define blockIdx blockIdx_
#define blockDim blockDim_
#define gridDim gridDim_

static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline T __ldcv(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * (uint64_t)b) >> 32); }
static inline void sincospif(float a, float* s, float* c) { *s = (float)std::sin(3.14159265358979323846 * (double)a); *c = (float)std::cos(3.14159265358979323846 * (double)a); }
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline bool __any_sync(unsigned, bool p) { return p; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
using std::min;
using std::max;
typedef int cudaError_t;

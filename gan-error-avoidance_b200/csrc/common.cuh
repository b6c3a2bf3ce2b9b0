// Shared helpers of the G-LIS sm_100a kernels: error reporting across the C ABI,
// geometry decoding, warp/block reductions.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "glis_b200.h"

namespace glis {

void set_error(const char* fmt, ...);

#define GLIS_REQUIRE(cond, code, ...)   \
  do {                                  \
    if (!(cond)) {                      \
      ::glis::set_error(__VA_ARGS__);   \
      return (code);                    \
    }                                   \
  } while (0)

// Checks the launch that was just enqueued (no synchronisation).
#define GLIS_CHECK_LAUNCH(what)                                                     \
  do {                                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      ::glis::set_error("%s: CUDA error %s", what, cudaGetErrorString(e__));        \
      return GLIS_E_CUDA;                                                           \
    }                                                                               \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum over a block of NT threads (NT multiple of 32, <= 1024). Result valid in thread 0
// (and broadcast to all when `bcast`).
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* smem /* >= 33 floats */, bool bcast = false) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = (lane < NT / 32) ? smem[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) smem[32] = t;
  }
  if (bcast) {
    __syncthreads();
    return smem[32];
  }
  return (threadIdx.x == 0) ? smem[32] : v;
}

int validate_geom(const glis_geom_t* g, const char* who);

}  // namespace glis

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

// Profiling experiments inside the tcgen05 kernels (GLIS_TC_DEBUG / GLIS_T2_DEBUG / GLIS_WG_DEBUG bit masks: no
// stores, no MMAs, no loads; GLIS_TC_TRACE: a globaltimer log of CTA 0) exist only in an instrumented build
// (`make EXTRA=-DGLIS_TC_INSTRUMENT`): in the release library the branches below are compile-time dead.
#ifdef GLIS_TC_INSTRUMENT
#define TC_DEBUG(P) ((P).debug)
#define TC_TRACE(P) ((P).trace)
#else
#define TC_DEBUG(P) 0
#define TC_TRACE(P) (static_cast<unsigned long long*>(nullptr))
#endif

// ------------------------------------------------------------------ programmatic dependent launch
// A kernel launched through launch_pdl() may start while its predecessor in the stream is still draining: its
// CTAs are scheduled as the predecessor's exit, run their prologue (barrier init, TMEM allocation, descriptor
// prefetch, index tables) and then block in pdl_wait() until the predecessor has COMPLETED and its writes are
// visible.  Rules every such kernel follows: pdl_launch_dependents() first, NO global-memory access (read or
// write) before pdl_wait(), and pdl_wait() on every path — a kernel that skipped it would let ITS dependents run
// ahead of the predecessor.  Both instructions are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel and process, under a mutex (entry points are
// re-entrant: two host threads may make a kernel's first launch at the same time).  Returns a cudaError_t.
cudaError_t ensure_max_dynamic_smem(const void* kernel, int bytes);

// SMs the launch planners of the persistent kernels fill: the device's count minus a reserve (GLIS_RESERVE_SMS /
// glis_set_reserved_sms: left free for a communication library's kernels under data parallelism).
int plan_sms();

int pdl_enabled();   // capi.cu: GLIS_PDL / glis_set_pdl(): 0 off, 1 small kernels only, 2 every kernel

// Mode 1 gives the attribute to SMALL launches only (at most four blocks per SM, modest shared memory): the losses,
// heads, LIS and TPReLU kernels that sit on the critical path between the big contractions.  A big persistent
// kernel scheduled early only parks 200 KB CTAs on SMs the other stream could have used.
static inline bool pdl_applies(dim3 grid, size_t smem) {
  const int mode = pdl_enabled();
  if (mode >= 2) return true;
  if (mode <= 0) return false;
  return (size_t)grid.x * grid.y * grid.z <= 148 * 4 && smem <= 100 * 1024;
}

template <typename... KP, typename... A>
inline cudaError_t launch_pdl(void (*kernel)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_applies(grid, smem) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KP>(args)...);
}

// `kernel<<<grid, block, smem, stream>>>(args)` with the attribute above; a launch error returns GLIS_E_CUDA.
#define GLIS_LAUNCH(kernel, grid, block, smem, st, ...)                                              \
  do {                                                                                               \
    cudaError_t le__ = ::glis::launch_pdl(kernel, grid, block, smem, st, __VA_ARGS__);               \
    if (le__ != cudaSuccess) {                                                                       \
      ::glis::set_error("%s: launch failed: %s", #kernel, cudaGetErrorString(le__));                 \
      return GLIS_E_CUDA;                                                                            \
    }                                                                                                \
  } while (0)

// The attribute for launches that build their own cudaLaunchConfig_t (cluster launches): attr[i] = pdl_attr().
static inline cudaLaunchAttribute pdl_attr() {
  cudaLaunchAttribute a;
  a.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  a.val.programmaticStreamSerializationAllowed = 1;
  return a;
}

}  // namespace glis

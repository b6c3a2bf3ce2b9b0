// Probe: cycles per tcgen05.mma (kind::f16, bf16, M = 128, cta_group::1, both operands in shared memory, K-major
// SWIZZLE_128B) as a function of N, issued back to back by one thread with nothing else running on the SM.
// Answers: is the SS-mode MMA paced by the tensor pipe (128 * N / 256 cycles) or by the operand fetch from shared
// memory ((128 + N) rows of 32 bytes per MMA)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../../gan-error-avoidance_b200/csrc -I../../include -o umma_rate_probe umma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include "sm100.cuh"

using namespace glis::sm100;

__global__ void __launch_bounds__(128, 1)
rate_kernel(int n_mma, int iters, int distinct, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(base + 160 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  fence_proxy_async();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, n_mma, 0, 0);
    const uint64_t da = umma_smem_desc(smem_u32(base), 16, 1024);             // A: 128 rows at offset 0
    const uint64_t db = umma_smem_desc(smem_u32(base + 32 * 1024), 16, 1024);  // B: up to 256 rows at 32 KB
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // `distinct`: walk different 32-byte K slices / tiles so that no operand fetch can be elided
      const uint32_t k = distinct ? (uint32_t)(i & 3) * 2u : 0u;
      const uint32_t tile = distinct ? (uint32_t)((i >> 2) & 1) * (64u * 1024u >> 4) : 0u;
      umma_bf16(tmem, da + k + (tile >> 2), db + k + tile, idesc, 1);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, sizeof(long long) * 148);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4096;
  for (int grid : {1, 148})
    for (int distinct : {0, 1})
      for (int n : {64, 112, 128, 208, 224, 256}) {
        rate_kernel<<<grid, 128, 170 * 1024>>>(n, iters, distinct, dout);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148];
        cudaMemcpy(h, dout, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("grid %3d distinct %d N %3d: %7.1f cycles / MMA   (pipe floor %5.1f, fetch at 64 B/clk %5.1f)\n", grid, distinct, n,
               (double)mx / iters, 128.0 * n / 256, (128.0 + n) * 32 / 64);
      }
  return 0;
}

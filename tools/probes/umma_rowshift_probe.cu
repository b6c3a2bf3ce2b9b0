// Probe: may the B operand of tcgen05.mma (K-major, SWIZZLE_128B) start at an arbitrary ROW of a TMA-written tile?
// A tile of 48 pixel rows x 64 bf16 is loaded once; for every row shift r in 0..15 one accumulator
// D_r = A[128 x 64] * B[r .. r+32)^T is computed with the descriptor's start address advanced by r * 128 bytes and
// the descriptor's base-offset field (bits 49-51) set to 0 (variant 0) or to (start >> 7) & 7 (variant 1).
// Prints, per (variant, r), the max abs error against the host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../../gan-error-avoidance_b200/csrc -I../../include -o probe umma_rowshift_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "sm100.cuh"

using namespace glis::sm100;

constexpr int ROWS_B = 48, N = 32, SHIFTS = 16;
#ifndef KE
#define KE 64
#endif
constexpr int ROWB = KE * 2;   // bytes per shared-memory row: 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out, int variant) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = base;                  // 128 rows x 128 B
  uint8_t* sb = base + 128 * 128;      // 48 rows
  uint64_t* bars = reinterpret_cast<uint64_t*>(sb + 64 * 128);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], 128 * ROWB + ROWS_B * ROWB);
    tma_load_3d(sa, &map_a, &bars[0], 0, 0, 0);
    tma_load_3d(sb, &map_b, &bars[0], 0, 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after_sync();
    const uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    const uint64_t swz = KE == 64 ? 0ull : ((4ull << 61) ^ (2ull << 61));   // SWIZZLE_64B = 4 (helper sets 2)
    const uint64_t da = umma_smem_desc(smem_u32(sa), 16, 8 * ROWB) ^ swz;
    for (int r = 0; r < SHIFTS; ++r) {
      const uint32_t start = smem_u32(sb) + r * ROWB;
      uint64_t db = umma_smem_desc(start, 16, 8 * ROWB) ^ swz;
      if (variant == 1) db |= (uint64_t)((start >> 7) & 7) << 49;
      for (int kk = 0; kk < KE / 16; ++kk) umma_bf16(tmem + r * N, da + 2 * kk, db + 2 * kk, idesc, kk > 0);
    }
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after_sync();
  // warp w reads TMEM lanes 32w .. 32w+31 (= rows of A)
  for (int r = 0; r < SHIFTS; ++r) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + r * N, v);
    tmem_ld_wait();
    for (int j = 0; j < N; ++j) out[((size_t)r * 128 + warp * 32 + lane) * N + j] = __uint_as_float(v[j]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap* m, void* ptr, uint64_t k, uint64_t rows, uint32_t box_rows) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 1;
  cuuint64_t d[3] = {k, rows, 1}, s[2] = {k * 2, k * rows * 2};
  cuuint32_t b[3] = {KE, box_rows, 1}, e[3] = {1, 1, 1};
  return ((EncodeTiledFn)p)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            KE == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

int main() {
  std::vector<__nv_bfloat16> ha(128 * KE), hb(ROWS_B * KE);
  std::vector<float> fa(128 * KE), fb(ROWS_B * KE);
  srand(1);
  for (size_t i = 0; i < ha.size(); ++i) { fa[i] = (float)(rand() % 17 - 8) / 8.f; ha[i] = __float2bfloat16(fa[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (float)(rand() % 13 - 6) / 4.f; hb[i] = __float2bfloat16(fb[i]); }
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, sizeof(float) * SHIFTS * 128 * N);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  if (make_map(&ma, da, KE, 128, 128) || make_map(&mb, db, KE, ROWS_B, ROWS_B)) { printf("tensor map failed\n"); return 1; }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> ho(SHIFTS * 128 * N);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dout, 0, sizeof(float) * ho.size());
    probe_kernel<<<1, 128, 48 * 1024>>>(ma, mb, dout, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(ho.data(), dout, sizeof(float) * ho.size(), cudaMemcpyDeviceToHost);
    for (int r = 0; r < SHIFTS; ++r) {
      double worst = 0;
      for (int m = 0; m < 128; ++m)
        for (int j = 0; j < N; ++j) {
          double ref = 0;
          for (int k = 0; k < KE; ++k) ref += (double)fa[m * KE + k] * fb[(r + j) * KE + k];
          worst = fmax(worst, fabs(ref - ho[((size_t)r * 128 + m) * N + j]));
        }
      printf("KE %d variant %d (base_offset %s) row shift %2d: max abs err %g %s\n", KE, variant, variant ? "set" : "0", r, worst,
             worst < 1e-3 ? "OK" : "WRONG");
    }
  }
  return 0;
}

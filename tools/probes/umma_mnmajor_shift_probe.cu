// Probe for the weight-gradient kernel's shared pixel box: with MN-major operands (a shared-memory row = one pixel's
// 64 contiguous channels, K runs along rows) may the B descriptor (1) start at an arbitrary ROW of a TMA-written
// tile and (2) describe two N-groups that OVERLAP, the second one being the first shifted by one row (leading byte
// offset = 128 B)?  If so, the two taps of a stride-parity class of a 4x4 / stride-2 convolution (pixel shift 1)
// multiply against ONE box in ONE MMA of N = 128.
// A: [KP pixels][128 channels] as two 64-channel groups, B: [KP + 16 pixels][64 channels].  For every start row r,
//   D_r[m][n] = sum_k A[k][m] * B[k + r + (n >= 64)][n % 64].
//   nvcc -gencode arch=compute_100a,code=sm_100a -I../../gan-error-avoidance_b200/csrc -I../../include -o mnprobe umma_mnmajor_shift_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "sm100.cuh"

using namespace glis::sm100;

constexpr int KP = 32, ROWS_B = KP + 16, SHIFTS = 4, N = 128;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t grp = KP * 128;           // one 64-channel group of A
  uint8_t* sa = base;                          // 2 groups
  uint8_t* sb = base + 2 * grp;                // ROWS_B rows
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
    mbar_arrive_expect_tx(&bars[0], 2 * grp + ROWS_B * 128);
    tma_load_3d(sa, &map_a, &bars[0], 0, 0, 0);
    tma_load_3d(sa + grp, &map_a, &bars[0], 64, 0, 0);
    tma_load_3d(sb, &map_b, &bars[0], 0, 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after_sync();
    const uint32_t idesc = umma_idesc_bf16(128, N, 1, 1);                  // both operands MN-major
    const uint64_t da = umma_smem_desc(smem_u32(sa), grp, 1024);           // 64-channel groups `grp` apart
    for (int r = 0; r < SHIFTS; ++r) {
      const uint64_t db = umma_smem_desc(smem_u32(sb) + r * 128, 128, 1024);   // second N-group = first one + 1 row
      for (int k16 = 0; k16 < KP / 16; ++k16) umma_bf16(tmem + r * N, da + 128 * k16, db + 128 * k16, idesc, k16 > 0);
    }
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after_sync();
  for (int r = 0; r < SHIFTS; ++r)
    for (int c = 0; c < N; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + r * N + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[((size_t)r * 128 + warp * 32 + lane) * N + c + j] = __uint_as_float(v[j]);
    }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap* m, void* ptr, uint64_t ch, uint64_t rows, uint32_t box_rows) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 1;
  cuuint64_t d[3] = {ch, rows, 1}, s[2] = {ch * 2, ch * rows * 2};
  cuuint32_t b[3] = {64, box_rows, 1}, e[3] = {1, 1, 1};
  return ((EncodeTiledFn)p)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS;
}

int main() {
  std::vector<__nv_bfloat16> ha(KP * 128), hb(ROWS_B * 64);
  std::vector<float> fa(KP * 128), fb(ROWS_B * 64);
  srand(1);
  for (size_t i = 0; i < ha.size(); ++i) { fa[i] = (float)(rand() % 17 - 8) / 8.f; ha[i] = __float2bfloat16(fa[i]); }
  for (size_t i = 0; i < hb.size(); ++i) { fb[i] = (float)(rand() % 13 - 6) / 4.f; hb[i] = __float2bfloat16(fb[i]); }
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, sizeof(float) * SHIFTS * 128 * N);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  if (make_map(&ma, da, 128, KP, KP) || make_map(&mb, db, 64, ROWS_B, ROWS_B)) { printf("tensor map failed\n"); return 1; }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> ho(SHIFTS * 128 * N);
  cudaMemset(dout, 0, sizeof(float) * ho.size());
  probe_kernel<<<1, 128, 48 * 1024>>>(ma, mb, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(ho.data(), dout, sizeof(float) * ho.size(), cudaMemcpyDeviceToHost);
  for (int r = 0; r < SHIFTS; ++r)
    for (int half = 0; half < 2; ++half) {
      double worst = 0;
      for (int m = 0; m < 128; ++m)
        for (int j = 0; j < 64; ++j) {
          double ref = 0;
          for (int k = 0; k < KP; ++k) ref += (double)fa[k * 128 + m] * fb[(k + r + half) * 64 + j];
          worst = fmax(worst, fabs(ref - ho[((size_t)r * 128 + m) * N + half * 64 + j]));
        }
      printf("MN-major start row %d, N-group %d (row shift %d): max abs err %g %s\n", r, half, r + half, worst,
             worst < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}

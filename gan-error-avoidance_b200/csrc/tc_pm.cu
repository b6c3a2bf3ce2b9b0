// Pixel-major tcgen05 product for the image-side layers:   D[pix, n] = sum_k X[pix, k] * W[n, k]
//
// The 3-channel side of D / R level 0 and G level 0 runs as 1x1 products over unfolded patches
// (csrc/image_side.cu): K = 48 or Cin, N = 64 or 48 output columns, hundreds of thousands of pixels.  One
// k-step of MMA per tile and 8-16 bytes of output per element: these launches are bound by their EPILOGUE
// and by HBM, and in the channel-major kernel (tc_conv.cu: M = channels) half of the TMEM lanes, half of the
// epilogue warps and 4-byte stores are all they get.  Here M = 128 PIXELS (every TMEM lane and all 16
// epilogue warps busy), N = all output columns, the weight matrix stays RESIDENT in shared memory for the
// whole launch, and a thread owns 16 consecutive channels of one pixel: 32-byte (STG.256) stores.
//
// Warp roles as in tc_conv.cu: warp 0 TMA producer (pixel tiles through a ring), warp 1 TMEM owner + MMA
// issuer, warps 2..17 epilogue; persistent CTAs, accumulator double-buffered in TMEM.
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace glis {

using namespace sm100;

int make_bf16_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                  const uint32_t* box);

constexpr int PM_BM = 128;       // pixels per tile (UMMA M)
constexpr int PM_BK = 64;
constexpr int PM_EPI_WARPS = 16;
constexpr int PM_THREADS = 64 + 32 * PM_EPI_WARPS;
constexpr int PM_MAX_STAGES = 6;
constexpr int PM_MAX_N = 256;

struct TcPmParams {
  long long M;        // pixels
  int N;              // output columns (multiple of 16, <= 256)
  int kblocks, passes, stages, tiles, tmem_cols;
  const float* bias; int act; const float* act_a; const float* act_b;
  float* preact; float* out_f32; __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
};

// 256-bit global stores (sm_100: STG.256): a lane writes whole 32-byte sectors.
__device__ __forceinline__ void st_v8(void* p, const uint32_t (&r)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void st_f32x16(float* p, const float (&v)[16]) {
  uint32_t r[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __float_as_uint(v[8 * h + j]);
    st_v8(p + 8 * h, r);
  }
}
__device__ __forceinline__ void st_bf16x16(__nv_bfloat16* p, const __nv_bfloat16 (&v)[16]) {
  uint32_t r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    r[j] = (uint32_t)__bfloat16_as_ushort(v[2 * j]) | ((uint32_t)__bfloat16_as_ushort(v[2 * j + 1]) << 16);
  st_v8(p, r);
}

// 16 consecutive channels of one pixel: bias, activation, stores (32-byte vectors).
template <int ACT, bool PREACT, bool F32, bool PLANES>
__device__ __forceinline__ void pm_store16(const uint32_t (&v)[16], long long off, const float* __restrict__ sb,
                                           const float* __restrict__ sa, const float* __restrict__ st,
                                           const TcPmParams& P, bool have_lo) {
  float y[16], o[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    y[j] = __uint_as_float(v[j]) + sb[j];
    o[j] = y[j];
    if (ACT == GLIS_ACT_TPRELU) { const float t = y[j] - st[j]; o[j] = (t > 0.f ? t : sa[j] * t) + st[j]; }
    if (ACT == GLIS_ACT_SIGMOID) o[j] = 1.f / (1.f + __expf(-y[j]));
  }
  if (PREACT) st_f32x16(P.preact + off, y);
  if (F32) st_f32x16(P.out_f32 + off, o);
  if (PLANES) {
    __nv_bfloat16 h[16], l[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) split_bf16(o[j], h[j], l[j]);
    st_bf16x16(P.out_hi + off, h);
    if (have_lo) st_bf16x16(P.out_lo + off, l);
  }
}

__global__ void __launch_bounds__(PM_THREADS, 1)
tc_pm_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
             const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
             const TcPmParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // [W hi: kblocks x N rows][W lo][stage s: X hi 128 rows, X lo 128 rows] ... barriers, epilogue parameters
  const uint32_t w_tile = (uint32_t)P.N * 128u;                 // one k-block of the weight matrix, one plane
  const uint32_t w_bytes = (uint32_t)P.kblocks * w_tile;        // one plane, all k-blocks
  const uint32_t w_total = ((P.passes == 3 ? 2u : 1u) * w_bytes + 1023u) & ~1023u;
  const uint32_t x_bytes = PM_BM * 128u;                        // one plane of a pixel tile
  const uint32_t stage_bytes = 2 * x_bytes;
  uint8_t* stage0 = base + w_total;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage0 + (size_t)P.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + PM_MAX_STAGES;
  uint64_t* w_bar = bars + 2 * PM_MAX_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * PM_MAX_STAGES + 1;    // [2]
  uint64_t* tmem_empty_bar = bars + 2 * PM_MAX_STAGES + 3;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * PM_MAX_STAGES + 5);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 4);   // [N]
  float* s_a = s_bias + PM_MAX_N;                            // [N] clamped slopes
  float* s_t = s_a + PM_MAX_N;                               // [N] TPReLU translations

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();   // the next kernel's prologue may overlap this one's tail (common.cuh)
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x_hi); tma_prefetch_desc(&map_w_hi);
    if (P.passes == 3) { tma_prefetch_desc(&map_x_lo); tma_prefetch_desc(&map_w_lo); }
    for (int s = 0; s < P.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(w_bar, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], PM_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  pdl_wait();   // first global-memory access below (bias / TPReLU parameters: the optimizer may be the predecessor)
  for (int c = threadIdx.x; c < P.N; c += PM_THREADS) {
    s_bias[c] = P.bias ? __ldg(P.bias + c) : 0.f;
    s_a[c] = P.act == GLIS_ACT_TPRELU ? fminf(fmaxf(__ldg(P.act_a + c), 0.f), 1.f) : 0.f;
    s_t[c] = P.act == GLIS_ACT_TPRELU ? __ldg(P.act_b + c) : 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_stride = (uint32_t)P.tmem_cols / 2;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      // the weight matrix, once
      mbar_arrive_expect_tx(w_bar, (P.passes == 3 ? 2u : 1u) * w_bytes);
      for (int kb = 0; kb < P.kblocks; ++kb) {
        tma_load_3d(base + (size_t)kb * w_tile, &map_w_hi, w_bar, kb * PM_BK, 0, 0);
        if (P.passes == 3) tma_load_3d(base + w_bytes + (size_t)kb * w_tile, &map_w_lo, w_bar, kb * PM_BK, 0, 0);
      }
      int s = 0; uint32_t parity = 0;
      const uint32_t tx = (P.passes == 3 ? 2u : 1u) * x_bytes;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
        for (int kb = 0; kb < P.kblocks; ++kb) {
          mbar_wait(&empty_bar[s], parity ^ 1);
          uint8_t* st = stage0 + (size_t)s * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[s], tx);
          tma_load_3d(st, &map_x_hi, &full_bar[s], kb * PM_BK, tile * PM_BM, 0);
          if (P.passes == 3) tma_load_3d(st + x_bytes, &map_x_lo, &full_bar[s], kb * PM_BK, tile * PM_BM, 0);
          if (++s == P.stages) { s = 0; parity ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(PM_BM, P.N, 0, 0);
      const uint64_t wdesc = umma_smem_desc(smem_u32(base), 16, 1024);
      const uint64_t xdesc = umma_smem_desc(smem_u32(stage0), 16, 1024);
      mbar_wait(w_bar, 0);
      tc_fence_after_sync();
      int s = 0; uint32_t parity = 0;
      uint32_t acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[acc], ((acc_phase >> acc) & 1u) ^ 1u);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * acc_stride;
        uint32_t accumulate = 0;
        for (int kb = 0; kb < P.kblocks; ++kb) {
          mbar_wait(&full_bar[s], parity);
          tc_fence_after_sync();
          const uint64_t dxh = xdesc + (uint64_t)(((uint32_t)s * stage_bytes) >> 4);
          const uint64_t dxl = dxh + (x_bytes >> 4);
          const uint64_t dwh = wdesc + (uint64_t)(((uint32_t)kb * w_tile) >> 4);
          const uint64_t dwl = dwh + (w_bytes >> 4);
          if (P.passes == 3) {
#pragma unroll
            for (int kk = 0; kk < PM_BK / 16; ++kk) {
              umma_bf16(tmem_d, dxh + 2 * kk, dwl + 2 * kk, idesc, accumulate);
              umma_bf16(tmem_d, dxl + 2 * kk, dwh + 2 * kk, idesc, 1);
              umma_bf16(tmem_d, dxh + 2 * kk, dwh + 2 * kk, idesc, 1);
              accumulate = 1;
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < PM_BK / 16; ++kk) {
              umma_bf16(tmem_d, dxh + 2 * kk, dwh + 2 * kk, idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit(&empty_bar[s]);
          if (++s == P.stages) { s = 0; parity ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);
        acc_phase ^= (1u << acc);
        acc ^= 1u;
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const int q = warp & 3;                 // TMEM lane quarter = pixels 32q .. 32q+31 of the tile
    const int part = (warp - 2) >> 2;       // takes every fourth 16-column chunk
    const bool have_lo = P.out_lo != nullptr;
    uint32_t acc = 0, full_phase = 0;
    for (int tile = blockIdx.x; tile < P.tiles; tile += gridDim.x) {
      const long long pix = (long long)tile * PM_BM + q * 32 + lane;
      const bool ok = pix < P.M;
      mbar_wait(&tmem_full_bar[acc], (full_phase >> acc) & 1u);
      tc_fence_after_sync();
      const uint32_t tmem_d = tmem_base + acc * acc_stride + ((uint32_t)(q * 32) << 16);
      for (int c0 = part * 16; c0 < P.N; c0 += 64) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_d + (uint32_t)c0, v);
        tmem_ld_wait();
        if (ok) {
          const long long off = pix * P.N + c0;
          const float* sb = s_bias + c0; const float* sa = s_a + c0; const float* st = s_t + c0;
          const bool pre = P.preact != nullptr, f32 = P.out_f32 != nullptr, pl = P.out_hi != nullptr;
          if (P.act == GLIS_ACT_TPRELU && pre && !f32 && pl) pm_store16<GLIS_ACT_TPRELU, true, false, true>(v, off, sb, sa, st, P, have_lo);
          else if (P.act == GLIS_ACT_TPRELU && !pre && !f32 && pl) pm_store16<GLIS_ACT_TPRELU, false, false, true>(v, off, sb, sa, st, P, have_lo);
          else if (P.act == GLIS_ACT_TPRELU && pre && f32 && pl) pm_store16<GLIS_ACT_TPRELU, true, true, true>(v, off, sb, sa, st, P, have_lo);
          else if (P.act == GLIS_ACT_NONE && !pre && f32 && !pl) pm_store16<GLIS_ACT_NONE, false, true, false>(v, off, sb, sa, st, P, have_lo);
          else {
            // generic combination (rare): element by element
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float y = __uint_as_float(v[j]) + sb[j];
              float o = y;
              if (P.act == GLIS_ACT_TPRELU) { const float t = y - st[j]; o = (t > 0.f ? t : sa[j] * t) + st[j]; }
              else if (P.act == GLIS_ACT_SIGMOID) o = 1.f / (1.f + __expf(-y));
              if (pre) P.preact[off + j] = y;
              if (f32) P.out_f32[off + j] = o;
              if (pl) {
                __nv_bfloat16 h, l;
                split_bf16(o, h, l);
                P.out_hi[off + j] = h;
                if (have_lo) P.out_lo[off + j] = l;
              }
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      full_phase ^= (1u << acc);
      acc ^= 1u;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// ------------------------------------------------------------------ host side
// 1x1, stride 1, no padding, few output columns: the image-side products.
int tc_pm_supported(const glis_geom_t* g) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("GLIS_TC_PM"); off = (e && atoi(e) == 0) ? 1 : 0; }   // GLIS_TC_PM=0: A/B knob
  if (off) return 0;
  if (g->relation != GLIS_CONV || g->KH != 1 || g->KW != 1 || g->stride_h != 1 || g->stride_w != 1) return 0;
  if (g->pad_h != 0 || g->pad_w != 0 || g->dil_h != 1 || g->dil_w != 1) return 0;
  if (g->Hi != g->Ho || g->Wi != g->Wo) return 0;
  if (g->Co % 16 != 0 || g->Co > 64) return 0;          // wider outputs fill the channel-major kernel's lanes
  if (g->Ci % 8 != 0 || g->Ci < 16) return 0;
  const int kblocks = (g->Ci + PM_BK - 1) / PM_BK;
  if ((long long)kblocks * g->Co * 256 > 96 * 1024) return 0;   // the resident weight matrix (hi + lo)
  return 1;
}

int tc_pm_forward(const glis_geom_t* g, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
                  const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, const glis_epilogue_t* ep, float* out_f32,
                  __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int precision, cudaStream_t st) {
  GLIS_REQUIRE(tc_pm_supported(g), GLIS_E_UNSUPPORTED, "glis_conv_forward_bf16(pixel-major): unsupported geometry");
  const int passes = precision == GLIS_PREC_BF16X3 ? 3 : 1;
  GLIS_REQUIRE(x_hi && w_hi && (passes == 1 || (x_lo && w_lo)), GLIS_E_BADARG,
               "glis_conv_forward_bf16: missing hi/lo operand planes");
  TcPmParams P;
  P.M = (long long)g->N * g->Ho * g->Wo;
  P.N = g->Co;
  P.kblocks = (g->Ci + PM_BK - 1) / PM_BK;
  P.passes = passes;
  P.tiles = (int)((P.M + PM_BM - 1) / PM_BM);
  P.tmem_cols = 32;
  while (P.tmem_cols < 2 * P.N) P.tmem_cols *= 2;
  P.bias = ep->bias; P.act = ep->act; P.act_a = ep->act_a; P.act_b = ep->act_b; P.preact = ep->preact;
  P.out_f32 = out_f32; P.out_hi = out_hi; P.out_lo = out_lo;
  GLIS_REQUIRE(ep->act_channels == 0 || ep->act_channels == g->Co, GLIS_E_UNSUPPORTED,
               "glis_conv_forward_bf16(pixel-major): TPReLU parameters must be per output channel");
  const size_t w_total = (((size_t)(passes == 3 ? 2 : 1) * P.kblocks * P.N * 128) + 1023) & ~(size_t)1023;
  const size_t stage_bytes = 2 * (size_t)PM_BM * 128;
  const size_t fixed = 1024 /*alignment*/ + w_total + 256 /*barriers*/ + 3 * PM_MAX_N * sizeof(float);
  int stages = (int)((220 * 1024 - fixed) / stage_bytes);
  if (stages > PM_MAX_STAGES) stages = PM_MAX_STAGES;
  GLIS_REQUIRE(stages >= 2, GLIS_E_UNSUPPORTED, "glis_conv_forward_bf16(pixel-major): does not fit shared memory");
  P.stages = stages;

  CUtensorMap mx_hi, mx_lo, mw_hi, mw_lo;
  {
    // pixels: rows of Ci elements; 3-D (Ci, M, 1) so that one load primitive serves both operands
    const uint64_t dims[3] = {(uint64_t)g->Ci, (uint64_t)P.M, 1};
    const uint64_t strides[2] = {(uint64_t)g->Ci * 2, (uint64_t)g->Ci * 2 * (uint64_t)P.M};
    const uint32_t box[3] = {PM_BK, PM_BM, 1};
    int rc = make_bf16_map(&mx_hi, x_hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_bf16_map(&mx_lo, passes == 3 ? x_lo : x_hi, 3, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)g->Ci, (uint64_t)g->Co, 1};
    const uint64_t strides[2] = {(uint64_t)g->Ci * 2, (uint64_t)g->Ci * g->Co * 2};
    const uint32_t box[3] = {PM_BK, (uint32_t)P.N, 1};
    int rc = make_bf16_map(&mw_hi, w_hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_bf16_map(&mw_lo, passes == 3 ? w_lo : w_hi, 3, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = fixed + (size_t)stages * stage_bytes;
  {
    cudaError_t e = ensure_max_dynamic_smem(reinterpret_cast<const void*>(tc_pm_kernel), 227 * 1024);
    GLIS_REQUIRE(e == cudaSuccess, GLIS_E_CUDA, "cudaFuncSetAttribute(tc_pm_kernel): %s", cudaGetErrorString(e));
  }
  const int num_sms = plan_sms();
  const int grid = P.tiles < num_sms ? P.tiles : num_sms;
  cudaError_t le = launch_pdl(tc_pm_kernel, dim3(grid), dim3(PM_THREADS), smem, st, mx_hi, mx_lo, mw_hi, mw_lo, P);
  GLIS_REQUIRE(le == cudaSuccess, GLIS_E_CUDA, "glis_conv_forward_bf16(pixel-major): launch failed: %s", cudaGetErrorString(le));
  GLIS_CHECK_LAUNCH("glis_conv_forward_bf16(pixel-major)");
  return GLIS_OK;
}

}  // namespace glis

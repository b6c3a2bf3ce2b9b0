// Image-side layers (3 colour channels on one side: D / R level 0, G level 0 and their gradients)
// on the tensor-core kernels.
//
// A 4x4 / stride 2 / pad 1 (transposed) convolution with C <= 4 channels on its image side is a
// plain matrix product once the image side is unfolded into 16*C columns per coarse pixel:
//
//   conv 3 -> Co      : y[pix][co]   = sum_j  P[pix][j] * E[co][j]          P = unfold(x)
//   its weight grad   : G[co][j]     = sum_pix dy[pix][co] * P[pix][j]      (master layout as is)
//   convT Ci -> 3     : cols[pix][j] = sum_ci x[pix][ci] * E[ci][j] ,  y = fold(cols)
//   its data grad     : dx[pix][ci]  = sum_j  unfold(dy)[pix][j] * E[ci][j]
//   its weight grad   : G[ci][j]     = sum_pix x[pix][ci] * unfold(dy)[pix][j]
//
// with j = c*16 + kh*4 + kw — exactly the memory order of the master weights (conv: [co][c][kh][kw],
// transposed: [ci][c][kh][kw]), so E is the effective weight matrix in master order and the weight
// gradients land in master layout without a permutation.  The products themselves are 1x1
// "convolutions" on tc_conv_kernel / tc_wgrad_kernel (K = 48 padded to one 64-wide block by TMA
// zero fill); this file holds the three small pointwise kernels around them:
//   unfold : fp32 NHWC image  -> bf16 hi/lo planes [N, H/2, W/2, 16*C]
//   fold   : fp32 cols [N, Hi, Wi, 16*C] -> fp32 NHWC image [N, 2Hi, 2Wi, C] (+ bias, activation)
//   pack   : master weights -> E (A x J) and E^T (J x A) as K-major bf16 hi/lo packs
#include "common.cuh"
#include "sm100.cuh"

namespace glis {

constexpr int IS_NT = 256;

// One thread per (coarse pixel, kh): reads the 4 fine pixels of its kernel row (4*C contiguous
// floats), writes 4 consecutive columns per channel — the four kh-lanes of a pixel fill whole
// 32-byte sectors of its row.
__global__ void __launch_bounds__(IS_NT)
unfold4x4s2_kernel(const float* __restrict__ x, int N, int H, int W, int C, __nv_bfloat16* __restrict__ hi,
                   __nv_bfloat16* __restrict__ lo) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, J = 16 * C;
  const int64_t total = (int64_t)N * Ho * Wo * 4;
  for (int64_t i = (int64_t)blockIdx.x * IS_NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * IS_NT) {
    const int kh = (int)(i & 3);
    const int64_t pix = i >> 2;
    const int p32 = (int)pix;   // < 2^31 pixels (checked by the host): 32-bit divisions
    const int ox = p32 % Wo; const int t = p32 / Wo; const int oy = t % Ho; const int n = t / Ho;
    const int iy = 2 * oy - 1 + kh;
    const bool row_ok = iy >= 0 && iy < H;
    const float* row = x + ((int64_t)n * H + (row_ok ? iy : 0)) * W * C;
    __nv_bfloat16* dh = hi + pix * J + kh * 4;
    __nv_bfloat16* dl = lo ? lo + pix * J + kh * 4 : nullptr;
    for (int c = 0; c < C; ++c) {
      __align__(8) __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int kw = 0; kw < 4; ++kw) {
        const int ix = 2 * ox - 1 + kw;
        const float v = (row_ok && ix >= 0 && ix < W) ? __ldg(row + (int64_t)ix * C + c) : 0.f;
        sm100::split_bf16(v, h[kw], l[kw]);
      }
      *reinterpret_cast<uint2*>(dh + c * 16) = *reinterpret_cast<const uint2*>(h);
      if (dl) *reinterpret_cast<uint2*>(dl + c * 16) = *reinterpret_cast<const uint2*>(l);
    }
  }
}

// One thread per fine output pixel: the 2x2 taps that reach it, C channels.
__global__ void __launch_bounds__(IS_NT)
fold4x4s2_kernel(const float* __restrict__ cols, int N, int Hi, int Wi, int C, const float* __restrict__ bias, int act,
                 float* __restrict__ out) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int Ho = 2 * Hi, Wo = 2 * Wi, J = 16 * C;
  const int64_t total = (int64_t)N * Ho * Wo;
  for (int64_t pix = (int64_t)blockIdx.x * IS_NT + threadIdx.x; pix < total; pix += (int64_t)gridDim.x * IS_NT) {
    const int p32 = (int)pix;   // < 2^31 pixels (checked by the host): 32-bit divisions
    const int ox = p32 % Wo; const int t = p32 / Wo; const int oy = t % Ho; const int n = t / Ho;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const int kh0 = (oy + 1) & 1, kw0 = (ox + 1) & 1;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int kh = kh0 + 2 * a, iy = (oy + 1 - kh) >> 1;   // oy = 2*iy - 1 + kh
      if (oy + 1 - kh < 0 || iy >= Hi) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int kw = kw0 + 2 * b, ix = (ox + 1 - kw) >> 1;
        if (ox + 1 - kw < 0 || ix >= Wi) continue;
        const float* src = cols + (((int64_t)n * Hi + iy) * Wi + ix) * J + kh * 4 + kw;
        for (int c = 0; c < C; ++c) acc[c] += __ldg(src + c * 16);
      }
    }
    for (int c = 0; c < C; ++c) {
      float y = acc[c] + (bias ? __ldg(bias + c) : 0.f);
      if (act == GLIS_ACT_SIGMOID) y = 1.f / (1.f + expf(-y));
      out[pix * C + c] = y;
    }
  }
}

// E[a][j] = w_master[a*J + j] * scale[o]/norm[o] with o = a (out_axis 0: conv, rows are output channels)
// or o = j / T (out_axis 1: transposed, the column group is the output channel).
__global__ void __launch_bounds__(IS_NT)
pack_matrix_bf16_kernel(const float* __restrict__ w, const float* __restrict__ scale, const float* __restrict__ norm,
                        int out_axis, int A, int J, int T, __nv_bfloat16* __restrict__ e_hi,
                        __nv_bfloat16* __restrict__ e_lo, __nv_bfloat16* __restrict__ et_hi,
                        __nv_bfloat16* __restrict__ et_lo) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int total = A * J;
  for (int i = blockIdx.x * IS_NT + threadIdx.x; i < total; i += gridDim.x * IS_NT) {
    const int a = i / J, j = i - a * J;
    const int o = out_axis == 0 ? a : j / T;
    const float v = __ldg(w + i) * (scale ? __ldg(scale + o) : 1.f) / __ldg(norm + o);
    __nv_bfloat16 h, l;
    sm100::split_bf16(v, h, l);
    if (e_hi) { e_hi[i] = h; if (e_lo) e_lo[i] = l; }
    if (et_hi) { et_hi[(size_t)j * A + a] = h; if (et_lo) et_lo[(size_t)j * A + a] = l; }
  }
}

static int is_blocks(int64_t work) {
  int64_t b = (work + IS_NT - 1) / IS_NT;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace glis

using namespace glis;

extern "C" int glis_unfold4x4s2_bf16(const float* x, int N, int H, int W, int C, void* hi, void* lo, void* stream) {
  GLIS_REQUIRE(x && hi, GLIS_E_BADARG, "glis_unfold4x4s2_bf16: NULL pointer");
  GLIS_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && C <= 4 && H % 2 == 0 && W % 2 == 0, GLIS_E_BADARG,
               "glis_unfold4x4s2_bf16: bad shape (N=%d H=%d W=%d C=%d)", N, H, W, C);
  GLIS_REQUIRE((int64_t)N * H * W < ((int64_t)1 << 31), GLIS_E_UNSUPPORTED, "glis_unfold4x4s2_bf16: more than 2^31 pixels");
  GLIS_LAUNCH(unfold4x4s2_kernel, dim3(is_blocks((int64_t)N * (H / 2) * (W / 2) * 4)), dim3(IS_NT), 0, (cudaStream_t)((cudaStream_t)stream), 
      x, N, H, W, C, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  GLIS_CHECK_LAUNCH("glis_unfold4x4s2_bf16");
  return GLIS_OK;
}

extern "C" int glis_fold4x4s2(const float* cols, int N, int Hi, int Wi, int C, const float* bias, int act, float* out,
                              void* stream) {
  GLIS_REQUIRE(cols && out, GLIS_E_BADARG, "glis_fold4x4s2: NULL pointer");
  GLIS_REQUIRE(N > 0 && Hi > 0 && Wi > 0 && C > 0 && C <= 4, GLIS_E_BADARG, "glis_fold4x4s2: bad shape");
  GLIS_REQUIRE((int64_t)N * Hi * Wi * 4 < ((int64_t)1 << 31), GLIS_E_UNSUPPORTED, "glis_fold4x4s2: more than 2^31 pixels");
  GLIS_REQUIRE(act == GLIS_ACT_NONE || act == GLIS_ACT_SIGMOID, GLIS_E_UNSUPPORTED, "glis_fold4x4s2: activation %d", act);
  GLIS_LAUNCH(fold4x4s2_kernel, dim3(is_blocks((int64_t)N * Hi * Wi * 4)), dim3(IS_NT), 0, (cudaStream_t)((cudaStream_t)stream), cols, N, Hi, Wi, C, bias, act,
                                                                                         out);
  GLIS_CHECK_LAUNCH("glis_fold4x4s2");
  return GLIS_OK;
}

extern "C" int glis_wn_pack_matrix_bf16(const float* w, const float* scale, const float* norm, int out_axis, int A, int J,
                                        int T, void* e_hi, void* e_lo, void* et_hi, void* et_lo, void* stream) {
  GLIS_REQUIRE(w && norm && (e_hi || et_hi), GLIS_E_BADARG, "glis_wn_pack_matrix_bf16: NULL pointer");
  GLIS_REQUIRE(A > 0 && J > 0 && T > 0 && J % T == 0 && (out_axis == 0 || out_axis == 1), GLIS_E_BADARG,
               "glis_wn_pack_matrix_bf16: bad shape");
  GLIS_LAUNCH(pack_matrix_bf16_kernel, dim3(is_blocks((int64_t)A * J)), dim3(IS_NT), 0, (cudaStream_t)((cudaStream_t)stream), 
      w, scale, norm, out_axis, A, J, T, (__nv_bfloat16*)e_hi, (__nv_bfloat16*)e_lo, (__nv_bfloat16*)et_hi,
      (__nv_bfloat16*)et_lo);
  GLIS_CHECK_LAUNCH("glis_wn_pack_matrix_bf16");
  return GLIS_OK;
}

// ---------------------------------------------------------------------------- input augmentation
// The reference augments every training image on the HOST with imgaug (g_lis/main.py:176-231: horizontal flip,
// additive Gaussian noise, brightness multiply, contrast normalisation, affine scale / rotate / translate), one PIL
// image at a time inside its synchronous loader loop.  Here the decoded batch is augmented on the DEVICE in one
// pass: per image a 12-float parameter row (drawn on the host, a few hundred bytes per batch)
//   [a00 a01 a02 a10 a11 a12  mul  alpha  sigma  flip  border  _]
// with (a..) the INVERSE affine map in pixel coordinates (output pixel centre -> source position), bilinear
// sampling, border 0 = constant black, 1 = symmetric reflection; then  v = v * mul;  v = 0.5 + alpha (v - 0.5);
// v += sigma * N(0, 1) (Philox, keyed by seed and the element);  clamp to [0, 1].  NCHW in (what the decoder
// delivers), NHWC out (what the kernels read).
namespace glis {

__device__ __forceinline__ float aug_fetch(const float* __restrict__ img, int H, int W, int y, int x, int border) {
  if (border) {            // symmetric: ... 1 0 | 0 1 2 ... W-1 | W-1 W-2 ...
    const int pw = 2 * W, ph = 2 * H;
    x = ((x % pw) + pw) % pw; if (x >= W) x = pw - 1 - x;
    y = ((y % ph) + ph) % ph; if (y >= H) y = ph - 1 - y;
    return __ldg(img + (size_t)y * W + x);
  }
  return (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(img + (size_t)y * W + x) : 0.f;
}

__device__ __forceinline__ void aug_philox(uint64_t seed, uint64_t ctr, uint32_t (&out)[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0x61756721u, 0u};   // (a stream of its own)
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0], hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}

__global__ void __launch_bounds__(IS_NT)
augment_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ params, int N, int C, int H,
               int W, uint64_t seed) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int64_t total = (int64_t)N * H * W;
  for (int64_t p = (int64_t)blockIdx.x * IS_NT + threadIdx.x; p < total; p += (int64_t)gridDim.x * IS_NT) {
    const int x = (int)(p % W), y = (int)((p / W) % H), n = (int)(p / ((int64_t)W * H));
    const float* q = params + (size_t)n * 12;
    const float fx = q[9] != 0.f ? (float)(W - 1 - x) : (float)x;      // horizontal flip of the OUTPUT
    const float sx = q[0] * fx + q[1] * (float)y + q[2], sy = q[3] * fx + q[4] * (float)y + q[5];
    const float x0f = floorf(sx), y0f = floorf(sy);
    const int x0 = (int)x0f, y0 = (int)y0f, border = q[10] != 0.f;
    const float wx = sx - x0f, wy = sy - y0f;
    uint32_t r[4] = {0, 0, 0, 0};
    if (q[8] > 0.f) aug_philox(seed, (uint64_t)p, r);
    for (int c = 0; c < C; ++c) {
      const float* img = in + ((size_t)n * C + c) * H * W;
      float v = (1.f - wy) * ((1.f - wx) * aug_fetch(img, H, W, y0, x0, border) + wx * aug_fetch(img, H, W, y0, x0 + 1, border)) +
                wy * ((1.f - wx) * aug_fetch(img, H, W, y0 + 1, x0, border) + wx * aug_fetch(img, H, W, y0 + 1, x0 + 1, border));
      v *= q[6];
      v = 0.5f + q[7] * (v - 0.5f);
      if (q[8] > 0.f) {        // one Gaussian per PIXEL, shared by the channels (imgaug per_channel=False)
        const float u1 = 1.f - (float)(r[0] >> 8) * (1.f / 16777216.f), u2 = (float)(r[1] >> 8) * (1.f / 16777216.f);
        float s, co;
        sincospif(2.f * u2, &s, &co);
        v += q[8] * sqrtf(-2.f * logf(u1)) * co;
      }
      out[(size_t)p * C + c] = fminf(fmaxf(v, 0.f), 1.f);
    }
  }
}

}  // namespace glis

extern "C" int glis_augment(const float* in_nchw, float* out_nhwc, const float* params, int N, int C, int H, int W,
                            uint64_t seed, void* stream) {
  using namespace glis;
  GLIS_REQUIRE(in_nchw && out_nhwc && params, GLIS_E_BADARG, "glis_augment: NULL pointer");
  GLIS_REQUIRE(N > 0 && C > 0 && C <= 4 && H > 0 && W > 0, GLIS_E_BADARG, "glis_augment: bad shape (N=%d C=%d H=%d W=%d)", N, C, H, W);
  GLIS_LAUNCH(augment_kernel, dim3(is_blocks((int64_t)N * H * W)), dim3(IS_NT), 0, (cudaStream_t)((cudaStream_t)stream), in_nchw, out_nhwc, params, N, C, H, W, seed);
  GLIS_CHECK_LAUNCH("glis_augment");
  return GLIS_OK;
}

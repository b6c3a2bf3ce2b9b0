// Weight normalisation, forward and backward, on the master parameter layouts.
//   forward : n[o] = sqrt(c*|w[o]|^2 + 1e-6),  w_hat[o] = w[o]*s[o]/n[o], packed for the GEMM kernels
//   backward: dw[o] = (s/n)(G[o] - c w[o] <G[o],w[o]>/n^2),  ds[o] = <G[o],w[o]>/n
// (closed forms: SURVEY.md App. E; reference: WeightNormalizedConv.py:29-49, WeightNormalizedLinear.py:30-39)
#include "common.cuh"
#include "sm100.cuh"

namespace glis {

// element (o, i, t) of a master weight: out_axis 0 -> (o*Cin + i)*T + t ; out_axis 1 -> (i*Cout + o)*T + t
__device__ __forceinline__ int64_t master_index(int out_axis, int Cout, int Cin, int T, int o, int i, int t) {
  return out_axis == 0 ? ((int64_t)o * Cin + i) * T + t : ((int64_t)i * Cout + o) * T + t;
}

constexpr int WN_NT = 256;

// One block per output channel.  NT = 1024 serves the few-channel / long-row heads (Cout = 1,
// R = 12800), where one 256-thread block would be a long serial chain of loads.
template <int NT>
__global__ void __launch_bounds__(NT)
wn_norm_kernel(const float* __restrict__ w, int out_axis, int Cout, int Cin, int T, float c,
               float* __restrict__ norm) {
  __shared__ float red[33];
  const int o = blockIdx.x;
  const int R = Cin * T;
  float ss = 0.f;
  if (out_axis == 0) {
    const float* row = w + (int64_t)o * R;
#pragma unroll 8
    for (int r = threadIdx.x; r < R; r += NT) { const float v = __ldg(row + r); ss = fmaf(v, v, ss); }
  } else {
#pragma unroll 8
    for (int r = threadIdx.x; r < R; r += NT) {
      const int i = r / T, t = r - i * T;
      const float v = __ldg(w + ((int64_t)i * Cout + o) * T + t);
      ss = fmaf(v, v, ss);
    }
  }
  ss = block_sum<NT>(ss, red);
  if (threadIdx.x == 0) norm[o] = sqrtf(ss * c + 1e-6f);
}

// One WARP per output channel, 8 channels per block: the short-row / many-channel layers
// (linears: R = 256, Cout up to 12800), where a block per channel is all launch overhead.
__global__ void __launch_bounds__(WN_NT)
wn_norm_warp_kernel(const float* __restrict__ w, int out_axis, int Cout, int Cin, int T, float c,
                    float* __restrict__ norm) {
  const int o = blockIdx.x * (WN_NT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (o >= Cout) return;
  const int R = Cin * T;
  float ss = 0.f;
#pragma unroll 8
  for (int r = lane; r < R; r += 32) {
    const int i = r / T, t = r - i * T;
    const float v = __ldg(w + master_index(out_axis, Cout, Cin, T, o, i, t));
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) norm[o] = sqrtf(ss * c + 1e-6f);
}

static void launch_wn_norm(const float* w, int out_axis, int Cout, int Cin, int T, float c, float* norm,
                           cudaStream_t st) {
  const int64_t R = (int64_t)Cin * T;
  if (R <= 512 && Cout >= 64)
    wn_norm_warp_kernel<<<(Cout + WN_NT / 32 - 1) / (WN_NT / 32), WN_NT, 0, st>>>(w, out_axis, Cout, Cin, T, c, norm);
  else if (R >= 8192 || (Cout <= 16 && R >= 4096))
    wn_norm_kernel<1024><<<Cout, 1024, 0, st>>>(w, out_axis, Cout, Cin, T, c, norm);
  else if (R >= 4096)
    wn_norm_kernel<512><<<Cout, 512, 0, st>>>(w, out_axis, Cout, Cin, T, c, norm);
  else
    wn_norm_kernel<WN_NT><<<Cout, WN_NT, 0, st>>>(w, out_axis, Cout, Cin, T, c, norm);
}

// One thread per packed element; writes coalesced, reads gathered through L2.
// Pack row o' -> master output channel (identity, or the NHWC feature order of glis_wn_prepare_perm).
__device__ __forceinline__ int master_channel(int o, int perm_c, int perm_p) {
  return perm_c ? (o % perm_c) * perm_p + o / perm_c : o;
}

// IDX = uint32_t whenever the tensor has fewer than 2^31 elements: the per-element index
// decomposition is a chain of divisions, and 64-bit integer division costs ~10x the 32-bit one
// (the kernel was bound by it, not by memory).
template <typename IDX>
__global__ void wn_pack_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                               const float* __restrict__ norm, int out_axis, int Cout, int Cin, int T,
                               float* __restrict__ pack_io, float* __restrict__ pack_oi, int perm_c, int perm_p) {
  const IDX total = (IDX)T * (IDX)Cin * (IDX)Cout;
  for (IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (IDX)gridDim.x * blockDim.x) {
    if (pack_io) {  // [t][i][o]
      const int o = master_channel((int)(e % (IDX)Cout), perm_c, perm_p);
      const IDX r = e / (IDX)Cout; const int i = (int)(r % (IDX)Cin); const int t = (int)(r / (IDX)Cin);
      const float a = (scale ? __ldg(scale + o) : 1.f) / __ldg(norm + o);
      pack_io[e] = __ldg(w + master_index(out_axis, Cout, Cin, T, o, i, t)) * a;
    }
    if (pack_oi) {  // [t][o][i]
      const int i = (int)(e % (IDX)Cin); const IDX r = e / (IDX)Cin;
      const int o = master_channel((int)(r % (IDX)Cout), perm_c, perm_p);
      const int t = (int)(r / (IDX)Cout);
      const float a = (scale ? __ldg(scale + o) : 1.f) / __ldg(norm + o);
      pack_oi[e] = __ldg(w + master_index(out_axis, Cout, Cin, T, o, i, t)) * a;
    }
  }
}

// bf16 hi/lo packs for the tensor-core kernels (K-major operands):
//   fwd [t][o][i] for the launch that reads Cin and writes Cout, bwd [t][i][o] for its data gradient.
template <typename IDX>
__global__ void wn_pack_bf16_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                    const float* __restrict__ norm, int out_axis, int Cout, int Cin, int T,
                                    __nv_bfloat16* __restrict__ fwd_hi, __nv_bfloat16* __restrict__ fwd_lo,
                                    __nv_bfloat16* __restrict__ bwd_hi, __nv_bfloat16* __restrict__ bwd_lo,
                                    int perm_c, int perm_p) {
  const IDX total = (IDX)T * (IDX)Cin * (IDX)Cout;
  for (IDX e = (IDX)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (IDX)gridDim.x * blockDim.x) {
    if (fwd_hi) {  // [t][o][i]
      const int i = (int)(e % (IDX)Cin); const IDX r = e / (IDX)Cin;
      const int o = master_channel((int)(r % (IDX)Cout), perm_c, perm_p);
      const int t = (int)(r / (IDX)Cout);
      const float a = (scale ? __ldg(scale + o) : 1.f) / __ldg(norm + o);
      __nv_bfloat16 h, l;
      sm100::split_bf16(__ldg(w + master_index(out_axis, Cout, Cin, T, o, i, t)) * a, h, l);
      fwd_hi[e] = h;
      if (fwd_lo) fwd_lo[e] = l;
    }
    if (bwd_hi) {  // [t][i][o]
      const int o = master_channel((int)(e % (IDX)Cout), perm_c, perm_p);
      const IDX r = e / (IDX)Cout; const int i = (int)(r % (IDX)Cin); const int t = (int)(r / (IDX)Cin);
      const float a = (scale ? __ldg(scale + o) : 1.f) / __ldg(norm + o);
      __nv_bfloat16 h, l;
      sm100::split_bf16(__ldg(w + master_index(out_axis, Cout, Cin, T, o, i, t)) * a, h, l);
      bwd_hi[e] = h;
      if (bwd_lo) bwd_lo[e] = l;
    }
  }
}

// One block per output channel, at most PER elements of the row per thread: G and w are read ONCE,
// with every load in flight together, and stay in registers between the dot product and the update.
template <int NT, int PER>
__global__ void __launch_bounds__(NT)
wn_project_kernel(const float* __restrict__ G, const float* __restrict__ w, const float* __restrict__ scale,
                  const float* __restrict__ norm, int out_axis, int Cout, int Cin, int T, float c,
                  float* __restrict__ dw, float* __restrict__ dscale, int accumulate) {
  __shared__ float red[33];
  const int o = blockIdx.x;
  const int R = Cin * T;
  float gv[PER], wv[PER], old[PER];
  float dot = 0.f;
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int r = threadIdx.x + u * NT;
    gv[u] = wv[u] = old[u] = 0.f;
    if (r < R) {
      const int i = r / T, t = r - i * T;
      const int64_t idx = master_index(out_axis, Cout, Cin, T, o, i, t);
      gv[u] = __ldg(G + idx);
      wv[u] = __ldg(w + idx);
      if (accumulate) old[u] = dw[idx];
    }
  }
#pragma unroll
  for (int u = 0; u < PER; ++u) dot = fmaf(gv[u], wv[u], dot);
  dot = block_sum<NT>(dot, red, true);
  const float n = __ldg(norm + o), s = scale ? __ldg(scale + o) : 1.f;
  const float a = s / n, k = c * dot / (n * n);
#pragma unroll
  for (int u = 0; u < PER; ++u) {
    const int r = threadIdx.x + u * NT;
    if (r < R) {
      const int i = r / T, t = r - i * T;
      dw[master_index(out_axis, Cout, Cin, T, o, i, t)] = old[u] + a * (gv[u] - k * wv[u]);
    }
  }
  if (dscale && threadIdx.x == 0) dscale[o] = accumulate ? dscale[o] + dot / n : dot / n;
}

// Rows longer than 16 x 1024 elements: two passes over memory.
__global__ void __launch_bounds__(1024)
wn_project_long_kernel(const float* __restrict__ G, const float* __restrict__ w, const float* __restrict__ scale,
                       const float* __restrict__ norm, int out_axis, int Cout, int Cin, int T, float c,
                       float* __restrict__ dw, float* __restrict__ dscale, int accumulate) {
  constexpr int NT = 1024;
  __shared__ float red[33];
  const int o = blockIdx.x;
  const int R = Cin * T;
  float dot = 0.f;
#pragma unroll 4
  for (int r = threadIdx.x; r < R; r += NT) {
    const int i = r / T, t = r - i * T;
    const int64_t idx = master_index(out_axis, Cout, Cin, T, o, i, t);
    dot = fmaf(__ldg(G + idx), __ldg(w + idx), dot);
  }
  dot = block_sum<NT>(dot, red, true);
  const float n = __ldg(norm + o), s = scale ? __ldg(scale + o) : 1.f;
  const float a = s / n, k = c * dot / (n * n);
#pragma unroll 4
  for (int r = threadIdx.x; r < R; r += NT) {
    const int i = r / T, t = r - i * T;
    const int64_t idx = master_index(out_axis, Cout, Cin, T, o, i, t);
    const float v = a * (__ldg(G + idx) - k * __ldg(w + idx));
    dw[idx] = accumulate ? dw[idx] + v : v;
  }
  if (dscale && threadIdx.x == 0) dscale[o] = accumulate ? dscale[o] + dot / n : dot / n;
}

// One warp per output channel (see wn_norm_warp_kernel).
__global__ void __launch_bounds__(WN_NT)
wn_project_warp_kernel(const float* __restrict__ G, const float* __restrict__ w, const float* __restrict__ scale,
                       const float* __restrict__ norm, int out_axis, int Cout, int Cin, int T, float c,
                       float* __restrict__ dw, float* __restrict__ dscale, int accumulate) {
  const int o = blockIdx.x * (WN_NT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (o >= Cout) return;
  const int R = Cin * T;
  float dot = 0.f;
#pragma unroll 8
  for (int r = lane; r < R; r += 32) {
    const int i = r / T, t = r - i * T;
    const int64_t idx = master_index(out_axis, Cout, Cin, T, o, i, t);
    dot = fmaf(__ldg(G + idx), __ldg(w + idx), dot);
  }
  dot = warp_sum(dot);
  const float n = __ldg(norm + o), s = scale ? __ldg(scale + o) : 1.f;
  const float a = s / n, k = c * dot / (n * n);
#pragma unroll 8
  for (int r = lane; r < R; r += 32) {
    const int i = r / T, t = r - i * T;
    const int64_t idx = master_index(out_axis, Cout, Cin, T, o, i, t);
    const float v = a * (__ldg(G + idx) - k * __ldg(w + idx));
    dw[idx] = accumulate ? dw[idx] + v : v;
  }
  if (dscale && lane == 0) dscale[o] = accumulate ? dscale[o] + dot / n : dot / n;
}

}  // namespace glis

using namespace glis;

extern "C" int glis_wn_prepare(const float* w, const float* scale, int out_axis, int Cout, int Cin, int T,
                               float c, float* norm, float* pack_io, float* pack_oi, void* stream) {
  return glis_wn_prepare_perm(w, scale, out_axis, Cout, Cin, T, c, norm, pack_io, pack_oi, 0, 0, stream);
}

static int check_perm(int out_axis, int Cout, int T, int perm_c, int perm_p, const char* who) {
  GLIS_REQUIRE((perm_c == 0 && perm_p == 0) ||
                   (perm_c > 0 && perm_p > 0 && T == 1 && out_axis == 0 && (int64_t)perm_c * perm_p == Cout),
               GLIS_E_BADARG, "%s: bad row permutation (C=%d P=%d for Cout=%d T=%d axis=%d)", who, perm_c, perm_p, Cout, T,
               out_axis);
  return GLIS_OK;
}

extern "C" int glis_wn_prepare_perm(const float* w, const float* scale, int out_axis, int Cout, int Cin, int T,
                                    float c, float* norm, float* pack_io, float* pack_oi, int perm_c, int perm_p,
                                    void* stream) {
  GLIS_REQUIRE(w && norm, GLIS_E_BADARG, "glis_wn_prepare: w/norm is NULL");
  if (int rc = check_perm(out_axis, Cout, T, perm_c, perm_p, "glis_wn_prepare_perm")) return rc;
  GLIS_REQUIRE(Cout > 0 && Cin > 0 && T > 0 && (out_axis == 0 || out_axis == 1), GLIS_E_BADARG,
               "glis_wn_prepare: bad shape (Cout=%d Cin=%d T=%d axis=%d)", Cout, Cin, T, out_axis);
  cudaStream_t st = (cudaStream_t)stream;
  launch_wn_norm(w, out_axis, Cout, Cin, T, c, norm, st);
  GLIS_CHECK_LAUNCH("glis_wn_prepare(norm)");
  if (pack_io || pack_oi) {
    const int64_t total = (int64_t)T * Cin * Cout;
    const int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
    if (total < ((int64_t)1 << 31))
      wn_pack_kernel<uint32_t><<<blocks, 256, 0, st>>>(w, scale, norm, out_axis, Cout, Cin, T, pack_io, pack_oi, perm_c, perm_p);
    else
      wn_pack_kernel<int64_t><<<blocks, 256, 0, st>>>(w, scale, norm, out_axis, Cout, Cin, T, pack_io, pack_oi, perm_c, perm_p);
    GLIS_CHECK_LAUNCH("glis_wn_prepare(pack)");
  }
  return GLIS_OK;
}

extern "C" int glis_wn_prepare_bf16(const float* w, const float* scale, int out_axis, int Cout, int Cin, int T,
                                    float c, float* norm, void* fwd_hi, void* fwd_lo, void* bwd_hi, void* bwd_lo,
                                    void* stream) {
  return glis_wn_prepare_bf16_perm(w, scale, out_axis, Cout, Cin, T, c, norm, fwd_hi, fwd_lo, bwd_hi, bwd_lo, 0, 0, stream);
}

extern "C" int glis_wn_prepare_bf16_perm(const float* w, const float* scale, int out_axis, int Cout, int Cin, int T,
                                         float c, float* norm, void* fwd_hi, void* fwd_lo, void* bwd_hi, void* bwd_lo,
                                         int perm_c, int perm_p, void* stream) {
  GLIS_REQUIRE(w && norm, GLIS_E_BADARG, "glis_wn_prepare_bf16: w/norm is NULL");
  if (int rc = check_perm(out_axis, Cout, T, perm_c, perm_p, "glis_wn_prepare_bf16_perm")) return rc;
  GLIS_REQUIRE(Cout > 0 && Cin > 0 && T > 0 && (out_axis == 0 || out_axis == 1), GLIS_E_BADARG,
               "glis_wn_prepare_bf16: bad shape (Cout=%d Cin=%d T=%d axis=%d)", Cout, Cin, T, out_axis);
  GLIS_REQUIRE((fwd_hi || !fwd_lo) && (bwd_hi || !bwd_lo), GLIS_E_BADARG, "glis_wn_prepare_bf16: lo plane without hi");
  cudaStream_t st = (cudaStream_t)stream;
  launch_wn_norm(w, out_axis, Cout, Cin, T, c, norm, st);
  GLIS_CHECK_LAUNCH("glis_wn_prepare_bf16(norm)");
  if (fwd_hi || bwd_hi) {
    const int64_t total = (int64_t)T * Cin * Cout;
    const int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
    if (total < ((int64_t)1 << 31))
      wn_pack_bf16_kernel<uint32_t><<<blocks, 256, 0, st>>>(w, scale, norm, out_axis, Cout, Cin, T, (__nv_bfloat16*)fwd_hi,
                                                           (__nv_bfloat16*)fwd_lo, (__nv_bfloat16*)bwd_hi,
                                                           (__nv_bfloat16*)bwd_lo, perm_c, perm_p);
    else
      wn_pack_bf16_kernel<int64_t><<<blocks, 256, 0, st>>>(w, scale, norm, out_axis, Cout, Cin, T, (__nv_bfloat16*)fwd_hi,
                                                          (__nv_bfloat16*)fwd_lo, (__nv_bfloat16*)bwd_hi,
                                                          (__nv_bfloat16*)bwd_lo, perm_c, perm_p);
    GLIS_CHECK_LAUNCH("glis_wn_prepare_bf16(pack)");
  }
  return GLIS_OK;
}

extern "C" int glis_wn_project(const float* G, const float* w, const float* scale, const float* norm,
                               int out_axis, int Cout, int Cin, int T, float c, float* dw, float* dscale,
                               int accumulate, void* stream) {
  GLIS_REQUIRE(G && w && norm && dw, GLIS_E_BADARG, "glis_wn_project: NULL pointer");
  GLIS_REQUIRE(Cout > 0 && Cin > 0 && T > 0 && (out_axis == 0 || out_axis == 1), GLIS_E_BADARG,
               "glis_wn_project: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t R = (int64_t)Cin * T;
#define WN_PROJECT_ARGS G, w, scale, norm, out_axis, Cout, Cin, T, c, dw, dscale, accumulate
  if (R <= 512 && Cout >= 64)
    wn_project_warp_kernel<<<(Cout + WN_NT / 32 - 1) / (WN_NT / 32), WN_NT, 0, st>>>(WN_PROJECT_ARGS);
  else if (R <= 8 * 128) wn_project_kernel<128, 8><<<Cout, 128, 0, st>>>(WN_PROJECT_ARGS);
  else if (R <= 16 * 128 && Cout >= 148) wn_project_kernel<128, 16><<<Cout, 128, 0, st>>>(WN_PROJECT_ARGS);
  else if (R <= 8 * 256) wn_project_kernel<256, 8><<<Cout, 256, 0, st>>>(WN_PROJECT_ARGS);
  else if (R <= 16 * 256 && Cout >= 148) wn_project_kernel<256, 16><<<Cout, 256, 0, st>>>(WN_PROJECT_ARGS);
  else if (R <= 8 * 512) wn_project_kernel<512, 8><<<Cout, 512, 0, st>>>(WN_PROJECT_ARGS);
  else if (R <= 16 * 512 && Cout >= 148) wn_project_kernel<512, 16><<<Cout, 512, 0, st>>>(WN_PROJECT_ARGS);
  else if (R <= 8 * 1024) wn_project_kernel<1024, 8><<<Cout, 1024, 0, st>>>(WN_PROJECT_ARGS);
  else if (R <= 16 * 1024) wn_project_kernel<1024, 16><<<Cout, 1024, 0, st>>>(WN_PROJECT_ARGS);
  else wn_project_long_kernel<<<Cout, 1024, 0, st>>>(WN_PROJECT_ARGS);
#undef WN_PROJECT_ARGS
  GLIS_CHECK_LAUNCH("glis_wn_project");
  return GLIS_OK;
}

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

__global__ void __launch_bounds__(WN_NT)
wn_norm_kernel(const float* __restrict__ w, int out_axis, int Cout, int Cin, int T, float c,
               float* __restrict__ norm) {
  __shared__ float red[33];
  const int o = blockIdx.x;
  const int R = Cin * T;
  float ss = 0.f;
  if (out_axis == 0) {
    const float* row = w + (int64_t)o * R;
    for (int r = threadIdx.x; r < R; r += WN_NT) { const float v = __ldg(row + r); ss = fmaf(v, v, ss); }
  } else {
    for (int r = threadIdx.x; r < R; r += WN_NT) {
      const int i = r / T, t = r - i * T;
      const float v = __ldg(w + ((int64_t)i * Cout + o) * T + t);
      ss = fmaf(v, v, ss);
    }
  }
  ss = block_sum<WN_NT>(ss, red);
  if (threadIdx.x == 0) norm[o] = sqrtf(ss * c + 1e-6f);
}

// One thread per packed element; writes coalesced, reads gathered through L2.
__global__ void wn_pack_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                               const float* __restrict__ norm, int out_axis, int Cout, int Cin, int T,
                               float* __restrict__ pack_io, float* __restrict__ pack_oi) {
  const int64_t total = (int64_t)T * Cin * Cout;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    if (pack_io) {  // [t][i][o]
      const int o = (int)(e % Cout); const int64_t r = e / Cout; const int i = (int)(r % Cin); const int t = (int)(r / Cin);
      const float a = (scale ? __ldg(scale + o) : 1.f) / __ldg(norm + o);
      pack_io[e] = __ldg(w + master_index(out_axis, Cout, Cin, T, o, i, t)) * a;
    }
    if (pack_oi) {  // [t][o][i]
      const int i = (int)(e % Cin); const int64_t r = e / Cin; const int o = (int)(r % Cout); const int t = (int)(r / Cout);
      const float a = (scale ? __ldg(scale + o) : 1.f) / __ldg(norm + o);
      pack_oi[e] = __ldg(w + master_index(out_axis, Cout, Cin, T, o, i, t)) * a;
    }
  }
}

// bf16 hi/lo packs for the tensor-core kernels (K-major operands):
//   fwd [t][o][i] for the launch that reads Cin and writes Cout, bwd [t][i][o] for its data gradient.
__global__ void wn_pack_bf16_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                    const float* __restrict__ norm, int out_axis, int Cout, int Cin, int T,
                                    __nv_bfloat16* __restrict__ fwd_hi, __nv_bfloat16* __restrict__ fwd_lo,
                                    __nv_bfloat16* __restrict__ bwd_hi, __nv_bfloat16* __restrict__ bwd_lo) {
  const int64_t total = (int64_t)T * Cin * Cout;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    if (fwd_hi) {  // [t][o][i]
      const int i = (int)(e % Cin); const int64_t r = e / Cin; const int o = (int)(r % Cout); const int t = (int)(r / Cout);
      const float a = (scale ? __ldg(scale + o) : 1.f) / __ldg(norm + o);
      __nv_bfloat16 h, l;
      sm100::split_bf16(__ldg(w + master_index(out_axis, Cout, Cin, T, o, i, t)) * a, h, l);
      fwd_hi[e] = h;
      if (fwd_lo) fwd_lo[e] = l;
    }
    if (bwd_hi) {  // [t][i][o]
      const int o = (int)(e % Cout); const int64_t r = e / Cout; const int i = (int)(r % Cin); const int t = (int)(r / Cin);
      const float a = (scale ? __ldg(scale + o) : 1.f) / __ldg(norm + o);
      __nv_bfloat16 h, l;
      sm100::split_bf16(__ldg(w + master_index(out_axis, Cout, Cin, T, o, i, t)) * a, h, l);
      bwd_hi[e] = h;
      if (bwd_lo) bwd_lo[e] = l;
    }
  }
}

__global__ void __launch_bounds__(WN_NT)
wn_project_kernel(const float* __restrict__ G, const float* __restrict__ w, const float* __restrict__ scale,
                  const float* __restrict__ norm, int out_axis, int Cout, int Cin, int T, float c,
                  float* __restrict__ dw, float* __restrict__ dscale, int accumulate) {
  __shared__ float red[33];
  const int o = blockIdx.x;
  const int R = Cin * T;
  float dot = 0.f;
  for (int r = threadIdx.x; r < R; r += WN_NT) {
    const int i = r / T, t = r - i * T;
    const int64_t idx = master_index(out_axis, Cout, Cin, T, o, i, t);
    dot = fmaf(__ldg(G + idx), __ldg(w + idx), dot);
  }
  dot = block_sum<WN_NT>(dot, red, true);
  const float n = __ldg(norm + o), s = scale ? __ldg(scale + o) : 1.f;
  const float a = s / n, k = c * dot / (n * n);
  for (int r = threadIdx.x; r < R; r += WN_NT) {
    const int i = r / T, t = r - i * T;
    const int64_t idx = master_index(out_axis, Cout, Cin, T, o, i, t);
    const float v = a * (__ldg(G + idx) - k * __ldg(w + idx));
    dw[idx] = accumulate ? dw[idx] + v : v;
  }
  if (dscale && threadIdx.x == 0) dscale[o] = accumulate ? dscale[o] + dot / n : dot / n;
}

}  // namespace glis

using namespace glis;

extern "C" int glis_wn_prepare(const float* w, const float* scale, int out_axis, int Cout, int Cin, int T,
                               float c, float* norm, float* pack_io, float* pack_oi, void* stream) {
  GLIS_REQUIRE(w && norm, GLIS_E_BADARG, "glis_wn_prepare: w/norm is NULL");
  GLIS_REQUIRE(Cout > 0 && Cin > 0 && T > 0 && (out_axis == 0 || out_axis == 1), GLIS_E_BADARG,
               "glis_wn_prepare: bad shape (Cout=%d Cin=%d T=%d axis=%d)", Cout, Cin, T, out_axis);
  cudaStream_t st = (cudaStream_t)stream;
  wn_norm_kernel<<<Cout, WN_NT, 0, st>>>(w, out_axis, Cout, Cin, T, c, norm);
  GLIS_CHECK_LAUNCH("glis_wn_prepare(norm)");
  if (pack_io || pack_oi) {
    const int64_t total = (int64_t)T * Cin * Cout;
    const int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
    wn_pack_kernel<<<blocks, 256, 0, st>>>(w, scale, norm, out_axis, Cout, Cin, T, pack_io, pack_oi);
    GLIS_CHECK_LAUNCH("glis_wn_prepare(pack)");
  }
  return GLIS_OK;
}

extern "C" int glis_wn_prepare_bf16(const float* w, const float* scale, int out_axis, int Cout, int Cin, int T,
                                    float c, float* norm, void* fwd_hi, void* fwd_lo, void* bwd_hi, void* bwd_lo,
                                    void* stream) {
  GLIS_REQUIRE(w && norm, GLIS_E_BADARG, "glis_wn_prepare_bf16: w/norm is NULL");
  GLIS_REQUIRE(Cout > 0 && Cin > 0 && T > 0 && (out_axis == 0 || out_axis == 1), GLIS_E_BADARG,
               "glis_wn_prepare_bf16: bad shape (Cout=%d Cin=%d T=%d axis=%d)", Cout, Cin, T, out_axis);
  GLIS_REQUIRE((fwd_hi || !fwd_lo) && (bwd_hi || !bwd_lo), GLIS_E_BADARG, "glis_wn_prepare_bf16: lo plane without hi");
  cudaStream_t st = (cudaStream_t)stream;
  wn_norm_kernel<<<Cout, WN_NT, 0, st>>>(w, out_axis, Cout, Cin, T, c, norm);
  GLIS_CHECK_LAUNCH("glis_wn_prepare_bf16(norm)");
  if (fwd_hi || bwd_hi) {
    const int64_t total = (int64_t)T * Cin * Cout;
    const int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
    wn_pack_bf16_kernel<<<blocks, 256, 0, st>>>(w, scale, norm, out_axis, Cout, Cin, T, (__nv_bfloat16*)fwd_hi,
                                               (__nv_bfloat16*)fwd_lo, (__nv_bfloat16*)bwd_hi, (__nv_bfloat16*)bwd_lo);
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
  wn_project_kernel<<<Cout, WN_NT, 0, (cudaStream_t)stream>>>(G, w, scale, norm, out_axis, Cout, Cin, T, c, dw,
                                                             dscale, accumulate);
  GLIS_CHECK_LAUNCH("glis_wn_project");
  return GLIS_OK;
}

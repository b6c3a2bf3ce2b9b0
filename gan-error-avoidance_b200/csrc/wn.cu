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
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
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
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
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

static int launch_wn_norm(const float* w, int out_axis, int Cout, int Cin, int T, float c, float* norm,
                           cudaStream_t st) {
  const int64_t R = (int64_t)Cin * T;
  if (R <= 512 && Cout >= 64)
    GLIS_LAUNCH(wn_norm_warp_kernel, dim3((Cout + WN_NT / 32 - 1) / (WN_NT / 32)), dim3(WN_NT), 0, (cudaStream_t)(st), w, out_axis, Cout, Cin, T, c, norm);
  else if (R >= 8192 || (Cout <= 16 && R >= 4096))
    GLIS_LAUNCH((wn_norm_kernel<1024>), dim3(Cout), dim3(1024), 0, (cudaStream_t)(st), w, out_axis, Cout, Cin, T, c, norm);
  else if (R >= 4096)
    GLIS_LAUNCH((wn_norm_kernel<512>), dim3(Cout), dim3(512), 0, (cudaStream_t)(st), w, out_axis, Cout, Cin, T, c, norm);
  else
    GLIS_LAUNCH((wn_norm_kernel<WN_NT>), dim3(Cout), dim3(WN_NT), 0, (cudaStream_t)(st), w, out_axis, Cout, Cin, T, c, norm);
  return GLIS_OK;
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
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
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
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
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

// Raw gradient element idx: the sum of the K-split slabs of a deterministic weight-gradient launch, always in slab
// order (n_slabs = 1: the plain buffer).
__device__ __forceinline__ float slab_sum(const float* __restrict__ G, int64_t idx, int n_slabs, int64_t slab_stride) {
  float v = __ldg(G + idx);
  for (int s = 1; s < n_slabs; ++s) v += __ldg(G + (int64_t)s * slab_stride + idx);
  return v;
}

// One block per output channel, at most PER elements of the row per thread: G and w are read ONCE,
// with every load in flight together, and stay in registers between the dot product and the update.
template <int NT, int PER>
__global__ void __launch_bounds__(NT)
wn_project_kernel(const float* __restrict__ G, const float* __restrict__ w, const float* __restrict__ scale,
                  const float* __restrict__ norm, int out_axis, int Cout, int Cin, int T, float c,
                  float* __restrict__ dw, float* __restrict__ dscale, int accumulate, int n_slabs, int64_t slab_stride) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
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
      gv[u] = slab_sum(G, idx, n_slabs, slab_stride);
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
                       float* __restrict__ dw, float* __restrict__ dscale, int accumulate, int n_slabs, int64_t slab_stride) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  constexpr int NT = 1024;
  __shared__ float red[33];
  const int o = blockIdx.x;
  const int R = Cin * T;
  float dot = 0.f;
#pragma unroll 4
  for (int r = threadIdx.x; r < R; r += NT) {
    const int i = r / T, t = r - i * T;
    const int64_t idx = master_index(out_axis, Cout, Cin, T, o, i, t);
    dot = fmaf(slab_sum(G, idx, n_slabs, slab_stride), __ldg(w + idx), dot);
  }
  dot = block_sum<NT>(dot, red, true);
  const float n = __ldg(norm + o), s = scale ? __ldg(scale + o) : 1.f;
  const float a = s / n, k = c * dot / (n * n);
#pragma unroll 4
  for (int r = threadIdx.x; r < R; r += NT) {
    const int i = r / T, t = r - i * T;
    const int64_t idx = master_index(out_axis, Cout, Cin, T, o, i, t);
    const float v = a * (slab_sum(G, idx, n_slabs, slab_stride) - k * __ldg(w + idx));
    dw[idx] = accumulate ? dw[idx] + v : v;
  }
  if (dscale && threadIdx.x == 0) dscale[o] = accumulate ? dscale[o] + dot / n : dot / n;
}

// One warp per output channel (see wn_norm_warp_kernel).
__global__ void __launch_bounds__(WN_NT)
wn_project_warp_kernel(const float* __restrict__ G, const float* __restrict__ w, const float* __restrict__ scale,
                       const float* __restrict__ norm, int out_axis, int Cout, int Cin, int T, float c,
                       float* __restrict__ dw, float* __restrict__ dscale, int accumulate, int n_slabs, int64_t slab_stride) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int o = blockIdx.x * (WN_NT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (o >= Cout) return;
  const int R = Cin * T;
  float dot = 0.f;
#pragma unroll 8
  for (int r = lane; r < R; r += 32) {
    const int i = r / T, t = r - i * T;
    const int64_t idx = master_index(out_axis, Cout, Cin, T, o, i, t);
    dot = fmaf(slab_sum(G, idx, n_slabs, slab_stride), __ldg(w + idx), dot);
  }
  dot = warp_sum(dot);
  const float n = __ldg(norm + o), s = scale ? __ldg(scale + o) : 1.f;
  const float a = s / n, k = c * dot / (n * n);
#pragma unroll 8
  for (int r = lane; r < R; r += 32) {
    const int i = r / T, t = r - i * T;
    const int64_t idx = master_index(out_axis, Cout, Cin, T, o, i, t);
    const float v = a * (slab_sum(G, idx, n_slabs, slab_stride) - k * __ldg(w + idx));
    dw[idx] = accumulate ? dw[idx] + v : v;
  }
  if (dscale && lane == 0) dscale[o] = accumulate ? dscale[o] + dot / n : dot / n;
}

static int check_perm(int out_axis, int Cout, int T, int perm_c, int perm_p, const char* who) {
  GLIS_REQUIRE((perm_c == 0 && perm_p == 0) ||
                   (perm_c > 0 && perm_p > 0 && T == 1 && out_axis == 0 && (int64_t)perm_c * perm_p == Cout),
               GLIS_E_BADARG, "%s: bad row permutation (C=%d P=%d for Cout=%d T=%d axis=%d)", who, perm_c, perm_p, Cout, T,
               out_axis);
  return GLIS_OK;
}

// ---------------------------------------------------------------------------- multi-tensor prepare
// All layers of a network in TWO launches (norms, then packs) instead of a launch pair per layer: the descriptors
// travel as a kernel parameter (captured by value into a CUDA graph), a block finds its layer by scanning the
// block-offset prefix.  ~30 launches of 3-20 us become 2 per network and update.
constexpr int WN_MULTI_MAX = 24;
struct WnMultiParams {
  int n;
  int norm_block_begin[WN_MULTI_MAX + 1];   // norm kernel: blocks of layer l = [begin[l], begin[l+1])
  int pack_block_begin[WN_MULTI_MAX + 1];   // pack kernel
  glis_wn_layer_t L[WN_MULTI_MAX];
};

__global__ void __launch_bounds__(WN_NT)
wn_norm_multi_kernel(const __grid_constant__ WnMultiParams P) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ float red[33];
  int l = 0;
  while (l + 1 < P.n && (int)blockIdx.x >= P.norm_block_begin[l + 1]) ++l;
  const glis_wn_layer_t& Y = P.L[l];
  const int b = blockIdx.x - P.norm_block_begin[l];
  const int R = Y.Cin * Y.T;
  const float* __restrict__ w = Y.w;
  if (R <= 512) {                      // one warp per output channel, 8 channels per block
    const int o = b * (WN_NT / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (o >= Y.Cout) return;
    float ss = 0.f;
#pragma unroll 8
    for (int r = lane; r < R; r += 32) {
      const int i = r / Y.T, t = r - i * Y.T;
      const float v = __ldg(w + master_index(Y.out_axis, Y.Cout, Y.Cin, Y.T, o, i, t));
      ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) Y.norm[o] = sqrtf(ss * Y.c + 1e-6f);
    return;
  }
  const int o = b;                     // one block per output channel
  float ss = 0.f;
#pragma unroll 8
  for (int r = threadIdx.x; r < R; r += WN_NT) {
    const int i = r / Y.T, t = r - i * Y.T;
    const float v = __ldg(w + master_index(Y.out_axis, Y.Cout, Y.Cin, Y.T, o, i, t));
    ss = fmaf(v, v, ss);
  }
  ss = block_sum<WN_NT>(ss, red);
  if (threadIdx.x == 0) Y.norm[o] = sqrtf(ss * Y.c + 1e-6f);
}

constexpr int WN_PACK_PER_BLOCK = WN_NT * 8;   // packed elements per block of the multi-tensor pack kernel

__global__ void __launch_bounds__(WN_NT)
wn_pack_multi_kernel(const __grid_constant__ WnMultiParams P) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  int l = 0;
  while (l + 1 < P.n && (int)blockIdx.x >= P.pack_block_begin[l + 1]) ++l;
  const glis_wn_layer_t& Y = P.L[l];
  const uint32_t Cout = (uint32_t)Y.Cout, Cin = (uint32_t)Y.Cin, T = (uint32_t)Y.T;
  const uint32_t total = T * Cin * Cout;
  const uint32_t e0 = (uint32_t)(blockIdx.x - P.pack_block_begin[l]) * WN_PACK_PER_BLOCK;
  const float* __restrict__ w = Y.w;
  const float* __restrict__ scale = Y.scale;
  const float* __restrict__ norm = Y.norm;
  __nv_bfloat16* fwd_hi = (__nv_bfloat16*)Y.fwd_hi; __nv_bfloat16* fwd_lo = (__nv_bfloat16*)Y.fwd_lo;
  __nv_bfloat16* bwd_hi = (__nv_bfloat16*)Y.bwd_hi; __nv_bfloat16* bwd_lo = (__nv_bfloat16*)Y.bwd_lo;
  __nv_bfloat16* e_hi = (__nv_bfloat16*)Y.mat_hi; __nv_bfloat16* e_lo = (__nv_bfloat16*)Y.mat_lo;
  __nv_bfloat16* et_hi = (__nv_bfloat16*)Y.matt_hi; __nv_bfloat16* et_lo = (__nv_bfloat16*)Y.matt_lo;
#pragma unroll 2
  for (uint32_t e = e0 + threadIdx.x; e < total && e < e0 + WN_PACK_PER_BLOCK; e += WN_NT) {
    if (Y.pack_oi || fwd_hi) {   // [t][o][i]
      const int i = (int)(e % Cin); const uint32_t r = e / Cin;
      const int o = master_channel((int)(r % Cout), Y.perm_c, Y.perm_p);
      const int t = (int)(r / Cout);
      const float v = __ldg(w + master_index(Y.out_axis, Y.Cout, Y.Cin, Y.T, o, i, t)) *
                      ((scale ? __ldg(scale + o) : 1.f) / __ldg(norm + o));
      if (Y.pack_oi) Y.pack_oi[e] = v;
      if (fwd_hi) {
        __nv_bfloat16 h, lo_;
        sm100::split_bf16(v, h, lo_);
        fwd_hi[e] = h;
        if (fwd_lo) fwd_lo[e] = lo_;
      }
    }
    if (Y.pack_io || bwd_hi) {   // [t][i][o]
      const int o = master_channel((int)(e % Cout), Y.perm_c, Y.perm_p);
      const uint32_t r = e / Cout; const int i = (int)(r % Cin); const int t = (int)(r / Cin);
      const float v = __ldg(w + master_index(Y.out_axis, Y.Cout, Y.Cin, Y.T, o, i, t)) *
                      ((scale ? __ldg(scale + o) : 1.f) / __ldg(norm + o));
      if (Y.pack_io) Y.pack_io[e] = v;
      if (bwd_hi) {
        __nv_bfloat16 h, lo_;
        sm100::split_bf16(v, h, lo_);
        bwd_hi[e] = h;
        if (bwd_lo) bwd_lo[e] = lo_;
      }
    }
    if (e_hi || et_hi) {         // E[a][j] over the master order (image-side layers), E^T[j][a]
      const uint32_t J = total / (uint32_t)Y.mat_rows;
      const uint32_t a = e / J, j = e - a * J;
      const int o = Y.out_axis == 0 ? (int)a : (int)(j / T);
      const float v = __ldg(w + e) * ((scale ? __ldg(scale + o) : 1.f) / __ldg(norm + o));
      __nv_bfloat16 h, lo_;
      sm100::split_bf16(v, h, lo_);
      if (e_hi) { e_hi[e] = h; if (e_lo) e_lo[e] = lo_; }
      if (et_hi) { et_hi[(size_t)j * Y.mat_rows + a] = h; if (et_lo) et_lo[(size_t)j * Y.mat_rows + a] = lo_; }
    }
  }
}

// ---------------------------------------------------------------------------- multi-tensor projection
struct WnProjMultiParams {
  int n;
  int block_begin[WN_MULTI_MAX + 1];
  glis_wn_proj_t L[WN_MULTI_MAX];
};

// Two passes over a row (the second one hits L1 / L2): dot = <G_o, w_o>, then dw_o (+)= (s/n)(G_o - c w_o dot / n^2).
// Rows of <= 512 elements: one warp per output channel, 8 per block; longer rows: one block per channel.
__global__ void __launch_bounds__(WN_NT)
wn_project_multi_kernel(const __grid_constant__ WnProjMultiParams P) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ float red[33];
  int l = 0;
  while (l + 1 < P.n && (int)blockIdx.x >= P.block_begin[l + 1]) ++l;
  const glis_wn_proj_t& Y = P.L[l];
  const int b = blockIdx.x - P.block_begin[l];
  const int R = Y.Cin * Y.T;
  const float* __restrict__ G = Y.G;
  const float* __restrict__ w = Y.w;
  float* __restrict__ dw = Y.dw;
  const bool warp_rows = R <= 512;
  const int o = warp_rows ? b * (WN_NT / 32) + (threadIdx.x >> 5) : b;
  const int r0 = warp_rows ? (threadIdx.x & 31) : threadIdx.x, rs = warp_rows ? 32 : WN_NT;
  if (o >= Y.Cout) return;             // (whole warps: the block-wide sum below is only reached by full blocks)
  float dot = 0.f;
#pragma unroll 4
  for (int r = r0; r < R; r += rs) {
    const int i = r / Y.T, t = r - i * Y.T;
    const int64_t idx = master_index(Y.out_axis, Y.Cout, Y.Cin, Y.T, o, i, t);
    dot = fmaf(__ldg(G + idx), __ldg(w + idx), dot);
  }
  dot = warp_rows ? warp_sum(dot) : block_sum<WN_NT>(dot, red, true);
  const float n = __ldg(Y.norm + o), sc = Y.scale ? __ldg(Y.scale + o) : 1.f;
  const float a = sc / n, k = Y.c * dot / (n * n);
#pragma unroll 4
  for (int r = r0; r < R; r += rs) {
    const int i = r / Y.T, t = r - i * Y.T;
    const int64_t idx = master_index(Y.out_axis, Y.Cout, Y.Cin, Y.T, o, i, t);
    const float v = a * (__ldg(G + idx) - k * __ldg(w + idx));
    dw[idx] = Y.accumulate ? dw[idx] + v : v;
  }
  if (Y.dscale && r0 == 0) Y.dscale[o] = Y.accumulate ? Y.dscale[o] + dot / n : dot / n;
}

// out[i] = slabs[0][i] + slabs[1][i] + ... in slab order: the K-split partial sums of a deterministic weight-gradient
// launch.  Pure streaming (float4, every slab's load of an element in flight together).
__global__ void __launch_bounds__(WN_NT)
slab_reduce_kernel(const float* __restrict__ slabs, int n_slabs, int64_t slab_stride, float* __restrict__ out, int64_t n4) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * WN_NT + threadIdx.x; i < n4; i += (int64_t)gridDim.x * WN_NT) {
    float4 acc = __ldg(reinterpret_cast<const float4*>(slabs) + i);
    int s = 1;
    for (; s + 3 < n_slabs; s += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(slabs + (int64_t)s * slab_stride) + i);
      const float4 b = __ldg(reinterpret_cast<const float4*>(slabs + (int64_t)(s + 1) * slab_stride) + i);
      const float4 c = __ldg(reinterpret_cast<const float4*>(slabs + (int64_t)(s + 2) * slab_stride) + i);
      const float4 d = __ldg(reinterpret_cast<const float4*>(slabs + (int64_t)(s + 3) * slab_stride) + i);
      acc.x = (((acc.x + a.x) + b.x) + c.x) + d.x; acc.y = (((acc.y + a.y) + b.y) + c.y) + d.y;
      acc.z = (((acc.z + a.z) + b.z) + c.z) + d.z; acc.w = (((acc.w + a.w) + b.w) + c.w) + d.w;
    }
    for (; s < n_slabs; ++s) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(slabs + (int64_t)s * slab_stride) + i);
      acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

}  // namespace glis

using namespace glis;

extern "C" int glis_wn_project_multi(const glis_wn_proj_t* items, int n, void* stream) {
  GLIS_REQUIRE(items && n >= 0, GLIS_E_BADARG, "glis_wn_project_multi: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int base = 0; base < n; base += WN_MULTI_MAX) {
    WnProjMultiParams P;
    P.n = n - base < WN_MULTI_MAX ? n - base : WN_MULTI_MAX;
    int nb = 0;
    for (int l = 0; l < P.n; ++l) {
      const glis_wn_proj_t& Y = items[base + l];
      GLIS_REQUIRE(Y.G && Y.w && Y.norm && Y.dw && Y.Cout > 0 && Y.Cin > 0 && Y.T > 0 && (Y.out_axis == 0 || Y.out_axis == 1),
                   GLIS_E_BADARG, "glis_wn_project_multi: item %d: bad descriptor", base + l);
      P.L[l] = Y;
      P.block_begin[l] = nb;
      nb += (int64_t)Y.Cin * Y.T <= 512 ? (Y.Cout + WN_NT / 32 - 1) / (WN_NT / 32) : Y.Cout;
    }
    P.block_begin[P.n] = nb;
    if (nb > 0) {
      GLIS_LAUNCH(wn_project_multi_kernel, dim3(nb), dim3(WN_NT), 0, (cudaStream_t)(st), P);
      GLIS_CHECK_LAUNCH("glis_wn_project_multi");
    }
  }
  return GLIS_OK;
}


extern "C" int glis_wn_prepare_multi(const glis_wn_layer_t* layers, int n, void* stream) {
  GLIS_REQUIRE(layers && n >= 0, GLIS_E_BADARG, "glis_wn_prepare_multi: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int base = 0; base < n; base += WN_MULTI_MAX) {
    WnMultiParams P;
    P.n = n - base < WN_MULTI_MAX ? n - base : WN_MULTI_MAX;
    int nb = 0, pb = 0;
    bool any_pack = false, any_norm = false;
    for (int l = 0; l < P.n; ++l) {
      const glis_wn_layer_t& Y = layers[base + l];
      GLIS_REQUIRE(Y.w && Y.norm && Y.Cout > 0 && Y.Cin > 0 && Y.T > 0 && (Y.out_axis == 0 || Y.out_axis == 1), GLIS_E_BADARG,
                   "glis_wn_prepare_multi: layer %d: bad descriptor", base + l);
      GLIS_REQUIRE((int64_t)Y.T * Y.Cin * Y.Cout < ((int64_t)1 << 31), GLIS_E_UNSUPPORTED,
                   "glis_wn_prepare_multi: layer %d has 2^31 or more weights", base + l);
      if (int rc = check_perm(Y.out_axis, Y.Cout, Y.T, Y.perm_c, Y.perm_p, "glis_wn_prepare_multi")) return rc;
      GLIS_REQUIRE((Y.fwd_hi || !Y.fwd_lo) && (Y.bwd_hi || !Y.bwd_lo) && (Y.mat_hi || !Y.mat_lo) && (Y.matt_hi || !Y.matt_lo),
                   GLIS_E_BADARG, "glis_wn_prepare_multi: layer %d: lo plane without hi", base + l);
      const bool mats = Y.mat_hi || Y.matt_hi;
      GLIS_REQUIRE(!mats || (Y.mat_rows > 0 && ((int64_t)Y.T * Y.Cin * Y.Cout) % Y.mat_rows == 0 && Y.perm_c == 0), GLIS_E_BADARG,
                   "glis_wn_prepare_multi: layer %d: bad matrix pack shape", base + l);
      P.L[l] = Y;
      P.norm_block_begin[l] = nb;
      if (Y.need_norm) {
        const int64_t R = (int64_t)Y.Cin * Y.T;
        nb += R <= 512 ? (Y.Cout + WN_NT / 32 - 1) / (WN_NT / 32) : Y.Cout;
        any_norm = true;
      }
      P.pack_block_begin[l] = pb;
      if (Y.pack_io || Y.pack_oi || Y.fwd_hi || Y.bwd_hi || mats) {
        const int64_t total = (int64_t)Y.T * Y.Cin * Y.Cout;
        pb += (int)((total + WN_PACK_PER_BLOCK - 1) / WN_PACK_PER_BLOCK);
        any_pack = true;
      }
    }
    P.norm_block_begin[P.n] = nb;
    P.pack_block_begin[P.n] = pb;
    // (a layer without work owns an empty block range: the prefix scan skips it)
    if (any_norm && nb > 0) {
      GLIS_LAUNCH(wn_norm_multi_kernel, dim3(nb), dim3(WN_NT), 0, (cudaStream_t)(st), P);
      GLIS_CHECK_LAUNCH("glis_wn_prepare_multi(norm)");
    }
    if (any_pack && pb > 0) {
      GLIS_LAUNCH(wn_pack_multi_kernel, dim3(pb), dim3(WN_NT), 0, (cudaStream_t)(st), P);
      GLIS_CHECK_LAUNCH("glis_wn_prepare_multi(pack)");
    }
  }
  return GLIS_OK;
}


extern "C" int glis_wn_prepare(const float* w, const float* scale, int out_axis, int Cout, int Cin, int T,
                               float c, float* norm, float* pack_io, float* pack_oi, void* stream) {
  return glis_wn_prepare_perm(w, scale, out_axis, Cout, Cin, T, c, norm, pack_io, pack_oi, 0, 0, stream);
}


extern "C" int glis_wn_prepare_perm(const float* w, const float* scale, int out_axis, int Cout, int Cin, int T,
                                    float c, float* norm, float* pack_io, float* pack_oi, int perm_c, int perm_p,
                                    void* stream) {
  GLIS_REQUIRE(w && norm, GLIS_E_BADARG, "glis_wn_prepare: w/norm is NULL");
  if (int rc = check_perm(out_axis, Cout, T, perm_c, perm_p, "glis_wn_prepare_perm")) return rc;
  GLIS_REQUIRE(Cout > 0 && Cin > 0 && T > 0 && (out_axis == 0 || out_axis == 1), GLIS_E_BADARG,
               "glis_wn_prepare: bad shape (Cout=%d Cin=%d T=%d axis=%d)", Cout, Cin, T, out_axis);
  cudaStream_t st = (cudaStream_t)stream;
  { const int rc_n = launch_wn_norm(w, out_axis, Cout, Cin, T, c, norm, st); if (rc_n != GLIS_OK) return rc_n; }
  GLIS_CHECK_LAUNCH("glis_wn_prepare(norm)");
  if (pack_io || pack_oi) {
    const int64_t total = (int64_t)T * Cin * Cout;
    const int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
    if (total < ((int64_t)1 << 31))
      GLIS_LAUNCH((wn_pack_kernel<uint32_t>), dim3(blocks), dim3(256), 0, (cudaStream_t)(st), w, scale, norm, out_axis, Cout, Cin, T, pack_io, pack_oi, perm_c, perm_p);
    else
      GLIS_LAUNCH((wn_pack_kernel<int64_t>), dim3(blocks), dim3(256), 0, (cudaStream_t)(st), w, scale, norm, out_axis, Cout, Cin, T, pack_io, pack_oi, perm_c, perm_p);
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
  { const int rc_n = launch_wn_norm(w, out_axis, Cout, Cin, T, c, norm, st); if (rc_n != GLIS_OK) return rc_n; }
  GLIS_CHECK_LAUNCH("glis_wn_prepare_bf16(norm)");
  if (fwd_hi || bwd_hi) {
    const int64_t total = (int64_t)T * Cin * Cout;
    const int blocks = (int)min((int64_t)148 * 16, (total + 255) / 256);
    if (total < ((int64_t)1 << 31))
      GLIS_LAUNCH((wn_pack_bf16_kernel<uint32_t>), dim3(blocks), dim3(256), 0, (cudaStream_t)(st), w, scale, norm, out_axis, Cout, Cin, T, (__nv_bfloat16*)fwd_hi,
                                                           (__nv_bfloat16*)fwd_lo, (__nv_bfloat16*)bwd_hi,
                                                           (__nv_bfloat16*)bwd_lo, perm_c, perm_p);
    else
      GLIS_LAUNCH((wn_pack_bf16_kernel<int64_t>), dim3(blocks), dim3(256), 0, (cudaStream_t)(st), w, scale, norm, out_axis, Cout, Cin, T, (__nv_bfloat16*)fwd_hi,
                                                          (__nv_bfloat16*)fwd_lo, (__nv_bfloat16*)bwd_hi,
                                                          (__nv_bfloat16*)bwd_lo, perm_c, perm_p);
    GLIS_CHECK_LAUNCH("glis_wn_prepare_bf16(pack)");
  }
  return GLIS_OK;
}

extern "C" int glis_wn_project(const float* G, const float* w, const float* scale, const float* norm,
                               int out_axis, int Cout, int Cin, int T, float c, float* dw, float* dscale,
                               int accumulate, void* stream) {
  return glis_wn_project_slabs(G, 1, 0, w, scale, norm, out_axis, Cout, Cin, T, c, dw, dscale, accumulate, stream);
}

extern "C" int glis_wn_project_slabs(const float* G, int n_slabs, int64_t slab_stride, const float* w, const float* scale,
                                     const float* norm, int out_axis, int Cout, int Cin, int T, float c, float* dw,
                                     float* dscale, int accumulate, void* stream) {
  GLIS_REQUIRE(G && w && norm && dw, GLIS_E_BADARG, "glis_wn_project: NULL pointer");
  GLIS_REQUIRE(n_slabs >= 1 && (n_slabs == 1 || slab_stride >= (int64_t)Cout * Cin * T), GLIS_E_BADARG,
               "glis_wn_project_slabs: bad slab layout");
  GLIS_REQUIRE(Cout > 0 && Cin > 0 && T > 0 && (out_axis == 0 || out_axis == 1), GLIS_E_BADARG,
               "glis_wn_project: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t R = (int64_t)Cin * T;
#define WN_PROJECT_ARGS G, w, scale, norm, out_axis, Cout, Cin, T, c, dw, dscale, accumulate, n_slabs, slab_stride
  if (R <= 512 && Cout >= 64)
    GLIS_LAUNCH(wn_project_warp_kernel, dim3((Cout + WN_NT / 32 - 1) / (WN_NT / 32)), dim3(WN_NT), 0, (cudaStream_t)(st), WN_PROJECT_ARGS);
  else if (R <= 8 * 128) GLIS_LAUNCH((wn_project_kernel<128, 8>), dim3(Cout), dim3(128), 0, (cudaStream_t)(st), WN_PROJECT_ARGS);
  else if (R <= 16 * 128 && Cout >= 148) GLIS_LAUNCH((wn_project_kernel<128, 16>), dim3(Cout), dim3(128), 0, (cudaStream_t)(st), WN_PROJECT_ARGS);
  else if (R <= 8 * 256) GLIS_LAUNCH((wn_project_kernel<256, 8>), dim3(Cout), dim3(256), 0, (cudaStream_t)(st), WN_PROJECT_ARGS);
  else if (R <= 16 * 256 && Cout >= 148) GLIS_LAUNCH((wn_project_kernel<256, 16>), dim3(Cout), dim3(256), 0, (cudaStream_t)(st), WN_PROJECT_ARGS);
  else if (R <= 8 * 512) GLIS_LAUNCH((wn_project_kernel<512, 8>), dim3(Cout), dim3(512), 0, (cudaStream_t)(st), WN_PROJECT_ARGS);
  else if (R <= 16 * 512 && Cout >= 148) GLIS_LAUNCH((wn_project_kernel<512, 16>), dim3(Cout), dim3(512), 0, (cudaStream_t)(st), WN_PROJECT_ARGS);
  else if (R <= 8 * 1024) GLIS_LAUNCH((wn_project_kernel<1024, 8>), dim3(Cout), dim3(1024), 0, (cudaStream_t)(st), WN_PROJECT_ARGS);
  else if (R <= 16 * 1024) GLIS_LAUNCH((wn_project_kernel<1024, 16>), dim3(Cout), dim3(1024), 0, (cudaStream_t)(st), WN_PROJECT_ARGS);
  else GLIS_LAUNCH(wn_project_long_kernel, dim3(Cout), dim3(1024), 0, (cudaStream_t)(st), WN_PROJECT_ARGS);
#undef WN_PROJECT_ARGS
  GLIS_CHECK_LAUNCH("glis_wn_project");
  return GLIS_OK;
}

extern "C" int glis_slab_reduce(const float* slabs, int n_slabs, int64_t slab_stride, float* out, int64_t numel, void* stream) {
  GLIS_REQUIRE(slabs && out && n_slabs >= 1 && numel >= 0 && slab_stride >= numel, GLIS_E_BADARG, "glis_slab_reduce: bad arguments");
  GLIS_REQUIRE(numel % 4 == 0 && slab_stride % 4 == 0 &&
                   ((reinterpret_cast<uintptr_t>(slabs) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               GLIS_E_UNSUPPORTED, "glis_slab_reduce: sizes must be multiples of 4 floats and buffers 16-byte aligned");
  if (numel == 0) return GLIS_OK;
  const int64_t n4 = numel / 4;
  int64_t blocks = (n4 + WN_NT - 1) / WN_NT;
  if (blocks > 148 * 8) blocks = 148 * 8;
  GLIS_LAUNCH(slab_reduce_kernel, dim3((int)blocks), dim3(WN_NT), 0, (cudaStream_t)((cudaStream_t)stream), slabs, n_slabs, slab_stride, out, n4);
  GLIS_CHECK_LAUNCH("glis_slab_reduce");
  return GLIS_OK;
}

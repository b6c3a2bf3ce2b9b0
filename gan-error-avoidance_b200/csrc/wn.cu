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

// Float offset of quad q (4 consecutive floats) of output channel o's row; needs R % 4 == 0 and, along axis 1, T % 4 == 0.
__device__ __forceinline__ int64_t master_quad(int out_axis, int Cout, int Cin, int T, int o, int q) {
  if (out_axis == 0) return (int64_t)o * Cin * T + 4 * (int64_t)q;
  const int r = 4 * q, i = r / T, t = r - i * T;
  return ((int64_t)i * Cout + o) * T + t;
}
__device__ __forceinline__ bool quads_ok(const void* p0, const void* p1, const void* p2, int out_axis, int R, int T) {
  return (R & 3) == 0 && (out_axis == 0 || (T & 3) == 0) &&
         ((reinterpret_cast<uintptr_t>(p0) | reinterpret_cast<uintptr_t>(p1) | reinterpret_cast<uintptr_t>(p2)) & 15) == 0;
}

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
  float dot = 0.f;
  if (n_slabs == 1 && quads_ok(G, w, dw, out_axis, R, T)) {
    // 16-byte form: PER / 4 quads per thread, every load in flight before the dot product
    constexpr int PQ = PER / 4;
    const int R4 = R >> 2;
    float4 gq[PQ], wq[PQ], oq[PQ];
#pragma unroll
    for (int u = 0; u < PQ; ++u) {
      const int q = threadIdx.x + u * NT;
      gq[u] = wq[u] = oq[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (q < R4) {
        const int64_t idx = master_quad(out_axis, Cout, Cin, T, o, q);
        gq[u] = __ldg(reinterpret_cast<const float4*>(G + idx));
        wq[u] = __ldg(reinterpret_cast<const float4*>(w + idx));
        if (accumulate) oq[u] = *reinterpret_cast<const float4*>(dw + idx);
      }
    }
#pragma unroll
    for (int u = 0; u < PQ; ++u) {
      dot = fmaf(gq[u].x, wq[u].x, dot); dot = fmaf(gq[u].y, wq[u].y, dot);
      dot = fmaf(gq[u].z, wq[u].z, dot); dot = fmaf(gq[u].w, wq[u].w, dot);
    }
    dot = block_sum<NT>(dot, red, true);
    const float n = __ldg(norm + o), s = scale ? __ldg(scale + o) : 1.f;
    const float a = s / n, k = c * dot / (n * n);
#pragma unroll
    for (int u = 0; u < PQ; ++u) {
      const int q = threadIdx.x + u * NT;
      if (q < R4) {
        float4 r4;
        r4.x = oq[u].x + a * (gq[u].x - k * wq[u].x); r4.y = oq[u].y + a * (gq[u].y - k * wq[u].y);
        r4.z = oq[u].z + a * (gq[u].z - k * wq[u].z); r4.w = oq[u].w + a * (gq[u].w - k * wq[u].w);
        *reinterpret_cast<float4*>(dw + master_quad(out_axis, Cout, Cin, T, o, q)) = r4;
      }
    }
    if (dscale && threadIdx.x == 0) dscale[o] = accumulate ? dscale[o] + dot / n : dot / n;
    return;
  }
  float gv[PER], wv[PER], old[PER];
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
    if (quads_ok(w, nullptr, nullptr, Y.out_axis, R, Y.T)) {
      const int R4 = R >> 2;
#pragma unroll 4
      for (int q = lane; q < R4; q += 32) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(w + master_quad(Y.out_axis, Y.Cout, Y.Cin, Y.T, o, q)));
        ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
      }
    } else {
#pragma unroll 8
      for (int r = lane; r < R; r += 32) {
        const int i = r / Y.T, t = r - i * Y.T;
        const float v = __ldg(w + master_index(Y.out_axis, Y.Cout, Y.Cin, Y.T, o, i, t));
        ss = fmaf(v, v, ss);
      }
    }
    ss = warp_sum(ss);
    if (lane == 0) Y.norm[o] = sqrtf(ss * Y.c + 1e-6f);
    return;
  }
  const int o = b;                     // one block per output channel
  float ss = 0.f;
  if (quads_ok(w, nullptr, nullptr, Y.out_axis, R, Y.T)) {   // 16-byte loads, 8 of them in flight per thread
    const int R4 = R >> 2;
#pragma unroll 8
    for (int q = threadIdx.x; q < R4; q += WN_NT) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(w + master_quad(Y.out_axis, Y.Cout, Y.Cin, Y.T, o, q)));
      ss = fmaf(v.x, v.x, ss); ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
    }
  } else {
#pragma unroll 8
    for (int r = threadIdx.x; r < R; r += WN_NT) {
      const int i = r / Y.T, t = r - i * Y.T;
      const float v = __ldg(w + master_index(Y.out_axis, Y.Cout, Y.Cin, Y.T, o, i, t));
      ss = fmaf(v, v, ss);
    }
  }
  ss = block_sum<WN_NT>(ss, red);
  if (threadIdx.x == 0) Y.norm[o] = sqrtf(ss * Y.c + 1e-6f);
}

constexpr int WN_PACK_PER_BLOCK = WN_NT * 2;   // packed elements per block of the multi-tensor pack kernel (what is left to it
                                               // are small layers: short blocks, both of a thread's elements in flight)

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

// ---------------------------------------------------------------------------- wide bf16 packs
// What bounds the element-wise pack above is neither bytes nor the gather but the NUMBER OF STORE INSTRUCTIONS: it
// issues four 2-byte stores per weight (hi / lo of both packs), one load in flight per thread.  2.1 M weights take
// 18 us whatever the access pattern (tools/pack_microbench.py; a shared-memory-tiled variant that kept the 2-byte
// stores had the same floor).  Here every store carries FOUR bf16 (8 bytes per lane), every load is 16 bytes, and a
// thread has 4..16 loads in flight:
//   phase F (the pack whose innermost axis is the master's inner channel axis): a thread owns 4 consecutive inner
//            channels x T taps = 4T contiguous floats, and writes one bf16 quad per tap and plane;
//   phase C (the pack whose innermost axis is the master's OUTER axis): a thread owns 4 consecutive outer channels
//            x T taps of one inner channel (four T-float rows, a master row pitch apart) — or, for one-tap (linear)
//            layers, a 4 x 4 block that it transposes in registers.
// No shared memory; master order [outer][inner][T]: out_axis 0 -> (outer, inner) = (Cout, Cin), 1 -> (Cin, Cout).
struct WnWideItem { int layer, phase; };     // phase 0 = F, 1 = C
struct WnWideParams {
  int n_items;
  int block_begin[2 * WN_MULTI_MAX + 1];
  WnWideItem item[2 * WN_MULTI_MAX];
  glis_wn_layer_t L[WN_MULTI_MAX];
};

static bool wn_wide_applies(const glis_wn_layer_t& Y) {
  if (Y.pack_io || Y.pack_oi || Y.mat_hi || Y.matt_hi) return false;      // fp32 packs / matrix packs: element-wise kernel
  if (!Y.fwd_hi && !Y.bwd_hi) return false;
  if ((Y.Cout & 3) || (Y.Cin & 3)) return false;                             // bf16 quads along either axis
  if (Y.T == 1 && ((Y.Cout & 7) || (Y.Cin & 7))) return false;               // linear layers: bf16 octets
  if (Y.T != 1 && Y.T != 4 && Y.T != 9 && Y.T != 16) return false;           // (registers: 4 T floats per thread)
  if ((reinterpret_cast<uintptr_t>(Y.w) & 15) != 0) return false;            // 16-byte loads
  if (((reinterpret_cast<uintptr_t>(Y.fwd_hi) | reinterpret_cast<uintptr_t>(Y.fwd_lo) | reinterpret_cast<uintptr_t>(Y.bwd_hi) |
        reinterpret_cast<uintptr_t>(Y.bwd_lo)) & 15) != 0) return false;     // 8- / 16-byte stores
  return true;
}

// threads of one phase of a layer
static int64_t wn_wide_threads(const glis_wn_layer_t& Y, int phase) {
  const int64_t pairs = (int64_t)Y.Cout * Y.Cin;
  if (Y.T == 1) return phase == 0 ? pairs / 8 : pairs / 32;   // 8 inner channels per thread / an 8 (outer) x 4 (inner) block
  return pairs / 4;
}

__device__ __forceinline__ void store_quad(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t e, const float (&x)[4]) {
  __nv_bfloat16 h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) sm100::split_bf16(x[k], h[k], l[k]);
  uint2 qh, ql;
  qh.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
  qh.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
  *reinterpret_cast<uint2*>(hi + e) = qh;
  if (lo) {
    ql.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    ql.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
    *reinterpret_cast<uint2*>(lo + e) = ql;
  }
}

// eight consecutive bf16 per plane: one 16-byte store each
__device__ __forceinline__ void store_octet(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t e, const float (&x)[8]) {
  uint32_t ph[4], pl[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    __nv_bfloat16 h0, l0, h1, l1;
    sm100::split_bf16(x[2 * k], h0, l0);
    sm100::split_bf16(x[2 * k + 1], h1, l1);
    ph[k] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    pl[k] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  *reinterpret_cast<uint4*>(hi + e) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
  if (lo) *reinterpret_cast<uint4*>(lo + e) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
}

__device__ __forceinline__ float wn_factor(const glis_wn_layer_t& Y, int o_master) {
  return (Y.scale ? __ldg(Y.scale + o_master) : 1.f) / __ldg(Y.norm + o_master);
}

template <int T>
__device__ __forceinline__ void wn_wide_phase_f(const glis_wn_layer_t& Y, int64_t g) {
  const int n_outer = Y.out_axis == 0 ? Y.Cout : Y.Cin, n_inner = Y.out_axis == 0 ? Y.Cin : Y.Cout;
  __nv_bfloat16* d_hi = (__nv_bfloat16*)(Y.out_axis == 0 ? Y.fwd_hi : Y.bwd_hi);
  __nv_bfloat16* d_lo = (__nv_bfloat16*)(Y.out_axis == 0 ? Y.fwd_lo : Y.bwd_lo);
  if (T == 1) {   // linear layers: a converting copy, eight inner channels per thread
    const int o_inner = n_inner >> 3;
    const int outer = (int)(g / o_inner), inner = (int)(g - (int64_t)outer * o_inner) * 8;
    if (outer >= n_outer) return;
    const int m_outer = Y.out_axis == 0 ? master_channel(outer, Y.perm_c, Y.perm_p) : outer;
    const float4* src = reinterpret_cast<const float4*>(Y.w + (size_t)m_outer * n_inner + inner);
    const float4 q0 = __ldg(src), q1 = __ldg(src + 1);
    float x[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    const float a_row = Y.out_axis == 0 ? wn_factor(Y, m_outer) : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] *= Y.out_axis == 0 ? a_row : wn_factor(Y, inner + k);
    store_octet(d_hi, d_lo, (size_t)outer * n_inner + inner, x);
    return;
  }
  const int q_inner = n_inner >> 2;
  const int outer = (int)(g / q_inner), inner = (int)(g - (int64_t)outer * q_inner) * 4;
  if (outer >= n_outer) return;
  const int m_outer = Y.out_axis == 0 ? master_channel(outer, Y.perm_c, Y.perm_p) : outer;
  // 4 T contiguous floats, 16-byte aligned (inner and n_inner are multiples of 4)
  const float4* src = reinterpret_cast<const float4*>(Y.w + ((size_t)m_outer * n_inner + inner) * T);
  float v[4 * T];
#pragma unroll
  for (int j = 0; j < T; ++j) {
    const float4 q = __ldg(src + j);
    v[4 * j] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
  }
  float a[4];
  if (Y.out_axis == 0) {
    a[0] = a[1] = a[2] = a[3] = wn_factor(Y, m_outer);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = wn_factor(Y, inner + k);
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    float x[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) x[k] = v[k * T + t] * a[k];
    store_quad(d_hi, d_lo, ((size_t)t * n_outer + outer) * n_inner + inner, x);
  }
}

template <int T>
__device__ __forceinline__ void wn_wide_phase_c(const glis_wn_layer_t& Y, int64_t g) {
  const int n_outer = Y.out_axis == 0 ? Y.Cout : Y.Cin, n_inner = Y.out_axis == 0 ? Y.Cin : Y.Cout;
  const int q_outer = n_outer >> 2;
  __nv_bfloat16* c_hi = (__nv_bfloat16*)(Y.out_axis == 0 ? Y.bwd_hi : Y.fwd_hi);
  __nv_bfloat16* c_lo = (__nv_bfloat16*)(Y.out_axis == 0 ? Y.bwd_lo : Y.fwd_lo);
  if (T == 1) {
    // 8 x 4 block: rows outer .. outer + 7 (pack order), columns inner .. inner + 3; transposed in registers
    const int q_inner = n_inner >> 2, o_outer = n_outer >> 3;
    const int i4 = (int)(g / o_outer), outer = (int)(g - (int64_t)i4 * o_outer) * 8, inner = i4 * 4;
    if (i4 >= q_inner) return;
    float m[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int m_outer = Y.out_axis == 0 ? master_channel(outer + k, Y.perm_c, Y.perm_p) : outer + k;
      const float4 q = __ldg(reinterpret_cast<const float4*>(Y.w + (size_t)m_outer * n_inner + inner));
      const float a = Y.out_axis == 0 ? wn_factor(Y, m_outer) : 1.f;
      m[k][0] = q.x * a; m[k][1] = q.y * a; m[k][2] = q.z * a; m[k][3] = q.w * a;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = Y.out_axis == 1 ? wn_factor(Y, inner + j) : 1.f;
      const float x[8] = {m[0][j] * a, m[1][j] * a, m[2][j] * a, m[3][j] * a, m[4][j] * a, m[5][j] * a, m[6][j] * a, m[7][j] * a};
      store_octet(c_hi, c_lo, (size_t)(inner + j) * n_outer + outer, x);
    }
    return;
  }
  const int inner = (int)(g / q_outer), outer = (int)(g - (int64_t)inner * q_outer) * 4;
  if (inner >= n_inner) return;
  float v[4][T], a[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int m_outer = Y.out_axis == 0 ? master_channel(outer + k, Y.perm_c, Y.perm_p) : outer + k;
    const float* src = Y.w + ((size_t)m_outer * n_inner + inner) * T;
    if (T % 4 == 0) {
#pragma unroll
      for (int t = 0; t < T; t += 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(src + t));
        v[k][t] = q.x; v[k][t + 1] = q.y; v[k][t + 2] = q.z; v[k][t + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int t = 0; t < T; ++t) v[k][t] = __ldg(src + t);
    }
    a[k] = Y.out_axis == 0 ? wn_factor(Y, m_outer) : 0.f;
  }
  if (Y.out_axis == 1) a[0] = a[1] = a[2] = a[3] = wn_factor(Y, inner);
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float x[4] = {v[0][t] * a[0], v[1][t] * a[1], v[2][t] * a[2], v[3][t] * a[3]};
    store_quad(c_hi, c_lo, ((size_t)t * n_inner + inner) * n_outer + outer, x);
  }
}

template <int T>
__device__ __forceinline__ void wn_wide_dispatch(const glis_wn_layer_t& Y, int phase, int64_t g) {
  if (phase == 0) wn_wide_phase_f<T>(Y, g); else wn_wide_phase_c<T>(Y, g);
}

__global__ void __launch_bounds__(WN_NT)
wn_pack_wide_kernel(const __grid_constant__ WnWideParams P) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  int it = 0;
  while (it + 1 < P.n_items && (int)blockIdx.x >= P.block_begin[it + 1]) ++it;
  const glis_wn_layer_t& Y = P.L[P.item[it].layer];
  const int phase = P.item[it].phase;
  const int64_t g = (int64_t)(blockIdx.x - P.block_begin[it]) * WN_NT + threadIdx.x;
  switch (Y.T) {
    case 1: wn_wide_dispatch<1>(Y, phase, g); break;
    case 4: wn_wide_dispatch<4>(Y, phase, g); break;
    case 9: wn_wide_dispatch<9>(Y, phase, g); break;
    case 16: wn_wide_dispatch<16>(Y, phase, g); break;
    default: break;   // (wn_wide_applies admits no other tap count)
  }
}

// Appends the phases of layer `l` of Q (already stored in Q.L[l]) to the launch list; returns the new block count.
static int wn_wide_add(WnWideParams& Q, int l, int blocks) {
  const glis_wn_layer_t& Y = Q.L[l];
  for (int phase = 0; phase < 2; ++phase) {
    const bool direct_is_fwd = Y.out_axis == 0;
    const void* dst = phase == 0 ? (direct_is_fwd ? Y.fwd_hi : Y.bwd_hi) : (direct_is_fwd ? Y.bwd_hi : Y.fwd_hi);
    if (!dst) continue;
    Q.item[Q.n_items].layer = l; Q.item[Q.n_items].phase = phase;
    Q.block_begin[Q.n_items++] = blocks;
    blocks += (int)((wn_wide_threads(Y, phase) + WN_NT - 1) / WN_NT);
  }
  return blocks;
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
    WnWideParams Q;
    Q.n_items = 0;
    int qb = 0, ql = 0;
    static int wide_cfg = -1;
    if (wide_cfg < 0) {
      const char* e = getenv("GLIS_PACK_WIDE");   // 0: every pack on the element-wise kernel
      wide_cfg = (e && atoi(e) == 0) ? 0 : 1;
    }
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
      if (wide_cfg && wn_wide_applies(Y)) {
        Q.L[ql] = Y;
        qb = wn_wide_add(Q, ql++, qb);
      } else if (Y.pack_io || Y.pack_oi || Y.fwd_hi || Y.bwd_hi || mats) {
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
    if (Q.n_items > 0) {
      Q.block_begin[Q.n_items] = qb;
      GLIS_LAUNCH(wn_pack_wide_kernel, dim3(qb), dim3(WN_NT), 0, (cudaStream_t)(st), Q);
      GLIS_CHECK_LAUNCH("glis_wn_prepare_multi(wide pack)");
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
    {
      glis_wn_layer_t Y = {};
      Y.w = w; Y.scale = scale; Y.norm = norm; Y.fwd_hi = fwd_hi; Y.fwd_lo = fwd_lo; Y.bwd_hi = bwd_hi; Y.bwd_lo = bwd_lo;
      Y.out_axis = out_axis; Y.Cout = Cout; Y.Cin = Cin; Y.T = T; Y.perm_c = perm_c; Y.perm_p = perm_p; Y.c = c;
      const char* e = getenv("GLIS_PACK_WIDE");
      if (!(e && atoi(e) == 0) && wn_wide_applies(Y) && total < ((int64_t)1 << 31)) {
        WnWideParams Q;
        Q.n_items = 0; Q.L[0] = Y;
        const int qb = wn_wide_add(Q, 0, 0);
        Q.block_begin[Q.n_items] = qb;
        GLIS_LAUNCH(wn_pack_wide_kernel, dim3(qb), dim3(WN_NT), 0, st, Q);
        GLIS_CHECK_LAUNCH("glis_wn_prepare_bf16(wide pack)");
        return GLIS_OK;
      }
    }
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

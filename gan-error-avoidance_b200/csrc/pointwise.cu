// HBM-bound pieces of the G-LIS step: TPReLU (standalone form), per-channel reductions,
// the BCE / MSE losses with their gradients, fused RMSprop, and the Philox generators.
// All are grid-stride kernels sized to a multiple of the 148 SMs.
#include "common.cuh"
#include "sm100.cuh"

namespace glis {

constexpr int PW_NT = 256;
static inline int pw_blocks(int64_t numel, int per_thread = 4) {
  int64_t b = (numel + (int64_t)PW_NT * per_thread - 1) / ((int64_t)PW_NT * per_thread);
  const int64_t cap = 148 * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

__device__ __forceinline__ float clamp01(float a) { return fminf(fmaxf(a, 0.f), 1.f); }

__global__ void __launch_bounds__(PW_NT)
tprelu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ a_raw, const float* __restrict__ b,
                  float* __restrict__ out, int64_t numel, int C, int inner) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * PW_NT + threadIdx.x; i < numel; i += (int64_t)gridDim.x * PW_NT) {
    const int c = (int)((i / inner) % C);
    const float bb = __ldg(b + c), a = clamp01(__ldg(a_raw + c));
    const float t = x[i] - bb;
    out[i] = (t > 0.f ? t : a * t) + bb;
  }
}

// NHWC TPReLU forward writing fp32 and/or bf16 hi/lo planes, four consecutive channels per thread.
__global__ void __launch_bounds__(PW_NT)
tprelu_fwd_planes_kernel(const float* __restrict__ x, int nslabs, int64_t slab_stride, const float* __restrict__ a_raw,
                         const float* __restrict__ b, float* __restrict__ preact, float* __restrict__ out,
                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t numel, int C, int CA) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int64_t n4 = numel >> 2;   // C % 4 == 0: a group of four never straddles a pixel
  for (int64_t q = (int64_t)blockIdx.x * PW_NT + threadIdx.x; q < n4; q += (int64_t)gridDim.x * PW_NT) {
    float4 v = reinterpret_cast<const float4*>(x)[q];
    for (int s = 1; s < nslabs; ++s) {     // the partial sums of a split-K launch, always in this order
      const float4 w = reinterpret_cast<const float4*>(x + (int64_t)s * slab_stride)[q];
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    if (preact) reinterpret_cast<float4*>(preact)[q] = v;
    const int c0 = (int)((q * 4) % C);
    float o[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (c0 + j) % CA;
      const float bb = __ldg(b + c), a = clamp01(__ldg(a_raw + c));
      const float t = o[j] - bb;
      o[j] = (t > 0.f ? t : a * t) + bb;
      sm100::split_bf16(o[j], h[j], l[j]);
    }
    if (out) reinterpret_cast<float4*>(out)[q] = make_float4(o[0], o[1], o[2], o[3]);
    if (hi) reinterpret_cast<uint2*>(hi)[q] = *reinterpret_cast<uint2*>(h);
    if (lo) reinterpret_cast<uint2*>(lo)[q] = *reinterpret_cast<uint2*>(l);
  }
}

// dy = dout * s * (1 - s): the backward of a sigmoid fused into a contraction's epilogue.
__global__ void __launch_bounds__(PW_NT)
sigmoid_bwd_kernel(const float* __restrict__ s, const float* __restrict__ dout, float* __restrict__ dy, int64_t numel) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  for (int64_t i = (int64_t)blockIdx.x * PW_NT + threadIdx.x; i < numel; i += (int64_t)gridDim.x * PW_NT) {
    const float v = s[i];
    dy[i] = dout[i] * v * (1.f - v);
  }
}

// Adds `v` into acc[c]; when the whole warp targets one channel the warp reduces first.
__device__ __forceinline__ void channel_accumulate(float* acc, int c, float v, bool active) {
  const unsigned full = 0xffffffffu;
  const int c0 = __shfl_sync(full, c, 0);
  const bool uniform = __all_sync(full, (!active) || c == c0) && __shfl_sync(full, (int)active, 0);
  if (uniform) {
    const float s = warp_sum(active ? v : 0.f);
    if ((threadIdx.x & 31) == 0) atomicAdd(acc + c0, s);
  } else if (active) {
    atomicAdd(acc + c, v);
  }
}

constexpr int SMEM_CH = 2048;  // channels staged in shared memory per accumulator

__global__ void __launch_bounds__(PW_NT)
tprelu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ a_raw, const float* __restrict__ b,
                  const float* __restrict__ dout, float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_hi,
                  __nv_bfloat16* __restrict__ dx_lo, float* __restrict__ da, float* __restrict__ db, int64_t numel,
                  int C, int inner) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ float sa[SMEM_CH], sb[SMEM_CH];
  const bool staged = C <= SMEM_CH;
  if (staged) {
    for (int c = threadIdx.x; c < C; c += PW_NT) { sa[c] = 0.f; sb[c] = 0.f; }
    __syncthreads();
  }
  float* acc_a = staged ? sa : da;
  float* acc_b = staged ? sb : db;
  const bool sums = da != nullptr;   // da and db come together (or not at all: frozen TPReLU parameters)
  const int64_t stride = (int64_t)gridDim.x * PW_NT;
  const int64_t rounds = (numel + stride - 1) / stride;
  for (int64_t r = 0; r < rounds; ++r) {
    const int64_t i = r * stride + (int64_t)blockIdx.x * PW_NT + threadIdx.x;
    const bool active = i < numel;
    int c = 0; float ga = 0.f, gb = 0.f;
    if (active) {
      c = (int)((i / inner) % C);
      const float ar = __ldg(a_raw + c), bb = __ldg(b + c), a = clamp01(ar);
      const float t = x[i] - bb, g = dout[i];
      const bool neg = !(t > 0.f);
      const float d = neg ? a * g : g;
      if (dx) dx[i] = d;
      if (dx_hi) {
        __nv_bfloat16 h, l;
        sm100::split_bf16(d, h, l);
        dx_hi[i] = h;
        if (dx_lo) dx_lo[i] = l;
      }
      if (neg) {
        ga = (ar >= 0.f && ar <= 1.f) ? g * t : 0.f;
        gb = g * (1.f - a);
      }
    }
    if (sums) {
      channel_accumulate(acc_a, c, ga, active);
      channel_accumulate(acc_b, c, gb, active);
    }
  }
  if (staged && sums) {
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += PW_NT) {
      if (sa[c] != 0.f) atomicAdd(da + c, sa[c]);
      if (sb[c] != 0.f) atomicAdd(db + c, sb[c]);
    }
  }
}

// NHWC fast path (inner == 1, C % 4 == 0, C/4 <= 256): a thread owns four fixed channels and
// walks rows, so the per-channel sums live in registers and meet global memory once per thread.
__global__ void __launch_bounds__(PW_NT, 4)
tprelu_bwd_nhwc_kernel(const float* __restrict__ x, const float* __restrict__ a_raw, const float* __restrict__ b,
                       const float* __restrict__ dout, float* __restrict__ dx, __nv_bfloat16* __restrict__ dx_hi,
                       __nv_bfloat16* __restrict__ dx_lo, float* __restrict__ da, float* __restrict__ db,
                       int64_t rows, int C) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ float s_a[1024], s_b[1024];
  for (int c = threadIdx.x; c < C; c += PW_NT) { s_a[c] = 0.f; s_b[c] = 0.f; }
  __syncthreads();
  const int c4n = C >> 2;                       // float4 groups per row
  const int rows_per_iter = PW_NT / c4n;        // rows handled by the block per iteration
  const int cg = threadIdx.x % c4n, rl = threadIdx.x / c4n;
  if (rl < rows_per_iter) {
    const float4 ar = __ldg(reinterpret_cast<const float4*>(a_raw) + cg);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + cg);
    const float av[4] = {clamp01(ar.x), clamp01(ar.y), clamp01(ar.z), clamp01(ar.w)};
    const float arv[4] = {ar.x, ar.y, ar.z, ar.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
    float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
    // two row groups per trip, all four loads requested before the first is used: a thread walks only a handful of
    // rows, so the loop is a chain of memory round trips unless they overlap
    constexpr int U = 2;
    const int64_t step = (int64_t)gridDim.x * rows_per_iter;
    for (int64_t r0 = (int64_t)blockIdx.x * rows_per_iter + rl; r0 < rows; r0 += U * step) {
      float4 xq[U], gq[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + u * step;
        if (r < rows) {
          xq[u] = reinterpret_cast<const float4*>(x)[r * c4n + cg];
          gq[u] = reinterpret_cast<const float4*>(dout)[r * c4n + cg];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t r = r0 + u * step;
        if (r >= rows) break;
        const int64_t i4 = r * c4n + cg;
        const float xs[4] = {xq[u].x, xq[u].y, xq[u].z, xq[u].w}, gs[4] = {gq[u].x, gq[u].y, gq[u].z, gq[u].w};
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float t = xs[j] - bv[j];
          const bool neg = !(t > 0.f);
          d[j] = neg ? av[j] * gs[j] : gs[j];
          if (neg) { sa[j] = fmaf(gs[j], t, sa[j]); sb[j] += gs[j]; }
        }
        if (dx) reinterpret_cast<float4*>(dx)[i4] = make_float4(d[0], d[1], d[2], d[3]);
        if (dx_hi) {
          __nv_bfloat16 h[4], l[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) sm100::split_bf16(d[j], h[j], l[j]);
          reinterpret_cast<uint2*>(dx_hi)[i4] = *reinterpret_cast<uint2*>(h);
          if (dx_lo) reinterpret_cast<uint2*>(dx_lo)[i4] = *reinterpret_cast<uint2*>(l);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // block-level: the row groups of this block meet in shared memory
      if (arv[j] >= 0.f && arv[j] <= 1.f && sa[j] != 0.f) atomicAdd(&s_a[cg * 4 + j], sa[j]);
      if (sb[j] != 0.f) atomicAdd(&s_b[cg * 4 + j], sb[j] * (1.f - av[j]));
    }
  }
  __syncthreads();
  if (da == nullptr) return;   // frozen TPReLU parameters: only dx is wanted
  for (int c = threadIdx.x; c < C; c += PW_NT) {   // one global atomic per channel per block
    if (s_a[c] != 0.f) atomicAdd(da + c, s_a[c]);
    if (s_b[c] != 0.f) atomicAdd(db + c, s_b[c]);
  }
}

__global__ void __launch_bounds__(PW_NT)
channel_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t numel, int C, int inner) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ float sa[SMEM_CH];
  const bool staged = C <= SMEM_CH;
  if (staged) {
    for (int c = threadIdx.x; c < C; c += PW_NT) sa[c] = 0.f;
    __syncthreads();
  }
  float* acc = staged ? sa : out;
  const int64_t stride = (int64_t)gridDim.x * PW_NT;
  const int64_t rounds = (numel + stride - 1) / stride;
  for (int64_t r = 0; r < rounds; ++r) {
    const int64_t i = r * stride + (int64_t)blockIdx.x * PW_NT + threadIdx.x;
    const bool active = i < numel;
    const int c = active ? (int)((i / inner) % C) : 0;
    channel_accumulate(acc, c, active ? x[i] : 0.f, active);
  }
  if (staged) {
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += PW_NT)
      if (sa[c] != 0.f) atomicAdd(out + c, sa[c]);
  }
}

// Channel sums of an NHWC tensor with C <= 4 (the image-side bias gradient): a thread owns whole
// pixels, keeps C partial sums in registers, warps reduce, one atomic per warp and channel.
__global__ void __launch_bounds__(PW_NT)
channel_sum_small_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t rows, int C) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t r = (int64_t)blockIdx.x * PW_NT + threadIdx.x; r < rows; r += (int64_t)gridDim.x * PW_NT) {
    const float* p = x + r * C;
#pragma unroll
    for (int c = 0; c < 4; ++c) if (c < C) acc[c] += __ldg(p + c);
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c < C) {
      const float s = warp_sum(acc[c]);
      if ((threadIdx.x & 31) == 0 && s != 0.f) atomicAdd(out + c, s);
    }
  }
}

// One block: B is the batch (tens to thousands of logits).
__global__ void __launch_bounds__(PW_NT)
bce_logits_kernel(const float* __restrict__ logit, float target, int B, float gscale, float* __restrict__ loss,
                  float* __restrict__ dlogit, float* __restrict__ prob) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ float red[33];
  float acc = 0.f;
  for (int i = threadIdx.x; i < B; i += PW_NT) {
    const float l = logit[i];
    const float p = 1.f / (1.f + expf(-l));
    // nn.BCELoss on p with log clamped at -100 (modern torch)
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
    acc -= target * lp + (1.f - target) * l1p;
    if (dlogit) {
      // d/dl of the clamped form: the clamped branch has zero slope
      float g = 0.f;
      if (target != 0.f && lp > -100.f) g -= target * (1.f - p);
      if (target != 1.f && l1p > -100.f) g += (1.f - target) * p;
      dlogit[i] = gscale * g / (float)B;
    }
    if (prob) prob[i] = p;
  }
  acc = block_sum<PW_NT>(acc, red);
  if (threadIdx.x == 0) loss[0] = acc / (float)B;
}

// --ls: nn.MSELoss() on D's sigmoid output against a constant target (g_lis/main.py:308-311):
// L = mean((p - t)^2), dL/dl = 2 (p - t) p (1 - p) / B.  One block, as bce_logits_kernel.
__global__ void __launch_bounds__(PW_NT)
lsq_logits_kernel(const float* __restrict__ logit, float target, int B, float gscale, float* __restrict__ loss,
                  float* __restrict__ dlogit, float* __restrict__ prob) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ float red[33];
  float acc = 0.f;
  for (int i = threadIdx.x; i < B; i += PW_NT) {
    const float p = 1.f / (1.f + expf(-logit[i]));
    const float d = p - target;
    acc = fmaf(d, d, acc);
    if (dlogit) dlogit[i] = gscale * 2.f * d * p * (1.f - p) / (float)B;
    if (prob) prob[i] = p;
  }
  acc = block_sum<PW_NT>(acc, red);
  if (threadIdx.x == 0) loss[0] = acc / (float)B;
}

__global__ void __launch_bounds__(PW_NT)
mse_scaled_kernel(const float* __restrict__ u, const float* __restrict__ z, int64_t numel, float lambda,
                  float* __restrict__ loss, float* __restrict__ du, int accumulate) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ float red[33];
  float acc = 0.f;
  const float k = 2.f * lambda / (float)numel;
  for (int64_t i = (int64_t)blockIdx.x * PW_NT + threadIdx.x; i < numel; i += (int64_t)gridDim.x * PW_NT) {
    const float d = u[i] - z[i];
    acc = fmaf(d, d, acc);
    if (du) du[i] = (accumulate & 1) ? du[i] + k * d : k * d;
  }
  acc = block_sum<PW_NT>(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, acc * ((accumulate & 2) ? 1.f : lambda) / (float)numel);
}

__global__ void __launch_bounds__(PW_NT)
rmsprop_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v, int64_t numel,
               float lr, float alpha, float eps, float gscale) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int64_t n4 = numel >> 2;
  const bool vec = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)v)) & 15) == 0;
  const int64_t tid = (int64_t)blockIdx.x * PW_NT + threadIdx.x, stride = (int64_t)gridDim.x * PW_NT;
  if (vec) {
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* v4 = reinterpret_cast<float4*>(v);
    for (int64_t i = tid; i < n4; i += stride) {
      float4 pp = p4[i], vv = v4[i]; const float4 gg = g4[i];
      float* pa = &pp.x; float* va = &vv.x; const float* ga = &gg.x;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gr = ga[j] * gscale;
        va[j] = alpha * va[j] + (1.f - alpha) * gr * gr;
        pa[j] -= lr * gr / (sqrtf(va[j]) + eps);
      }
      p4[i] = pp; v4[i] = vv;
    }
    for (int64_t i = (n4 << 2) + tid; i < numel; i += stride) {
      const float gr = g[i] * gscale;
      const float vv = alpha * v[i] + (1.f - alpha) * gr * gr;
      v[i] = vv; p[i] -= lr * gr / (sqrtf(vv) + eps);
    }
  } else {
    for (int64_t i = tid; i < numel; i += stride) {
      const float gr = g[i] * gscale;
      const float vv = alpha * v[i] + (1.f - alpha) * gr * gr;
      v[i] = vv; p[i] -= lr * gr / (sqrtf(vv) + eps);
    }
  }
}

// ---- Philox4x32-10 (Salmon et al., SC'11), counter = (offset + element/4), key = seed
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
__device__ __forceinline__ void philox4(uint64_t seed, uint64_t ctr, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, k);
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * (1.f / 16777216.f); }  // [0,1)

template <bool NORMAL>
__global__ void __launch_bounds__(PW_NT)
philox_fill_kernel(float* __restrict__ out, int64_t numel, uint64_t seed, uint64_t offset) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int64_t quads = (numel + 3) >> 2;
  for (int64_t q = (int64_t)blockIdx.x * PW_NT + threadIdx.x; q < quads; q += (int64_t)gridDim.x * PW_NT) {
    uint32_t r[4];
    philox4(seed, offset + (uint64_t)q, r);
    float v[4];
    if (NORMAL) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float u1 = 1.f - u01(r[2 * h]);  // (0,1]
        const float u2 = u01(r[2 * h + 1]);
        const float rad = sqrtf(-2.f * logf(u1));
        float s, c;
        sincospif(2.f * u2, &s, &c);
        v[2 * h] = rad * c; v[2 * h + 1] = rad * s;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = u01(r[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (q * 4 + j < numel) out[q * 4 + j] = v[j];
  }
}

// ---- dropout (nn.Dropout / nn.Dropout2d, common/model.py:52-53, :344-346) -------------------------------------
// keep(e) for mask element e of a draw = u01(Philox4x32-10(key = seed, counter = {e / 4, stream})[e % 4]) >= p, with
// stream = *counter + call: `counter` lives in device memory and is advanced once per training iteration
// (glis_counter_add), `call` numbers the dropout calls inside an iteration — so a CUDA-graph replay draws a fresh
// mask every time, forward and backward of one call regenerate the SAME mask (nothing is stored), and the host can
// reproduce any mask from (seed, counter value, call).  Element mode: one mask element per tensor element (storage
// order); channel mode (Dropout2d): one per (n, c), tensor element i of an NHWC tensor -> (i / (HW*C)) * C + i % C.
__device__ __forceinline__ void philox4_stream(uint64_t seed, uint64_t quad, uint64_t stream, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)quad, (uint32_t)(quad >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, k);
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}

__global__ void __launch_bounds__(PW_NT)
dropout_elem_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t numel, float p, float keep_scale,
                    uint64_t seed, const uint64_t* __restrict__ counter, uint64_t call) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const uint64_t stream = (counter ? *counter : 0ull) + call;
  const int64_t quads = (numel + 3) >> 2;
  const bool vec = ((((uintptr_t)x) | ((uintptr_t)out)) & 15) == 0;
  for (int64_t q = (int64_t)blockIdx.x * PW_NT + threadIdx.x; q < quads; q += (int64_t)gridDim.x * PW_NT) {
    uint32_t r[4];
    philox4_stream(seed, (uint64_t)q, stream, r);
    if (vec && q * 4 + 3 < numel) {
      float4 v = reinterpret_cast<const float4*>(x)[q];
      v.x = u01(r[0]) >= p ? v.x * keep_scale : 0.f;
      v.y = u01(r[1]) >= p ? v.y * keep_scale : 0.f;
      v.z = u01(r[2]) >= p ? v.z * keep_scale : 0.f;
      v.w = u01(r[3]) >= p ? v.w * keep_scale : 0.f;
      reinterpret_cast<float4*>(out)[q] = v;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (q * 4 + j < numel) out[q * 4 + j] = u01(r[j]) >= p ? x[q * 4 + j] * keep_scale : 0.f;
    }
  }
}

// channel mode: `chan_stride` = elements per image (H*W*C for NHWC with inner == 1; C*inner images laid out NCHW use
// element -> (i / (C*inner)) * C + (i / inner) % C)
__global__ void __launch_bounds__(PW_NT)
dropout_channel_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t numel, int C, int64_t per_image,
                       int inner, float p, float keep_scale, uint64_t seed, const uint64_t* __restrict__ counter,
                       uint64_t call) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const uint64_t stream = (counter ? *counter : 0ull) + call;
  for (int64_t i = (int64_t)blockIdx.x * PW_NT + threadIdx.x; i < numel; i += (int64_t)gridDim.x * PW_NT) {
    const int64_t e = (i / per_image) * C + (i / inner) % C;
    uint32_t r[4];
    philox4_stream(seed, (uint64_t)(e >> 2), stream, r);
    const uint32_t w = (e & 3) == 0 ? r[0] : (e & 3) == 1 ? r[1] : (e & 3) == 2 ? r[2] : r[3];
    out[i] = u01(w) >= p ? x[i] * keep_scale : 0.f;
  }
}

__global__ void counter_add_kernel(uint64_t* counter, uint64_t inc) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait(); *counter += inc; }

}  // namespace glis

using namespace glis;

extern "C" int glis_tprelu_forward(const float* x, const float* a_raw, const float* b, float* out,
                                   int64_t numel, int C, int inner, void* stream) {
  GLIS_REQUIRE(x && a_raw && b && out, GLIS_E_BADARG, "glis_tprelu_forward: NULL pointer");
  GLIS_REQUIRE(numel >= 0 && C > 0 && inner > 0, GLIS_E_BADARG, "glis_tprelu_forward: bad sizes");
  if (numel == 0) return GLIS_OK;
  GLIS_LAUNCH(tprelu_fwd_kernel, dim3(pw_blocks(numel)), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), x, a_raw, b, out, numel, C, inner);
  GLIS_CHECK_LAUNCH("glis_tprelu_forward");
  return GLIS_OK;
}

extern "C" int glis_sigmoid_backward(const float* s, const float* dout, float* dy, int64_t numel, void* stream) {
  GLIS_REQUIRE(s && dout && dy && numel >= 0, GLIS_E_BADARG, "glis_sigmoid_backward: bad arguments");
  if (numel == 0) return GLIS_OK;
  GLIS_LAUNCH(sigmoid_bwd_kernel, dim3(pw_blocks(numel)), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), s, dout, dy, numel);
  GLIS_CHECK_LAUNCH("glis_sigmoid_backward");
  return GLIS_OK;
}

extern "C" int glis_tprelu_forward_planes(const float* x, const float* a_raw, const float* b, float* out,
                                          void* out_hi, void* out_lo, int64_t numel, int C, int act_channels,
                                          void* stream) {
  GLIS_REQUIRE(x && a_raw && b && (out || out_hi), GLIS_E_BADARG, "glis_tprelu_forward_planes: NULL pointer");
  GLIS_REQUIRE(numel >= 0 && C > 0 && act_channels >= 0, GLIS_E_BADARG, "glis_tprelu_forward_planes: bad sizes");
  GLIS_REQUIRE(C % 4 == 0 && numel % C == 0, GLIS_E_UNSUPPORTED,
               "glis_tprelu_forward_planes: C must be a multiple of 4 and numel a multiple of C");
  GLIS_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                   ((reinterpret_cast<uintptr_t>(out_hi) | reinterpret_cast<uintptr_t>(out_lo)) & 7) == 0,
               GLIS_E_BADARG, "glis_tprelu_forward_planes: misaligned buffer");
  if (numel == 0) return GLIS_OK;
  GLIS_LAUNCH(tprelu_fwd_planes_kernel, dim3(pw_blocks(numel, 8)), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), 
      x, 1, 0, a_raw, b, nullptr, out, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, numel, C,
      act_channels > 0 ? act_channels : C);
  GLIS_CHECK_LAUNCH("glis_tprelu_forward_planes");
  return GLIS_OK;
}

extern "C" int glis_tprelu_forward_planes_sum(const float* slabs, int nslabs, int64_t slab_stride, const float* a_raw,
                                              const float* b, float* preact, float* out, void* out_hi, void* out_lo,
                                              int64_t numel, int C, int act_channels, void* stream) {
  GLIS_REQUIRE(slabs && a_raw && b && (out || out_hi || preact), GLIS_E_BADARG,
               "glis_tprelu_forward_planes_sum: NULL pointer");
  GLIS_REQUIRE(numel >= 0 && C > 0 && act_channels >= 0 && nslabs >= 1 && (nslabs == 1 || slab_stride >= numel),
               GLIS_E_BADARG, "glis_tprelu_forward_planes_sum: bad sizes");
  GLIS_REQUIRE(C % 4 == 0 && numel % C == 0 && slab_stride % 4 == 0, GLIS_E_UNSUPPORTED,
               "glis_tprelu_forward_planes_sum: C and the slab stride must be multiples of 4, numel a multiple of C");
  GLIS_REQUIRE(((reinterpret_cast<uintptr_t>(slabs) | reinterpret_cast<uintptr_t>(out) |
                 reinterpret_cast<uintptr_t>(preact)) & 15) == 0 &&
                   ((reinterpret_cast<uintptr_t>(out_hi) | reinterpret_cast<uintptr_t>(out_lo)) & 7) == 0,
               GLIS_E_BADARG, "glis_tprelu_forward_planes_sum: misaligned buffer");
  if (numel == 0) return GLIS_OK;
  GLIS_LAUNCH(tprelu_fwd_planes_kernel, dim3(pw_blocks(numel, 8)), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), 
      slabs, nslabs, slab_stride, a_raw, b, preact, out, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, numel, C,
      act_channels > 0 ? act_channels : C);
  GLIS_CHECK_LAUNCH("glis_tprelu_forward_planes_sum");
  return GLIS_OK;
}

extern "C" int glis_tprelu_backward(const float* x, const float* a_raw, const float* b, const float* dout,
                                    float* dx, float* da, float* db, int64_t numel, int C, int inner,
                                    void* stream) {
  GLIS_REQUIRE(x && a_raw && b && dout && dx && da && db, GLIS_E_BADARG, "glis_tprelu_backward: NULL pointer");
  GLIS_REQUIRE(numel >= 0 && C > 0 && inner > 0, GLIS_E_BADARG, "glis_tprelu_backward: bad sizes");
  if (numel == 0) return GLIS_OK;
  GLIS_LAUNCH(tprelu_bwd_kernel, dim3(pw_blocks(numel, 8)), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), x, a_raw, b, dout, dx, nullptr, nullptr,
                                                                            da, db, numel, C, inner);
  GLIS_CHECK_LAUNCH("glis_tprelu_backward");
  return GLIS_OK;
}

extern "C" int glis_tprelu_backward_planes(const float* x, const float* a_raw, const float* b, const float* dout,
                                           float* dx, void* dx_hi, void* dx_lo, float* da, float* db, int64_t numel,
                                           int C, int inner, void* stream) {
  GLIS_REQUIRE(x && a_raw && b && dout && (dx || dx_hi) && ((da != nullptr) == (db != nullptr)), GLIS_E_BADARG,
               "glis_tprelu_backward_planes: NULL pointer");
  GLIS_REQUIRE(numel >= 0 && C > 0 && inner > 0, GLIS_E_BADARG, "glis_tprelu_backward_planes: bad sizes");
  if (numel == 0) return GLIS_OK;
  const bool aligned = ((((uintptr_t)x) | ((uintptr_t)dout) | ((uintptr_t)dx) | ((uintptr_t)a_raw) | ((uintptr_t)b)) & 15) == 0 &&
                       ((((uintptr_t)dx_hi) | ((uintptr_t)dx_lo)) & 7) == 0;
  if (inner == 1 && C % 4 == 0 && C / 4 <= PW_NT && aligned) {
    const int64_t rows = numel / C;
    const int rows_per_iter = PW_NT / (C / 4);
    // up to 8 row groups per thread (amortises the per-block sums) once that still gives every SM 4 blocks;
    // small tensors (the LIS module, the 5x5 maps) get one row group per thread: more, shorter blocks
    int64_t per = rows / ((int64_t)rows_per_iter * 148 * 4);
    per = per < 1 ? 1 : (per > 8 ? 8 : per);
    int64_t want = (rows + rows_per_iter * per - 1) / (rows_per_iter * per);
    const int blocks = (int)(want < 1 ? 1 : (want > 148 * 4 ? 148 * 4 : want));
    GLIS_LAUNCH(tprelu_bwd_nhwc_kernel, dim3(blocks), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), 
        x, a_raw, b, dout, dx, (__nv_bfloat16*)dx_hi, (__nv_bfloat16*)dx_lo, da, db, rows, C);
    GLIS_CHECK_LAUNCH("glis_tprelu_backward_planes(nhwc)");
    return GLIS_OK;
  }
  GLIS_LAUNCH(tprelu_bwd_kernel, dim3(pw_blocks(numel, 8)), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), 
      x, a_raw, b, dout, dx, (__nv_bfloat16*)dx_hi, (__nv_bfloat16*)dx_lo, da, db, numel, C, inner);
  GLIS_CHECK_LAUNCH("glis_tprelu_backward_planes");
  return GLIS_OK;
}

extern "C" int glis_channel_sum(const float* x, float* out, int64_t numel, int C, int inner, int accumulate,
                                void* stream) {
  GLIS_REQUIRE(x && out, GLIS_E_BADARG, "glis_channel_sum: NULL pointer");
  GLIS_REQUIRE(numel >= 0 && C > 0 && inner > 0, GLIS_E_BADARG, "glis_channel_sum: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    if (cudaMemsetAsync(out, 0, sizeof(float) * C, st) != cudaSuccess) {
      set_error("glis_channel_sum: memset failed");
      return GLIS_E_CUDA;
    }
  }
  if (numel == 0) return GLIS_OK;
  if (inner == 1 && C <= 4 && numel % C == 0) {
    const int64_t rows = numel / C;
    GLIS_LAUNCH(channel_sum_small_kernel, dim3(pw_blocks(rows, 8)), dim3(PW_NT), 0, (cudaStream_t)(st), x, out, rows, C);
    GLIS_CHECK_LAUNCH("glis_channel_sum(small C)");
    return GLIS_OK;
  }
  GLIS_LAUNCH(channel_sum_kernel, dim3(pw_blocks(numel, 8)), dim3(PW_NT), 0, (cudaStream_t)(st), x, out, numel, C, inner);
  GLIS_CHECK_LAUNCH("glis_channel_sum");
  return GLIS_OK;
}

extern "C" int glis_bce_logits(const float* logit, float target, int B, float gscale, float* loss, float* dlogit,
                               float* prob, void* stream) {
  GLIS_REQUIRE(logit && loss, GLIS_E_BADARG, "glis_bce_logits: NULL pointer");
  GLIS_REQUIRE(B > 0, GLIS_E_BADARG, "glis_bce_logits: empty batch");
  GLIS_LAUNCH(bce_logits_kernel, dim3(1), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), logit, target, B, gscale, loss, dlogit, prob);
  GLIS_CHECK_LAUNCH("glis_bce_logits");
  return GLIS_OK;
}

extern "C" int glis_lsq_logits(const float* logit, float target, int B, float gscale, float* loss, float* dlogit,
                               float* prob, void* stream) {
  GLIS_REQUIRE(logit && loss, GLIS_E_BADARG, "glis_lsq_logits: NULL pointer");
  GLIS_REQUIRE(B > 0, GLIS_E_BADARG, "glis_lsq_logits: empty batch");
  GLIS_LAUNCH(lsq_logits_kernel, dim3(1), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), logit, target, B, gscale, loss, dlogit, prob);
  GLIS_CHECK_LAUNCH("glis_lsq_logits");
  return GLIS_OK;
}

extern "C" int glis_dropout(const float* x, float* out, int64_t numel, int C, int inner, int64_t per_image,
                            int channel_mode, float p, uint64_t seed, const void* counter, uint64_t call,
                            void* stream) {
  GLIS_REQUIRE(x && out, GLIS_E_BADARG, "glis_dropout: NULL pointer");
  GLIS_REQUIRE(numel >= 0 && p >= 0.f && p < 1.f, GLIS_E_BADARG, "glis_dropout: bad size or probability (0 <= p < 1)");
  if (numel == 0) return GLIS_OK;
  const float keep_scale = 1.f / (1.f - p);
  if (channel_mode) {
    GLIS_REQUIRE(C > 0 && inner > 0 && per_image > 0 && per_image % ((int64_t)C * inner) == 0, GLIS_E_BADARG,
                 "glis_dropout: channel mode needs C, inner and the elements per image");
    GLIS_LAUNCH(dropout_channel_kernel, dim3(pw_blocks(numel, 8)), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), 
        x, out, numel, C, per_image, inner, p, keep_scale, seed, (const uint64_t*)counter, call);
  } else {
    GLIS_LAUNCH(dropout_elem_kernel, dim3(pw_blocks(numel, 16)), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), 
        x, out, numel, p, keep_scale, seed, (const uint64_t*)counter, call);
  }
  GLIS_CHECK_LAUNCH("glis_dropout");
  return GLIS_OK;
}

extern "C" int glis_counter_add(void* counter, uint64_t inc, void* stream) {
  GLIS_REQUIRE(counter, GLIS_E_BADARG, "glis_counter_add: NULL pointer");
  GLIS_LAUNCH(counter_add_kernel, dim3(1), dim3(1), 0, (cudaStream_t)((cudaStream_t)stream), (uint64_t*)counter, inc);
  GLIS_CHECK_LAUNCH("glis_counter_add");
  return GLIS_OK;
}

extern "C" int glis_mse_scaled(const float* u, const float* z, int64_t numel, float lambda, float* loss, float* du,
                               int accumulate, void* stream) {
  GLIS_REQUIRE(u && z && loss, GLIS_E_BADARG, "glis_mse_scaled: NULL pointer");
  GLIS_REQUIRE(numel > 0, GLIS_E_BADARG, "glis_mse_scaled: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(loss, 0, sizeof(float), st) != cudaSuccess) {
    set_error("glis_mse_scaled: memset failed");
    return GLIS_E_CUDA;
  }
  GLIS_LAUNCH(mse_scaled_kernel, dim3(pw_blocks(numel)), dim3(PW_NT), 0, (cudaStream_t)(st), u, z, numel, lambda, loss, du, accumulate);
  GLIS_CHECK_LAUNCH("glis_mse_scaled");
  return GLIS_OK;
}

extern "C" int glis_rmsprop(float* p, const float* g, float* v, int64_t numel, float lr, float alpha, float eps,
                            float gscale, void* stream) {
  GLIS_REQUIRE(p && g && v, GLIS_E_BADARG, "glis_rmsprop: NULL pointer");
  GLIS_REQUIRE(numel >= 0, GLIS_E_BADARG, "glis_rmsprop: negative size");
  if (numel == 0) return GLIS_OK;
  GLIS_LAUNCH(rmsprop_kernel, dim3(pw_blocks(numel, 8)), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), p, g, v, numel, lr, alpha, eps, gscale);
  GLIS_CHECK_LAUNCH("glis_rmsprop");
  return GLIS_OK;
}

extern "C" int glis_randn(float* out, int64_t numel, uint64_t seed, uint64_t offset, void* stream) {
  GLIS_REQUIRE(out && numel >= 0, GLIS_E_BADARG, "glis_randn: bad arguments");
  if (numel == 0) return GLIS_OK;
  GLIS_LAUNCH((philox_fill_kernel<true>), dim3(pw_blocks(numel, 16)), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), out, numel, seed, offset);
  GLIS_CHECK_LAUNCH("glis_randn");
  return GLIS_OK;
}

extern "C" int glis_uniform(float* out, int64_t numel, uint64_t seed, uint64_t offset, void* stream) {
  GLIS_REQUIRE(out && numel >= 0, GLIS_E_BADARG, "glis_uniform: bad arguments");
  if (numel == 0) return GLIS_OK;
  GLIS_LAUNCH((philox_fill_kernel<false>), dim3(pw_blocks(numel, 16)), dim3(PW_NT), 0, (cudaStream_t)((cudaStream_t)stream), out, numel, seed, offset);
  GLIS_CHECK_LAUNCH("glis_uniform");
  return GLIS_OK;
}

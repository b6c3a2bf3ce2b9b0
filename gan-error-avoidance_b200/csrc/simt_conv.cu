// fp32 FFMA gather-GEMM kernels: the exact-arithmetic path of every convolution-shaped
// contraction of the G-LIS step (conv / transposed conv / linear, forward, data gradient
// and weight gradient).  They serve (i) as the fp32 reference the tensor-core kernels are
// checked against on the device, and (ii) as the product kernels of the layers whose
// contraction is too thin for tcgen05 tiles (Cin=3, Cout=3, Cout=1 heads, batch-64 linears).
//
// Forward launch (glis_conv_forward):
//   out[n,oy,ox,co] = sum_{tap,ci} in[n, iy(oy,tap), ix(ox,tap), ci] * wpack[tap][ci][co]
// as a 64(pixels) x 64(channels) block tile with a 4x4 register micro-tile per thread and
// K = (tap,ci) walked in chunks of 16.  A transposed convolution is decomposed into
// stride_h*stride_w output phases (blockIdx.z) so that no multiply is spent on zeros.
#include "common.cuh"
#include "sm100.cuh"

namespace glis {

struct PhaseInfo {
  int ry, rx;    // first output row/col of this phase
  int Hq, Wq;    // number of output rows/cols of this phase
  int nth, ntw;  // taps of this phase along h / w
  int py, px;    // phase index (first tap) along h / w
};

__device__ __forceinline__ PhaseInfo decode_phase(const glis_geom_t& g, int z) {
  PhaseInfo p;
  if (g.relation == GLIS_CONV) {
    p.ry = p.rx = 0; p.Hq = g.Ho; p.Wq = g.Wo; p.nth = g.KH; p.ntw = g.KW; p.py = p.px = 0;
  } else {
    p.py = z / g.stride_w; p.px = z % g.stride_w;
    p.ry = ((p.py - g.pad_h) % g.stride_h + g.stride_h) % g.stride_h;
    p.rx = ((p.px - g.pad_w) % g.stride_w + g.stride_w) % g.stride_w;
    p.Hq = g.Ho > p.ry ? (g.Ho - p.ry + g.stride_h - 1) / g.stride_h : 0;
    p.Wq = g.Wo > p.rx ? (g.Wo - p.rx + g.stride_w - 1) / g.stride_w : 0;
    p.nth = g.KH > p.py ? (g.KH - p.py + g.stride_h - 1) / g.stride_h : 0;
    p.ntw = g.KW > p.px ? (g.KW - p.px + g.stride_w - 1) / g.stride_w : 0;
  }
  return p;
}

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

__global__ void __launch_bounds__(NT)
gather_gemm_fwd(const glis_geom_t g, const float* __restrict__ in, const float* __restrict__ wp,
                const glis_epilogue_t ep, float* __restrict__ out, int ksplit, int k_per_split) {
  // blockIdx.z = phase * ksplit + split; with ksplit > 1 the partial sums are added atomically
  // into a zero-filled `out` (only used when the epilogue is bias-only).
  const int split = blockIdx.z % ksplit;
  const PhaseInfo ph = decode_phase(g, blockIdx.z / ksplit);
  const int P = g.N * ph.Hq * ph.Wq;  // output pixels of this phase
  const int m0 = blockIdx.x * BM;
  if (m0 >= P) return;
  const int n0 = blockIdx.y * BN;
  const int ntaps = ph.nth * ph.ntw;
  const int Kall = ntaps * g.Ci;
  const int kbeg = split * k_per_split;
  const int K = min(Kall, kbeg + k_per_split);
  if (kbeg >= K && ksplit > 1) return;

  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  // ---- A-load role: k = tid % 16, pixels tid/16 + 16 j
  const int a_k = tid & 15;
  int a_n[4], a_oy[4], a_ox[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int m = m0 + (tid >> 4) + 16 * j;
    if (m < P) {
      int qx = m % ph.Wq; int t = m / ph.Wq; int qy = t % ph.Hq; a_n[j] = t / ph.Hq;
      if (g.relation == GLIS_CONV) { a_oy[j] = qy; a_ox[j] = qx; }
      else { a_oy[j] = qy * g.stride_h + ph.ry; a_ox[j] = qx * g.stride_w + ph.rx; }
    } else { a_n[j] = -1; a_oy[j] = a_ox[j] = 0; }
  }
  // ---- B-load role: col = tid % 64, rows tid/64 + 4 j
  const int b_c = tid & 63, b_r = tid >> 6;

  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < K; k0 += BK) {
    {  // A tile
      const int kk = k0 + a_k;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (kk < K) {
        const int tap = kk / g.Ci, ci = kk - tap * g.Ci;
        const int jh = tap / ph.ntw, jw = tap - jh * ph.ntw;
        int kh, kw;
        if (g.relation == GLIS_CONV) { kh = jh; kw = jw; }
        else { kh = ph.py + jh * g.stride_h; kw = ph.px + jw * g.stride_w; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (a_n[j] < 0) continue;
          int iy, ix;
          if (g.relation == GLIS_CONV) {
            iy = a_oy[j] * g.stride_h - g.pad_h + kh * g.dil_h;
            ix = a_ox[j] * g.stride_w - g.pad_w + kw * g.dil_w;
          } else {
            iy = (a_oy[j] + g.pad_h - kh) / g.stride_h;  // exact inside a phase
            ix = (a_ox[j] + g.pad_w - kw) / g.stride_w;
            if (a_oy[j] + g.pad_h - kh < 0) iy = -1;
            if (a_ox[j] + g.pad_w - kw < 0) ix = -1;
          }
          if (iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi)
            v[j] = __ldg(in + (((int64_t)a_n[j] * g.Hi + iy) * g.Wi + ix) * g.Ci + ci);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) As[a_k][(tid >> 4) + 16 * j] = v[j];
    }
    {  // B tile
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = b_r + 4 * j, kk = k0 + r, co = n0 + b_c;
        float v = 0.f;
        if (kk < K && co < g.Co) {
          const int tap = kk / g.Ci, ci = kk - tap * g.Ci;
          const int jh = tap / ph.ntw, jw = tap - jh * ph.ntw;
          int kh, kw;
          if (g.relation == GLIS_CONV) { kh = jh; kw = jw; }
          else { kh = ph.py + jh * g.stride_h; kw = ph.px + jw * g.stride_w; }
          v = __ldg(wp + ((int64_t)(kh * g.KW + kw) * g.Ci + ci) * g.Co + co);
        }
        Bs[r][b_c] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue: bias, activation, optional pre-activation save
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= P) continue;
    int qx = m % ph.Wq; int t = m / ph.Wq; int qy = t % ph.Hq; int n = t / ph.Hq;
    int oy = qy, ox = qx;
    if (g.relation == GLIS_TCONV) { oy = qy * g.stride_h + ph.ry; ox = qx * g.stride_w + ph.rx; }
    const int64_t base = (((int64_t)n * g.Ho + oy) * g.Wo + ox) * g.Co;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= g.Co) continue;
      float y = acc[i][j];
      if (ksplit > 1) {
        if (ep.bias && split == 0) y += __ldg(ep.bias + co);
        atomicAdd(out + base + co, y);
        continue;
      }
      if (ep.bias) y += __ldg(ep.bias + co);
      if (ep.preact) ep.preact[base + co] = y;
      float o = y;
      if (ep.act == GLIS_ACT_TPRELU) {
        const float b = __ldg(ep.act_b + co), a = fminf(fmaxf(__ldg(ep.act_a + co), 0.f), 1.f);
        const float tt = y - b;
        o = (tt > 0.f ? tt : a * tt) + b;
      } else if (ep.act == GLIS_ACT_SIGMOID) {
        o = 1.f / (1.f + expf(-y));
      }
      out[base + co] = o;
      if (ep.out_hi) {
        __nv_bfloat16 h, l;
        sm100::split_bf16(o, h, l);
        reinterpret_cast<__nv_bfloat16*>(ep.out_hi)[base + co] = h;
        if (ep.out_lo) reinterpret_cast<__nv_bfloat16*>(ep.out_lo)[base + co] = l;
      }
    }
  }
}

// Forward launch for image-side layers with at most 4 output channels and many pixels
// (G's last transposed conv 64 -> 3, D's first-layer data gradient 64 -> 3).  The 64x64 tile
// above would idle 61 of 64 columns; here 8 lanes share one output pixel, each owning an
// interleaved quarter-cache-line slice of the input channels (coalesced 128-byte reads per
// tap), the partial sums meet in three shuffles and lane 0 applies the epilogue.
constexpr int SC_NT = 256;
__global__ void __launch_bounds__(SC_NT)
small_cout_fwd(const glis_geom_t g, const float* __restrict__ in, const float* __restrict__ wp,
               const glis_epilogue_t ep, float* __restrict__ out) {
  // [tap][j][c4] -> (co0..co3) of input channel 4*c4 + j: the 8 lanes of a pixel group read 8
  // consecutive float4 (bank-conflict free), all groups on the same tap read the same ones.
  extern __shared__ float4 w4[];
  const int T = g.KH * g.KW, C4 = g.Ci / 4;
  for (int i = threadIdx.x; i < T * g.Ci; i += SC_NT) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* src = wp + (int64_t)i * g.Co;
    v.x = __ldg(src);
    if (g.Co > 1) v.y = __ldg(src + 1);
    if (g.Co > 2) v.z = __ldg(src + 2);
    if (g.Co > 3) v.w = __ldg(src + 3);
    const int tap = i / g.Ci, ci = i - tap * g.Ci;
    w4[(tap * 4 + (ci & 3)) * C4 + (ci >> 2)] = v;
  }
  __syncthreads();
  const int sub = threadIdx.x & 7;                   // lane within the pixel group
  const int64_t P = (int64_t)g.N * g.Ho * g.Wo;
  const bool phase_major = g.relation == GLIS_TCONV && g.Ho % g.stride_h == 0 && g.Wo % g.stride_w == 0;
  const int64_t groups = (int64_t)gridDim.x * (SC_NT / 8);
  for (int64_t pix = (int64_t)blockIdx.x * (SC_NT / 8) + (threadIdx.x >> 3);; pix += groups) {
    // all 8 lanes of a group share `pix`; whole warps leave together only when every group is done
    const bool live = pix < P;
    if (__all_sync(0xffffffffu, !live)) break;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int n = 0, oy = 0, ox = 0;
    if (live) {
      if (phase_major) {
        // pixels of one output phase are consecutive: every group of a warp then walks the same
        // taps (shared-memory weight reads become broadcasts) and neighbouring input pixels
        const int Hq = g.Ho / g.stride_h, Wq = g.Wo / g.stride_w;
        const int64_t per_phase = (int64_t)g.N * Hq * Wq;
        const int phase = (int)(pix / per_phase);
        const int64_t r = pix - (int64_t)phase * per_phase;
        const int qx = (int)(r % Wq); const int64_t t = r / Wq; const int qy = (int)(t % Hq); n = (int)(t / Hq);
        oy = qy * g.stride_h + phase / g.stride_w; ox = qx * g.stride_w + phase % g.stride_w;
      } else {
        ox = (int)(pix % g.Wo); const int64_t t = pix / g.Wo; oy = (int)(t % g.Ho); n = (int)(t / g.Ho);
      }
      int kh0 = 0, kw0 = 0, sth = 1, stw = 1;
      if (g.relation == GLIS_TCONV) {
        kh0 = (oy + g.pad_h) % g.stride_h; kw0 = (ox + g.pad_w) % g.stride_w; sth = g.stride_h; stw = g.stride_w;
      }
      for (int kh = kh0; kh < g.KH; kh += sth) {
        int iy;
        if (g.relation == GLIS_CONV) iy = oy * g.stride_h - g.pad_h + kh * g.dil_h;
        else iy = (oy + g.pad_h - kh) / g.stride_h;  // numerator >= 0 is checked below
        if (g.relation == GLIS_TCONV && oy + g.pad_h - kh < 0) continue;
        if (iy < 0 || iy >= g.Hi) continue;
        for (int kw = kw0; kw < g.KW; kw += stw) {
          int ix;
          if (g.relation == GLIS_CONV) ix = ox * g.stride_w - g.pad_w + kw * g.dil_w;
          else ix = (ox + g.pad_w - kw) / g.stride_w;
          if (g.relation == GLIS_TCONV && ox + g.pad_w - kw < 0) continue;
          if (ix < 0 || ix >= g.Wi) continue;
          const float4* src = reinterpret_cast<const float4*>(in + (((int64_t)n * g.Hi + iy) * g.Wi + ix) * g.Ci);
          const float4* wt = w4 + (kh * g.KW + kw) * g.Ci;
          for (int c4 = sub; c4 < C4; c4 += 8) {
            const float4 x = __ldg(src + c4);
            const float4 w0 = wt[c4], w1 = wt[C4 + c4], w2 = wt[2 * C4 + c4], w3 = wt[3 * C4 + c4];
            a0 = fmaf(x.x, w0.x, a0); a1 = fmaf(x.x, w0.y, a1); a2 = fmaf(x.x, w0.z, a2); a3 = fmaf(x.x, w0.w, a3);
            a0 = fmaf(x.y, w1.x, a0); a1 = fmaf(x.y, w1.y, a1); a2 = fmaf(x.y, w1.z, a2); a3 = fmaf(x.y, w1.w, a3);
            a0 = fmaf(x.z, w2.x, a0); a1 = fmaf(x.z, w2.y, a1); a2 = fmaf(x.z, w2.z, a2); a3 = fmaf(x.z, w2.w, a3);
            a0 = fmaf(x.w, w3.x, a0); a1 = fmaf(x.w, w3.y, a1); a2 = fmaf(x.w, w3.z, a2); a3 = fmaf(x.w, w3.w, a3);
          }
        }
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
    if (live && sub == 0) {
      const float acc[4] = {a0, a1, a2, a3};
      const int64_t base = (((int64_t)n * g.Ho + oy) * g.Wo + ox) * g.Co;
      for (int co = 0; co < g.Co; ++co) {
        float y = acc[co];
        if (ep.bias) y += __ldg(ep.bias + co);
        if (ep.preact) ep.preact[base + co] = y;
        float o = y;
        if (ep.act == GLIS_ACT_TPRELU) {
          const float b = __ldg(ep.act_b + co), a = fminf(fmaxf(__ldg(ep.act_a + co), 0.f), 1.f);
          const float tt = y - b;
          o = (tt > 0.f ? tt : a * tt) + b;
        } else if (ep.act == GLIS_ACT_SIGMOID) {
          o = 1.f / (1.f + expf(-y));
        }
        out[base + co] = o;
      }
    }
  }
}

// The case that matters (G level 0 / D level-0 data gradient): 4x4 kernel, stride 2, transposed
// relation, CI input channels.  Every output pixel has exactly 2x2 taps; all of a lane's input
// loads (CI/32 float4 per tap) are issued before the first FMA so that one pixel keeps
// 4*CI/32 128-bit loads in flight.
template <int CI>
__global__ void __launch_bounds__(SC_NT)
small_cout_tconv4x4s2(const glis_geom_t g, const float* __restrict__ in, const float* __restrict__ wp,
                      const glis_epilogue_t ep, float* __restrict__ out) {
  constexpr int C4 = CI / 4, R = CI / 32;
  __shared__ float4 w4[16 * CI];   // [tap][j][c4]
  for (int i = threadIdx.x; i < 16 * CI; i += SC_NT) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* src = wp + (int64_t)i * g.Co;
    v.x = __ldg(src);
    if (g.Co > 1) v.y = __ldg(src + 1);
    if (g.Co > 2) v.z = __ldg(src + 2);
    if (g.Co > 3) v.w = __ldg(src + 3);
    const int tap = i / CI, ci = i - tap * CI;
    w4[(tap * 4 + (ci & 3)) * C4 + (ci >> 2)] = v;
  }
  __syncthreads();
  const int sub = threadIdx.x & 7;
  const int Hq = g.Ho / 2, Wq = g.Wo / 2;
  const int64_t per_phase = (int64_t)g.N * Hq * Wq, P = 4 * per_phase;
  const int64_t groups = (int64_t)gridDim.x * (SC_NT / 8);
  for (int64_t pix = (int64_t)blockIdx.x * (SC_NT / 8) + (threadIdx.x >> 3);; pix += groups) {
    const bool live = pix < P;
    if (__all_sync(0xffffffffu, !live)) break;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int n = 0, oy = 0, ox = 0;
    if (live) {
      const int phase = (int)(pix / per_phase);
      const int64_t r = pix - (int64_t)phase * per_phase;
      const int qx = (int)(r % Wq); const int64_t t = r / Wq; const int qy = (int)(t % Hq); n = (int)(t / Hq);
      oy = 2 * qy + (phase >> 1); ox = 2 * qx + (phase & 1);
      const int kh0 = (oy + g.pad_h) & 1, kw0 = (ox + g.pad_w) & 1;
      const int iy0 = (oy + g.pad_h - kh0) >> 1, ix0 = (ox + g.pad_w - kw0) >> 1;   // taps kh0, kw0
      float4 xv[4][R];
      int tapi[4];
#pragma unroll
      for (int th = 0; th < 2; ++th)
#pragma unroll
        for (int tw = 0; tw < 2; ++tw) {
          const int iy = iy0 - th, ix = ix0 - tw;     // taps kh0 + 2*th, kw0 + 2*tw
          const bool ok = iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi;
          const float4* src = reinterpret_cast<const float4*>(in + (((int64_t)n * g.Hi + iy) * g.Wi + ix) * CI);
          tapi[th * 2 + tw] = (kh0 + 2 * th) * 4 + (kw0 + 2 * tw);
#pragma unroll
          for (int rr = 0; rr < R; ++rr)
            xv[th * 2 + tw][rr] = ok ? __ldg(src + sub + 8 * rr) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int tp = 0; tp < 4; ++tp) {
        const float4* wt = w4 + tapi[tp] * CI;
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
          const int c4 = sub + 8 * rr;
          const float4 x = xv[tp][rr];
          const float4 w0 = wt[c4], w1 = wt[C4 + c4], w2 = wt[2 * C4 + c4], w3 = wt[3 * C4 + c4];
          a0 = fmaf(x.x, w0.x, a0); a1 = fmaf(x.x, w0.y, a1); a2 = fmaf(x.x, w0.z, a2); a3 = fmaf(x.x, w0.w, a3);
          a0 = fmaf(x.y, w1.x, a0); a1 = fmaf(x.y, w1.y, a1); a2 = fmaf(x.y, w1.z, a2); a3 = fmaf(x.y, w1.w, a3);
          a0 = fmaf(x.z, w2.x, a0); a1 = fmaf(x.z, w2.y, a1); a2 = fmaf(x.z, w2.z, a2); a3 = fmaf(x.z, w2.w, a3);
          a0 = fmaf(x.w, w3.x, a0); a1 = fmaf(x.w, w3.y, a1); a2 = fmaf(x.w, w3.z, a2); a3 = fmaf(x.w, w3.w, a3);
        }
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o);
      a2 += __shfl_xor_sync(0xffffffffu, a2, o); a3 += __shfl_xor_sync(0xffffffffu, a3, o);
    }
    if (live && sub < g.Co) {   // lane c of the group finishes output channel c
      const float acc = sub == 0 ? a0 : (sub == 1 ? a1 : (sub == 2 ? a2 : a3));
      const int64_t idx = (((int64_t)n * g.Ho + oy) * g.Wo + ox) * g.Co + sub;
      float y = acc;
      if (ep.bias) y += __ldg(ep.bias + sub);
      if (ep.preact) ep.preact[idx] = y;
      float o = y;
      if (ep.act == GLIS_ACT_TPRELU) {
        const float b = __ldg(ep.act_b + sub), a = fminf(fmaxf(__ldg(ep.act_a + sub), 0.f), 1.f);
        const float tt = y - b;
        o = (tt > 0.f ? tt : a * tt) + b;
      } else if (ep.act == GLIS_ACT_SIGMOID) {
        o = 1.f / (1.f + expf(-y));
      }
      out[idx] = o;
    }
  }
}

// Forward launch for the image-side conv with <= 4 input channels (D / R level 0: 3 -> 64).
// K = taps * Cin is only 48, so the layer is bound by writing its output.  One thread owns one
// output pixel and all (<= 64) output channels: the 48 input values sit in registers, the
// weights are read from shared memory as warp-wide broadcasts, and the thread writes whole
// 256-byte runs of its pixel (fp32 output, pre-activation, bf16 planes).
constexpr int SI_NT = 128, SI_MAXK = 64, SI_MAXCO = 64;
// TKH/TKW/TCI > 0: kernel extent known at compile time (the 4x4x3 case) so the input gather
// indexes registers directly; 0: runtime extent.
template <int TKH, int TKW, int TCI>
__global__ void __launch_bounds__(SI_NT)
small_cin_fwd(const glis_geom_t g, const float* __restrict__ in, const float* __restrict__ wp,
              const glis_epilogue_t ep, float* __restrict__ out) {
  __shared__ __align__(16) float Ws[SI_MAXK * SI_MAXCO];   // [k = tap*Ci + ci][co]
  __shared__ float s_bias[SI_MAXCO], s_a[SI_MAXCO], s_b[SI_MAXCO];
  const int K = g.KH * g.KW * g.Ci;
  for (int i = threadIdx.x; i < K * g.Co; i += SI_NT) Ws[(i / g.Co) * SI_MAXCO + (i % g.Co)] = __ldg(wp + i);
  for (int c = threadIdx.x; c < g.Co; c += SI_NT) {
    s_bias[c] = ep.bias ? __ldg(ep.bias + c) : 0.f;
    s_a[c] = ep.act == GLIS_ACT_TPRELU ? fminf(fmaxf(__ldg(ep.act_a + c), 0.f), 1.f) : 0.f;
    s_b[c] = ep.act == GLIS_ACT_TPRELU ? __ldg(ep.act_b + c) : 0.f;
  }
  __syncthreads();
  const int64_t P = (int64_t)g.N * g.Ho * g.Wo;
  __shared__ __align__(16) float s_tile[SI_NT / 32][32 * 20];
  const int64_t pix = (int64_t)blockIdx.x * SI_NT + threadIdx.x;
  const bool live = pix < P;   // dead lanes still help their warp store
  const int64_t pc = live ? pix : P - 1;
  const int ox = (int)(pc % g.Wo); const int64_t t = pc / g.Wo; const int oy = (int)(t % g.Ho); const int n = (int)(t / g.Ho);

  float x[SI_MAXK];
#pragma unroll
  for (int k = 0; k < SI_MAXK; ++k) x[k] = 0.f;
  if (TKH > 0) {
#pragma unroll
    for (int kh = 0; kh < TKH; ++kh) {
      const int iy = oy * g.stride_h - g.pad_h + kh * g.dil_h;
#pragma unroll
      for (int kw = 0; kw < TKW; ++kw) {
        const int ix = ox * g.stride_w - g.pad_w + kw * g.dil_w;
        const bool ok = iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi;
        const float* src = in + (((int64_t)n * g.Hi + iy) * g.Wi + ix) * TCI;
#pragma unroll
        for (int ci = 0; ci < TCI; ++ci) x[(kh * TKW + kw) * TCI + ci] = ok ? __ldg(src + ci) : 0.f;
      }
    }
  } else {
    int k = 0;
    for (int kh = 0; kh < g.KH; ++kh) {
      const int iy = oy * g.stride_h - g.pad_h + kh * g.dil_h;
      for (int kw = 0; kw < g.KW; ++kw) {
        const int ix = ox * g.stride_w - g.pad_w + kw * g.dil_w;
        const bool ok = iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi;
        const float* src = in + (((int64_t)n * g.Hi + iy) * g.Wi + ix) * g.Ci;
        for (int ci = 0; ci < g.Ci; ++ci, ++k) {
          const float v = ok ? __ldg(src + ci) : 0.f;
          // registers need compile-time indices: select into the unrolled array
#pragma unroll
          for (int kk = 0; kk < SI_MAXK; ++kk) if (kk == k) x[kk] = v;
        }
      }
    }
  }
  for (int c0 = 0; c0 < g.Co; c0 += 16) {       // 16 channels at a time keeps the accumulators in registers
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < SI_MAXK; ++k) {
      if (k < K) {
        const float4* wrow = reinterpret_cast<const float4*>(Ws + k * SI_MAXCO + c0);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 w = wrow[q];
          acc[4 * q + 0] = fmaf(x[k], w.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(x[k], w.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(x[k], w.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(x[k], w.w, acc[4 * q + 3]);
        }
      }
    }
    float y[16], o[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      y[j] = acc[j] + s_bias[c0 + j];
      o[j] = y[j];
      if (ep.act == GLIS_ACT_TPRELU) { const float tt = y[j] - s_b[c0 + j]; o[j] = (tt > 0.f ? tt : s_a[c0 + j] * tt) + s_b[c0 + j]; }
      else if (ep.act == GLIS_ACT_SIGMOID) o[j] = 1.f / (1.f + expf(-y[j]));
    }
    // ---- coalesced stores: the warp's 32 pixels x 16 channels go through a padded shared tile so
    // that 4 consecutive lanes write one pixel's 64 contiguous bytes (full 32-byte sectors)
    float* tile = s_tile[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int64_t wbase = (pix - lane) * g.Co + c0;     // first pixel of this warp
    const int64_t wpix_left = P - (pix - lane);          // valid pixels in this warp
    if (ep.preact) {
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(tile + lane * 20)[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = i * 32 + lane, p = idx >> 2, q = idx & 3;
        if (p < wpix_left) reinterpret_cast<float4*>(ep.preact + wbase + (int64_t)p * g.Co)[q] = reinterpret_cast<const float4*>(tile + p * 20)[q];
      }
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(tile + lane * 20)[q] = make_float4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = i * 32 + lane, p = idx >> 2, q = idx & 3;
      if (p < wpix_left) reinterpret_cast<float4*>(out + wbase + (int64_t)p * g.Co)[q] = reinterpret_cast<const float4*>(tile + p * 20)[q];
    }
    if (ep.out_hi) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int idx = i * 32 + lane, p = idx >> 1, h8 = idx & 1;   // 8 channels = 16 bytes of bf16 per lane
        if (p < wpix_left) {
          const float4 v0 = reinterpret_cast<const float4*>(tile + p * 20)[2 * h8], v1 = reinterpret_cast<const float4*>(tile + p * 20)[2 * h8 + 1];
          const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          __align__(16) __nv_bfloat16 hh[8], ll[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) sm100::split_bf16(vv[j], hh[j], ll[j]);
          const int64_t off = wbase + (int64_t)p * g.Co + h8 * 8;
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out_hi) + off) = *reinterpret_cast<uint4*>(hh);
          if (ep.out_lo) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out_lo) + off) = *reinterpret_cast<uint4*>(ll);
        }
      }
    }
  }
}

// Weight gradient: G[a][b][tap] += sum_pix small[pix][a] * big[gather(pix,tap)][b]
// Block tile 64 (a) x 64 (flattened tap*Cb + b), K = pixels of `small`, split over blockIdx.z.
__global__ void __launch_bounds__(NT)
gather_gemm_wgrad(const glis_geom_t g, const float* __restrict__ small, const float* __restrict__ big,
                  float* __restrict__ G, int pix_per_split) {
  const int Ca = g.Co, Cb = g.Ci, T = g.KH * g.KW;
  const int P = g.N * g.Ho * g.Wo;
  const int a0 = blockIdx.x * BM, c0 = blockIdx.y * BN;
  const int p_begin = blockIdx.z * pix_per_split;
  const int p_end = min(P, p_begin + pix_per_split);
  if (p_begin >= p_end) return;

  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int l_c = tid & 63, l_r = tid >> 6;  // both tiles: col = tid%64, rows tid/64 + 4j
  // B column decode (fixed for the block's lifetime)
  const int nn = c0 + l_c;
  const bool b_ok = nn < T * Cb;
  const int tap = b_ok ? nn / Cb : 0, bch = b_ok ? nn - tap * Cb : 0;
  const int kh = tap / g.KW, kw = tap - kh * g.KW;
  const bool a_ok = (a0 + l_c) < Ca;

  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int p0 = p_begin; p0 < p_end; p0 += BK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = l_r + 4 * j, pix = p0 + r;
      float va = 0.f, vb = 0.f;
      if (pix < p_end) {
        if (a_ok) va = __ldg(small + (int64_t)pix * Ca + a0 + l_c);
        if (b_ok) {
          const int ox = pix % g.Wo; const int t = pix / g.Wo; const int oy = t % g.Ho; const int n = t / g.Ho;
          const int iy = oy * g.stride_h - g.pad_h + kh * g.dil_h;
          const int ix = ox * g.stride_w - g.pad_w + kw * g.dil_w;
          if (iy >= 0 && iy < g.Hi && ix >= 0 && ix < g.Wi)
            vb = __ldg(big + (((int64_t)n * g.Hi + iy) * g.Wi + ix) * Cb + bch);
        }
      }
      As[r][l_c] = va;
      Bs[r][l_c] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = a0 + ty * 4 + i;
    if (a >= Ca) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c >= T * Cb) continue;
      const int tp = c / Cb, b = c - tp * Cb;
      atomicAdd(G + ((int64_t)a * Cb + b) * T + tp, acc[i][j]);
    }
  }
}

int validate_geom(const glis_geom_t* g, const char* who) {
  GLIS_REQUIRE(g != nullptr, GLIS_E_BADARG, "%s: geometry is NULL", who);
  GLIS_REQUIRE(g->relation == GLIS_CONV || g->relation == GLIS_TCONV, GLIS_E_BADARG,
               "%s: unknown relation %d", who, g->relation);
  GLIS_REQUIRE(g->N > 0 && g->Hi > 0 && g->Wi > 0 && g->Ci > 0 && g->Ho > 0 && g->Wo > 0 && g->Co > 0,
               GLIS_E_BADARG, "%s: non-positive tensor extent", who);
  GLIS_REQUIRE(g->KH > 0 && g->KW > 0 && g->stride_h > 0 && g->stride_w > 0 && g->dil_h > 0 && g->dil_w > 0 &&
                   g->pad_h >= 0 && g->pad_w >= 0,
               GLIS_E_BADARG, "%s: bad kernel/stride/pad/dilation", who);
  if (g->relation == GLIS_TCONV)
    GLIS_REQUIRE(g->dil_h == 1 && g->dil_w == 1, GLIS_E_UNSUPPORTED, "%s: dilated transposed relation", who);
  const int64_t big = (int64_t)1 << 31;
  GLIS_REQUIRE((int64_t)g->N * g->Hi * g->Wi < big && (int64_t)g->N * g->Ho * g->Wo < big, GLIS_E_UNSUPPORTED,
               "%s: more than 2^31 pixels", who);
  return GLIS_OK;
}

bool is_linear_geom(const glis_geom_t* g);
bool is_head_dgrad_geom(const glis_geom_t* g);
int simt_linear_forward(const glis_geom_t* g, const float* in, const float* wpack, const glis_epilogue_t* ep,
                        float* out, cudaStream_t st);

int simt_conv_forward(const glis_geom_t* g, const float* in, const float* wpack, const glis_epilogue_t* ep,
                      float* out, cudaStream_t st) {
  if (is_linear_geom(g) || is_head_dgrad_geom(g)) {
    const int rc = simt_linear_forward(g, in, wpack, ep, out, st);
    if (rc != GLIS_E_UNSUPPORTED) return rc;
  }
  GLIS_REQUIRE(ep->act_channels <= 0, GLIS_E_UNSUPPORTED,
               "glis_conv_forward: act_channels is only implemented by the linear and tensor-core kernels");
  if (g->relation == GLIS_CONV && g->Ci <= 4 && g->KH * g->KW * g->Ci <= SI_MAXK && g->Co <= SI_MAXCO &&
      g->Co % 16 == 0 && (int64_t)g->N * g->Ho * g->Wo >= 4096 &&
      ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(ep->preact) |
        reinterpret_cast<uintptr_t>(ep->out_hi) | reinterpret_cast<uintptr_t>(ep->out_lo)) & 15) == 0) {
    const int64_t pixels = (int64_t)g->N * g->Ho * g->Wo;
    const unsigned blocks = (unsigned)((pixels + SI_NT - 1) / SI_NT);
    if (g->KH == 4 && g->KW == 4 && g->Ci == 3)
      small_cin_fwd<4, 4, 3><<<blocks, SI_NT, 0, st>>>(*g, in, wpack, *ep, out);
    else
      small_cin_fwd<0, 0, 0><<<blocks, SI_NT, 0, st>>>(*g, in, wpack, *ep, out);
    GLIS_CHECK_LAUNCH("glis_conv_forward(fp32, small Cin)");
    return GLIS_OK;
  }
  {
    const int64_t pixels = (int64_t)g->N * g->Ho * g->Wo;
    const size_t wsmem = (size_t)g->KH * g->KW * g->Ci * sizeof(float4);
    if (g->Co <= 4 && g->Ci % 32 == 0 && pixels >= 4096 && wsmem <= 96 * 1024 &&
        (reinterpret_cast<uintptr_t>(in) & 15) == 0) {
      (void)ensure_max_dynamic_smem(reinterpret_cast<const void*>(small_cout_fwd), 96 * 1024);
      int64_t want = (pixels + SC_NT / 8 - 1) / (SC_NT / 8);
      const int blocks = (int)(want < 148 * 8 ? want : 148 * 8);
      if (g->relation == GLIS_TCONV && g->KH == 4 && g->KW == 4 && g->stride_h == 2 && g->stride_w == 2 &&
          g->Ci == 64 && g->Ho % 2 == 0 && g->Wo % 2 == 0 && g->pad_h <= 2 && g->pad_w <= 2) {
        small_cout_tconv4x4s2<64><<<blocks, SC_NT, 0, st>>>(*g, in, wpack, *ep, out);
        GLIS_CHECK_LAUNCH("glis_conv_forward(fp32, small Cout 4x4s2)");
        return GLIS_OK;
      }
      small_cout_fwd<<<blocks, SC_NT, wsmem, st>>>(*g, in, wpack, *ep, out);
      GLIS_CHECK_LAUNCH("glis_conv_forward(fp32, small Cout)");
      return GLIS_OK;
    }
  }
  int nphase = 1, maxP = g->N * g->Ho * g->Wo;
  if (g->relation == GLIS_TCONV) {
    nphase = g->stride_h * g->stride_w;
    const int Hq = cdiv(g->Ho, g->stride_h), Wq = cdiv(g->Wo, g->stride_w);
    maxP = g->N * Hq * Wq;
  }
  // Thin grids with a long contraction (batch-sized linears, the 1-channel head): split K.
  int ksplit = 1, k_per_split = 1 << 30;
  const int blocks = cdiv(maxP, BM) * cdiv(g->Co, BN) * nphase;
  const int Kmax = cdiv(g->KH, g->relation == GLIS_TCONV ? g->stride_h : 1) *
                   cdiv(g->KW, g->relation == GLIS_TCONV ? g->stride_w : 1) * g->Ci;
  if (blocks < 74 && Kmax >= 1024 && ep->act == GLIS_ACT_NONE && !ep->preact) {
    ksplit = min(cdiv(Kmax, 256), cdiv(148 * 2, blocks));
    k_per_split = cdiv(cdiv(Kmax, ksplit), BK) * BK;
    ksplit = cdiv(Kmax, k_per_split);
    if (ksplit > 1) {
      cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)g->N * g->Ho * g->Wo * g->Co, st);
      GLIS_REQUIRE(e == cudaSuccess, GLIS_E_CUDA, "glis_conv_forward: memset failed: %s", cudaGetErrorString(e));
    } else {
      k_per_split = 1 << 30;
    }
  }
  dim3 grid(cdiv(maxP, BM), cdiv(g->Co, BN), nphase * ksplit);
  gather_gemm_fwd<<<grid, NT, 0, st>>>(*g, in, wpack, *ep, out, ksplit, k_per_split);
  GLIS_CHECK_LAUNCH("glis_conv_forward(fp32)");
  return GLIS_OK;
}

int simt_linear_wgrad(const glis_geom_t* g, const float* small, const float* big, float* G, cudaStream_t st);

int simt_conv_wgrad(const glis_geom_t* g, const float* small, const float* big, float* G, cudaStream_t st) {
  if (g->relation == GLIS_CONV && is_linear_geom(g)) {
    const int rc = simt_linear_wgrad(g, small, big, G, st);
    if (rc != GLIS_E_UNSUPPORTED) return rc;
  }
  const int T = g->KH * g->KW;
  const int P = g->N * g->Ho * g->Wo;
  const int tiles = cdiv(g->Co, BM) * cdiv((int64_t)T * g->Ci, BN);
  // enough K-splits to fill the machine a few times over, at least 64 pixels each
  int splits = (148 * 4 + tiles - 1) / tiles;
  splits = max(1, min(splits, cdiv(P, 64)));
  int per = cdiv(P, splits);
  per = cdiv(per, BK) * BK;
  splits = cdiv(P, per);
  dim3 grid(cdiv(g->Co, BM), cdiv((int64_t)T * g->Ci, BN), splits);
  gather_gemm_wgrad<<<grid, NT, 0, st>>>(*g, small, big, G, per);
  GLIS_CHECK_LAUNCH("glis_conv_wgrad(fp32)");
  return GLIS_OK;
}

}  // namespace glis

// Pieces shared by the tcgen05 convolution kernels (tc_conv.cu: one box per tap, one CTA per tile; tc_conv_pair.cu:
// cta_group::2 pairs): tile constants, the phase decode of the transposed relation and the specialised epilogue chunk.
#pragma once
#include "common.cuh"
#include "sm100.cuh"

namespace glis {

using namespace sm100;

constexpr int TC_BM = 128;       // channels per CTA (UMMA M)
constexpr int TC_BK = 64;        // bf16 elements per 128-byte swizzle row
constexpr int TC_EPI_WARPS = 16; // 4 per TMEM lane quarter
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;  // TMA warp, MMA warp, epilogue warps
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_DEFAULT_CLUSTER = 1;

struct TcPhase { int ry, rx, Hq, Wq, nth, ntw, py, px; };

__device__ __forceinline__ TcPhase tc_phase(const glis_geom_t& g, int z) {
  TcPhase p;
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

__device__ __forceinline__ int floor_div(int a, int b) {  // b > 0
  int q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}

// Epilogue of one 32-column accumulator chunk for this lane's channel, specialised at compile
// time so that the per-column code is a handful of predicated instructions.  `rel` = element
// offsets of the chunk's columns relative to the tile origin (shared memory, built once per CTA:
// every tile of a launch has the same shape), `base` = the tile origin + this lane's channel.
template <int ACT, bool PREACT, bool F32, bool PLANES>
__device__ __forceinline__ void tc_epilogue_chunk(const uint32_t (&v)[32], int nvalid, long long base,
                                                  const uint32_t* __restrict__ rel, float bias, float ta,
                                                  float tb, float* __restrict__ preact, float* __restrict__ out_f32,
                                                  __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < nvalid) {
      const long long off = base + rel[j];
      const float y = __uint_as_float(v[j]) + bias;
      if (PREACT) preact[off] = y;
      float o = y;
      if (ACT == GLIS_ACT_TPRELU) { const float t = y - tb; o = (t > 0.f ? t : ta * t) + tb; }
      if (ACT == GLIS_ACT_SIGMOID) o = 1.f / (1.f + __expf(-y));
      if (F32) out_f32[off] = o;
      if (PLANES) {
        __nv_bfloat16 hi, lo;
        split_bf16(o, hi, lo);
        out_hi[off] = hi;
        if (out_lo) out_lo[off] = lo;
      }
    }
  }
}

}  // namespace glis

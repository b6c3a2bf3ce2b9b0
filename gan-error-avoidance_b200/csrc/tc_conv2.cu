// tcgen05 implicit-GEMM convolution / transposed convolution, HALO form (forward and data gradient).
//
// Same contraction and orientation as tc_conv.cu (M = 128 output channels, N = output pixels, K = (tap, 64-channel
// block), bf16 hi/lo planes, fp32 accumulators in TMEM) — but the taps of one stride-parity class no longer fetch
// their pixels separately.  In tc_conv.cu every tap pulls its own pixel box from L2: 16 boxes per 64-channel block
// of a 4x4 convolution, although the 4 taps of a parity class read the SAME plane shifted by one pixel; that made
// the kernel L2->SM bound (~85 KB per k-step against 1250 MMA cycles).  Here a class loads ONE box with a halo,
//
//      (tw + hx) x (th + hy) [x tn images]  pixels,   hx / hy = spread of the class's shifts (1 for 4x4 stride 2),
//
// and every tap of the class multiplies against that box through a UMMA descriptor whose start address is advanced
// by whole 128-byte rows: the 128-byte swizzle is a function of the absolute shared-memory address, so a K-major
// operand may start at ANY row of a TMA-written tile (tools/probes/umma_rowshift_probe.cu: exact for every shift,
// descriptor base-offset field 0).  To make one uniform shift per tap work, the accumulator keeps the PADDED
// layout of the box: column j = (n * (th + hy) + y) * (tw + hx) + x, so tap (dy, dx) reads box row
// j + dy * (tw + hx) + dx.  Columns with x >= tw or y >= th hold garbage (each accumulator column depends on its own
// B row only) and are skipped by the epilogue; they cost (tw + hx) / tw of MMA time, 5 % at 20 columns.
// Pixel traffic drops from 16 to 4 * 1.15 boxes per block (3.5x), bytes per k-step from 85 to ~47 KB.
//
// Two shared-memory rings: pixel boxes (2 slots) and weight tiles (2-4 slots); per (class, channel block) the
// producer loads one box and then one weight tile per tap.  Everything else — persistent CTAs, double-buffered
// accumulators, split K, the specialised epilogues — is tc_conv.cu's.
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace glis {

using namespace sm100;

int make_bf16_map_swz(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box, int swizzle_bytes);
int tc_conv_supported(const glis_geom_t* g);

constexpr int T2_BM = 128;
constexpr int T2_EPI_WARPS = 16;
constexpr int T2_THREADS = 64 + 32 * T2_EPI_WARPS;   // warp 0: TMA producer, 1: MMA issuer, 2..17: epilogue
constexpr int T2_NX = 2;          // pixel-box slots
constexpr int T2_MAX_NW = 8;      // weight-tile slots
constexpr int T2_MAX_CLASSES = 16;
constexpr int T2_MAX_TAPS = 36;
constexpr uint32_t T2_SKIP = 0xffffffffu;
constexpr int T2_TAIL = 2048;     // readable padding behind the pixel ring

struct T2Class { short ox, oy, parx, pary, tap_begin, tap_count; };   // box origin relative to the tile origin
struct T2Tap { short tap, shift; };                                   // weight tap (kh * KW + kw), box-row shift

struct T2Params {
  glis_geom_t g;
  int tw, th, tn;        // valid pixel tile on the (phase) output grid; tw spans the full width
  int pw, ph;            // padded tile: tw + hx, th + hy
  int n_mma, tmem_cols, kblocks, passes, nw, a_rows;
  int bk;                // channels per stage: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B)
  int x_rows;            // rows per pixel slot (multiple of 8)
  int box_rows;          // rows one TMA box delivers: pw * ph * tn
  int tiles_h, tiles_x, tiles_co, total_tiles, n_groups, ksplit;
  int phase_cls_begin[17], n_phases;     // classes of phase z: [phase_cls_begin[z], phase_cls_begin[z + 1])
  T2Class cls[T2_MAX_CLASSES];
  T2Tap taps[T2_MAX_TAPS];
  const float* bias; int act; const float* act_a; const float* act_b; int act_channels;
  float* preact; float* out_f32; __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  int ep_mode;
  long long slab_stride;
  unsigned long long* trace;   // GLIS_TC_TRACE: CTA 0 logs globaltimer per event ([0,512) W issue, [512,1024) w_full seen,
                               // [1024,1088) X issue, [1088,1152) x_full seen, [1152,1216) epilogue start / end, [1216] start)
  int debug;             // GLIS_T2_DEBUG (experiments): 1 = all tap shifts 0, 2 = no MMA, 4 = no stores, 8 = shifts rounded to 8 rows
};

struct T2Tile {
  int z, ry, rx, Hq, Wq;       // phase, output residues (transposed relation), phase grid size
  int qy0, n0, co0, split, kb_beg, kb_end, cls_beg, cls_n;
  bool empty, ghost;
};

__device__ __forceinline__ T2Tile t2_tile(const T2Params& P, int item) {
  const glis_geom_t& g = P.g;
  T2Tile t;
  const int id = item / P.ksplit;
  t.split = item - id * P.ksplit;
  t.kb_beg = (int)((long long)P.kblocks * t.split / P.ksplit);
  t.kb_end = (int)((long long)P.kblocks * (t.split + 1) / P.ksplit);
  const int per_phase = P.tiles_x * P.tiles_co;
  t.z = id / per_phase;
  const int rem = id - t.z * per_phase;
  const int y = rem / P.tiles_x, x = rem - y * P.tiles_x;
  if (g.relation == GLIS_CONV) {
    t.ry = t.rx = 0; t.Hq = g.Ho; t.Wq = g.Wo;
  } else {
    const int py = t.z / g.stride_w, px = t.z % g.stride_w;
    t.ry = ((py - g.pad_h) % g.stride_h + g.stride_h) % g.stride_h;
    t.rx = ((px - g.pad_w) % g.stride_w + g.stride_w) % g.stride_w;
    t.Hq = g.Ho > t.ry ? (g.Ho - t.ry + g.stride_h - 1) / g.stride_h : 0;
    t.Wq = g.Wo > t.rx ? (g.Wo - t.rx + g.stride_w - 1) / g.stride_w : 0;
  }
  const int tile_h = x % P.tiles_h, tile_n = x / P.tiles_h;
  t.qy0 = tile_h * P.th;
  t.n0 = tile_n * P.tn;
  t.co0 = y * T2_BM;
  t.cls_beg = P.phase_cls_begin[t.z];
  t.cls_n = P.phase_cls_begin[t.z + 1] - t.cls_beg;
  t.empty = t.Hq <= 0 || t.Wq <= 0 || t.cls_n == 0 || t.kb_end == t.kb_beg;
  t.ghost = t.qy0 >= t.Hq || t.n0 >= g.N;
  return t;
}

template <int ACT, bool PREACT, bool F32, bool PLANES>
__device__ __forceinline__ void t2_chunk(const uint32_t (&v)[32], int nvalid, long long base,
                                         const uint32_t* __restrict__ rel, float bias, float ta, float tb,
                                         float* __restrict__ preact, float* __restrict__ out_f32,
                                         __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const uint32_t r = rel[j];
    if (j < nvalid && r != T2_SKIP) {
      const long long off = base + r;
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

template <int BK>   // channels per stage: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B)
__global__ void __launch_bounds__(T2_THREADS, 1)
tc_conv_halo_kernel(const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                    const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                    const __grid_constant__ T2Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const glis_geom_t& g = P.g;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t row_bytes = 2u * BK;
  const uint32_t x_plane = (uint32_t)P.x_rows * row_bytes, x_slot = 2 * x_plane;
  const uint32_t a_bytes = (uint32_t)P.a_rows * row_bytes, w_slot = 2 * a_bytes;
  // weight ring first: with 64-row weight tiles the 128-row MMA read of a lo tile runs 8 KB past it — into the next
  // weight slot or the pixel ring.  An MMA may also read up to 15 rows past the box of the LAST pixel slot (N is
  // rounded up to 16): T2_TAIL bytes of padding keep that inside the allocation.  Whatever those rows hold, they
  // only feed TMEM lanes / accumulator columns nobody stores.
  uint8_t* wring = base;
  uint8_t* xring = base + (size_t)P.nw * w_slot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xring + (size_t)T2_NX * x_slot + T2_TAIL);
  uint64_t* x_full = bars;                       // [T2_NX]
  uint64_t* x_empty = bars + T2_NX;              // [T2_NX]
  uint64_t* w_full = bars + 2 * T2_NX;           // [T2_MAX_NW]
  uint64_t* w_empty = w_full + T2_MAX_NW;        // [T2_MAX_NW]
  uint64_t* tmem_full_bar = w_empty + T2_MAX_NW;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint32_t* rel = tmem_slot + 4;   // [256] accumulator column -> element offset from the tile origin, or T2_SKIP

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();   // the next kernel's prologue may overlap this one's tail (common.cuh)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_x_hi);
    if (P.passes == 3) { tma_prefetch_desc(&map_w_lo); tma_prefetch_desc(&map_x_lo); }
    for (int s = 0; s < T2_NX; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], 1); }
    for (int s = 0; s < P.nw; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], T2_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  {
    const int sh = g.relation == GLIS_TCONV ? g.stride_h : 1, sw = g.relation == GLIS_TCONV ? g.stride_w : 1;
    const int per_img = P.pw * P.ph;
    for (int c = threadIdx.x; c < 256; c += T2_THREADS) {
      const int in_ = c / per_img, r = c - in_ * per_img, ih = r / P.pw, iw = r - ih * P.pw;
      const bool ok = iw < P.tw && ih < P.th && in_ < P.tn;
      rel[c] = ok ? (uint32_t)((((long long)in_ * g.Ho + (long long)ih * sh) * g.Wo + (long long)iw * sw) * g.Co) : T2_SKIP;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_stride = (uint32_t)P.tmem_cols / 2;
  const int n_ctas = (int)gridDim.x;
  pdl_wait();   // everything above touched shared memory, TMEM and kernel parameters only
  if (TC_TRACE(P) && blockIdx.x == 0 && threadIdx.x == 0) TC_TRACE(P)[1216] = global_timer_ns();

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (one thread feeds both rings.  A second producer thread, so that a box is requested a whole class ahead instead
    //  of queueing behind the previous class's weight tiles, was tried and measured: no faster — what bounds the main
    //  loop is the LATENCY of the weight tiles, 1.5-1.8 us under load with two of them in flight.)
    if (lane == 0) {
      const uint32_t planes = P.passes == 3 ? 2u : 1u;
      const uint32_t x_tx = planes * (uint32_t)P.box_rows * row_bytes, w_tx = planes * a_bytes;
      int xs = 0, ws = 0; uint32_t xpar = 0, wpar = 0;
      int trw = 0, trx = 1024;
      const bool tracing = TC_TRACE(P) && blockIdx.x == 0;
      for (int item = blockIdx.x; item < P.n_groups; item += n_ctas) {
        const T2Tile tl = t2_tile(P, item);
        if (tl.empty) continue;
        // every CTA walks the classes from a different starting point (L2 slices, accumulation order is free)
        const int rot = (int)(((uint32_t)blockIdx.x * 5u + (uint32_t)item * 3u) % (uint32_t)tl.cls_n);
        for (int c0 = 0; c0 < tl.cls_n; ++c0) {
          const T2Class& C = P.cls[tl.cls_beg + (c0 + rot) % tl.cls_n];
          for (int kb = tl.kb_beg; kb < tl.kb_end; ++kb) {
            mbar_wait(&x_empty[xs], xpar ^ 1);
            if (tracing && trx < 1088) TC_TRACE(P)[trx++] = global_timer_ns();
            uint8_t* xb = xring + (size_t)xs * x_slot;
            mbar_arrive_expect_tx(&x_full[xs], x_tx);
            if (g.relation == GLIS_CONV) {
              tma_load_5d(xb, &map_x_hi, &x_full[xs], C.parx * g.Ci + kb * BK, C.ox, C.pary, tl.qy0 + C.oy, tl.n0);
              if (P.passes == 3)
                tma_load_5d(xb + x_plane, &map_x_lo, &x_full[xs], C.parx * g.Ci + kb * BK, C.ox, C.pary, tl.qy0 + C.oy, tl.n0);
            } else {
              tma_load_4d(xb, &map_x_hi, &x_full[xs], kb * BK, C.ox, tl.qy0 + C.oy, tl.n0);
              if (P.passes == 3)
                tma_load_4d(xb + x_plane, &map_x_lo, &x_full[xs], kb * BK, C.ox, tl.qy0 + C.oy, tl.n0);
            }
            if (++xs == T2_NX) { xs = 0; xpar ^= 1; }
            for (int t = 0; t < C.tap_count; ++t) {
              const int tap = P.taps[C.tap_begin + t].tap;
              mbar_wait(&w_empty[ws], wpar ^ 1);
              if (tracing && trw < 512) TC_TRACE(P)[trw++] = global_timer_ns();
              uint8_t* wb = wring + (size_t)ws * w_slot;
              mbar_arrive_expect_tx(&w_full[ws], w_tx);
              tma_load_3d(wb, &map_w_hi, &w_full[ws], kb * BK, tl.co0, tap);
              if (P.passes == 3) tma_load_3d(wb + a_bytes, &map_w_lo, &w_full[ws], kb * BK, tl.co0, tap);
              if (++ws == P.nw) { ws = 0; wpar ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(T2_BM, P.n_mma, 0, 0);
      // K-major operands, 8-row groups 8 * row_bytes apart; layout field (bits 61-63): 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
      constexpr uint64_t layout_fix = BK == 64 ? 0ull : ((2ull << 61) ^ (4ull << 61));
      const uint64_t desc_x0 = umma_smem_desc(smem_u32(xring), 16, 8 * row_bytes) ^ layout_fix;
      const uint64_t desc_w0 = umma_smem_desc(smem_u32(wring), 16, 8 * row_bytes) ^ layout_fix;
      constexpr int nkk = BK / 16;
      int xs = 0, ws = 0; uint32_t xpar = 0, wpar = 0;
      uint32_t acc = 0, acc_phase = 0;
      int trw = 512, trx = 1088;
      const bool tracing = TC_TRACE(P) && blockIdx.x == 0;
      for (int item = blockIdx.x; item < P.n_groups; item += n_ctas) {
        const T2Tile tl = t2_tile(P, item);
        if (tl.empty) continue;
        mbar_wait(&tmem_empty_bar[acc], ((acc_phase >> acc) & 1u) ^ 1u);
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * acc_stride;
        uint32_t accumulate = 0;
        const int rot = (int)(((uint32_t)blockIdx.x * 5u + (uint32_t)item * 3u) % (uint32_t)tl.cls_n);
        for (int c0 = 0; c0 < tl.cls_n; ++c0) {
          const T2Class& C = P.cls[tl.cls_beg + (c0 + rot) % tl.cls_n];
          for (int kb = tl.kb_beg; kb < tl.kb_end; ++kb) {
            mbar_wait(&x_full[xs], xpar);
            tc_fence_after_sync();
            if (tracing && trx < 1152) TC_TRACE(P)[trx++] = global_timer_ns();
            const uint64_t dxh = desc_x0 + (uint64_t)(((uint32_t)xs * x_slot) >> 4);
            const uint64_t dxl = dxh + (x_plane >> 4);
            for (int t = 0; t < C.tap_count; ++t) {
              uint32_t shift8 = (uint32_t)P.taps[C.tap_begin + t].shift * (row_bytes >> 4);   // rows * row bytes >> 4
              if (TC_DEBUG(P) & 1) shift8 = 0;
              if (TC_DEBUG(P) & 8) shift8 &= ~63u;
              mbar_wait(&w_full[ws], wpar);
              tc_fence_after_sync();
              if (tracing && trw < 1024) TC_TRACE(P)[trw++] = global_timer_ns();
              const uint64_t dah = desc_w0 + (uint64_t)(((uint32_t)ws * w_slot) >> 4);
              const uint64_t dal = dah + (a_bytes >> 4);
              const uint64_t dbh = dxh + shift8, dbl = dxl + shift8;
              if (TC_DEBUG(P) & 2) {
              } else if (P.passes == 3) {
#pragma unroll
                for (int kk = 0; kk < nkk; ++kk) {
                  umma_bf16(tmem_d, dah + 2 * kk, dbl + 2 * kk, idesc, accumulate);
                  umma_bf16(tmem_d, dal + 2 * kk, dbh + 2 * kk, idesc, 1);
                  umma_bf16(tmem_d, dah + 2 * kk, dbh + 2 * kk, idesc, 1);
                  accumulate = 1;
                }
              } else {
#pragma unroll
                for (int kk = 0; kk < nkk; ++kk) {
                  umma_bf16(tmem_d, dah + 2 * kk, dbh + 2 * kk, idesc, accumulate);
                  accumulate = 1;
                }
              }
              umma_commit(&w_empty[ws]);
              if (++ws == P.nw) { ws = 0; wpar ^= 1; }
            }
            umma_commit(&x_empty[xs]);   // every MMA that reads this box has been issued before this commit
            if (++xs == T2_NX) { xs = 0; xpar ^= 1; }
          }
        }
        umma_commit(&tmem_full_bar[acc]);
        acc_phase ^= (1u << acc);
        acc ^= 1u;
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    const int sh = g.relation == GLIS_TCONV ? g.stride_h : 1;
    const int per_img = P.pw * P.ph;
    uint32_t acc = 0, full_phase = 0;
    int tr_e = 1152;
    for (int item = blockIdx.x; item < P.n_groups; item += n_ctas) {
      const T2Tile tl = t2_tile(P, item);
      if (tl.empty) continue;
      const int co = tl.co0 + q * 32 + lane;
      const bool ch_ok = co < g.Co && !tl.ghost && !(TC_DEBUG(P) & 4);
      float bias = 0.f, ta = 0.f, tb = 0.f;
      if (ch_ok) {
        if (P.bias && tl.split == 0) bias = __ldg(P.bias + co);
        if (P.act == GLIS_ACT_TPRELU) {
          const int ca = P.act_channels > 0 ? co % P.act_channels : co;
          ta = fminf(fmaxf(__ldg(P.act_a + ca), 0.f), 1.f); tb = __ldg(P.act_b + ca);
        }
      }
      const int oy0 = g.relation == GLIS_TCONV ? tl.qy0 * sh + tl.ry : tl.qy0;
      const int ox0 = g.relation == GLIS_TCONV ? tl.rx : 0;
      const long long base_off = (((long long)tl.n0 * g.Ho + oy0) * g.Wo + ox0) * g.Co + co;
      // valid columns: a prefix of the padded layout (ragged rows when tn == 1, ragged images otherwise) minus the
      // halo columns / rows the table marks T2_SKIP
      const int valid_cols = P.tn == 1 ? min(P.th, tl.Hq - tl.qy0) * P.pw : min(P.tn, g.N - tl.n0) * per_img;
      const int cols = P.tn == 1 ? (P.th - 1) * P.pw + P.tw : ((P.tn - 1) * P.ph + P.th - 1) * P.pw + P.tw;
      mbar_wait(&tmem_full_bar[acc], (full_phase >> acc) & 1u);
      tc_fence_after_sync();
      if (TC_TRACE(P) && blockIdx.x == 0 && warp == 2 && lane == 0 && tr_e < 1216) TC_TRACE(P)[tr_e++] = global_timer_ns();
      const uint32_t tmem_d = tmem_base + acc * acc_stride + ((uint32_t)(q * 32) << 16);
      const bool quarter_ok = tl.co0 + q * 32 < g.Co && !tl.ghost;
      for (int cb = part * 32; cb < cols && quarter_ok; cb += 128) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_d + (uint32_t)cb, v);
        tmem_ld_wait();
        const int nvalid = ch_ok ? valid_cols - cb : 0;
        const uint32_t* rc = rel + cb;
        switch (P.ep_mode) {
          case 1: t2_chunk<GLIS_ACT_NONE, false, true, false>(v, nvalid, base_off, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 2: t2_chunk<GLIS_ACT_TPRELU, true, false, true>(v, nvalid, base_off, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 3: t2_chunk<GLIS_ACT_NONE, false, false, true>(v, nvalid, base_off, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 4: t2_chunk<GLIS_ACT_TPRELU, true, true, false>(v, nvalid, base_off, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 5: t2_chunk<GLIS_ACT_TPRELU, true, true, true>(v, nvalid, base_off, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 7: t2_chunk<GLIS_ACT_TPRELU, false, false, true>(v, nvalid, base_off, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 8: {   // split K, deterministic form: this share's partial sums into its own slab
            float* slab = P.out_f32 + (long long)tl.split * P.slab_stride;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid && rc[j] != T2_SKIP) slab[base_off + rc[j]] = __uint_as_float(v[j]) + bias;
            break;
          }
          case 6: {   // split K: add this item's partial sums
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid && rc[j] != T2_SKIP) atomicAdd(P.out_f32 + base_off + rc[j], __uint_as_float(v[j]) + bias);
            break;
          }
          default: {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nvalid && rc[j] != T2_SKIP) {
                const long long off = base_off + rc[j];
                const float y = __uint_as_float(v[j]) + bias;
                if (P.preact) P.preact[off] = y;
                float o = y;
                if (P.act == GLIS_ACT_TPRELU) { const float t = y - tb; o = (t > 0.f ? t : ta * t) + tb; }
                else if (P.act == GLIS_ACT_SIGMOID) { o = 1.f / (1.f + __expf(-y)); }
                if (P.out_f32) P.out_f32[off] = o;
                if (P.out_hi) {
                  __nv_bfloat16 hi, lo;
                  split_bf16(o, hi, lo);
                  P.out_hi[off] = hi;
                  if (P.out_lo) P.out_lo[off] = lo;
                }
              }
            }
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (TC_TRACE(P) && blockIdx.x == 0 && warp == 2 && lane == 0 && tr_e < 1216) TC_TRACE(P)[tr_e++] = global_timer_ns();
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
static int t2_round_up(int a, int b) { return (a + b - 1) / b * b; }
static int t2_mod(int a, int b) { return ((a % b) + b) % b; }

static int t2_num_sms() { return plan_sms(); }

// Tap classes of the launch: which taps share a pixel box and where that box sits.  Returns the halo (hx, hy) shared
// by every class, or -1 when the halo form does not apply (no class with more than one tap, too many taps).
static int t2_classes(const glis_geom_t* g, T2Params& P, int& hx, int& hy) {
  int ncls = 0, ntaps = 0;
  hx = hy = 0;
  bool shares = false;
  // first pass fills classes with ABSOLUTE shifts in T2Tap::shift (dy * 1024 + dx packed later)
  struct Raw { int tap, sy, sx; };
  Raw raw[T2_MAX_TAPS];
  if (g->relation == GLIS_CONV) {
    P.n_phases = 1;
    P.phase_cls_begin[0] = 0;
    for (int pary = 0; pary < g->stride_h; ++pary)
      for (int parx = 0; parx < g->stride_w; ++parx) {
        int begin = ntaps, miny = 1 << 20, minx = 1 << 20, maxy = -(1 << 20), maxx = -(1 << 20);
        for (int kh = 0; kh < g->KH; ++kh) {
          const int ey = kh * g->dil_h - g->pad_h;
          if (t2_mod(ey, g->stride_h) != pary) continue;
          for (int kw = 0; kw < g->KW; ++kw) {
            const int ex = kw * g->dil_w - g->pad_w;
            if (t2_mod(ex, g->stride_w) != parx) continue;
            if (ntaps >= T2_MAX_TAPS) return -1;
            const int sy = (ey - pary) / g->stride_h, sx = (ex - parx) / g->stride_w;
            raw[ntaps++] = {kh * g->KW + kw, sy, sx};
            miny = sy < miny ? sy : miny; maxy = sy > maxy ? sy : maxy;
            minx = sx < minx ? sx : minx; maxx = sx > maxx ? sx : maxx;
          }
        }
        if (ntaps == begin) continue;
        if (ncls >= T2_MAX_CLASSES) return -1;
        P.cls[ncls++] = {(short)minx, (short)miny, (short)parx, (short)pary, (short)begin, (short)(ntaps - begin)};
        hx = maxx - minx > hx ? maxx - minx : hx;
        hy = maxy - miny > hy ? maxy - miny : hy;
        shares = shares || ntaps - begin > 1;
      }
    P.phase_cls_begin[1] = ncls;
  } else {
    P.n_phases = g->stride_h * g->stride_w;
    if (P.n_phases > 16) return -1;
    for (int z = 0; z < P.n_phases; ++z) {
      P.phase_cls_begin[z] = ncls;
      const int py = z / g->stride_w, px = z % g->stride_w;
      const int ry = t2_mod(py - g->pad_h, g->stride_h), rx = t2_mod(px - g->pad_w, g->stride_w);
      int begin = ntaps, miny = 1 << 20, minx = 1 << 20, maxy = -(1 << 20), maxx = -(1 << 20);
      for (int kh = py; kh < g->KH; kh += g->stride_h)
        for (int kw = px; kw < g->KW; kw += g->stride_w) {
          if (ntaps >= T2_MAX_TAPS) return -1;
          const int sy = (ry + g->pad_h - kh) / g->stride_h, sx = (rx + g->pad_w - kw) / g->stride_w;   // exact
          raw[ntaps++] = {kh * g->KW + kw, sy, sx};
          miny = sy < miny ? sy : miny; maxy = sy > maxy ? sy : maxy;
          minx = sx < minx ? sx : minx; maxx = sx > maxx ? sx : maxx;
        }
      if (ntaps == begin) continue;
      if (ncls >= T2_MAX_CLASSES) return -1;
      P.cls[ncls++] = {(short)minx, (short)miny, 0, 0, (short)begin, (short)(ntaps - begin)};
      hx = maxx - minx > hx ? maxx - minx : hx;
      hy = maxy - miny > hy ? maxy - miny : hy;
      shares = shares || ntaps - begin > 1;
    }
    P.phase_cls_begin[P.n_phases] = ncls;
  }
  if (!shares || hx > 3 || hy > 3) return -1;
  for (int c = 0; c < ncls; ++c)
    for (int t = P.cls[c].tap_begin; t < P.cls[c].tap_begin + P.cls[c].tap_count; ++t) {
      P.taps[t].tap = (short)raw[t].tap;
      P.taps[t].shift = (short)((raw[t].sy - P.cls[c].oy) * 1024 + (raw[t].sx - P.cls[c].ox));   // (dy, dx); rows set by the plan
    }
  return ncls;
}

static int t2_enabled() {
  const char* e = getenv("GLIS_TC_HALO");      // 0 = every launch on tc_conv.cu's one-box-per-tap kernel
  if (e && atoi(e) == 0) return 0;
  const char* c = getenv("GLIS_TC_CLUSTER");   // the weight-multicast experiment lives in tc_conv.cu
  if (c && atoi(c) > 1) return 0;
  return 1;
}

// Tile shape, rings and K split.  Returns GLIS_E_UNSUPPORTED when the halo form does not apply or does not fit.
static int t2_plan(const glis_geom_t* g, bool plain_out, T2Params& P) {
  if (!t2_enabled() || !tc_conv_supported(g)) return GLIS_E_UNSUPPORTED;
  P.g = *g;
  int hx, hy;
  if (t2_classes(g, P, hx, hy) < 0) return GLIS_E_UNSUPPORTED;
  int nphase = 1, Hq = g->Ho, Wq = g->Wo;
  if (g->relation == GLIS_TCONV) {
    nphase = g->stride_h * g->stride_w;
    Hq = (g->Ho + g->stride_h - 1) / g->stride_h;
    Wq = (g->Wo + g->stride_w - 1) / g->stride_w;
  }
  const int pw = Wq + hx;
  if (pw > 256) return GLIS_E_UNSUPPORTED;
  // The halo columns / rows are accumulator columns nobody stores: (Wq + hx)(th + hy) / (Wq th) of the MMA work.
  // Measured (tools/tc_microbench.py h0 h1): 5-7 % faster than one box per tap on 20- and 40-wide maps, equal at 10,
  // 20-30 % SLOWER on 5x5 maps (44 % padding) — narrow maps stay on tc_conv.cu.
  {
    const char* e = getenv("GLIS_TC_HALO_MINW");
    const int minw = e ? atoi(e) : 16;
    if (Wq < minw) return GLIS_E_UNSUPPORTED;
  }
  const int num_sms = t2_num_sms();
  const int co_tiles = (g->Co + T2_BM - 1) / T2_BM;
  int bk_force = 0;
  {
    const char* e = getenv("GLIS_TC_HALO_BK");   // 32 / 64: force the channels per stage
    bk_force = e ? atoi(e) : 0;
    if (bk_force != 32 && bk_force != 64) bk_force = 0;
  }
  const int a_rows_small = g->Co <= 64 ? 64 : T2_BM;
  int taps_max = 0, taps_total_max = 0;
  for (int z = 0; z < P.n_phases; ++z) {
    int tot = 0;
    for (int c = P.phase_cls_begin[z]; c < P.phase_cls_begin[z + 1]; ++c) {
      taps_max = P.cls[c].tap_count > taps_max ? P.cls[c].tap_count : taps_max;
      tot += P.cls[c].tap_count;
    }
    taps_total_max = tot > taps_total_max ? tot : taps_total_max;
  }
  int ksplit_max = 32;
  {
    const char* e = getenv("GLIS_TC_KSPLIT");
    if (e) ksplit_max = atoi(e);
    if (ksplit_max < 1) ksplit_max = 1;
  }
  const long smem_cap = 227 * 1024 - 1024 /*alignment*/ - 256 /*barriers*/ - 1024 /*column table*/ - T2_TAIL;
  long best_cost = -1;
  int best_th = 0, best_tn = 0, best_ks = 1, best_nw = 0, best_n = 0, best_ar = 0, best_bk = 64;
  auto consider = [&](int th, int tn) {
    const int ph = th + hy;
    const int cols = tn == 1 ? (th - 1) * pw + Wq : ((tn - 1) * ph + th - 1) * pw + Wq;
    const int n = t2_round_up(cols, 16);
    if (n > 256 || ph > 256) return;
    const int box_rows = pw * ph * tn;
    const int x_rows = t2_round_up(box_rows, 8);
    const int a_rows = n >= 64 ? a_rows_small : T2_BM;     // (a 128-row MMA read starting in the lo tile stays in the ring)
    // Channels per stage: 64.  Halving the stage (GLIS_TC_HALO_BK=32: 64-byte rows, SWIZZLE_64B, twice the weight
    // slots for the same bytes) was built to get more weight tiles in flight and measured SLOWER (47 vs 41 us on D's
    // level 1): every stage costs the single MMA-issuing thread a barrier wait, a fence and a commit, and at 72
    // cycles per tcgen05.mma issue (tools/probes/umma_rate_probe.cu) that thread has no slack to give.
    int bk = 64;
    long x_bytes = 2L * x_rows * 128, w_bytes = 2L * a_rows * 128;
    long nw = (smem_cap - T2_NX * x_bytes) / w_bytes;
    if (bk_force == 32) {
      bk = 32;
      x_bytes /= 2; w_bytes /= 2;
      nw = (smem_cap - T2_NX * x_bytes) / w_bytes;
    }
    if (nw > T2_MAX_NW) nw = T2_MAX_NW;
    if (nw < 2) return;
    const int kblocks = (g->Ci + bk - 1) / bk;
    const long tiles_x = (long)((Hq + th - 1) / th) * ((g->N + tn - 1) / tn);
    const long tiles = tiles_x * co_tiles * nphase;
    for (int ks = 1; ks <= ksplit_max && ks <= kblocks; ks *= 2) {
      if (ks > 1 && !plain_out) break;
      if (ks > 1 && kblocks % ks != 0 && kblocks < 4 * ks) break;
      const long waves = (tiles * ks + num_sms - 1) / num_sms;
      const long kb = (kblocks + ks - 1) / ks;
      // cycles per tap step: 12 (or 4) MMAs of n / 2 cycles each, against the bytes the step pulls from L2 at
      // ~40 B / cycle / SM (weights every step, the pixel box once per class)
      const long mma = 6L * n * bk / 64, load = (w_bytes + x_bytes / taps_max) / 40;
      const long step = mma > load ? mma : load;
      // + per work item: pipeline fill, accumulator hand-over and the part of the epilogue the next tile cannot hide
      long cost = waves * ((long)taps_total_max * kb * step + 2500 + 8 * n) + (ks > 1 ? waves * 2 * n + 1024 : 0);
      cost = cost * 1024 + n;
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost; best_th = th; best_tn = tn; best_ks = ks; best_nw = (int)nw; best_n = n; best_ar = a_rows;
        best_bk = bk;
      }
    }
  };
  for (int th = 1; th <= Hq; ++th) consider(th, 1);
  for (int tn = 2; tn <= g->N; ++tn) consider(Hq, tn);
  if (best_cost < 0) return GLIS_E_UNSUPPORTED;
  P.tw = Wq; P.th = best_th; P.tn = best_tn;
  P.pw = pw; P.ph = best_th + hy;
  P.n_mma = best_n;
  P.a_rows = best_ar;
  P.nw = best_nw;
  P.bk = best_bk;
  const int kblocks = (g->Ci + best_bk - 1) / best_bk;
  P.box_rows = P.pw * P.ph * P.tn;
  P.x_rows = t2_round_up(P.box_rows, 8);
  P.tmem_cols = 64;
  while (P.tmem_cols < 2 * P.n_mma) P.tmem_cols *= 2;
  P.kblocks = kblocks;
  P.tiles_h = (Hq + P.th - 1) / P.th;
  P.tiles_x = P.tiles_h * ((g->N + P.tn - 1) / P.tn);
  P.tiles_co = co_tiles;
  P.total_tiles = P.tiles_x * P.tiles_co * nphase;
  P.ksplit = best_ks;
  P.n_groups = P.total_tiles * best_ks;
  const int ntaps = P.cls[P.phase_cls_begin[P.n_phases] - 1].tap_begin + P.cls[P.phase_cls_begin[P.n_phases] - 1].tap_count;
  for (int t = 0; t < ntaps; ++t) {
    const int dy = P.taps[t].shift / 1024, dx = P.taps[t].shift % 1024;
    P.taps[t].shift = (short)(dy * P.pw + dx);
  }
  return GLIS_OK;
}

// 1 when a launch of this geometry takes the halo kernel (plain_out: the launch writes fp32 sums only).
int tc_conv_halo_applies(const glis_geom_t* g, int plain_out) {
  T2Params P;
  return t2_plan(g, plain_out != 0, P) == GLIS_OK ? 1 : 0;
}

int tc_conv_halo_ksplit(const glis_geom_t* g) {
  T2Params P;
  if (t2_plan(g, true, P) != GLIS_OK) return 0;
  return P.ksplit;
}

// out = {tw, th, tn, n_mma, tmem_cols, kblocks, ksplit, a_rows, weight slots, tiles_h, tiles_x, tiles_co,
//        total_tiles, n_groups, dynamic shared memory bytes, hx, hy, box rows, pixel slot rows, tap classes}
int tc_conv_halo_describe(const glis_geom_t* g, int plain_out, int out[20]) {
  T2Params P;
  int rc = t2_plan(g, plain_out != 0, P);
  if (rc != GLIS_OK) return rc;
  const size_t smem = (size_t)T2_NX * 2 * P.x_rows * 2 * P.bk + (size_t)P.nw * 2 * P.a_rows * 2 * P.bk + 1024 + 256 + 1024 + T2_TAIL;
  const int v[20] = {P.tw, P.th, P.tn, P.n_mma, P.tmem_cols, P.kblocks, P.ksplit, P.a_rows, P.nw, P.tiles_h, P.tiles_x,
                     P.tiles_co, P.total_tiles, P.n_groups, (int)smem, P.pw - P.tw, P.ph - P.th, P.box_rows, P.x_rows,
                     P.phase_cls_begin[P.n_phases] + 1000 * P.bk};
  for (int i = 0; i < 20; ++i) out[i] = v[i];
  return GLIS_OK;
}

int tc_conv_halo_forward(const glis_geom_t* g, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
                         const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, const glis_epilogue_t* ep, float* out_f32,
                         __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int precision, cudaStream_t st) {
  const int passes = precision == GLIS_PREC_BF16X3 ? 3 : 1;
  GLIS_REQUIRE(x_hi && w_hi && (passes == 1 || (x_lo && w_lo)), GLIS_E_BADARG,
               "glis_conv_forward_bf16: missing hi/lo operand planes");
  T2Params P;
  const bool plain = ep->act == GLIS_ACT_NONE && !ep->preact && out_f32 && !out_hi;
  int rc = t2_plan(g, plain, P);
  if (rc != GLIS_OK) return rc;
  P.passes = passes;
  P.bias = ep->bias; P.act = ep->act; P.act_a = ep->act_a; P.act_b = ep->act_b; P.preact = ep->preact;
  P.act_channels = ep->act_channels;
  P.out_f32 = out_f32; P.out_hi = out_hi; P.out_lo = out_lo;
  P.ep_mode = 0;
  if (ep->act == GLIS_ACT_NONE && !ep->preact && out_f32 && !out_hi) P.ep_mode = 1;
  else if (ep->act == GLIS_ACT_TPRELU && ep->preact && !out_f32 && out_hi) P.ep_mode = 2;
  else if (ep->act == GLIS_ACT_NONE && !ep->preact && !out_f32 && out_hi) P.ep_mode = 3;
  else if (ep->act == GLIS_ACT_TPRELU && ep->preact && out_f32 && !out_hi) P.ep_mode = 4;
  else if (ep->act == GLIS_ACT_TPRELU && ep->preact && out_f32 && out_hi) P.ep_mode = 5;
  else if (ep->act == GLIS_ACT_TPRELU && !ep->preact && !out_f32 && out_hi) P.ep_mode = 7;
  P.slab_stride = 0;
  {
    const char* dbg = getenv("GLIS_T2_DEBUG");
    P.debug = dbg ? atoi(dbg) : 0;
    const char* trc = getenv("GLIS_TC_TRACE");  // hex device address of a >= 1217-entry u64 buffer
    P.trace = trc ? (unsigned long long*)strtoull(trc, nullptr, 16) : nullptr;
  }
  if (P.ksplit > 1 && ep->split_slabs > 0) {
    GLIS_REQUIRE(ep->split_slabs >= P.ksplit, GLIS_E_BADARG, "glis_conv_forward_bf16: %d slabs for a %d-way K split",
                 ep->split_slabs, P.ksplit);
    P.ep_mode = 8;
    P.slab_stride = (long long)g->N * g->Ho * g->Wo * g->Co;
  } else if (P.ksplit > 1) {
    P.ep_mode = 6;
    cudaError_t me = cudaMemsetAsync(out_f32, 0, sizeof(float) * (size_t)g->N * g->Ho * g->Wo * g->Co, st);
    GLIS_REQUIRE(me == cudaSuccess, GLIS_E_CUDA, "glis_conv_forward_bf16: memset failed: %s", cudaGetErrorString(me));
  }

  CUtensorMap mw_hi, mw_lo, mx_hi, mx_lo;
  const int T = g->KH * g->KW;
  {
    const uint64_t dims[3] = {(uint64_t)g->Ci, (uint64_t)g->Co, (uint64_t)T};
    const uint64_t strides[2] = {(uint64_t)g->Ci * 2, (uint64_t)g->Ci * g->Co * 2};
    const uint32_t box[3] = {(uint32_t)P.bk, (uint32_t)P.a_rows, 1};
    rc = make_bf16_map_swz(&mw_hi, w_hi, 3, dims, strides, box, 2 * P.bk);
    if (rc) return rc;
    rc = make_bf16_map_swz(&mw_lo, passes == 3 ? w_lo : w_hi, 3, dims, strides, box, 2 * P.bk);
    if (rc) return rc;
  }
  if (g->relation == GLIS_CONV) {
    const uint64_t C = g->Ci, W = g->Wi, H = g->Hi, sw = g->stride_w, sh = g->stride_h;
    const uint64_t dims[5] = {sw * C, W / sw, sh, H / sh, (uint64_t)g->N};
    const uint64_t strides[4] = {sw * C * 2, W * C * 2, sh * W * C * 2, H * W * C * 2};
    const uint32_t box[5] = {(uint32_t)P.bk, (uint32_t)P.pw, 1, (uint32_t)P.ph, (uint32_t)P.tn};
    rc = make_bf16_map_swz(&mx_hi, x_hi, 5, dims, strides, box, 2 * P.bk);
    if (rc) return rc;
    rc = make_bf16_map_swz(&mx_lo, passes == 3 ? x_lo : x_hi, 5, dims, strides, box, 2 * P.bk);
    if (rc) return rc;
  } else {
    const uint64_t C = g->Ci, W = g->Wi, H = g->Hi;
    const uint64_t dims[4] = {C, W, H, (uint64_t)g->N};
    const uint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    const uint32_t box[4] = {(uint32_t)P.bk, (uint32_t)P.pw, (uint32_t)P.ph, (uint32_t)P.tn};
    rc = make_bf16_map_swz(&mx_hi, x_hi, 4, dims, strides, box, 2 * P.bk);
    if (rc) return rc;
    rc = make_bf16_map_swz(&mx_lo, passes == 3 ? x_lo : x_hi, 4, dims, strides, box, 2 * P.bk);
    if (rc) return rc;
  }

  const size_t smem = (size_t)T2_NX * 2 * P.x_rows * 2 * P.bk + (size_t)P.nw * 2 * P.a_rows * 2 * P.bk + 1024 + 256 + 1024 + T2_TAIL;
  GLIS_REQUIRE(P.tmem_cols <= 512 && smem <= 227 * 1024, GLIS_E_UNSUPPORTED, "glis_conv_forward_bf16: halo tile does not fit");
  {
    cudaError_t e = ensure_max_dynamic_smem(reinterpret_cast<const void*>(tc_conv_halo_kernel<64>), 227 * 1024);
    if (e == cudaSuccess) e = ensure_max_dynamic_smem(reinterpret_cast<const void*>(tc_conv_halo_kernel<32>), 227 * 1024);
    GLIS_REQUIRE(e == cudaSuccess, GLIS_E_CUDA, "cudaFuncSetAttribute(tc_conv_halo_kernel): %s", cudaGetErrorString(e));
  }
  const int num_sms = t2_num_sms();
  const int grid = P.n_groups < num_sms ? P.n_groups : num_sms;
  cudaError_t le = launch_pdl(P.bk == 64 ? tc_conv_halo_kernel<64> : tc_conv_halo_kernel<32>, dim3(grid), dim3(T2_THREADS),
                              smem, st, mw_hi, mw_lo, mx_hi, mx_lo, P);
  GLIS_REQUIRE(le == cudaSuccess, GLIS_E_CUDA, "glis_conv_forward_bf16(halo): launch failed: %s", cudaGetErrorString(le));
  GLIS_CHECK_LAUNCH("glis_conv_forward_bf16(halo)");
  return GLIS_OK;
}

}  // namespace glis

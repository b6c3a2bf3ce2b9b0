// tcgen05 implicit-GEMM convolution / transposed convolution (forward and data gradient).
//
//   D[co, pix] = sum_{tap} sum_{ci}  W[tap][co][ci] * X[gather(pix, tap)][ci]
//
// M = 128 output channels (TMEM lanes), N = one tile of output pixels (<= 256 TMEM columns),
// K walked as (tap, 64-channel block).  Both operands are K-major bf16 in 128-byte-swizzled
// shared memory, staged by TMA:
//   * weights  : 3-D map (Ci, Co, taps), box 64 x 128 x 1;
//   * pixels   : the NHWC activation tensor seen through a map whose box IS the im2col
//                gather of one tap — for the strided (GLIS_CONV) relation a 5-D view
//                (stride*C, W/stride, stride, H/stride, N) in which a tap is a fixed
//                (parity, shift) pair; for the transposed (GLIS_TCONV) relation the plain
//                4-D (C, W, H, N) view shifted by the tap, one launch slice per output phase.
//     Zero padding is TMA out-of-bounds fill; ragged channel counts likewise.
// fp32 fidelity comes from a bf16 hi/lo split of both operands: three MMAs per k-step
// (hi*hi + hi*lo + lo*hi), ~2^-16 relative (GLIS_PREC_BF16X3); one MMA in GLIS_PREC_BF16.
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..17 = epilogue (TMEM -> registers -> bias / TPReLU / sigmoid -> global, plus the
// optional bf16 hi/lo planes the next tensor-core layer consumes).
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"
#include "tc_common.cuh"

namespace glis {

struct TcConvParams {
  glis_geom_t g;
  int tw, th, tn;    // pixel tile on the (phase) output grid; tw spans the full width
  int n_mma;         // UMMA N: tw*th*tn rounded up to 16
  int tmem_cols;     // power of two >= n_mma
  int kblocks;       // ceil(Ci / 64)
  int passes;        // 3 (bf16x3) or 1
  int stages;
  int tiles_h;       // ceil(Hq_max / th)
  int tiles_x;       // pixel tiles per (phase, channel tile) = tiles_h * ceil(N / tn)
  int a_rows;        // weight rows staged per tile: 128, or 64 when Co <= 64 (the MMA still reads 128 rows: the
                     // upper 64 are whatever follows in the stage and only feed TMEM lanes nobody reads)
  int tiles_co;      // ceil(Co / 128)
  int total_tiles;   // tiles_x * tiles_co * phases
  int cluster;       // CTAs per cluster (1, 2, 4): the pixel tiles of a cluster share ONE weight tile, each
                     // CTA fetches 128/cluster of its rows and multicasts them (tiles_x is padded to a multiple)
  int n_groups;      // total_tiles / cluster (x ksplit)
  int ksplit;        // K (the 64-channel blocks of every tap) split across `ksplit` work items per tile; their
                     // partial sums are added into a zero-filled fp32 output (plain-output launches only)
  const float* bias; int act; const float* act_a; const float* act_b; int act_channels;
  float* preact; float* out_f32; __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  int ep_mode;       // compile-time specialised epilogue (0 = generic)
  long long slab_stride;   // split K without atomics: share s stores its sums at out_f32 + s * slab_stride
  unsigned long long* trace;  // optional timeline buffer (GLIS_TC_TRACE): CTA 0 logs globaltimer per event
  int debug;         // GLIS_TC_DEBUG bits (profiling experiments only): 1 = no stores, 2 = no MMA, 4 = no x loads
};

// One output tile of the persistent kernel.
struct TcTile {
  TcPhase ph;
  int qy0, n0, co0, ntaps, ksteps;
  int kb_beg, kb_end, split;   // this work item's share of the 64-channel blocks
  bool empty;   // nothing to contract (uniform over the tiles of a cluster)
  bool ghost;   // padding tile: takes part in loads and MMAs (zero pixels), stores nothing
};

__device__ __forceinline__ TcTile tc_tile(const TcConvParams& P, int item) {
  TcTile t;
  const int id = item / P.ksplit;
  t.split = item - id * P.ksplit;
  t.kb_beg = (int)((long long)P.kblocks * t.split / P.ksplit);
  t.kb_end = (int)((long long)P.kblocks * (t.split + 1) / P.ksplit);
  const int per_phase = P.tiles_x * P.tiles_co;
  const int z = id / per_phase, rem = id - z * per_phase;
  const int y = rem / P.tiles_x, x = rem - y * P.tiles_x;
  t.ph = tc_phase(P.g, z);
  const int tile_h = x % P.tiles_h, tile_n = x / P.tiles_h;
  t.qy0 = tile_h * P.th;
  t.n0 = tile_n * P.tn;
  t.co0 = y * TC_BM;
  t.ntaps = t.ph.nth * t.ph.ntw;
  t.ksteps = t.ntaps * (t.kb_end - t.kb_beg);
  t.empty = t.ph.Hq <= 0 || t.ph.Wq <= 0 || t.ksteps == 0;
  t.ghost = t.qy0 >= t.ph.Hq || t.n0 >= P.g.N;
  return t;
}

// Persistent: each CTA walks tiles id = blockIdx.x, blockIdx.x + gridDim.x, ...  The smem
// ring runs across tile boundaries and the accumulator is double-buffered in TMEM, so the
// epilogue of tile i overlaps the TMA + MMA main loop of tile i+1.
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
               const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
               const TcConvParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const glis_geom_t& g = P.g;

  // ---- shared memory carve-up (1024-byte aligned operand tiles)
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = (uint32_t)P.a_rows * 128, b_bytes = (uint32_t)P.n_mma * 128;
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)P.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + TC_MAX_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * TC_MAX_STAGES;       // [2]
  uint64_t* tmem_empty_bar = bars + 2 * TC_MAX_STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_MAX_STAGES + 4);
  uint32_t* rel = tmem_slot + 4;   // [256] column -> element offset from the tile origin

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();   // the next kernel's prologue may overlap this one's tail (common.cuh)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_x_hi);
    if (P.passes == 3) { tma_prefetch_desc(&map_w_lo); tma_prefetch_desc(&map_x_lo); }
    for (int s = 0; s < P.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], (uint32_t)P.cluster); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], TC_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)P.tmem_cols);
  {
    const int sh = g.relation == GLIS_TCONV ? g.stride_h : 1, sw = g.relation == GLIS_TCONV ? g.stride_w : 1;
    const int per_img = P.tw * P.th;
    for (int c = threadIdx.x; c < 256; c += TC_THREADS) {
      const int in_ = c / per_img, r = c - in_ * per_img, ih = r / P.tw, iw = r - ih * P.tw;
      rel[c] = (uint32_t)((((long long)in_ * g.Ho + (long long)ih * sh) * g.Wo + (long long)iw * sw) * g.Co);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (P.cluster > 1) cluster_sync_all();   // every CTA's barriers exist before a neighbour signals them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_stride = (uint32_t)P.tmem_cols / 2;
  const int cs = P.cluster;
  const int crank = cs > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = (int)blockIdx.x / cs, n_clusters = (int)gridDim.x / cs;
  const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
  pdl_wait();   // everything above touched shared memory, TMEM and kernel parameters only
  if (TC_TRACE(P) && blockIdx.x == 0 && threadIdx.x == 0) TC_TRACE(P)[1087] = global_timer_ns();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t box_rows = (uint32_t)(P.tw * P.th * P.tn);
      const uint32_t tx_bytes = (P.passes == 3 ? 2u : 1u) * (a_bytes + ((TC_DEBUG(P) & 4) ? 0u : box_rows * 128u));
      int s = 0; uint32_t parity = 0;
      int tr_n = 0;
      const uint32_t w_rows = (uint32_t)(P.a_rows / cs), w_slice = w_rows * 128u;   // this CTA's share of the weight tile
      for (int grp = cluster_id; grp < P.n_groups; grp += n_clusters) {
        const int id = grp * cs + crank;
        const TcTile tl = tc_tile(P, id);
        if (tl.empty) continue;
        // Every cluster walks the taps from a different starting point: otherwise all 148 SMs pull the
        // same weight tile from the same L2 slices in lock step (accumulation order is free).
        const int rot = (int)(((uint32_t)cluster_id * 5u + (uint32_t)grp * 3u) % (uint32_t)tl.ntaps);
        for (int t0 = 0; t0 < tl.ntaps; ++t0) {
          const int t = (t0 + rot) % tl.ntaps;
          const int jh = t / tl.ph.ntw, jw = t - jh * tl.ph.ntw;
          int kh, kw, cpar = 0, c1, c2 = 0, c3;
          if (g.relation == GLIS_CONV) {
            kh = jh; kw = jw;
            const int ey = kh * g.dil_h - g.pad_h, ex = kw * g.dil_w - g.pad_w;
            const int pary = ((ey % g.stride_h) + g.stride_h) % g.stride_h;
            const int parx = ((ex % g.stride_w) + g.stride_w) % g.stride_w;
            cpar = parx * g.Ci;
            c1 = (ex - parx) / g.stride_w;
            c2 = pary;
            c3 = tl.qy0 + (ey - pary) / g.stride_h;
          } else {
            kh = tl.ph.py + jh * g.stride_h; kw = tl.ph.px + jw * g.stride_w;
            c1 = (tl.ph.rx + g.pad_w - kw) / g.stride_w;   // exact inside a phase
            c3 = tl.qy0 + (tl.ph.ry + g.pad_h - kh) / g.stride_h;
          }
          const int tap = kh * g.KW + kw;
          for (int kb = tl.kb_beg; kb < tl.kb_end; ++kb) {
            mbar_wait(&empty_bar[s], parity ^ 1);
            if (TC_TRACE(P) && blockIdx.x == 0 && tr_n < 512) TC_TRACE(P)[tr_n++] = global_timer_ns();
            uint8_t* st = base + (size_t)s * stage_bytes;
            mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
            if (cs == 1) tma_load_3d(st, &map_w_hi, &full_bar[s], kb * TC_BK, tl.co0, tap);
            else tma_load_3d_mc(st + crank * w_slice, &map_w_hi, &full_bar[s], kb * TC_BK, tl.co0 + crank * (int)w_rows, tap, cmask);
            if (TC_DEBUG(P) & 4) { /* expect_tx below was reduced accordingly */ }
            else if (g.relation == GLIS_CONV)
              tma_load_5d(st + 2 * a_bytes, &map_x_hi, &full_bar[s], cpar + kb * TC_BK, c1, c2, c3, tl.n0);
            else
              tma_load_4d(st + 2 * a_bytes, &map_x_hi, &full_bar[s], kb * TC_BK, c1, c3, tl.n0);
            if (P.passes == 3) {
              if (cs == 1) tma_load_3d(st + a_bytes, &map_w_lo, &full_bar[s], kb * TC_BK, tl.co0, tap);
              else tma_load_3d_mc(st + a_bytes + crank * w_slice, &map_w_lo, &full_bar[s], kb * TC_BK, tl.co0 + crank * (int)w_rows, tap, cmask);
              if (TC_DEBUG(P) & 4) {}
              else if (g.relation == GLIS_CONV)
                tma_load_5d(st + 2 * a_bytes + b_bytes, &map_x_lo, &full_bar[s], cpar + kb * TC_BK, c1, c2, c3, tl.n0);
              else
                tma_load_4d(st + 2 * a_bytes + b_bytes, &map_x_lo, &full_bar[s], kb * TC_BK, c1, c3, tl.n0);
            }
            if (++s == P.stages) { s = 0; parity ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TC_BM, P.n_mma, 0, 0);
      const uint64_t desc0 = umma_smem_desc(smem_u32(base), 16, 1024);  // K-major SW128, stage 0, A_hi
      int s = 0; uint32_t parity = 0;
      uint32_t acc = 0, acc_phase = 0;  // bit a of acc_phase = parity of tmem_empty_bar[a] to wait for
      int tr_n = 512;
      for (int grp = cluster_id; grp < P.n_groups; grp += n_clusters) {
        const int id = grp * cs + crank;
        const TcTile tl = tc_tile(P, id);
        if (tl.empty) continue;
        mbar_wait(&tmem_empty_bar[acc], ((acc_phase >> acc) & 1u) ^ 1u);  // epilogue drained this accumulator
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * acc_stride;
        uint32_t accumulate = 0;
        for (int ks = 0; ks < tl.ksteps; ++ks) {
          mbar_wait(&full_bar[s], parity);
          tc_fence_after_sync();
          if (TC_TRACE(P) && blockIdx.x == 0 && tr_n < 1024) TC_TRACE(P)[tr_n++] = global_timer_ns();
          // descriptors differ only in their 14-bit start-address field: one add per MMA operand
          const uint64_t dah0 = desc0 + (uint64_t)(((uint32_t)s * stage_bytes) >> 4);
          const uint64_t dal0 = dah0 + (a_bytes >> 4);
          const uint64_t dbh0 = dah0 + ((2 * a_bytes) >> 4);
          const uint64_t dbl0 = dbh0 + (b_bytes >> 4);
          if (!(TC_DEBUG(P) & 2)) {
            if (P.passes == 3) {
#pragma unroll
              for (int kk = 0; kk < TC_BK / 16; ++kk) {  // +2 = 32 bytes = 16 bf16 along K in the swizzled row
                umma_bf16(tmem_d, dah0 + 2 * kk, dbl0 + 2 * kk, idesc, accumulate);
                umma_bf16(tmem_d, dal0 + 2 * kk, dbh0 + 2 * kk, idesc, 1);
                umma_bf16(tmem_d, dah0 + 2 * kk, dbh0 + 2 * kk, idesc, 1);
                accumulate = 1;
              }
            } else {
#pragma unroll
              for (int kk = 0; kk < TC_BK / 16; ++kk) {
                umma_bf16(tmem_d, dah0 + 2 * kk, dbh0 + 2 * kk, idesc, accumulate);
                accumulate = 1;
              }
            }
          }
          // frees the stage once these MMAs have read it — in EVERY CTA of the cluster, whose producers
          // multicast weight rows into this CTA's stage as well
          if (cs == 1) umma_commit(&empty_bar[s]); else umma_commit_mc(&empty_bar[s], cmask);
          if (++s == P.stages) { s = 0; parity ^= 1; }
        }
        umma_commit(&tmem_full_bar[acc]);
        acc_phase ^= (1u << acc);
        acc ^= 1u;
      }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    // Four warps per TMEM lane quarter; they take every fourth 32-column chunk of the tile.
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int part = (warp - 2) >> 2;       // 0..3
    const int cols = P.tw * P.th * P.tn;
    const int sh = g.relation == GLIS_TCONV ? g.stride_h : 1;
    uint32_t acc = 0, full_phase = 0;
    int tr_e = 1024;
    for (int grp = cluster_id; grp < P.n_groups; grp += n_clusters) {
      const int id = grp * cs + crank;
      const TcTile tl = tc_tile(P, id);
      if (tl.empty) continue;
      const int co = tl.co0 + q * 32 + lane;
      const bool ch_ok = co < g.Co && !(TC_DEBUG(P) & 1) && !tl.ghost;
      float bias = 0.f, ta = 0.f, tb = 0.f;
      if (ch_ok) {
        if (P.bias && tl.split == 0) bias = __ldg(P.bias + co);
        if (P.act == GLIS_ACT_TPRELU) {
          const int ca = P.act_channels > 0 ? co % P.act_channels : co;
          ta = fminf(fmaxf(__ldg(P.act_a + ca), 0.f), 1.f); tb = __ldg(P.act_b + ca);
        }
      }
      const int oy0 = g.relation == GLIS_TCONV ? tl.qy0 * sh + tl.ph.ry : tl.qy0;
      const int ox0 = g.relation == GLIS_TCONV ? tl.ph.rx : 0;
      const long long base = (((long long)tl.n0 * g.Ho + oy0) * g.Wo + ox0) * g.Co + co;
      const int valid_cols = P.tn == 1 ? min(P.th, tl.ph.Hq - tl.qy0) * P.tw : min(P.tn, g.N - tl.n0) * P.th * P.tw;
      mbar_wait(&tmem_full_bar[acc], (full_phase >> acc) & 1u);
      tc_fence_after_sync();
      if (TC_TRACE(P) && blockIdx.x == 0 && warp == 2 && lane == 0 && tr_e < 1024 + 64) TC_TRACE(P)[tr_e++] = global_timer_ns();
      const uint32_t tmem_d = tmem_base + acc * acc_stride + ((uint32_t)(q * 32) << 16);
      // a quarter whose 32 channels are all out of range (Cout <= 64 or 96) has nothing to store
      const bool quarter_ok = tl.co0 + q * 32 < g.Co && !(TC_DEBUG(P) & 1) && !tl.ghost;
      for (int cb = part * 32; cb < cols && quarter_ok; cb += 128) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_d + (uint32_t)cb, v);
        tmem_ld_wait();
        // valid columns form a prefix of the tile: ragged rows (tn == 1) or ragged images (th == Hq) come last
        const int nvalid = ch_ok ? valid_cols - cb : 0;
        const uint32_t* rc = rel + cb;
        switch (P.ep_mode) {
          case 1: tc_epilogue_chunk<GLIS_ACT_NONE, false, true, false>(v, nvalid, base, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 2: tc_epilogue_chunk<GLIS_ACT_TPRELU, true, false, true>(v, nvalid, base, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 3: tc_epilogue_chunk<GLIS_ACT_NONE, false, false, true>(v, nvalid, base, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 4: tc_epilogue_chunk<GLIS_ACT_TPRELU, true, true, false>(v, nvalid, base, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 5: tc_epilogue_chunk<GLIS_ACT_TPRELU, true, true, true>(v, nvalid, base, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 7: tc_epilogue_chunk<GLIS_ACT_TPRELU, false, false, true>(v, nvalid, base, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 8: {   // split K, deterministic form: this share's partial sums into its own slab
            float* slab = P.out_f32 + (long long)tl.split * P.slab_stride;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) slab[base + rc[j]] = __uint_as_float(v[j]) + bias;
            break;
          }
          case 6: {   // split K: add this item's partial sums (one 128-byte reduction per warp and column)
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nvalid) atomicAdd(P.out_f32 + base + rc[j], __uint_as_float(v[j]) + bias);
            break;
          }
          default: {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nvalid) {
                const long long off = base + rc[j];
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
      if (TC_TRACE(P) && blockIdx.x == 0 && warp == 2 && lane == 0 && tr_e < 1024 + 64) TC_TRACE(P)[tr_e++] = global_timer_ns();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      full_phase ^= (1u << acc);
      acc ^= 1u;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (P.cluster > 1) cluster_sync_all();   // nobody leaves while a neighbour may still signal its barriers
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// fp32 -> bf16 hi/lo planes (same element order)
__global__ void split_planes_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, int64_t numel) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t n4 = numel >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    __nv_bfloat16 h[4], l[4];
    split_bf16(v.x, h[0], l[0]); split_bf16(v.y, h[1], l[1]); split_bf16(v.z, h[2], l[2]); split_bf16(v.w, h[3], l[3]);
    reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
    if (lo) reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel;
       i += (int64_t)gridDim.x * blockDim.x) {
    __nv_bfloat16 h, l;
    split_bf16(x[i], h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;  // benign race: every thread resolves the same pointer
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 tensor map, 128-byte swizzle, zero fill out of bounds. dims/box innermost first;
// strides in bytes for dims 1..rank-1.
int make_bf16_map_swz(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box, int swizzle_bytes);

int make_bf16_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                  const uint32_t* box) {
  return make_bf16_map_swz(m, ptr, rank, dims, strides, box, 128);
}

// swizzle_bytes: 128 (64 bf16 per shared-memory row) or 64 (32 per row)
int make_bf16_map_swz(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                      const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  GLIS_REQUIRE(enc != nullptr, GLIS_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides[i];
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), d, s, b, e,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GLIS_REQUIRE(r == CUDA_SUCCESS, GLIS_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
  return GLIS_OK;
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Reasons the tensor-core path does not apply (the caller then uses the fp32 kernel).
int tc_conv_supported(const glis_geom_t* g) {
  // K blocks of 64 channels; a ragged last block is TMA zero fill on the weight side (whatever the
  // pixel box reads beyond Ci is multiplied by zero), at most half of a block wasted
  if (g->Ci % 8 != 0 || (g->Ci % TC_BK != 0 && g->Ci < 32)) return 0;
  if (g->Co < 32) return 0;                               // 3- and 1-channel outputs: < 25 % of the 128 MMA rows
  if (g->dil_h != 1 || g->dil_w != 1) return 0;
  if (g->relation == GLIS_CONV) {
    if (g->Hi % g->stride_h != 0 || g->Wi % g->stride_w != 0) return 0;  // parity view of the input
    if ((int64_t)g->stride_w * g->Ci > 0x7fffffff) return 0;
  }
  int Wq = g->Wo;
  if (g->relation == GLIS_TCONV) Wq = (g->Wo + g->stride_w - 1) / g->stride_w;
  if (Wq > 256) return 0;                                 // one tile row must fit the MMA N
  if (g->relation == GLIS_TCONV && g->Wo % g->stride_w != 0) return 0;  // all phases equally wide
  return 1;
}

static int tc_num_sms() { return plan_sms(); }

// Tile shape, K split and work-item counts of one launch (everything that does not depend on the
// operand pointers).  `plain_out`: the launch writes fp32 sums only (no activation, pre-activation or
// planes), which is what allows a K split.
static int tc_plan(const glis_geom_t* g, bool plain_out, TcConvParams& P) {
  P.g = *g;
  int nphase = 1, Hq = g->Ho, Wq = g->Wo;
  if (g->relation == GLIS_TCONV) {
    nphase = g->stride_h * g->stride_w;
    Hq = (g->Ho + g->stride_h - 1) / g->stride_h;
    Wq = (g->Wo + g->stride_w - 1) / g->stride_w;
  }
  // ---- pixel tile: full rows (tn = 1, th rows) or whole images (th = Hq, tn images), <= 256 columns.
  // The main loop is bound by shared-memory traffic — per k-step the MMAs read 128 weight rows plus
  // N pixel rows — and by how evenly the tiles fill the SMs, so pick the shape that minimises
  //   waves(tiles / #SMs) x (N + 128).
  const int num_sms = tc_num_sms();
  static int nmax_cfg = 0;
  if (!nmax_cfg) {
    const char* e = getenv("GLIS_TC_NMAX");   // tuning knob: upper bound on columns per tile (16..256)
    nmax_cfg = e ? atoi(e) : 256;
    if (nmax_cfg < 16 || nmax_cfg > 256) nmax_cfg = 256;
  }
  const int NMAX = nmax_cfg;
  const int co_tiles = (g->Co + TC_BM - 1) / TC_BM;
  int cs = 1;
  {
    const char* e = getenv("GLIS_TC_CLUSTER");   // CTAs sharing one multicast weight tile: 1, 2 or 4
    cs = e ? atoi(e) : TC_DEFAULT_CLUSTER;
    if (cs != 1 && cs != 2 && cs != 4) cs = 1;
  }
  P.tw = Wq;
  // Co <= 64: stage 64 weight rows instead of 128 (smaller stages -> a deeper TMA pipeline)
  int a_rows = (g->Co <= 64 && cs == 1) ? 64 : TC_BM;
  int depth_penalty = 0;
  {
    const char* e = getenv("GLIS_TC_AROWS");          // tuning knob: 128 = always stage full weight tiles
    if (e && atoi(e) == 128) a_rows = TC_BM;
    const char* d = getenv("GLIS_TC_DEPTH_PENALTY");  // tuning knob: % cost added to tiles that leave only 2 stages
    depth_penalty = d ? atoi(d) : 0;
  }
  const int kblocks = (g->Ci + TC_BK - 1) / TC_BK;
  int ksplit_max = 32;
  {
    const char* e = getenv("GLIS_TC_KSPLIT");   // tuning knob: upper bound on the K split (1 = off)
    if (e) ksplit_max = atoi(e);
    if (ksplit_max < 1) ksplit_max = 1;
  }
  int best_ks = 1;
  {
    // cost ~ waves x (k-steps per item + fixed prologue / epilogue, ~4 k-steps) x shared-memory bytes per
    // k-step (n + 128 rows); splitting K pays one extra pass of reductions over the output and a memset
    long best_cost = -1;
    int best_th = 1, best_tn = 1;
    const int ntaps_max = g->relation == GLIS_CONV ? g->KH * g->KW
                          : ((g->KH + g->stride_h - 1) / g->stride_h) * ((g->KW + g->stride_w - 1) / g->stride_w);
    auto consider = [&](int th, int tn) {
      const int n = round_up(Wq * th * tn, 16);
      if (n > NMAX || n > 256) return;
      const long tiles_x = ((long)((Hq + th - 1) / th) * ((g->N + tn - 1) / tn) + cs - 1) / cs * cs;
      const long tiles = tiles_x * co_tiles * nphase;
      const long slots = num_sms / cs * cs;
      for (int ks = 1; ks <= ksplit_max && ks <= kblocks; ks *= 2) {
        if (ks > 1 && (!plain_out || cs > 1)) break;
        // uneven shares are fine (tc_tile cuts the blocks proportionally) once every share has a few blocks
        if (ks > 1 && kblocks % ks != 0 && kblocks < 4 * ks) break;
        const long waves = (tiles * ks + slots - 1) / slots;
        const long steps = (long)ntaps_max * ((kblocks + ks - 1) / ks) + 4;
        long cost = waves * steps * (n + 128 / cs) + (ks > 1 ? waves * 2 * n + 1024 : 0);
        const int ar = n >= 64 ? a_rows : TC_BM;
        if ((219 * 1024) / (256 * (ar + n)) < 3) cost += cost * depth_penalty / 100;
        cost = cost * 1024 + n;   // tie-break: smaller tiles
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_th = th; best_tn = tn; best_ks = ks; }
      }
    };
    for (int th = 1; th <= Hq && Wq * th <= 256; ++th) consider(th, 1);
    for (int tn = 2; tn <= g->N && Wq * Hq * tn <= 256; ++tn) consider(Hq, tn);
    GLIS_REQUIRE(best_cost >= 0, GLIS_E_UNSUPPORTED, "glis_conv_forward_bf16: no pixel tile fits");
    P.th = best_th;
    P.tn = best_tn;
  }
  P.n_mma = round_up(P.tw * P.th * P.tn, 16);
  P.a_rows = P.n_mma >= 64 ? a_rows : TC_BM;   // (a 128-row MMA read starting in the lo tile must stay inside the stage)
  P.tmem_cols = 64;  // two accumulators of tmem_cols / 2 columns each
  while (P.tmem_cols < 2 * P.n_mma) P.tmem_cols *= 2;
  P.kblocks = (g->Ci + TC_BK - 1) / TC_BK;
  P.tiles_h = (Hq + P.th - 1) / P.th;
  const int tiles_n = (g->N + P.tn - 1) / P.tn;
  P.tiles_x = (P.tiles_h * tiles_n + cs - 1) / cs * cs;   // padded: the tiles of a cluster share (phase, channel tile)
  P.tiles_co = (g->Co + TC_BM - 1) / TC_BM;
  P.total_tiles = P.tiles_x * P.tiles_co * nphase;
  P.cluster = cs;
  P.ksplit = best_ks;
  P.n_groups = P.total_tiles / cs * best_ks;
  return GLIS_OK;
}

// The K split a plain-output launch of this geometry would use (1 = none): lets the host decide to run a
// fused-epilogue layer as split-K sums + one pointwise pass when its tiles alone cannot fill the machine.
int tc_conv_halo_applies(const glis_geom_t* g, int plain_out);
int tc_conv_halo_ksplit(const glis_geom_t* g);
int tc_conv_halo_forward(const glis_geom_t* g, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
                         const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, const glis_epilogue_t* ep, float* out_f32,
                         __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int precision, cudaStream_t st);

int tc_conv_pair_applies(const glis_geom_t* g, int plain_out);
int tc_conv_pair_ksplit(const glis_geom_t* g);
int tc_conv_pair_forward(const glis_geom_t* g, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
                         const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, const glis_epilogue_t* ep, float* out_f32,
                         __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int precision, cudaStream_t st);

int tc_conv_plan_ksplit(const glis_geom_t* g) {
  if (!tc_conv_supported(g)) return 1;
  if (tc_conv_halo_applies(g, 1)) return tc_conv_halo_ksplit(g);   // the kernel that would run
  if (tc_conv_pair_applies(g, 1)) return tc_conv_pair_ksplit(g);
  TcConvParams P;
  if (tc_plan(g, true, P) != GLIS_OK) return 1;
  return P.ksplit;
}

// The whole plan, for inspection (tests/test_host_cpu.py checks its invariants over many geometries):
// out = {tw, th, tn, n_mma, tmem_cols, kblocks, ksplit, a_rows, stages, tiles_h, tiles_x, tiles_co, total_tiles,
//        n_groups, dynamic shared memory bytes}
int tc_conv_plan_describe(const glis_geom_t* g, int plain_out, int out[15]) {
  if (!tc_conv_supported(g)) return GLIS_E_UNSUPPORTED;
  TcConvParams P;
  int rc = tc_plan(g, plain_out != 0, P);
  if (rc != GLIS_OK) return rc;
  const size_t stage_bytes = 2 * (size_t)P.a_rows * 128 + 2 * (size_t)P.n_mma * 128;
  int stages = (int)((219 * 1024) / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256 + 1024;
  const int v[15] = {P.tw, P.th, P.tn, P.n_mma, P.tmem_cols, P.kblocks, P.ksplit, P.a_rows, stages, P.tiles_h,
                     P.tiles_x, P.tiles_co, P.total_tiles, P.n_groups, (int)smem};
  for (int i = 0; i < 15; ++i) out[i] = v[i];
  return GLIS_OK;
}

int tc_conv_forward(const glis_geom_t* g, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
                    const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, const glis_epilogue_t* ep, float* out_f32,
                    __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int precision, cudaStream_t st) {
  GLIS_REQUIRE(tc_conv_supported(g), GLIS_E_UNSUPPORTED, "glis_conv_forward_bf16: geometry not tileable for tcgen05");
  {
    // layers whose taps share pixel boxes (4x4 stride 2, 3x3 stride 1) run on the halo kernel (tc_conv2.cu)
    const bool plain = ep->act == GLIS_ACT_NONE && !ep->preact && out_f32 && !out_hi;
    if (tc_conv_halo_applies(g, plain))
      return tc_conv_halo_forward(g, x_hi, x_lo, w_hi, w_lo, ep, out_f32, out_hi, out_lo, precision, st);
    // >= 256 output channels on maps too narrow for the halo form: CTA pairs (tc_conv_pair.cu)
    if (tc_conv_pair_applies(g, plain))
      return tc_conv_pair_forward(g, x_hi, x_lo, w_hi, w_lo, ep, out_f32, out_hi, out_lo, precision, st);
  }
  const int passes = precision == GLIS_PREC_BF16X3 ? 3 : 1;
  GLIS_REQUIRE(x_hi && w_hi && (passes == 1 || (x_lo && w_lo)), GLIS_E_BADARG,
               "glis_conv_forward_bf16: missing hi/lo operand planes");
  TcConvParams P;
  {
    const bool plain = ep->act == GLIS_ACT_NONE && !ep->preact && out_f32 && !out_hi;
    int rc = tc_plan(g, plain, P);
    if (rc != GLIS_OK) return rc;
  }
  const int cs = P.cluster;
  const int best_ks = P.ksplit;
  const int num_sms = tc_num_sms();
  P.passes = passes;
  const size_t stage_bytes = 2 * (size_t)P.a_rows * 128 + 2 * (size_t)P.n_mma * 128;
  int stages = (int)((219 * 1024) / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  GLIS_REQUIRE(stages >= 2, GLIS_E_UNSUPPORTED, "glis_conv_forward_bf16: tile does not fit shared memory");
  P.stages = stages;
  P.bias = ep->bias; P.act = ep->act; P.act_a = ep->act_a; P.act_b = ep->act_b; P.preact = ep->preact;
  P.act_channels = ep->act_channels;
  P.out_f32 = out_f32; P.out_hi = out_hi; P.out_lo = out_lo;
  P.ep_mode = 0;
  if (ep->act == GLIS_ACT_NONE && !ep->preact && out_f32 && !out_hi) P.ep_mode = 1;
  else if (ep->act == GLIS_ACT_TPRELU && ep->preact && !out_f32 && out_hi) P.ep_mode = 2;
  else if (ep->act == GLIS_ACT_NONE && !ep->preact && !out_f32 && out_hi) P.ep_mode = 3;
  else if (ep->act == GLIS_ACT_TPRELU && ep->preact && out_f32 && !out_hi) P.ep_mode = 4;
  else if (ep->act == GLIS_ACT_TPRELU && ep->preact && out_f32 && out_hi) P.ep_mode = 5;
  else if (ep->act == GLIS_ACT_TPRELU && !ep->preact && !out_f32 && out_hi) P.ep_mode = 7;   // no-grad forward
  P.slab_stride = 0;
  if (best_ks > 1 && ep->split_slabs > 0) {
    GLIS_REQUIRE(ep->split_slabs >= best_ks, GLIS_E_BADARG, "glis_conv_forward_bf16: %d slabs for a %d-way K split",
                 ep->split_slabs, best_ks);
    P.ep_mode = 8;
    P.slab_stride = (long long)g->N * g->Ho * g->Wo * g->Co;
  } else if (best_ks > 1) {
    P.ep_mode = 6;
    cudaError_t me = cudaMemsetAsync(out_f32, 0, sizeof(float) * (size_t)g->N * g->Ho * g->Wo * g->Co, st);
    GLIS_REQUIRE(me == cudaSuccess, GLIS_E_CUDA, "glis_conv_forward_bf16: memset failed: %s", cudaGetErrorString(me));
  }
  {
    const char* dbg = getenv("GLIS_TC_DEBUG");
    P.debug = dbg ? atoi(dbg) : 0;
    const char* trc = getenv("GLIS_TC_TRACE");  // hex device address of a >= 1088-entry u64 buffer
    P.trace = trc ? (unsigned long long*)strtoull(trc, nullptr, 16) : nullptr;
  }

  // ---- tensor maps
  CUtensorMap mw_hi, mw_lo, mx_hi, mx_lo;
  const int T = g->KH * g->KW;
  {
    const uint64_t dims[3] = {(uint64_t)g->Ci, (uint64_t)g->Co, (uint64_t)T};
    const uint64_t strides[2] = {(uint64_t)g->Ci * 2, (uint64_t)g->Ci * g->Co * 2};
    const uint32_t box[3] = {TC_BK, (uint32_t)(P.a_rows / cs), 1};   // one CTA's share of the (multicast) weight tile
    int rc = make_bf16_map(&mw_hi, w_hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_bf16_map(&mw_lo, passes == 3 ? w_lo : w_hi, 3, dims, strides, box);
    if (rc) return rc;
  }
  if (g->relation == GLIS_CONV) {
    const uint64_t C = g->Ci, W = g->Wi, H = g->Hi, sw = g->stride_w, sh = g->stride_h;
    const uint64_t dims[5] = {sw * C, W / sw, sh, H / sh, (uint64_t)g->N};
    const uint64_t strides[4] = {sw * C * 2, W * C * 2, sh * W * C * 2, H * W * C * 2};
    const uint32_t box[5] = {TC_BK, (uint32_t)P.tw, 1, (uint32_t)P.th, (uint32_t)P.tn};
    int rc = make_bf16_map(&mx_hi, x_hi, 5, dims, strides, box);
    if (rc) return rc;
    rc = make_bf16_map(&mx_lo, passes == 3 ? x_lo : x_hi, 5, dims, strides, box);
    if (rc) return rc;
  } else {
    const uint64_t C = g->Ci, W = g->Wi, H = g->Hi;
    const uint64_t dims[4] = {C, W, H, (uint64_t)g->N};
    const uint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    const uint32_t box[4] = {TC_BK, (uint32_t)P.tw, (uint32_t)P.th, (uint32_t)P.tn};
    int rc = make_bf16_map(&mx_hi, x_hi, 4, dims, strides, box);
    if (rc) return rc;
    rc = make_bf16_map(&mx_lo, passes == 3 ? x_lo : x_hi, 4, dims, strides, box);
    if (rc) return rc;
  }

  const size_t smem = (size_t)stages * stage_bytes + 1024 /*alignment slack*/ + 256 /*barriers*/ + 1024 /*column offsets*/;
  GLIS_REQUIRE(P.tmem_cols <= 512, GLIS_E_UNSUPPORTED, "glis_conv_forward_bf16: accumulators exceed TMEM");
  {
    cudaError_t e = ensure_max_dynamic_smem(reinterpret_cast<const void*>(tc_conv_kernel), 227 * 1024);
    GLIS_REQUIRE(e == cudaSuccess, GLIS_E_CUDA, "cudaFuncSetAttribute(tc_conv_kernel): %s", cudaGetErrorString(e));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1] = pdl_attr();
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent: as many clusters as the device can hold at once (one CTA per SM), at most one per group
  static int max_clusters[5] = {0, 0, 0, 0, 0};
  if (!max_clusters[cs]) {
    int nc = 0;
    cfg.gridDim = dim3(num_sms / cs * cs);
    if (cs == 1 || cudaOccupancyMaxActiveClusters(&nc, tc_conv_kernel, &cfg) != cudaSuccess || nc <= 0) nc = num_sms / cs;
    (void)cudaGetLastError();
    max_clusters[cs] = nc;
  }
  const int n_clusters = P.n_groups < max_clusters[cs] ? P.n_groups : max_clusters[cs];
  cfg.gridDim = dim3(n_clusters * cs);
  cfg.numAttrs = pdl_applies(cfg.gridDim, smem) ? 2 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, tc_conv_kernel, mw_hi, mw_lo, mx_hi, mx_lo, P);
  GLIS_REQUIRE(le == cudaSuccess, GLIS_E_CUDA, "glis_conv_forward_bf16: launch failed: %s", cudaGetErrorString(le));
  GLIS_CHECK_LAUNCH("glis_conv_forward_bf16");
  return GLIS_OK;
}

int split_planes(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t numel, cudaStream_t st) {
  int blocks = (int)((numel / 4 + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaError_t le = launch_pdl(split_planes_kernel, dim3(blocks), dim3(256), 0, st, x, hi, lo, numel);
  GLIS_REQUIRE(le == cudaSuccess, GLIS_E_CUDA, "glis_split_bf16: launch failed: %s", cudaGetErrorString(le));
  GLIS_CHECK_LAUNCH("glis_split_bf16");
  return GLIS_OK;
}

}  // namespace glis

// tcgen05 implicit-GEMM convolution / transposed convolution on CTA PAIRS (cta_group::2), forward and data gradient.
//
// Same contraction as tc_conv.cu — D[co, pix] = sum_tap sum_ci W[tap][co][ci] * X[gather(pix, tap)][ci], bf16 hi/lo
// planes, three MMAs per k-step, fp32 accumulators in TMEM — for layers with >= 256 output channels.  A cluster of
// two CTAs works on ONE tile of 256 channels x N pixels with ONE tcgen05.mma.cta_group::2 of M = 256 per k-step part:
//
//   * each CTA stages ITS 128 weight rows (channels co0 + 128 * rank) and HALF of the pixel tile (N / 2 rows of B),
//   * the even-ranked CTA (the leader) issues every MMA; the tensor cores of both SMs execute it, each reading the
//     operands out of its own shared memory, and each accumulates the 128 TMEM lanes of its own channels,
//   * both CTAs run their own epilogue over all N columns of their lanes.
//
// What it buys: per 64-channel k-step a CTA of tc_conv.cu pulls 32 KB of weights + N * 256 B of pixels from L2 and its
// MMAs read all of it out of shared memory three times; a CTA of a pair pulls (and re-reads) only half of the pixel
// tile — 64 instead of 96 KB per k-step at N = 256, so 3 stages fit where 2 did, and the shared-memory read
// traffic of the MMAs drops by a third (the pixel-side operand is what the three passes re-read most).
//
// Protocol (PTX forms as in the vendored CUTLASS headers, see sm100.cuh):
//   full[s]       lives in the LEADER: its producer thread does arrive.expect_tx(bytes of BOTH CTAs); both CTAs' TMA
//                 loads (cp.async.bulk.tensor...cta_group::2) land in their own shared memory and complete_tx on it.
//   empty[s]      one per CTA, freed by the leader's tcgen05.commit.cta_group::2 ... multicast::cluster 0b11.
//   tmem_full[a]  one per CTA, same multicast commit after the last k-step of a tile: both epilogues start.
//   tmem_empty[a] lives in the leader, 2 x 16 arrivals: the follower's epilogue warps arrive through a mapa'd address.
//   TMEM is allocated / released with the cta_group::2 forms by warp 1 of both CTAs.
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"
#include "tc_common.cuh"

namespace glis {

int make_bf16_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                  const uint32_t* box);
int tc_conv_supported(const glis_geom_t* g);

constexpr uint32_t TP_SKIP = 0xffffffffu;

struct PairParams {
  glis_geom_t g;
  int tw, th, tn;        // the pair's pixel tile on the (phase) output grid; tw spans the full width
  int hh, hn;            // one half of it: hh rows x hn images (th x tn split along rows or along images)
  int n_half;            // valid pixel rows of a half: tw * hh * hn
  int n_half_pad;        // rounded up to 8: B rows a CTA stages = N / 2 of the MMA
  int n_mma;             // 2 * n_half_pad
  int tmem_cols;         // two accumulators of tmem_cols / 2 columns
  int kblocks, passes, stages;
  int tiles_h, tiles_x, tiles_cp, total_tiles, n_groups, ksplit;
  const float* bias; int act; const float* act_a; const float* act_b; int act_channels;
  float* preact; float* out_f32; __nv_bfloat16* out_hi; __nv_bfloat16* out_lo;
  int ep_mode;
  long long slab_stride;
};

struct PairTile {
  TcPhase ph;
  int qy0, n0, co0, ntaps, ksteps, kb_beg, kb_end, split;
  bool empty;
};

__device__ __forceinline__ PairTile pair_tile(const PairParams& P, int item) {
  PairTile t;
  const int id = item / P.ksplit;
  t.split = item - id * P.ksplit;
  t.kb_beg = (int)((long long)P.kblocks * t.split / P.ksplit);
  t.kb_end = (int)((long long)P.kblocks * (t.split + 1) / P.ksplit);
  const int per_phase = P.tiles_x * P.tiles_cp;
  const int z = id / per_phase, rem = id - z * per_phase;
  const int y = rem / P.tiles_x, x = rem - y * P.tiles_x;
  t.ph = tc_phase(P.g, z);
  const int tile_h = x % P.tiles_h, tile_n = x / P.tiles_h;
  t.qy0 = tile_h * P.th;
  t.n0 = tile_n * P.tn;
  t.co0 = y * 256;
  t.ntaps = t.ph.nth * t.ph.ntw;
  t.ksteps = t.ntaps * (t.kb_end - t.kb_beg);
  t.empty = t.ph.Hq <= 0 || t.ph.Wq <= 0 || t.ksteps == 0;
  return t;
}

// One 32-column chunk of this lane's channel; `ok` bit j = column j is a pixel of the tensor (halves are padded to 8
// rows and tiles may hang over the last image row / the last image).
template <int ACT, bool PREACT, bool F32, bool PLANES>
__device__ __forceinline__ void pair_chunk(const uint32_t (&v)[32], uint32_t ok, long long base,
                                           const uint32_t* __restrict__ rel, float bias, float ta, float tb,
                                           float* __restrict__ preact, float* __restrict__ out_f32,
                                           __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if ((ok >> j) & 1u) {
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

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_pair_kernel(const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                    const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                    const __grid_constant__ PairParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const glis_geom_t& g = P.g;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr uint32_t a_bytes = 128u * 128u;                        // this CTA's 128 weight rows, one plane
  const uint32_t b_bytes = (uint32_t)P.n_half_pad * 128u;          // this CTA's half of the pixel tile, one plane
  const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;          // [A hi][A lo][B hi][B lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)P.stages * stage_bytes);
  uint64_t* full_bar = bars;                                       // (used in the leader)
  uint64_t* empty_bar = bars + TC_MAX_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * TC_MAX_STAGES;              // [2]
  uint64_t* tmem_empty_bar = bars + 2 * TC_MAX_STAGES + 2;         // [2] (used in the leader)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_MAX_STAGES + 4);
  uint32_t* rel = tmem_slot + 4;      // [256] accumulator column -> element offset from the tile origin, or TP_SKIP
  uint32_t* pos = rel + 256;          // [256] (image offset << 16) | row offset of the column inside the tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();      // 0 = leader
  const bool leader = rank == 0;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_w_hi); tma_prefetch_desc(&map_x_hi);
    if (P.passes == 3) { tma_prefetch_desc(&map_w_lo); tma_prefetch_desc(&map_x_lo); }
    for (int s = 0; s < P.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 2 * TC_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, (uint32_t)P.tmem_cols);
  {
    // split along rows: hn == tn, halves are row blocks of hh rows; split along images: hh == th, halves of hn images
    const int sh = g.relation == GLIS_TCONV ? g.stride_h : 1, sw = g.relation == GLIS_TCONV ? g.stride_w : 1;
    const bool by_rows = P.hn == P.tn;
    const int per_img = P.tw * P.hh;
    for (int c = threadIdx.x; c < 256; c += TC_THREADS) {
      const int half = c / P.n_half_pad, j = c - half * P.n_half_pad;
      uint32_t r = TP_SKIP, ps = 0;
      if (c < P.n_mma && j < P.n_half) {
        const int in_ = j / per_img, q = j - in_ * per_img, ih = q / P.tw, iw = q - ih * P.tw;
        const int img = in_ + (by_rows ? 0 : half * P.hn), row = ih + (by_rows ? half * P.hh : 0);
        r = (uint32_t)((((long long)img * g.Ho + (long long)row * sh) * g.Wo + (long long)iw * sw) * g.Co);
        ps = ((uint32_t)img << 16) | (uint32_t)row;
      }
      rel[c] = r; pos[c] = ps;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers exist before anybody signals them
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_stride = (uint32_t)P.tmem_cols / 2;
  const int pair_id = (int)blockIdx.x >> 1, n_pairs = (int)gridDim.x >> 1;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      const uint32_t planes = P.passes == 3 ? 2u : 1u;
      const uint32_t tx_pair = 2u * planes * (a_bytes + (uint32_t)P.n_half * 128u);   // both CTAs' boxes
      const bool by_rows = P.hn == P.tn;
      const int half_row = by_rows ? (int)rank * P.hh : 0, half_img = by_rows ? 0 : (int)rank * P.hn;
      int s = 0; uint32_t parity = 0;
      for (int grp = pair_id; grp < P.n_groups; grp += n_pairs) {
        const PairTile tl = pair_tile(P, grp);
        if (tl.empty) continue;
        const int rot = (int)(((uint32_t)pair_id * 5u + (uint32_t)grp * 3u) % (uint32_t)tl.ntaps);
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
            c3 = tl.qy0 + half_row + (ey - pary) / g.stride_h;
          } else {
            kh = tl.ph.py + jh * g.stride_h; kw = tl.ph.px + jw * g.stride_w;
            c1 = (tl.ph.rx + g.pad_w - kw) / g.stride_w;
            c3 = tl.qy0 + half_row + (tl.ph.ry + g.pad_h - kh) / g.stride_h;
          }
          const int tap = kh * g.KW + kw;
          const int co = tl.co0 + 128 * (int)rank, n0 = tl.n0 + half_img;
          for (int kb = tl.kb_beg; kb < tl.kb_end; ++kb) {
            mbar_wait(&empty_bar[s], parity ^ 1);
            uint8_t* st = base + (size_t)s * stage_bytes;
            if (leader) mbar_arrive_expect_tx(&full_bar[s], tx_pair);
            const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);      // the LEADER's full barrier
            tma_load_3d_pair(st, &map_w_hi, fb, kb * TC_BK, co, tap);
            if (g.relation == GLIS_CONV) tma_load_5d_pair(st + 2 * a_bytes, &map_x_hi, fb, cpar + kb * TC_BK, c1, c2, c3, n0);
            else tma_load_4d_pair(st + 2 * a_bytes, &map_x_hi, fb, kb * TC_BK, c1, c3, n0);
            if (P.passes == 3) {
              tma_load_3d_pair(st + a_bytes, &map_w_lo, fb, kb * TC_BK, co, tap);
              if (g.relation == GLIS_CONV)
                tma_load_5d_pair(st + 2 * a_bytes + b_bytes, &map_x_lo, fb, cpar + kb * TC_BK, c1, c2, c3, n0);
              else
                tma_load_4d_pair(st + 2 * a_bytes + b_bytes, &map_x_lo, fb, kb * TC_BK, c1, c3, n0);
            }
            if (++s == P.stages) { s = 0; parity ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (lane == 0 && leader) {
      const uint32_t idesc = umma_idesc_bf16(256, P.n_mma, 0, 0);
      const uint64_t desc0 = umma_smem_desc(smem_u32(base), 16, 1024);
      int s = 0; uint32_t parity = 0;
      uint32_t acc = 0, acc_phase = 0;
      for (int grp = pair_id; grp < P.n_groups; grp += n_pairs) {
        const PairTile tl = pair_tile(P, grp);
        if (tl.empty) continue;
        mbar_wait(&tmem_empty_bar[acc], ((acc_phase >> acc) & 1u) ^ 1u);   // both epilogues drained this accumulator
        tc_fence_after_sync();
        const uint32_t tmem_d = tmem_base + acc * acc_stride;
        uint32_t accumulate = 0;
        for (int ks = 0; ks < tl.ksteps; ++ks) {
          mbar_wait(&full_bar[s], parity);
          tc_fence_after_sync();
          const uint64_t dah0 = desc0 + (uint64_t)(((uint32_t)s * stage_bytes) >> 4);
          const uint64_t dal0 = dah0 + (a_bytes >> 4);
          const uint64_t dbh0 = dah0 + ((2 * a_bytes) >> 4);
          const uint64_t dbl0 = dbh0 + (b_bytes >> 4);
          if (P.passes == 3) {
#pragma unroll
            for (int kk = 0; kk < TC_BK / 16; ++kk) {
              umma_bf16_pair(tmem_d, dah0 + 2 * kk, dbl0 + 2 * kk, idesc, accumulate);
              umma_bf16_pair(tmem_d, dal0 + 2 * kk, dbh0 + 2 * kk, idesc, 1);
              umma_bf16_pair(tmem_d, dah0 + 2 * kk, dbh0 + 2 * kk, idesc, 1);
              accumulate = 1;
            }
          } else {
#pragma unroll
            for (int kk = 0; kk < TC_BK / 16; ++kk) {
              umma_bf16_pair(tmem_d, dah0 + 2 * kk, dbh0 + 2 * kk, idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit_pair(&empty_bar[s], 0x3);        // frees the stage in both CTAs
          if (++s == P.stages) { s = 0; parity ^= 1; }
        }
        umma_commit_pair(&tmem_full_bar[acc], 0x3);    // both epilogues
        acc_phase ^= (1u << acc);
        acc ^= 1u;
      }
    }
  } else {
    // ===================== epilogue (warps 2..17, both CTAs: their own 128 channels, all N columns) ==========
    const int q = warp & 3;
    const int part = (warp - 2) >> 2;
    const int sh = g.relation == GLIS_TCONV ? g.stride_h : 1;
    uint32_t acc = 0, full_phase = 0;
    const uint32_t te0 = mapa_u32(smem_u32(&tmem_empty_bar[0]), 0), te1 = mapa_u32(smem_u32(&tmem_empty_bar[1]), 0);
    for (int grp = pair_id; grp < P.n_groups; grp += n_pairs) {
      const PairTile tl = pair_tile(P, grp);
      if (tl.empty) continue;
      const int co = tl.co0 + 128 * (int)rank + q * 32 + lane;
      const bool ch_ok = co < g.Co;
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
      const long long obase = (((long long)tl.n0 * g.Ho + oy0) * g.Wo + ox0) * g.Co + co;
      const int lim_img = g.N - tl.n0, lim_row = tl.ph.Hq - tl.qy0;      // (<= 0 for a ghost tile: nothing stored)
      mbar_wait(&tmem_full_bar[acc], (full_phase >> acc) & 1u);
      tc_fence_after_sync();
      const uint32_t tmem_d = tmem_base + acc * acc_stride + ((uint32_t)(q * 32) << 16);
      const bool quarter_ok = tl.co0 + 128 * (int)rank + q * 32 < g.Co && lim_img > 0 && lim_row > 0;
      for (int cb = part * 32; cb < P.n_mma && quarter_ok; cb += 128) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_d + (uint32_t)cb, v);
        tmem_ld_wait();
        const uint32_t* rc = rel + cb;
        uint32_t ok = 0;
        if (ch_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const uint32_t ps = pos[cb + j];
            if (rc[j] != TP_SKIP && (int)(ps >> 16) < lim_img && (int)(ps & 0xffffu) < lim_row) ok |= 1u << j;
          }
        }
        switch (P.ep_mode) {
          case 1: pair_chunk<GLIS_ACT_NONE, false, true, false>(v, ok, obase, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 2: pair_chunk<GLIS_ACT_TPRELU, true, false, true>(v, ok, obase, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 3: pair_chunk<GLIS_ACT_NONE, false, false, true>(v, ok, obase, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 4: pair_chunk<GLIS_ACT_TPRELU, true, true, false>(v, ok, obase, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 5: pair_chunk<GLIS_ACT_TPRELU, true, true, true>(v, ok, obase, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 7: pair_chunk<GLIS_ACT_TPRELU, false, false, true>(v, ok, obase, rc, bias, ta, tb, P.preact, P.out_f32, P.out_hi, P.out_lo); break;
          case 8: {   // split K, deterministic form: this share's partial sums into its own slab
            float* slab = P.out_f32 + (long long)tl.split * P.slab_stride;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if ((ok >> j) & 1u) slab[obase + rc[j]] = __uint_as_float(v[j]) + bias;
            break;
          }
          case 6: {   // split K: add this item's partial sums
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if ((ok >> j) & 1u) atomicAdd(P.out_f32 + obase + rc[j], __uint_as_float(v[j]) + bias);
            break;
          }
          default: {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if ((ok >> j) & 1u) {
                const long long off = obase + rc[j];
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
      if (lane == 0) mbar_arrive_cluster(acc ? te1 : te0);     // the LEADER's barrier (its own address for the leader)
      full_phase ^= (1u << acc);
      acc ^= 1u;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();          // nobody leaves while the peer's MMAs read its shared memory or signal its barriers
  if (warp == 1) tmem_dealloc_pair(tmem_base, (uint32_t)P.tmem_cols);
}

// ------------------------------------------------------------------ host side
static int pair_round_up(int a, int b) { return (a + b - 1) / b * b; }

// GLIS_TC_PAIR=1 routes the eligible launches here.  Off by default — measured (B200, config 2): bit-identical sums
// to the one-CTA kernel, 2-9 % faster on the fused-epilogue launches of the 2B-image pass (D level 3: 51.7 -> 48.6 us),
// 3-4 % SLOWER on the split-K data gradients at batch 64, and the whole iteration 1.67 vs 1.65 ms: these layers are
// bound by how 128 tiles fill 148 SMs and by the MMA pipe itself, not by operand traffic.
static int pair_enabled() {
  const char* e = getenv("GLIS_TC_PAIR");
  return e && atoi(e) != 0;
}

static int pair_num_sms() { return plan_sms(); }

// Tile shape and K split, or GLIS_E_UNSUPPORTED when the pair form does not apply.
static int pair_plan(const glis_geom_t* g, bool plain_out, PairParams& P) {
  if (!pair_enabled() || !tc_conv_supported(g)) return GLIS_E_UNSUPPORTED;
  if (g->Co < 256 || g->Co % 256 != 0) return GLIS_E_UNSUPPORTED;
  if (g->Ci % TC_BK != 0) return GLIS_E_UNSUPPORTED;
  P.g = *g;
  int nphase = 1, Hq = g->Ho, Wq = g->Wo;
  if (g->relation == GLIS_TCONV) {
    nphase = g->stride_h * g->stride_w;
    Hq = (g->Ho + g->stride_h - 1) / g->stride_h;
    Wq = (g->Wo + g->stride_w - 1) / g->stride_w;
  }
  const int slots = pair_num_sms() / 2;            // pairs resident at once
  const int co_pairs = g->Co / 256;
  const int kblocks = g->Ci / TC_BK;
  const int ntaps_max = g->relation == GLIS_CONV ? g->KH * g->KW
                        : ((g->KH + g->stride_h - 1) / g->stride_h) * ((g->KW + g->stride_w - 1) / g->stride_w);
  int ksplit_max = 32;
  {
    const char* e = getenv("GLIS_TC_KSPLIT");
    if (e) ksplit_max = atoi(e);
    if (ksplit_max < 1) ksplit_max = 1;
  }
  long best_cost = -1;
  int best_th = 0, best_tn = 0, best_hh = 0, best_hn = 0, best_ks = 1;
  auto consider = [&](int th, int tn, int hh, int hn) {
    const int n_half = Wq * hh * hn, n_half_pad = pair_round_up(n_half, 8);
    if (n_half_pad > 128 || n_half < 8) return;
    const int n = 2 * n_half_pad;
    const long tiles = (long)((Hq + th - 1) / th) * ((g->N + tn - 1) / tn) * co_pairs * nphase;
    for (int ks = 1; ks <= ksplit_max && ks <= kblocks; ks *= 2) {
      if (ks > 1 && !plain_out) break;
      if (ks > 1 && kblocks % ks != 0 && kblocks < 4 * ks) break;
      const long waves = (tiles * ks + slots - 1) / slots;
      const long steps = (long)ntaps_max * ((kblocks + ks - 1) / ks) + 4;
      // per CTA and k-step: 128 weight rows + half of the pixel tile
      long cost = waves * steps * (n_half_pad + 128) + (ks > 1 ? waves * 2 * n + 1024 : 0);
      cost = cost * 1024 + n;
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost; best_th = th; best_tn = tn; best_hh = hh; best_hn = hn; best_ks = ks;
      }
    }
  };
  for (int th = 2; th <= Hq + 1; th += 2) consider(th, 1, th / 2, 1);            // rows, split in two row blocks
  for (int tn = 2; tn <= g->N + 1; tn += 2) consider(Hq, tn, Hq, tn / 2);        // whole images, split in two groups
  if (best_cost < 0) return GLIS_E_UNSUPPORTED;
  P.tw = Wq; P.th = best_th; P.tn = best_tn; P.hh = best_hh; P.hn = best_hn;
  P.n_half = Wq * best_hh * best_hn;
  P.n_half_pad = pair_round_up(P.n_half, 8);
  P.n_mma = 2 * P.n_half_pad;
  P.tmem_cols = 64;
  while (P.tmem_cols < 2 * P.n_mma) P.tmem_cols *= 2;
  P.kblocks = kblocks;
  P.tiles_h = (Hq + P.th - 1) / P.th;
  P.tiles_x = P.tiles_h * ((g->N + P.tn - 1) / P.tn);
  P.tiles_cp = co_pairs;
  P.total_tiles = P.tiles_x * co_pairs * nphase;
  P.ksplit = best_ks;
  P.n_groups = P.total_tiles * best_ks;
  const size_t stage_bytes = 2 * 128 * 128 + 2 * (size_t)P.n_half_pad * 128;
  int stages = (int)((219 * 1024) / stage_bytes);
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  if (stages < 2) return GLIS_E_UNSUPPORTED;
  P.stages = stages;
  return GLIS_OK;
}

int tc_conv_pair_applies(const glis_geom_t* g, int plain_out) {
  PairParams P;
  return pair_plan(g, plain_out != 0, P) == GLIS_OK ? 1 : 0;
}

int tc_conv_pair_ksplit(const glis_geom_t* g) {
  PairParams P;
  if (pair_plan(g, true, P) != GLIS_OK) return 0;
  return P.ksplit;
}

// out = {tw, th, tn, hh, hn, n_half, n_mma, tmem_cols, kblocks, ksplit, stages, tiles_h, tiles_x, co pairs, total_tiles, n_groups}
int tc_conv_pair_describe(const glis_geom_t* g, int plain_out, int out[16]) {
  PairParams P;
  int rc = pair_plan(g, plain_out != 0, P);
  if (rc != GLIS_OK) return rc;
  const int v[16] = {P.tw, P.th, P.tn, P.hh, P.hn, P.n_half, P.n_mma, P.tmem_cols, P.kblocks, P.ksplit, P.stages, P.tiles_h,
                     P.tiles_x, P.tiles_cp, P.total_tiles, P.n_groups};
  for (int i = 0; i < 16; ++i) out[i] = v[i];
  return GLIS_OK;
}

int tc_conv_pair_forward(const glis_geom_t* g, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
                         const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, const glis_epilogue_t* ep, float* out_f32,
                         __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int precision, cudaStream_t st) {
  const int passes = precision == GLIS_PREC_BF16X3 ? 3 : 1;
  GLIS_REQUIRE(x_hi && w_hi && (passes == 1 || (x_lo && w_lo)), GLIS_E_BADARG,
               "glis_conv_forward_bf16: missing hi/lo operand planes");
  PairParams P;
  const bool plain = ep->act == GLIS_ACT_NONE && !ep->preact && out_f32 && !out_hi;
  int rc = pair_plan(g, plain, P);
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
    const uint32_t box[3] = {TC_BK, 128, 1};
    rc = make_bf16_map(&mw_hi, w_hi, 3, dims, strides, box);
    if (rc) return rc;
    rc = make_bf16_map(&mw_lo, passes == 3 ? w_lo : w_hi, 3, dims, strides, box);
    if (rc) return rc;
  }
  if (g->relation == GLIS_CONV) {
    const uint64_t C = g->Ci, W = g->Wi, H = g->Hi, sw = g->stride_w, sh = g->stride_h;
    const uint64_t dims[5] = {sw * C, W / sw, sh, H / sh, (uint64_t)g->N};
    const uint64_t strides[4] = {sw * C * 2, W * C * 2, sh * W * C * 2, H * W * C * 2};
    const uint32_t box[5] = {TC_BK, (uint32_t)P.tw, 1, (uint32_t)P.hh, (uint32_t)P.hn};
    rc = make_bf16_map(&mx_hi, x_hi, 5, dims, strides, box);
    if (rc) return rc;
    rc = make_bf16_map(&mx_lo, passes == 3 ? x_lo : x_hi, 5, dims, strides, box);
    if (rc) return rc;
  } else {
    const uint64_t C = g->Ci, W = g->Wi, H = g->Hi;
    const uint64_t dims[4] = {C, W, H, (uint64_t)g->N};
    const uint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    const uint32_t box[4] = {TC_BK, (uint32_t)P.tw, (uint32_t)P.hh, (uint32_t)P.hn};
    rc = make_bf16_map(&mx_hi, x_hi, 4, dims, strides, box);
    if (rc) return rc;
    rc = make_bf16_map(&mx_lo, passes == 3 ? x_lo : x_hi, 4, dims, strides, box);
    if (rc) return rc;
  }

  const size_t stage_bytes = 2 * 128 * 128 + 2 * (size_t)P.n_half_pad * 128;
  const size_t smem = (size_t)P.stages * stage_bytes + 1024 /*alignment*/ + 256 /*barriers*/ + 2048 /*column tables*/;
  GLIS_REQUIRE(P.tmem_cols <= 512 && smem <= 227 * 1024, GLIS_E_UNSUPPORTED, "glis_conv_forward_bf16: pair tile does not fit");
  {
    cudaError_t e = ensure_max_dynamic_smem(reinterpret_cast<const void*>(tc_conv_pair_kernel), 227 * 1024);
    GLIS_REQUIRE(e == cudaSuccess, GLIS_E_CUDA, "cudaFuncSetAttribute(tc_conv_pair_kernel): %s", cudaGetErrorString(e));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1] = pdl_attr();
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() >= 2 ? 2 : 1;
  static int max_pairs = 0;
  if (!max_pairs) {
    int nc = 0;
    cfg.gridDim = dim3(pair_num_sms() / 2 * 2);
    if (cudaOccupancyMaxActiveClusters(&nc, tc_conv_pair_kernel, &cfg) != cudaSuccess || nc <= 0) nc = pair_num_sms() / 2;
    (void)cudaGetLastError();
    max_pairs = nc;
  }
  const int n_pairs = P.n_groups < max_pairs ? P.n_groups : max_pairs;
  cfg.gridDim = dim3(2 * n_pairs);
  cudaError_t le = cudaLaunchKernelEx(&cfg, tc_conv_pair_kernel, mw_hi, mw_lo, mx_hi, mx_lo, P);
  GLIS_REQUIRE(le == cudaSuccess, GLIS_E_CUDA, "glis_conv_forward_bf16(pair): launch failed: %s", cudaGetErrorString(le));
  GLIS_CHECK_LAUNCH("glis_conv_forward_bf16(pair)");
  return GLIS_OK;
}

}  // namespace glis

// tcgen05 weight gradient of a (transposed) convolution:
//
//   G[a][b][tap] += sum_{pix} S[pix][a] * B[pix*stride - pad + tap][b]
//
// S = the tensor on the coarse grid (dy of a conv layer / x of a transposed layer), B = the
// tensor on the fine grid.  Per CTA: one tap, 128 a-channels (TMEM lanes) x NB <= 128
// b-channels (TMEM columns); the contraction runs over pixels, so BOTH operands are
// "MN-major": a shared-memory row is one pixel's 64 contiguous channels (exactly what an
// NHWC TMA box delivers), 8-pixel groups are 1024 B apart, 64-channel groups one tile apart.
// The gather of B for the tap is again just a shifted box of the stride-parity view.
// K is split across blockIdx.x; partial sums are added atomically into G (master layout).
#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace glis {

using namespace sm100;

int make_bf16_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
                  const uint32_t* box);

constexpr int WG_THREADS = 192;
constexpr int WG_MAX_STAGES = 4;
constexpr int WG_MAX_TPC = 4;      // taps per CTA (N-groups of one accumulator)

struct TcWgradParams {
  glis_geom_t g;
  int tw, th, tn;       // pixel tile on the coarse grid (full rows)
  int rows;             // tw*th*tn  (valid smem rows per 64-channel group)
  int kp;               // rows rounded up to 16 (smem rows per group; tail rows are zero)
  int nb;               // b-channels per tap per CTA: 64 or 128
  int n_btiles;         // ceil(Cb / nb)
  int tpc;              // taps per CTA: their B tiles sit side by side as N-groups, so ONE MMA of
                        // N = tpc*nb columns covers them all and reads the S tile once
  int passes, stages;
  int tiles_h, tiles_total, tiles_per_split;
  float* G;
  long long slab_stride;   // > 0: deterministic form — K split i STORES its partial sums at G + i * slab_stride (one slab per
                           // split, summed in a fixed order by glis_wn_project_slabs) instead of adding them atomically
  int debug;         // GLIS_WG_DEBUG bits (profiling experiments only): 1 = no stores, 2 = no MMA, 4 = no B loads, 8 = no S loads
};

__global__ void __launch_bounds__(WG_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_s_hi, const __grid_constant__ CUtensorMap map_s_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                const TcWgradParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const glis_geom_t& g = P.g;
  const int Ca = g.Co, Cb = g.Ci, T = g.KH * g.KW;
  const int tap0 = blockIdx.z * P.tpc;
  const int a_tile = blockIdx.y / P.n_btiles, b_tile = blockIdx.y - a_tile * P.n_btiles;
  const int a0 = a_tile * 128, b0 = b_tile * P.nb;
  const int t_beg = blockIdx.x * P.tiles_per_split;
  const int t_end = min(P.tiles_total, t_beg + P.tiles_per_split);
  const int ksteps = t_end - t_beg;

  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t grp_bytes = (uint32_t)P.kp * 128;   // one 64-channel group of one plane
  const int a_groups = 2, b_groups = P.tpc * (P.nb / 64);
  const int n_cols = P.tpc * P.nb;
  const uint32_t plane_a = a_groups * grp_bytes, plane_b = b_groups * grp_bytes;
  const uint32_t stage_bytes = 2 * (plane_a + plane_b);   // [A_hi][A_lo][B_hi][B_lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (size_t)P.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WG_MAX_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * WG_MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WG_MAX_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = n_cols <= 64 ? 64 : (n_cols <= 128 ? 128 : 256);
  pdl_launch_dependents();   // the next kernel's prologue may overlap this one's tail (common.cuh)

  // zero the tail rows (never written by TMA) of every group so they contribute nothing
  if (P.kp > P.rows) {
    const int tail16 = (P.kp - P.rows) * 8;  // 16-byte chunks per group
    const int groups_per_stage = 2 * (a_groups + b_groups);
    for (int i = threadIdx.x; i < P.stages * groups_per_stage * tail16; i += WG_THREADS) {
      const int gi = i / tail16, c = i - gi * tail16;
      uint4* p = reinterpret_cast<uint4*>(base + (size_t)gi * grp_bytes + (size_t)P.rows * 128) + c;
      *p = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_s_hi); tma_prefetch_desc(&map_b_hi);
    if (P.passes == 3) { tma_prefetch_desc(&map_s_lo); tma_prefetch_desc(&map_b_lo); }
    for (int s = 0; s < P.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above touched shared memory, TMEM and kernel parameters only

  if (ksteps > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const uint32_t tx_bytes = (P.passes == 3 ? 2u : 1u) *
                                  (uint32_t)(((TC_DEBUG(P) & 8) ? 0 : a_groups) + ((TC_DEBUG(P) & 4) ? 0 : b_groups)) *
                                  (uint32_t)P.rows * 128u;
        const int gpt = P.nb / 64;   // 64-channel groups per tap
        // tap -> (parity, shift) of the fine-grid gather, decoded ONCE: the integer divisions must not
        // sit in the per-stage issue loop of this single thread
        int bc0[WG_MAX_TPC], bfx[WG_MAX_TPC], bpy[WG_MAX_TPC], bfy[WG_MAX_TPC];
#pragma unroll
        for (int tl = 0; tl < WG_MAX_TPC; ++tl) {
          const int tap = tap0 + (tl < P.tpc ? tl : 0), kh = tap / g.KW, kw = tap - kh * g.KW;
          const int ey = kh * g.dil_h - g.pad_h, ex = kw * g.dil_w - g.pad_w;
          const int pary = ((ey % g.stride_h) + g.stride_h) % g.stride_h;
          const int parx = ((ex % g.stride_w) + g.stride_w) % g.stride_w;
          bc0[tl] = parx * Cb + b0; bpy[tl] = pary;
          bfy[tl] = (ey - pary) / g.stride_h; bfx[tl] = (ex - parx) / g.stride_w;
        }
        const int ntl = (TC_DEBUG(P) & 4) ? 0 : P.tpc, nag = (TC_DEBUG(P) & 8) ? 0 : a_groups;
        const int npl = P.passes == 3 ? 2 : 1;
        int s = 0; uint32_t parity = 0;
        int tile_h = t_beg % P.tiles_h, tile_n = t_beg / P.tiles_h;
        for (int t = t_beg; t < t_end; ++t) {
          const int y0 = tile_h * P.th, n0 = tile_n * P.tn;
          if (++tile_h == P.tiles_h) { tile_h = 0; ++tile_n; }
          mbar_wait(&empty_bar[s], parity ^ 1);
          uint8_t* st = base + (size_t)s * stage_bytes;
          mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
          for (int pl = 0; pl < npl; ++pl) {
            const CUtensorMap* ms = pl ? &map_s_lo : &map_s_hi;
            const CUtensorMap* mb = pl ? &map_b_lo : &map_b_hi;
            uint8_t* sa = st + pl * plane_a;
            uint8_t* sb = st + 2 * plane_a + pl * plane_b;
            for (int gi = 0; gi < nag; ++gi)
              tma_load_4d(sa + gi * grp_bytes, ms, &full_bar[s], a0 + gi * 64, 0, y0, n0);
#pragma unroll
            for (int tl = 0; tl < WG_MAX_TPC; ++tl) {
              if (tl < ntl) {
                for (int gi = 0; gi < gpt; ++gi)
                  tma_load_5d(sb + (tl * gpt + gi) * grp_bytes, mb, &full_bar[s], bc0[tl] + gi * 64, bfx[tl], bpy[tl],
                              y0 + bfy[tl], n0);
              }
            }
          }
          if (++s == P.stages) { s = 0; parity ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, n_cols, 1, 1);  // both operands MN-major
        const uint64_t desc0 = umma_smem_desc(smem_u32(base), grp_bytes, 1024);  // stage 0, S_hi
        int s = 0; uint32_t parity = 0, accumulate = 0;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&full_bar[s], parity);
          tc_fence_after_sync();
          const uint64_t dah0 = desc0 + (uint64_t)(((uint32_t)s * stage_bytes) >> 4);
          const uint64_t dal0 = dah0 + (plane_a >> 4);
          const uint64_t dbh0 = dah0 + ((2 * plane_a) >> 4);
          const uint64_t dbl0 = dbh0 + (plane_b >> 4);
          const int nk16 = P.kp / 16;
          if (TC_DEBUG(P) & 2) {
          } else if (P.passes == 3) {
            for (int k16 = 0; k16 < nk16; ++k16) {  // 16 pixel rows of 128 B = 2048 B = 128 descriptor units
              umma_bf16(tmem_base, dah0 + 128 * k16, dbl0 + 128 * k16, idesc, accumulate);
              umma_bf16(tmem_base, dal0 + 128 * k16, dbh0 + 128 * k16, idesc, 1);
              umma_bf16(tmem_base, dah0 + 128 * k16, dbh0 + 128 * k16, idesc, 1);
              accumulate = 1;
            }
          } else {
            for (int k16 = 0; k16 < nk16; ++k16) {
              umma_bf16(tmem_base, dah0 + 128 * k16, dbh0 + 128 * k16, idesc, accumulate);
              accumulate = 1;
            }
          }
          umma_commit(&empty_bar[s]);
          if (++s == P.stages) { s = 0; parity ^= 1; }
        }
        umma_commit(tmem_full_bar);
      }
    } else {
      const int q = warp & 3;
      const int a = a0 + q * 32 + lane;
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after_sync();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16);
      const bool st_ok = a < Ca && !(TC_DEBUG(P) & 1);
      float* row = P.G + (size_t)blockIdx.x * (size_t)P.slab_stride + (size_t)a * Cb * T + tap0;
      const bool slabs = P.slab_stride > 0;
      if (P.tpc == 4) {
        // columns = [tap][b]: gather the 4 taps of 16 b-channels, then ONE 16-byte reduction per (a, b)
        // (taps are the contiguous axis of the master layout; tap0 is a multiple of 4)
        for (int bb = 0; bb < P.nb; bb += 16) {
          uint32_t v[4][16];
#pragma unroll
          for (int tl = 0; tl < 4; ++tl) tmem_ld_32x16(tbase + (uint32_t)(tl * P.nb + bb), v[tl]);
          tmem_ld_wait();
          if (st_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int b = b0 + bb + j;
              if (b < Cb) {
                if (slabs)
                  *reinterpret_cast<float4*>(row + (size_t)b * T) = make_float4(__uint_as_float(v[0][j]), __uint_as_float(v[1][j]),
                                                                                __uint_as_float(v[2][j]), __uint_as_float(v[3][j]));
                else
                  red_add_v4(row + (size_t)b * T, __uint_as_float(v[0][j]), __uint_as_float(v[1][j]),
                             __uint_as_float(v[2][j]), __uint_as_float(v[3][j]));
              }
            }
          }
        }
      } else if (P.tpc == 2) {
        for (int bb = 0; bb < P.nb; bb += 32) {
          uint32_t v0[32], v1[32];
          tmem_ld_32x32(tbase + (uint32_t)bb, v0);
          tmem_ld_32x32(tbase + (uint32_t)(P.nb + bb), v1);
          tmem_ld_wait();
          if (st_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int b = b0 + bb + j;
              if (b < Cb) {
                if (slabs) *reinterpret_cast<float2*>(row + (size_t)b * T) = make_float2(__uint_as_float(v0[j]), __uint_as_float(v1[j]));
                else red_add_v2(row + (size_t)b * T, __uint_as_float(v0[j]), __uint_as_float(v1[j]));
              }
            }
          }
        }
      } else if (T == 1 && (Cb & 3) == 0) {
        // 1x1: consecutive columns are consecutive addresses of a master row (16-byte aligned: Cb % 4 == 0)
        for (int cb = 0; cb < n_cols; cb += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tbase + (uint32_t)cb, v);
          tmem_ld_wait();
          if (st_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const int b = b0 + cb + j;
              if (b < Cb) {
                if (slabs)
                  *reinterpret_cast<float4*>(row + b) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                    __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                else
                  red_add_v4(row + b, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                             __uint_as_float(v[j + 3]));
              }
            }
          }
        }
      } else {
        for (int cb = 0; cb < n_cols; cb += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tbase + (uint32_t)cb, v);
          tmem_ld_wait();
          const int tl = cb / P.nb, bb = b0 + (cb - tl * P.nb);   // a 32-column chunk never straddles a tap
          if (st_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int b = bb + j;
              if (b < Cb) {
                if (slabs) row[(size_t)b * T + tl] = __uint_as_float(v[j]);
                else atomicAdd(row + (size_t)b * T + tl, __uint_as_float(v[j]));
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

int tc_wgrad_supported(const glis_geom_t* g) {
  if (g->relation != GLIS_CONV) return 0;
  if (g->Co % 8 != 0 || g->Ci % 8 != 0 || g->Ci < 32 || g->Co < 64) return 0;   // ragged 64-channel groups: TMA zero fill
  if (g->Hi % g->stride_h != 0 || g->Wi % g->stride_w != 0) return 0;
  if (g->Wo > 64) return 0;  // one coarse row per K tile at least
  return 1;
}

// Tile shape, taps per CTA, K split and stage count.  Returns the number of K splits (grid.x) or a negative error.
static int wg_plan(const glis_geom_t* g, int passes, TcWgradParams& P, int& n_atiles_out) {
  P.g = *g;
  const int KMAX = 64;
  P.tw = g->Wo;
  if (g->Wo * g->Ho <= KMAX) {
    P.th = g->Ho;
    P.tn = KMAX / (g->Wo * g->Ho);
    if (P.tn > g->N) P.tn = g->N;
  } else {
    P.tn = 1;
    int best = 1; double best_eff = -1;
    for (int th = 1; th <= KMAX / g->Wo && th <= g->Ho; ++th) {
      const int kp = (g->Wo * th + 15) / 16 * 16;
      const double eff = (double)g->Ho * g->Wo / ((double)((g->Ho + th - 1) / th) * (kp + 8));
      if (eff > best_eff) { best_eff = eff; best = th; }
    }
    P.th = best;
  }
  P.rows = P.tw * P.th * P.tn;
  P.kp = (P.rows + 15) / 16 * 16;
  P.nb = g->Ci > 64 ? 128 : 64;
  P.n_btiles = (g->Ci + P.nb - 1) / P.nb;
  const int T_all = g->KH * g->KW;
  int tpc_cfg;
  {
    const char* e = getenv("GLIS_WG_COLS");   // tuning knob: accumulator columns per CTA (64..256)
    tpc_cfg = e ? atoi(e) : 256;
    if (tpc_cfg < 64 || tpc_cfg > 256) tpc_cfg = 256;
  }
  P.tpc = tpc_cfg / P.nb;                   // e.g. 2 taps of 64 channels, or 1 tap of 128
  if (P.tpc < 1) P.tpc = 1;
  if (P.tpc == 3) P.tpc = 2;
  if (P.tpc > WG_MAX_TPC) P.tpc = WG_MAX_TPC;
  while (P.tpc > 1 && T_all % P.tpc != 0) P.tpc /= 2;
  const int n_atiles = (g->Co + 127) / 128;
  P.passes = passes;
  P.tiles_h = (g->Ho + P.th - 1) / P.th;
  P.tiles_total = P.tiles_h * ((g->N + P.tn - 1) / P.tn);
  const int T = g->KH * g->KW;
  const int ctas = n_atiles * P.n_btiles * (T / P.tpc);
  // Split K so that the CTAs fill whole waves of the machine (one CTA per SM: the stages take most of
  // the shared memory): cost = waves x (K tiles per CTA + the fixed prologue / reduction epilogue,
  // worth about EPI tiles); fewer splits also mean fewer reductions into G.
  const int num_sms = plan_sms();
  int epi = 10;
  {
    const char* e = getenv("GLIS_WG_EPI");
    if (e) epi = atoi(e);
  }
  int max_splits = (P.tiles_total + 3) / 4;  // at least 4 K-tiles per CTA
  if (max_splits < 1) max_splits = 1;
  if (max_splits > 4 * num_sms) max_splits = 4 * num_sms;
  int splits = 1;
  long best_cost = -1;
  for (int sp = 1; sp <= max_splits; ++sp) {
    const int tps = (P.tiles_total + sp - 1) / sp;
    const int sp_eff = (P.tiles_total + tps - 1) / tps;
    if (sp_eff != sp) continue;
    const long waves = ((long)ctas * sp + num_sms - 1) / num_sms;
    const long cost = waves * (tps + epi);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; splits = sp; }
  }
  P.tiles_per_split = (P.tiles_total + splits - 1) / splits;
  splits = (P.tiles_total + P.tiles_per_split - 1) / P.tiles_per_split;
  const size_t stage_bytes = 2 * (size_t)(2 + P.tpc * (P.nb / 64)) * P.kp * 128;
  int stages = (int)((220 * 1024) / stage_bytes);
  if (stages > WG_MAX_STAGES) stages = WG_MAX_STAGES;
  GLIS_REQUIRE(stages >= 2, GLIS_E_UNSUPPORTED, "glis_conv_wgrad_bf16: tile does not fit shared memory");
  P.stages = stages;
  n_atiles_out = n_atiles;
  return splits;
}

// Number of K splits a launch of this geometry uses = slabs glis_conv_wgrad_bf16_slabs needs (0 if unsupported).
int tc_wgrad_splits(const glis_geom_t* g) {
  if (!tc_wgrad_supported(g)) return 0;
  TcWgradParams P;
  int n_atiles;
  const int splits = wg_plan(g, 3, P, n_atiles);
  return splits > 0 ? splits : 0;
}

int tc_wgrad(const glis_geom_t* g, const __nv_bfloat16* s_hi, const __nv_bfloat16* s_lo, const __nv_bfloat16* b_hi,
             const __nv_bfloat16* b_lo, float* G, int n_slabs, int precision, cudaStream_t st) {
  GLIS_REQUIRE(tc_wgrad_supported(g), GLIS_E_UNSUPPORTED, "glis_conv_wgrad_bf16: geometry not tileable for tcgen05");
  const int passes = precision == GLIS_PREC_BF16X3 ? 3 : 1;
  GLIS_REQUIRE(s_hi && b_hi && (passes == 1 || (s_lo && b_lo)), GLIS_E_BADARG,
               "glis_conv_wgrad_bf16: missing hi/lo operand planes");
  TcWgradParams P;
  int n_atiles;
  int splits = wg_plan(g, passes, P, n_atiles);
  if (splits < 0) return splits;
  const int T = g->KH * g->KW;
  P.passes = passes;
  P.G = G;
  P.slab_stride = 0;
  if (n_slabs > 0) {
    GLIS_REQUIRE(n_slabs == splits, GLIS_E_BADARG, "glis_conv_wgrad_bf16_slabs: %d slabs for %d K splits (glis_wgrad_tc_splits)",
                 n_slabs, splits);
    P.slab_stride = (long long)g->Co * g->Ci * T;
  }
  {
    const char* dbg = getenv("GLIS_WG_DEBUG");
    P.debug = dbg ? atoi(dbg) : 0;
  }
  const size_t stage_bytes = 2 * (size_t)(2 + P.tpc * (P.nb / 64)) * P.kp * 128;
  const int stages = P.stages;

  CUtensorMap ms_hi, ms_lo, mb_hi, mb_lo;
  {
    const uint64_t C = g->Co, W = g->Wo, H = g->Ho;
    const uint64_t dims[4] = {C, W, H, (uint64_t)g->N};
    const uint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    const uint32_t box[4] = {64, (uint32_t)P.tw, (uint32_t)P.th, (uint32_t)P.tn};
    int rc = make_bf16_map(&ms_hi, s_hi, 4, dims, strides, box);
    if (rc) return rc;
    rc = make_bf16_map(&ms_lo, passes == 3 ? s_lo : s_hi, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    const uint64_t C = g->Ci, W = g->Wi, H = g->Hi, sw = g->stride_w, sh = g->stride_h;
    const uint64_t dims[5] = {sw * C, W / sw, sh, H / sh, (uint64_t)g->N};
    const uint64_t strides[4] = {sw * C * 2, W * C * 2, sh * W * C * 2, H * W * C * 2};
    const uint32_t box[5] = {64, (uint32_t)P.tw, 1, (uint32_t)P.th, (uint32_t)P.tn};
    int rc = make_bf16_map(&mb_hi, b_hi, 5, dims, strides, box);
    if (rc) return rc;
    rc = make_bf16_map(&mb_lo, passes == 3 ? b_lo : b_hi, 5, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 256;
  {
    cudaError_t e = ensure_max_dynamic_smem(reinterpret_cast<const void*>(tc_wgrad_kernel), 227 * 1024);
    GLIS_REQUIRE(e == cudaSuccess, GLIS_E_CUDA, "cudaFuncSetAttribute(tc_wgrad_kernel): %s", cudaGetErrorString(e));
  }
  dim3 grid(splits, n_atiles * P.n_btiles, T / P.tpc);
  cudaError_t le = launch_pdl(tc_wgrad_kernel, grid, dim3(WG_THREADS), smem, st, ms_hi, ms_lo, mb_hi, mb_lo, P);
  GLIS_REQUIRE(le == cudaSuccess, GLIS_E_CUDA, "glis_conv_wgrad_bf16: launch failed: %s", cudaGetErrorString(le));
  GLIS_CHECK_LAUNCH("glis_conv_wgrad_bf16");
  return GLIS_OK;
}

}  // namespace glis

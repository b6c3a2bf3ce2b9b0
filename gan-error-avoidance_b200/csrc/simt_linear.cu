// Batch-sized ("skinny") fp32 GEMMs of the G-LIS step: the LIS linears, G's initial linear and
// its data gradient, the discriminator / reverser heads, and the weight gradients of all of them.
//
//   forward / data gradient :  out[M, N] = X[M, K] * Wp[K, N],   M = batch (64-128 rows)
//   weight gradient         :  G[n][j]  += sum_m dy[m][n] * x[m][j]
//
// The work is tiny (a few MFLOP) and bound by latency and by reading the weights once, so the
// kernels are organised to put MANY short blocks on the machine:
//   * forward: a block owns a 64 x 32 output tile and ONE 64-deep K chunk; the K chunks of a tile
//     form a thread-block cluster (up to 8 blocks along z) whose partial tiles meet through
//     distributed shared memory — block r of the cluster finishes rows [r*64/cs, (r+1)*64/cs) and
//     applies the epilogue (bias, TPReLU, pre-activation, bf16 planes).  Contractions deeper than
//     one cluster (K > 512: the 12800-deep ones) add the cluster results atomically into a
//     zero-filled output (bias-only epilogues).
//   * weight gradient: a block owns 16 output features x 256 input features and walks the batch
//     in 64-row chunks staged in shared memory; every output has one owner, no atomics.
#include <cooperative_groups.h>

#include <stdlib.h>

#include "common.cuh"
#include "sm100.cuh"

namespace cg = cooperative_groups;

namespace glis {

constexpr int LC_TM = 64, LC_KC = 64, LC_NT = 256;
constexpr int LC_LD = LC_TM + 4;   // padded row of the k-major X tile (keeps float4 alignment)
constexpr int LC_MAXCS = 8;

// CPT = output columns per thread (tile = 64 rows x 32*CPT columns): 1 for narrow outputs, where the
// blocks are spread over K instead; 2 when N alone provides enough blocks (halves the re-reads of X).
// `chunks` consecutive 64-deep K chunks per block, the next chunk's global loads in flight (in
// registers) while the current one is multiplied out of shared memory.
template <int CPT>
__global__ void __launch_bounds__(LC_NT)
linear_cluster_kernel(const float* __restrict__ X, const float* __restrict__ Wp, int M, int N, int K,
                      const glis_epilogue_t ep, float* __restrict__ out, int kgroups, int chunks) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  constexpr int TN = 32 * CPT;
  __shared__ __align__(16) float Xs[LC_KC * LC_LD];   // [k][row]
  __shared__ __align__(16) float Ws[LC_KC * TN];      // [k][col]; afterwards this block's partial tile [row][col]
  static_assert(LC_KC == LC_TM, "the partial tile reuses the weight tile");
  float* part = Ws;
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int m0 = blockIdx.y * LC_TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x;
  const int c = tid & 31, rg = tid >> 5;   // column (and column + 32) of the tile, group of 8 rows (= warp)

  const bool wvec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(Wp) & 15) == 0);
  const bool xvec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  constexpr int WV = LC_KC * (TN / 4) / LC_NT;        // float4 per thread
  constexpr int XV = LC_TM * (LC_KC / 4) / LC_NT;
  float4 wv[WV], xv[XV];

  auto load_chunk = [&](int k0) {
    const int kc = max(0, min(LC_KC, K - k0));
#pragma unroll
    for (int u = 0; u < WV; ++u) {
      const int i = tid + u * LC_NT;
      const int k = i / (TN / 4), nn = n0 + 4 * (i % (TN / 4));
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < kc && nn < N) {
        const float* src = Wp + (size_t)(k0 + k) * N + nn;
        if (wvec && nn + 3 < N) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          v.x = __ldg(src);
          if (nn + 1 < N) v.y = __ldg(src + 1);
          if (nn + 2 < N) v.z = __ldg(src + 2);
          if (nn + 3 < N) v.w = __ldg(src + 3);
        }
      }
      wv[u] = v;
    }
#pragma unroll
    for (int u = 0; u < XV; ++u) {   // consecutive lanes take consecutive ROWS, each reads 16 bytes along k
      const int i = tid + u * LC_NT;
      const int r = i % LC_TM, k = 4 * (i / LC_TM);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M && k < kc) {
        const float* src = X + (size_t)(m0 + r) * K + k0 + k;
        if (xvec && k + 3 < kc) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          v.x = __ldg(src);
          if (k + 1 < kc) v.y = __ldg(src + 1);
          if (k + 2 < kc) v.z = __ldg(src + 2);
          if (k + 3 < kc) v.w = __ldg(src + 3);
        }
      }
      xv[u] = v;
    }
  };

  float acc[CPT][8];
#pragma unroll
  for (int q = 0; q < CPT; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[q][i] = 0.f;

  const int kbase = blockIdx.z * chunks * LC_KC;
  load_chunk(kbase);
  for (int ch = 0; ch < chunks; ++ch) {
    if (ch > 0) __syncthreads();   // the previous chunk has been consumed
#pragma unroll
    for (int u = 0; u < XV; ++u) {   // transposed, conflict-free shared stores
      const int i = tid + u * LC_NT;
      const int r = i % LC_TM, k = 4 * (i / LC_TM);
      Xs[(k + 0) * LC_LD + r] = xv[u].x; Xs[(k + 1) * LC_LD + r] = xv[u].y;
      Xs[(k + 2) * LC_LD + r] = xv[u].z; Xs[(k + 3) * LC_LD + r] = xv[u].w;
    }
#pragma unroll
    for (int u = 0; u < WV; ++u) reinterpret_cast<float4*>(Ws)[tid + u * LC_NT] = wv[u];
    __syncthreads();
    if (ch + 1 < chunks) load_chunk(kbase + (ch + 1) * LC_KC);   // in flight during the multiply
#pragma unroll 16
    for (int j = 0; j < LC_KC; ++j) {
      const float4* xr = reinterpret_cast<const float4*>(Xs + j * LC_LD + rg * 8);   // warp-wide broadcast
      const float4 x0 = xr[0], x1 = xr[1];
#pragma unroll
      for (int q = 0; q < CPT; ++q) {
        const float w = Ws[j * TN + c + 32 * q];
        acc[q][0] = fmaf(x0.x, w, acc[q][0]); acc[q][1] = fmaf(x0.y, w, acc[q][1]);
        acc[q][2] = fmaf(x0.z, w, acc[q][2]); acc[q][3] = fmaf(x0.w, w, acc[q][3]);
        acc[q][4] = fmaf(x1.x, w, acc[q][4]); acc[q][5] = fmaf(x1.y, w, acc[q][5]);
        acc[q][6] = fmaf(x1.z, w, acc[q][6]); acc[q][7] = fmaf(x1.w, w, acc[q][7]);
      }
    }
  }
  __syncthreads();   // everyone is done with the weight tile
#pragma unroll
  for (int q = 0; q < CPT; ++q)
#pragma unroll
    for (int i = 0; i < 8; ++i) part[(rg * 8 + i) * TN + c + 32 * q] = acc[q][i];
  if (cs > 1) cluster.sync(); else __syncthreads();

  // ---- block `rank` finishes its share of the rows: sum the cluster's partial tiles
  const int rows_per = LC_TM / cs;
  const float* remote[LC_MAXCS];
#pragma unroll
  for (int q = 0; q < LC_MAXCS; ++q) remote[q] = q < cs ? cluster.map_shared_rank(part, q) : part;
  for (int i = tid; i < rows_per * TN; i += LC_NT) {
    const int r = rank * rows_per + i / TN, cc = i % TN;
    const int m = m0 + r, nn = n0 + cc;
    float y = 0.f;
#pragma unroll
    for (int q = 0; q < LC_MAXCS; ++q) if (q < cs) y += remote[q][r * TN + cc];
    if (m >= M || nn >= N) continue;
    const size_t idx = (size_t)m * N + nn;
    if (kgroups > 1) {
      if (ep.bias && blockIdx.z < (unsigned)cs) y += __ldg(ep.bias + nn);
      atomicAdd(out + idx, y);
      continue;
    }
    if (ep.bias) y += __ldg(ep.bias + nn);
    if (ep.preact) ep.preact[idx] = y;
    float o = y;
    if (ep.act == GLIS_ACT_TPRELU) {
      const int ca = ep.act_channels > 0 ? nn % ep.act_channels : nn;
      const float b = __ldg(ep.act_b + ca), a = fminf(fmaxf(__ldg(ep.act_a + ca), 0.f), 1.f);
      const float t = y - b;
      o = (t > 0.f ? t : a * t) + b;
    } else if (ep.act == GLIS_ACT_SIGMOID) {
      o = 1.f / (1.f + expf(-y));
    }
    out[idx] = o;
    if (ep.out_hi) {
      __nv_bfloat16 h, l;
      sm100::split_bf16(o, h, l);
      reinterpret_cast<__nv_bfloat16*>(ep.out_hi)[idx] = h;
      if (ep.out_lo) reinterpret_cast<__nv_bfloat16*>(ep.out_lo)[idx] = l;
    }
  }
  if (cs > 1) cluster.sync();   // nobody leaves while a neighbour still reads its partial tile
}

// Does this launch reduce to a plain [M,K] x [K,N] product over the NHWC-flattened input?
// (linear layers; a "valid" conv whose kernel covers the whole input, e.g. the D / R heads)
bool is_linear_geom(const glis_geom_t* g) {
  if (g->pad_h != 0 || g->pad_w != 0 || g->dil_h != 1 || g->dil_w != 1 || g->Ho != 1 || g->Wo != 1) return false;
  if (g->relation == GLIS_CONV) return g->KH == g->Hi && g->KW == g->Wi;
  // the data gradient of a linear layer arrives as the (degenerate) transposed relation
  return g->KH == 1 && g->KW == 1 && g->Hi == 1 && g->Wi == 1 && g->stride_h == 1 && g->stride_w == 1;
}

// The data gradient of a one-channel "valid" head (D's final conv): out[n][tap][co] = in[n] * Wp[tap][0][co],
// a rank-1 product — the K = 1 case of the linear kernel with Wp read as one row of KH*KW*Co columns.
bool is_head_dgrad_geom(const glis_geom_t* g) {
  return g->relation == GLIS_TCONV && g->Ci == 1 && g->Hi == 1 && g->Wi == 1 && g->pad_h == 0 && g->pad_w == 0 &&
         g->dil_h == 1 && g->dil_w == 1 && g->Ho == g->KH && g->Wo == g->KW;
}

// ---- the two degenerate shapes of a one-channel head (D's final conv and its data gradient), as what they are: a
// matrix-vector product and an outer product.  The tiled kernel above spent 12.7 / 16.5 us on them (6.5 MB each at
// batch 128) — on the critical path of both backward passes, with the machine otherwise idle.
constexpr int HV_NT = 256;

// out[m] = sum_k X[m][k] * w[k] (+ bias): one block per row, every load of a thread in flight before the sum
__global__ void __launch_bounds__(HV_NT)
head_gemv_kernel(const float* __restrict__ X, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                 int K) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ float red[33];
  const float* row = X + (size_t)blockIdx.x * K;
  float acc = 0.f;
  if ((K & 3) == 0 && ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(w)) & 15) == 0) {
    const float4* r4 = reinterpret_cast<const float4*>(row);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const int K4 = K >> 2;
#pragma unroll 4
    for (int q = threadIdx.x; q < K4; q += HV_NT) {
      const float4 a = __ldg(r4 + q), b = __ldg(w4 + q);
      acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    }
  } else {
    for (int k = threadIdx.x; k < K; k += HV_NT) acc = fmaf(__ldg(row + k), __ldg(w + k), acc);
  }
  acc = block_sum<HV_NT>(acc, red);
  if (threadIdx.x == 0) out[blockIdx.x] = acc + (bias ? __ldg(bias) : 0.f);
}

// out[m][n] = x[m] * w[n]
__global__ void __launch_bounds__(HV_NT)
head_outer_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ out, int M, int N) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  const int64_t total = (int64_t)M * N;
  if ((N & 3) == 0 && ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(w)) & 15) == 0) {
    const int N4 = N >> 2;
    const int64_t total4 = total >> 2;
    for (int64_t q = (int64_t)blockIdx.x * HV_NT + threadIdx.x; q < total4; q += (int64_t)gridDim.x * HV_NT) {
      const int m = (int)(q / N4), n4 = (int)(q - (int64_t)m * N4);
      const float xv = __ldg(x + m);
      const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + n4);
      reinterpret_cast<float4*>(out)[q] = make_float4(xv * wv.x, xv * wv.y, xv * wv.z, xv * wv.w);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * HV_NT + threadIdx.x; i < total; i += (int64_t)gridDim.x * HV_NT) {
      const int m = (int)(i / N);
      out[i] = __ldg(x + m) * __ldg(w + (i - (int64_t)m * N));
    }
  }
}

// Returns GLIS_E_UNSUPPORTED when the generic kernel should run instead.
int simt_linear_forward(const glis_geom_t* g, const float* in, const float* wpack, const glis_epilogue_t* ep,
                        float* out, cudaStream_t st) {
  int M = g->N, N = g->Co, K = g->KH * g->KW * g->Ci;
  if (is_head_dgrad_geom(g)) { N = g->KH * g->KW * g->Co; K = 1; }
  if (M > 4096) return GLIS_E_UNSUPPORTED;
  {
    const bool plain = ep->act == GLIS_ACT_NONE && !ep->preact && !ep->out_hi;
    static int head_fast = -1;
    if (head_fast < 0) {
      const char* e = getenv("GLIS_HEAD_FAST");   // 0: the tiled kernel for the one-channel head shapes too
      head_fast = (e && atoi(e) == 0) ? 0 : 1;
    }
    if (head_fast && plain && N == 1 && K >= 256) {
      GLIS_LAUNCH(head_gemv_kernel, dim3(M), dim3(HV_NT), 0, st, in, wpack, ep->bias, out, K);
      GLIS_CHECK_LAUNCH("glis_conv_forward(fp32, head)");
      return GLIS_OK;
    }
    if (head_fast && plain && K == 1 && !ep->bias && N >= 256) {
      int blocks = (int)(((int64_t)M * N / 4 + HV_NT - 1) / HV_NT);
      if (blocks > 148 * 16) blocks = 148 * 16;
      if (blocks < 1) blocks = 1;
      GLIS_LAUNCH(head_outer_kernel, dim3(blocks), dim3(HV_NT), 0, st, in, wpack, out, M, N);
      GLIS_CHECK_LAUNCH("glis_conv_forward(fp32, head data gradient)");
      return GLIS_OK;
    }
  }
  const int nk = (K + LC_KC - 1) / LC_KC;
  const int gy = (M + LC_TM - 1) / LC_TM;
  // wide tiles once N alone gives every SM a block; then spread over K until ~3 blocks per SM
  const int cpt = ((N + 63) / 64) * gy >= 148 ? 2 : 1;
  const int gx = (N + 32 * cpt - 1) / (32 * cpt);
  const int want_z = (3 * 148 + gx * gy - 1) / (gx * gy);
  int chunks = (nk + want_z - 1) / want_z;
  if (chunks < 1) chunks = 1;
  const int gz = (nk + chunks - 1) / chunks;
  int cs = 1;
  while (cs * 2 <= LC_MAXCS && cs * 2 <= gz) cs *= 2;
  const int kgroups = (gz + cs - 1) / cs;
  if (kgroups > 1 && (ep->act != GLIS_ACT_NONE || ep->preact || ep->out_hi)) return GLIS_E_UNSUPPORTED;
  if (kgroups > 1) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)M * N, st);
    GLIS_REQUIRE(e == cudaSuccess, GLIS_E_CUDA, "glis_conv_forward(linear): memset failed: %s", cudaGetErrorString(e));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(gx, gy, kgroups * cs);
  cfg.blockDim = dim3(LC_NT);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = cs;
  attr[1] = pdl_attr();
  cfg.attrs = attr;
  cfg.numAttrs = pdl_applies(cfg.gridDim, 0) ? 2 : 1;
  cudaError_t e = cpt == 2
      ? cudaLaunchKernelEx(&cfg, linear_cluster_kernel<2>, in, wpack, M, N, K, *ep, out, kgroups, chunks)
      : cudaLaunchKernelEx(&cfg, linear_cluster_kernel<1>, in, wpack, M, N, K, *ep, out, kgroups, chunks);
  GLIS_REQUIRE(e == cudaSuccess, GLIS_E_CUDA, "glis_conv_forward(fp32, linear): %s", cudaGetErrorString(e));
  GLIS_CHECK_LAUNCH("glis_conv_forward(fp32, linear)");
  return GLIS_OK;
}

// ---------------------------------------------------------------------------- weight gradient
constexpr int LW_TJ = 256, LW_MC = 32, LW_NT = 256;

// G[(a*Cb + b)*T + tap] += sum_m small[m][a] * big[m][j],  j = tap*Cb + b (the NHWC-flattened row of
// `big`).  Block = 8*APT a x 256 j, thread = APT a x 8 consecutive j (APT = 4 when the layer is wide
// enough to fill the machine with 32-row blocks, else 2); the batch is walked in 32-row chunks staged in
// shared memory, the next chunk's loads in flight during the multiply.  One owner per output element:
// no atomics.
template <int APT>
__global__ void __launch_bounds__(LW_NT)
linear_wgrad_kernel(const float* __restrict__ small, const float* __restrict__ big, float* __restrict__ G,
                    int M, int Ca, int Cb, int T, int perm_c, int perm_p) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  __shared__ __align__(16) float xs[LW_MC * LW_TJ];   // [m][j]
  constexpr int LW_TN = 8 * APT;
  __shared__ __align__(16) float ds[LW_MC * LW_TN];   // [m][a]
  const int J = Cb * T;
  const int j0 = blockIdx.x * LW_TJ, a0 = blockIdx.y * LW_TN;
  const int tid = threadIdx.x;
  const int tk = tid & 31, tn = tid >> 5;
  const bool xvec = (J % 4 == 0) && ((reinterpret_cast<uintptr_t>(big) & 15) == 0);
  const bool dvec = (Ca % 4 == 0) && ((reinterpret_cast<uintptr_t>(small) & 15) == 0);
  constexpr int XV = LW_MC * (LW_TJ / 4) / LW_NT, DV = (LW_MC * (LW_TN / 4) + LW_NT - 1) / LW_NT;
  float4 xv[XV], dv[DV];

  auto load_chunk = [&](int mb) {
#pragma unroll
    for (int u = 0; u < XV; ++u) {
      const int i = tid + u * LW_NT;
      const int r = i / (LW_TJ / 4), j = j0 + 4 * (i % (LW_TJ / 4));
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mb + r < M && j < J) {
        const float* src = big + (size_t)(mb + r) * J + j;
        if (xvec && j + 3 < J) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          v.x = __ldg(src);
          if (j + 1 < J) v.y = __ldg(src + 1);
          if (j + 2 < J) v.z = __ldg(src + 2);
          if (j + 3 < J) v.w = __ldg(src + 3);
        }
      }
      xv[u] = v;
    }
#pragma unroll
    for (int u = 0; u < DV; ++u) {
      const int i = tid + u * LW_NT;
      const int r = i / (LW_TN / 4), a = a0 + 4 * (i % (LW_TN / 4));
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < LW_MC && mb + r < M && a < Ca) {
        const float* src = small + (size_t)(mb + r) * Ca + a;
        if (dvec && a + 3 < Ca) v = __ldg(reinterpret_cast<const float4*>(src));
        else {
          v.x = __ldg(src);
          if (a + 1 < Ca) v.y = __ldg(src + 1);
          if (a + 2 < Ca) v.z = __ldg(src + 2);
          if (a + 3 < Ca) v.w = __ldg(src + 3);
        }
      }
      dv[u] = v;
    }
  };

  float acc[APT][8];
#pragma unroll
  for (int i = 0; i < APT; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[i][q] = 0.f;

  load_chunk(0);
  for (int mb = 0; mb < M; mb += LW_MC) {
    if (mb > 0) __syncthreads();
#pragma unroll
    for (int u = 0; u < XV; ++u) reinterpret_cast<float4*>(xs)[tid + u * LW_NT] = xv[u];
#pragma unroll
    for (int u = 0; u < DV; ++u)
      if (tid + u * LW_NT < LW_MC * (LW_TN / 4)) reinterpret_cast<float4*>(ds)[tid + u * LW_NT] = dv[u];
    __syncthreads();
    if (mb + LW_MC < M) load_chunk(mb + LW_MC);
#pragma unroll 4
    for (int m = 0; m < LW_MC; ++m) {
      const float* dr = ds + m * LW_TN + tn * APT;   // warp-wide broadcast
      const float4* xr = reinterpret_cast<const float4*>(xs + m * LW_TJ + tk * 8);
      const float4 x0 = xr[0], x1 = xr[1];
      float dd[APT];
      if (APT == 8) {
        const float4 d0 = reinterpret_cast<const float4*>(dr)[0], d1 = reinterpret_cast<const float4*>(dr)[1];
        dd[0] = d0.x; dd[1] = d0.y; dd[2 % APT] = d0.z; dd[3 % APT] = d0.w;
        dd[4 % APT] = d1.x; dd[5 % APT] = d1.y; dd[6 % APT] = d1.z; dd[7 % APT] = d1.w;
      } else {
#pragma unroll
        for (int i = 0; i < APT; ++i) dd[i] = dr[i];
      }
      const float xx[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int i = 0; i < APT; ++i)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[i][q] = fmaf(dd[i], xx[q], acc[i][q]);
    }
  }
#pragma unroll
  for (int i = 0; i < APT; ++i) {
    const int a = a0 + APT * tn + i;
    if (a >= Ca) continue;
    const int arow = perm_c ? (a % perm_c) * perm_p + a / perm_c : a;   // master row of this (permuted) feature
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = j0 + 8 * tk + q;
      if (j >= J) continue;
      const int tap = j / Cb, b = j - tap * Cb;
      float* dst = G + ((size_t)arow * Cb + b) * T + tap;   // one owner per element
      *dst += acc[i][q];
    }
  }
}

// Weight gradient of a linear-shaped layer (is_linear_geom, GLIS_CONV): adds into G.
static int launch_linear_wgrad(const float* small, const float* big, float* G, int M, int Ca, int Cb, int T,
                               int perm_c, int perm_p, cudaStream_t st) {
  if (M > 4096) return GLIS_E_UNSUPPORTED;
  const int gx = (Cb * T + LW_TJ - 1) / LW_TJ;
  if (gx * ((Ca + 31) / 32) >= 296)
    GLIS_LAUNCH((linear_wgrad_kernel<4>), dim3(dim3(gx, (Ca + 31) / 32)), dim3(LW_NT), 0, (cudaStream_t)(st), small, big, G, M, Ca, Cb, T, perm_c, perm_p);
  else
    GLIS_LAUNCH((linear_wgrad_kernel<2>), dim3(dim3(gx, (Ca + 15) / 16)), dim3(LW_NT), 0, (cudaStream_t)(st), small, big, G, M, Ca, Cb, T, perm_c, perm_p);
  GLIS_CHECK_LAUNCH("glis_conv_wgrad(fp32, linear)");
  return GLIS_OK;
}

// ---------------------------------------------------------------------------- weight gradient + projection
// Weight gradient of a weight-normalised LINEAR layer with the weight-norm projection in the same
// kernel: for master row o
//     G_o = sum_m dy[m][a(o)] x[m][:],   dw_o (+)= (s_o / n_o) (G_o - w_o <G_o, w_o> / n_o^2),   ds_o (+)= <G_o, w_o> / n_o
// The batch is small (64-256 rows) and a row short (<= 1024 inputs), so a WARP owns a whole row: x is staged in
// shared memory once per block, the row's 8 (16, 32) outputs per lane stay in registers through the dot product
// with w_o, and the raw gradient never exists in memory.  Replaces two launches (linear_wgrad_kernel +
// wn_project_warp_kernel: 51 + 26 us for G's 12800 x 256 initial linear) by one (~10 us), at the END of G's
// backward, where nothing else is left to overlap with.
constexpr int LP_NT = 256;
// rows a warp carries at once: every x value read from shared memory feeds that many FMAs (bounded by registers)
template <int PER, int LP_ROWS>   // outputs per lane: Cb <= 32 * PER; rows per warp and pass
__global__ void __launch_bounds__(LP_NT, (PER <= 8 ? 3 : 1))
linear_wgrad_project_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                            const float* __restrict__ scale, const float* __restrict__ norm, float* __restrict__ dw,
                            float* __restrict__ dscale, int M, int Ca, int Cb, int perm_c, int perm_p, int accumulate,
                            int rows_per_block, int row_begin, int row_count) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  extern __shared__ __align__(16) float lp_xs[];        // [M][Cb]
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // rows are MASTER rows o in [row_begin, row_begin + row_count): dw is written contiguously; the matching column of
  // dy is a(o) = (o % P) * C + o / P under the NHWC row permutation of glis_wn_prepare_perm
  const int a_beg = row_begin + blockIdx.x * rows_per_block, a_end = min(row_begin + row_count, a_beg + rows_per_block);
  // The kernel is a chain of dependent phases (stage x -> dy -> FMA loop -> w -> store) run by a few warps per SM, so
  // what it costs is the number of memory round trips on that chain, not bytes: the first rows' dy columns and
  // w rows are requested BEFORE x is staged, and x is staged with every load of a thread in flight at once.
  const int a_first = a_beg + wid * LP_ROWS;
  float dpre[2][LP_ROWS];      // dy[m0 + lane][a_first + r] for the first two 32-row chunks of the batch
#pragma unroll
  for (int c = 0; c < 2; ++c)
#pragma unroll
    for (int r = 0; r < LP_ROWS; ++r) {
      const int o = a_first + r, m = c * 32 + lane;
      const int a = perm_c ? (o % perm_p) * perm_c + o / perm_p : o;
      dpre[c][r] = (m < M && o < a_end) ? __ldg(dy + (size_t)m * Ca + a) : 0.f;
    }
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) {      // a misaligned view: scalar staging
    for (int i = tid; i < M * Cb; i += LP_NT) lp_xs[i] = __ldg(x + i);
  } else {
    const int total4 = (M * Cb) >> 2;                     // Cb % 4 == 0 (checked by the host): 16-byte staging
    const float4* x4 = reinterpret_cast<const float4*>(x);
    float4* s4 = reinterpret_cast<float4*>(lp_xs);
    for (int i0 = tid; i0 < total4; i0 += 8 * LP_NT) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u * LP_NT < total4) v[u] = __ldg(x4 + i0 + u * LP_NT);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (i0 + u * LP_NT < total4) s4[i0 + u * LP_NT] = v[u];
    }
  }
  __syncthreads();
  for (int a0 = a_first; a0 < a_end; a0 += (LP_NT / 32) * LP_ROWS) {
    float acc[LP_ROWS][PER];
#pragma unroll
    for (int r = 0; r < LP_ROWS; ++r)
#pragma unroll
      for (int u = 0; u < PER; ++u) acc[r][u] = 0.f;
    // dy[m][a0 .. a0 + 3]: lanes fetch 32 batch rows at a time, then broadcast
    for (int m0 = 0; m0 < M; m0 += 32) {
      float dmine[LP_ROWS];
      const bool pre = a0 == a_first && m0 < 64;
#pragma unroll
      for (int r = 0; r < LP_ROWS; ++r) {
        if (pre) {
          dmine[r] = m0 == 0 ? dpre[0][r] : dpre[1][r];
        } else {
          const int o = a0 + r;
          const int a = perm_c ? (o % perm_p) * perm_c + o / perm_p : o;
          dmine[r] = (m0 + lane < M && o < a_end) ? __ldg(dy + (size_t)(m0 + lane) * Ca + a) : 0.f;
        }
      }
      const int mc = min(32, M - m0);
#pragma unroll 4
      for (int mm = 0; mm < mc; ++mm) {
        float d[LP_ROWS];
#pragma unroll
        for (int r = 0; r < LP_ROWS; ++r) d[r] = __shfl_sync(0xffffffffu, dmine[r], mm);
        const float* xr = lp_xs + (size_t)(m0 + mm) * Cb;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
          const int j = lane + 32 * u;
          const float xv = j < Cb ? xr[j] : 0.f;
#pragma unroll
          for (int r = 0; r < LP_ROWS; ++r) acc[r][u] = fmaf(d[r], xv, acc[r][u]);
        }
      }
    }
    // every row's w load in flight before the first dot product needs it
    float wv[LP_ROWS][PER], nrm[LP_ROWS], scl[LP_ROWS];
#pragma unroll
    for (int r = 0; r < LP_ROWS; ++r) {
      const int o = a0 + r;
      const bool row_ok = o < a_end;
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int j = lane + 32 * u;
        wv[r][u] = (row_ok && j < Cb) ? __ldg(w + (size_t)o * Cb + j) : 0.f;
      }
      nrm[r] = row_ok ? __ldg(norm + o) : 1.f;
      scl[r] = (row_ok && scale) ? __ldg(scale + o) : 1.f;
    }
#pragma unroll
    for (int r = 0; r < LP_ROWS; ++r) {
      float dot = 0.f;
#pragma unroll
      for (int u = 0; u < PER; ++u) dot = fmaf(acc[r][u], wv[r][u], dot);
      dot = warp_sum(dot);
      const float n = nrm[r];
      const float k1 = scl[r] / n, k2 = dot / (n * n);
#pragma unroll
      for (int u = 0; u < PER; ++u) acc[r][u] = k1 * (acc[r][u] - k2 * wv[r][u]);
      nrm[r] = dot / n;    // the scale gradient of the row
    }
    // stores (read-modify-writes when accumulating: independent of each other, one round trip for all of them)
#pragma unroll
    for (int r = 0; r < LP_ROWS; ++r) {
      const int o = a0 + r;
      if (o >= a_end) break;                                             // (uniform across the warp)
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int j = lane + 32 * u;
        if (j < Cb) {
          float* dst = dw + (size_t)o * Cb + j;
          *dst = accumulate ? *dst + acc[r][u] : acc[r][u];
        }
      }
      if (dscale && lane == 0) dscale[o] = accumulate ? dscale[o] + nrm[r] : nrm[r];
    }
  }
}

int linear_wgrad_project_supported(int M, int Ca, int Cb) {
  return M >= 1 && M <= 512 && Cb >= 4 && Cb <= 1024 && (Cb & 3) == 0 && (size_t)M * Cb * sizeof(float) <= 200 * 1024 && Ca >= 1;
}

int linear_wgrad_project(const float* dy, const float* x, const float* w, const float* scale, const float* norm,
                         float* dw, float* dscale, int M, int Ca, int Cb, int perm_c, int perm_p, int accumulate,
                         int row_begin, int row_count, cudaStream_t st) {
  GLIS_REQUIRE(linear_wgrad_project_supported(M, Ca, Cb), GLIS_E_UNSUPPORTED,
               "glis_linear_wgrad_project: batch %d x %d inputs does not fit the fused kernel", M, Cb);
  const size_t smem = (size_t)M * Cb * sizeof(float);
  // enough blocks to give every SM a few, each with >= 8 rows (one per warp) so that staging x pays
  GLIS_REQUIRE(row_begin >= 0 && row_count >= 0 && row_begin + row_count <= Ca, GLIS_E_BADARG,
               "glis_linear_wgrad_project: rows [%d, %d) of %d", row_begin, row_begin + row_count, Ca);
  if (row_count == 0) return GLIS_OK;
  // blocks resident at once: shared memory (x staged per block) and 2048 threads bound the blocks per SM
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  // rows per warp and pass: as many as the registers hold (every x value read from shared memory then feeds that
  // many FMAs) — unless the layer is so small that blocks of 8 x that many rows leave most SMs idle (the LIS linears:
  // 256 rows; the kernel is a latency chain there, and more, shorter blocks shorten it)
  const int rows_max = Cb <= 256 ? 4 : (Cb <= 512 ? 2 : 1);
  const int rows = row_count >= 148 * 8 * rows_max ? rows_max : 1;
  const int gran = 8 * rows;
  int rows_per_block = (row_count + 148 * per_sm - 1) / (148 * per_sm);
  rows_per_block = (rows_per_block + gran - 1) / gran * gran;       // 8 warps x `rows` rows per pass
  if (rows_per_block < gran) rows_per_block = gran;
  const int blocks = (row_count + rows_per_block - 1) / rows_per_block;
#define LP_LAUNCH(PER) do { if (rows == 1) LP_LAUNCH2(PER, 1); else if (rows == 2) LP_LAUNCH2(PER, 2); else LP_LAUNCH2(PER, 4); } while (0)
#define LP_LAUNCH2(PER, ROWS)                                                                                       \
  do {                                                                                                              \
    {                                                                                                               \
      cudaError_t e = ensure_max_dynamic_smem(reinterpret_cast<const void*>(linear_wgrad_project_kernel<PER, ROWS>), 200 * 1024); \
      GLIS_REQUIRE(e == cudaSuccess, GLIS_E_CUDA, "cudaFuncSetAttribute(linear_wgrad_project): %s", cudaGetErrorString(e)); \
    }                                                                                                               \
    GLIS_LAUNCH((linear_wgrad_project_kernel<PER, ROWS>), dim3(blocks), dim3(LP_NT), smem, (cudaStream_t)(st), dy, x, w, scale, norm, dw, dscale, M, Ca, Cb, perm_c, \
                                                                  perm_p, accumulate, rows_per_block, row_begin,   \
                                                                  row_count);                                      \
  } while (0)
  if (Cb <= 128) LP_LAUNCH(4);
  else if (Cb <= 256) LP_LAUNCH(8);
  else if (Cb <= 512) LP_LAUNCH(16);
  else LP_LAUNCH(32);
#undef LP_LAUNCH
#undef LP_LAUNCH2
  GLIS_CHECK_LAUNCH("glis_linear_wgrad_project");
  return GLIS_OK;
}

int simt_linear_wgrad(const glis_geom_t* g, const float* small, const float* big, float* G, cudaStream_t st) {
  return launch_linear_wgrad(small, big, G, g->N, g->Co, g->Ci, g->KH * g->KW, 0, 0, st);
}

}  // namespace glis

extern "C" int glis_linear_wgrad(const float* dy, const float* x, float* G, int M, int Ca, int Cb, int perm_c, int perm_p,
                                 void* stream) {
  using namespace glis;
  GLIS_REQUIRE(dy && x && G && M > 0 && Ca > 0 && Cb > 0, GLIS_E_BADARG, "glis_linear_wgrad: bad arguments");
  GLIS_REQUIRE((perm_c == 0 && perm_p == 0) || (perm_c > 0 && perm_p > 0 && (int64_t)perm_c * perm_p == Ca), GLIS_E_BADARG,
               "glis_linear_wgrad: bad row permutation (C=%d P=%d for %d rows)", perm_c, perm_p, Ca);
  return launch_linear_wgrad(dy, x, G, M, Ca, Cb, 1, perm_c, perm_p, (cudaStream_t)stream);
}

extern "C" int glis_linear_wgrad_project_supported(int M, int Ca, int Cb) {
  return glis::linear_wgrad_project_supported(M, Ca, Cb);
}

extern "C" int glis_linear_wgrad_project(const float* dy, const float* x, const float* w, const float* scale,
                                         const float* norm, float* dw, float* dscale, int M, int Ca, int Cb, int perm_c,
                                         int perm_p, int accumulate, int row_begin, int row_count, void* stream) {
  using namespace glis;
  GLIS_REQUIRE(dy && x && w && norm && dw && M > 0 && Ca > 0 && Cb > 0, GLIS_E_BADARG,
               "glis_linear_wgrad_project: bad arguments");
  GLIS_REQUIRE((perm_c == 0 && perm_p == 0) || (perm_c > 0 && perm_p > 0 && (int64_t)perm_c * perm_p == Ca), GLIS_E_BADARG,
               "glis_linear_wgrad_project: bad row permutation (C=%d P=%d for %d rows)", perm_c, perm_p, Ca);
  return linear_wgrad_project(dy, x, w, scale, norm, dw, dscale, M, Ca, Cb, perm_c, perm_p, accumulate, row_begin,
                              row_count, (cudaStream_t)stream);
}

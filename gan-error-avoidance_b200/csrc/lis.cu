// The LIS module (common/model.py:176-192, :281-297) as ONE thread-block-cluster kernel per direction.
//
//   forward :  h = u W1^  ;  a = TPReLU(h)            ;  u' = u + a W2^
//   backward:  da = du' W2^ ;  dh = da * TPReLU'(h)    ;  du = du' + dh W1^      (+ the TPReLU parameter sums)
//
// (W^ = the weight-normalised matrices, read as the [K][N] packs glis_wn_prepare builds.)  Both directions are
// the same chain  GEMM -> per-element op -> GEMM -> residual  on a [B x code] tile with code <= 256: a few
// MFLOP, bound by launch and load latency, which is why it was five launches (two skinny GEMMs, a residual
// add, ...) of ~7 us each.  Here a cluster of code/32 CTAs owns 16 batch rows: CTA r computes columns
// [32r, 32r + 32) of the first product from the row tile in shared memory, the activated slices meet through
// DISTRIBUTED SHARED MEMORY, every CTA assembles the full intermediate row tile and computes its 32 columns of
// the second product; intermediates never touch HBM except as the tensors backward needs (h, a / dh).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace glis {

constexpr int LIS_ROWS = 16;      // batch rows per cluster
constexpr int LIS_COLS = 32;      // output columns per CTA
constexpr int LIS_NT = 256;       // thread = (row, column pair)
constexpr int LIS_MAX_CODE = 256; // cluster of <= 8 CTAs (portable size)

struct LisParams {
  const float* x;        // [B][code]  forward: u            backward: du' (gradient w.r.t. the module output)
  const float* p1;       // [code][code] first pack  ([k][n]):  forward io of linear 1, backward oi of linear 2
  const float* p2;       // second pack:                        forward io of linear 2, backward oi of linear 1
  const float* bias1;    // forward only, may be NULL (added to h)
  const float* bias2;    // forward only, may be NULL (added to the second product)
  const float* a_raw;    // [code] TPReLU slopes (clamped to [0, 1] here)
  const float* b_t;      // [code] TPReLU translations
  const float* h;        // backward: the pre-activations forward stored
  float* mid_pre;        // forward: h out (may be NULL under no_grad)
  float* mid;            // forward: a out (may be NULL under no_grad);  backward: dh out
  float* out;            // [B][code]  forward: u'           backward: du
  float* da;             // backward: TPReLU slope gradient sums (+=), may be NULL together with db
  float* db;
  int B, code;
};

// This CTA's 32 columns of a [code][code] pack -> ws[k][32], every load of a thread in flight at once (the
// kernel is latency-bound: a k-loop that fetched its weights from L2 as it went took 4x longer).
__device__ __forceinline__ void lis_stage_packs(float* __restrict__ ws, float* __restrict__ ws2, const float* __restrict__ P1,
                                                const float* __restrict__ P2, int code, int n0, int tid) {
  // code <= 256: at most 8 quads per thread and pack, ALL of BOTH packs requested before the first one is stored (a
  // load -> store loop whose trip count the compiler does not know serialises on the load latency, and the second
  // pack used to be fetched only after the first product: one more round trip on a kernel that is nothing but latency)
  constexpr int Q = LIS_MAX_CODE * (LIS_COLS / 4) / LIS_NT;
  const int total = code * (LIS_COLS / 4);
  float4 v[Q], v2[Q];
#pragma unroll
  for (int u = 0; u < Q; ++u) {
    const int i = tid + u * LIS_NT;
    if (i < total) {
      v[u] = __ldg(reinterpret_cast<const float4*>(P1 + (size_t)(i >> 3) * code + n0) + (i & 7));
      v2[u] = __ldg(reinterpret_cast<const float4*>(P2 + (size_t)(i >> 3) * code + n0) + (i & 7));
    }
  }
#pragma unroll
  for (int u = 0; u < Q; ++u) {
    const int i = tid + u * LIS_NT;
    if (i < total) {
      *reinterpret_cast<float4*>(ws + (i >> 3) * LIS_COLS + 4 * (i & 7)) = v[u];
      *reinterpret_cast<float4*>(ws2 + (i >> 3) * LIS_COLS + 4 * (i & 7)) = v2[u];
    }
  }
}

// acc[0..1] = sum_k xs[row][k] * ws[k][2 cp .. 2 cp + 1]
__device__ __forceinline__ void lis_dot(const float* __restrict__ xs, int ld, const float* __restrict__ ws, int code,
                                        int row, int cp, float (&acc)[2]) {
  const float* xr = xs + row * ld;
  const float2* wc = reinterpret_cast<const float2*>(ws) + cp;
  float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
  for (int k = 0; k < code; ++k) {
    const float2 w = wc[k * (LIS_COLS / 2)];
    const float v = xr[k];
    a0 = fmaf(v, w.x, a0);
    a1 = fmaf(v, w.y, a1);
  }
  acc[0] = a0; acc[1] = a1;
}

template <bool BACKWARD>
__global__ void __launch_bounds__(LIS_NT)
lis_chain_kernel(const LisParams P) {
  pdl_launch_dependents();   // programmatic dependent launch (common.cuh)
  pdl_wait();
  extern __shared__ __align__(16) float lis_smem[];
  float* xs = lis_smem;                                          // [16][code + 4] input row tile, later the full intermediate
  float* ws = xs + LIS_ROWS * (LIS_MAX_CODE + 4);                // [code][32] this CTA's columns of the current pack
  float* ws2 = ws + LIS_MAX_CODE * LIS_COLS;                     // [code][32] ... and of the second pack, fetched up front
  float* slice = ws2 + LIS_MAX_CODE * LIS_COLS;                  // [16][32] this CTA's columns of the intermediate
  float* red_a = slice + LIS_ROWS * LIS_COLS;                    // [16][32] TPReLU parameter partial sums
  float* red_b = red_a + LIS_ROWS * LIS_COLS;
  cg::cluster_group cluster = cg::this_cluster();
  const int code = P.code, ld = code + 4;
  const int r = (int)cluster.block_rank();            // column block
  const int m0 = blockIdx.y * LIS_ROWS;
  const int tid = threadIdx.x;
  const int row = tid >> 4, cp = tid & 15;
  const int col0 = r * LIS_COLS + 2 * cp;
  const int m = m0 + row;
  const bool row_ok = m < P.B;

  // ---- first pack slice + input row tile (zero rows beyond the batch)
  lis_stage_packs(ws, ws2, P.p1, P.p2, code, r * LIS_COLS, tid);
  {
    constexpr int XQ = LIS_ROWS * (LIS_MAX_CODE / 4) / LIS_NT;     // quads of the row tile per thread, all in flight
    float4 xv[XQ];
#pragma unroll
    for (int u = 0; u < XQ; ++u) {
      const int i = tid + u * LIS_NT;
      xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < LIS_ROWS * (code >> 2)) {
        const int rr = i / (code >> 2), c4 = i - rr * (code >> 2);
        if (m0 + rr < P.B) xv[u] = __ldg(reinterpret_cast<const float4*>(P.x + (size_t)(m0 + rr) * code) + c4);
      }
    }
#pragma unroll
    for (int u = 0; u < XQ; ++u) {
      const int i = tid + u * LIS_NT;
      if (i < LIS_ROWS * (code >> 2)) {
        const int rr = i / (code >> 2), c4 = i - rr * (code >> 2);
        *reinterpret_cast<float4*>(xs + rr * ld + 4 * c4) = xv[u];
      }
    }
  }
  __syncthreads();

  // ---- first product + per-element op
  float acc[2];
  lis_dot(xs, ld, ws, code, row, cp, acc);
  float mid[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int c = col0 + j;
    const float a = fminf(fmaxf(__ldg(P.a_raw + c), 0.f), 1.f), bt = __ldg(P.b_t + c);
    if (!BACKWARD) {
      const float hval = acc[j] + (P.bias1 ? __ldg(P.bias1 + c) : 0.f);
      const float t = hval - bt;
      mid[j] = (t > 0.f ? t : a * t) + bt;
      if (row_ok && P.mid_pre) P.mid_pre[(size_t)m * code + c] = hval;
      red_a[row * LIS_COLS + 2 * cp + j] = 0.f;
      red_b[row * LIS_COLS + 2 * cp + j] = 0.f;
    } else {
      const float hval = row_ok ? __ldg(P.h + (size_t)m * code + c) : 1.f;
      const float t = hval - bt;
      const bool neg = !(t > 0.f) && row_ok;
      const float g = acc[j];                     // d(loss)/d(activated)
      mid[j] = neg ? a * g : g;
      red_a[row * LIS_COLS + 2 * cp + j] = neg ? g * t : 0.f;
      red_b[row * LIS_COLS + 2 * cp + j] = neg ? g : 0.f;
    }
    slice[row * LIS_COLS + 2 * cp + j] = mid[j];
    if (row_ok && P.mid) P.mid[(size_t)m * code + c] = mid[j];
  }
  __syncthreads();
  if (BACKWARD && P.da && tid < LIS_COLS) {
    // TPReLU parameter gradients: da_raw += sum g*t over the negative side (only while 0 <= a_raw <= 1: the clamp
    // passes no gradient outside), db += (1 - a) * sum g over the negative side
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int rr = 0; rr < LIS_ROWS; ++rr) { sa += red_a[rr * LIS_COLS + tid]; sb += red_b[rr * LIS_COLS + tid]; }
    const int c = r * LIS_COLS + tid;
    const float ar = __ldg(P.a_raw + c);
    if (ar >= 0.f && ar <= 1.f && sa != 0.f) atomicAdd(P.da + c, sa);
    const float a = fminf(fmaxf(ar, 0.f), 1.f);
    if (sb != 0.f) atomicAdd(P.db + c, (1.f - a) * sb);
  }

  // ---- every CTA assembles the full intermediate row tile from its neighbours' slices
  cluster.sync();
  const int cs = (int)cluster.num_blocks();
  {
    // 16 x code floats per CTA, all distributed-shared-memory loads of a thread in flight together
    constexpr int G = LIS_ROWS * LIS_MAX_CODE / LIS_NT;
    float g[G];
#pragma unroll
    for (int u = 0; u < G; ++u) {
      const int i = tid + u * LIS_NT;
      if (i < LIS_ROWS * code) {
        const int rr = i / code, k = i - rr * code;
        const int src = k / LIS_COLS;
        const float* remote = src < cs ? cluster.map_shared_rank(slice, src) : slice;
        g[u] = remote[rr * LIS_COLS + (k - src * LIS_COLS)];
      }
    }
#pragma unroll
    for (int u = 0; u < G; ++u) {
      const int i = tid + u * LIS_NT;
      if (i < LIS_ROWS * code) { const int rr = i / code, k = i - rr * code; xs[rr * ld + k] = g[u]; }
    }
  }
  cluster.sync();     // all slices read (nobody may leave or overwrite before that), xs complete

  // ---- second product + residual
  lis_dot(xs, ld, ws2, code, row, cp, acc);
  if (row_ok) {
    float2 res = __ldg(reinterpret_cast<const float2*>(P.x + (size_t)m * code + col0));
    if (!BACKWARD && P.bias2) { res.x += __ldg(P.bias2 + col0); res.y += __ldg(P.bias2 + col0 + 1); }
    *reinterpret_cast<float2*>(P.out + (size_t)m * code + col0) = make_float2(res.x + acc[0], res.y + acc[1]);
  }
}

static int lis_launch(const LisParams& P, bool backward, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(P.code / LIS_COLS, (P.B + LIS_ROWS - 1) / LIS_ROWS, 1);
  cfg.blockDim = dim3(LIS_NT);
  const size_t smem = sizeof(float) * (LIS_ROWS * (LIS_MAX_CODE + 4) + 2 * LIS_MAX_CODE * LIS_COLS + 3 * LIS_ROWS * LIS_COLS);
  {
    cudaError_t e1 = ensure_max_dynamic_smem(reinterpret_cast<const void*>(lis_chain_kernel<true>), (int)smem);
    cudaError_t e2 = ensure_max_dynamic_smem(reinterpret_cast<const void*>(lis_chain_kernel<false>), (int)smem);
    GLIS_REQUIRE(e1 == cudaSuccess && e2 == cudaSuccess, GLIS_E_CUDA, "glis_lis: cudaFuncSetAttribute failed");
  }
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = P.code / LIS_COLS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1] = pdl_attr();
  cfg.attrs = attr;
  cfg.numAttrs = pdl_applies(cfg.gridDim, smem) ? 2 : 1;
  cudaError_t e = backward ? cudaLaunchKernelEx(&cfg, lis_chain_kernel<true>, P)
                           : cudaLaunchKernelEx(&cfg, lis_chain_kernel<false>, P);
  GLIS_REQUIRE(e == cudaSuccess, GLIS_E_CUDA, "glis_lis_%s: launch failed: %s", backward ? "backward" : "forward",
               cudaGetErrorString(e));
  return GLIS_OK;
}

static int lis_check(int B, int code, const char* who) {
  GLIS_REQUIRE(B > 0 && code > 0, GLIS_E_BADARG, "%s: empty problem (B=%d code=%d)", who, B, code);
  GLIS_REQUIRE(code % LIS_COLS == 0 && code <= LIS_MAX_CODE, GLIS_E_UNSUPPORTED,
               "%s: code size %d (needs a multiple of %d, at most %d)", who, code, LIS_COLS, LIS_MAX_CODE);
  return GLIS_OK;
}

}  // namespace glis

using namespace glis;

extern "C" int glis_lis_supported(int code) { return code > 0 && code % LIS_COLS == 0 && code <= LIS_MAX_CODE; }

extern "C" int glis_lis_forward(const float* u, const float* io1, const float* bias1, const float* a_raw,
                                const float* b_t, const float* io2, const float* bias2, int B, int code, float* h,
                                float* act, float* u_out, void* stream) {
  if (int rc = lis_check(B, code, "glis_lis_forward")) return rc;
  GLIS_REQUIRE(u && io1 && io2 && a_raw && b_t && u_out, GLIS_E_BADARG, "glis_lis_forward: NULL pointer");
  LisParams P = {};
  P.x = u; P.p1 = io1; P.p2 = io2; P.bias1 = bias1; P.bias2 = bias2; P.a_raw = a_raw; P.b_t = b_t;
  P.mid_pre = h; P.mid = act; P.out = u_out; P.B = B; P.code = code;
  return lis_launch(P, false, (cudaStream_t)stream);
}

extern "C" int glis_lis_backward(const float* du_out, const float* oi2, const float* h, const float* a_raw,
                                 const float* b_t, const float* oi1, int B, int code, float* dh, float* du_in,
                                 float* da, float* db, void* stream) {
  if (int rc = lis_check(B, code, "glis_lis_backward")) return rc;
  GLIS_REQUIRE(du_out && oi2 && oi1 && h && a_raw && b_t && dh && du_in, GLIS_E_BADARG,
               "glis_lis_backward: NULL pointer");
  GLIS_REQUIRE((da == nullptr) == (db == nullptr), GLIS_E_BADARG, "glis_lis_backward: da and db come together");
  LisParams P = {};
  P.x = du_out; P.p1 = oi2; P.p2 = oi1; P.a_raw = a_raw; P.b_t = b_t; P.h = h;
  P.mid = dh; P.out = du_in; P.da = da; P.db = db; P.B = B; P.code = code;
  return lis_launch(P, true, (cudaStream_t)stream);
}

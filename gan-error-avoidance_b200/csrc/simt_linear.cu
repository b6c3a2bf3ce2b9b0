// Batch-sized ("skinny") fp32 GEMMs of the G-LIS step: the LIS linears, G's initial linear and
// its data gradient, the discriminator / reverser heads.  out[M, N] = X[M, K] * Wp[K, N] with
// M = batch (64-128 rows).  The work is tiny and latency / weight-read bound: see the kernel
// comment.  K longer than a 256-deep chunk is split across blockIdx.z (atomicAdd into a
// zero-filled output; bias-only epilogues), which is also what gives the 12800-deep
// contractions enough blocks.
#include "common.cuh"
#include "sm100.cuh"

namespace glis {

constexpr int LN_TM = 64, LN_TN = 64, LN_KC = 256, LN_KQ = 4, LN_NT = LN_TN * LN_KQ;
constexpr int LN_LD = LN_TM + 4;   // padded row of the k-major X tile (keeps float4 alignment)

// Block = 64 output columns x 4 K-slices (256 threads).  The 64 x 256 chunk of X is staged k-major
// in shared memory (all loads in flight at once); a thread owns ONE output column and one quarter
// of the chunk: per k it reads its weight W[k][n] straight from global memory (coalesced across
// the 64 columns) and the 64 row values X[.][k] as 16 broadcast float4 from shared memory — 64 FMAs
// per 17 loads.  The four K-slices then meet in shared memory and the block applies the epilogue.
__global__ void __launch_bounds__(LN_NT)
linear_fwd_kernel(const float* __restrict__ X, const float* __restrict__ Wp, int M, int N, int K,
                  const glis_epilogue_t ep, float* __restrict__ out, int ksplit) {
  extern __shared__ __align__(16) float lsm[];   // Xs[LN_KC][LN_TM] then reused as partial sums [LN_KQ][LN_TM][LN_TN]
  const int m0 = blockIdx.y * LN_TM, n0 = blockIdx.x * LN_TN;
  const int k0 = blockIdx.z * LN_KC;
  const int kc = min(LN_KC, K - k0);
  const int tid = threadIdx.x;

  // ---- stage X[m0:m0+64, k0:k0+kc] transposed (k-major): Xs[k][r]
  // consecutive lanes take consecutive ROWS (conflict-free shared stores); each reads 16 bytes along k
  const bool vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && (kc % 4 == 0);
  for (int i = tid; i < LN_TM * (LN_KC / 4); i += LN_NT) {
    const int r = i % LN_TM, g4 = i / LN_TM, k = 4 * g4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + r < M && k < kc) {
      const float* src = X + (size_t)(m0 + r) * K + k0 + k;
      if (vec) v = __ldg(reinterpret_cast<const float4*>(src));
      else {
        v.x = __ldg(src);
        if (k + 1 < kc) v.y = __ldg(src + 1);
        if (k + 2 < kc) v.z = __ldg(src + 2);
        if (k + 3 < kc) v.w = __ldg(src + 3);
      }
    }
    lsm[(k + 0) * LN_LD + r] = v.x; lsm[(k + 1) * LN_LD + r] = v.y;
    lsm[(k + 2) * LN_LD + r] = v.z; lsm[(k + 3) * LN_LD + r] = v.w;
  }
  __syncthreads();

  const int c = tid & (LN_TN - 1), kq = tid / LN_TN;
  const int n = n0 + c;
  const int kper = LN_KC / LN_KQ;
  float acc[LN_TM];
#pragma unroll
  for (int r = 0; r < LN_TM; ++r) acc[r] = 0.f;
  const int kb = kq * kper, ke = min(kc, kb + kper);
  for (int k8 = kb; k8 < ke; k8 += 8) {
    float w[8];   // eight weight loads in flight before the first FMA needs one
#pragma unroll
    for (int j = 0; j < 8; ++j) w[j] = (n < N && k8 + j < ke) ? __ldg(Wp + (size_t)(k0 + k8 + j) * N + n) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4* xr = reinterpret_cast<const float4*>(lsm + (k8 + j) * LN_LD);
#pragma unroll
      for (int q = 0; q < LN_TM / 4; ++q) {
        const float4 x = xr[q];
        acc[4 * q + 0] = fmaf(x.x, w[j], acc[4 * q + 0]); acc[4 * q + 1] = fmaf(x.y, w[j], acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(x.z, w[j], acc[4 * q + 2]); acc[4 * q + 3] = fmaf(x.w, w[j], acc[4 * q + 3]);
      }
    }
  }
  __syncthreads();   // everyone is done reading Xs: reuse it for the K-slice partial sums
  float* part = lsm;  // [kq][r][c]
#pragma unroll
  for (int r = 0; r < LN_TM; ++r) part[(kq * LN_TM + r) * LN_TN + c] = acc[r];
  __syncthreads();

  for (int i = tid; i < LN_TM * LN_TN; i += LN_NT) {
    const int r = i / LN_TN, cc = i - r * LN_TN;
    const int m = m0 + r, nn = n0 + cc;
    if (m >= M || nn >= N) continue;
    float y = 0.f;
#pragma unroll
    for (int q = 0; q < LN_KQ; ++q) y += part[(q * LN_TM + r) * LN_TN + cc];
    const size_t idx = (size_t)m * N + nn;
    if (ksplit > 1) {
      if (ep.bias && blockIdx.z == 0) y += __ldg(ep.bias + nn);
      atomicAdd(out + idx, y);
      continue;
    }
    if (ep.bias) y += __ldg(ep.bias + nn);
    if (ep.preact) ep.preact[idx] = y;
    float o = y;
    if (ep.act == GLIS_ACT_TPRELU) {
      const float b = __ldg(ep.act_b + nn), a = fminf(fmaxf(__ldg(ep.act_a + nn), 0.f), 1.f);
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
}

// Does this launch reduce to a plain [M,K] x [K,N] product over the NHWC-flattened input?
// (linear layers; a "valid" conv whose kernel covers the whole input, e.g. the D / R heads)
bool is_linear_geom(const glis_geom_t* g) {
  if (g->pad_h != 0 || g->pad_w != 0 || g->dil_h != 1 || g->dil_w != 1 || g->Ho != 1 || g->Wo != 1) return false;
  if (g->relation == GLIS_CONV) return g->KH == g->Hi && g->KW == g->Wi;
  // the data gradient of a linear layer arrives as the (degenerate) transposed relation
  return g->KH == 1 && g->KW == 1 && g->Hi == 1 && g->Wi == 1 && g->stride_h == 1 && g->stride_w == 1;
}

// Returns GLIS_E_UNSUPPORTED when the generic kernel should run instead.
int simt_linear_forward(const glis_geom_t* g, const float* in, const float* wpack, const glis_epilogue_t* ep,
                        float* out, cudaStream_t st) {
  const int M = g->N, N = g->Co, K = g->KH * g->KW * g->Ci;
  if (M > 4096) return GLIS_E_UNSUPPORTED;
  int ksplit = (K + LN_KC - 1) / LN_KC;
  if (ksplit > 1 && (ep->act != GLIS_ACT_NONE || ep->preact || ep->out_hi)) return GLIS_E_UNSUPPORTED;
  const size_t smem = sizeof(float) * (size_t)LN_KC * LN_LD;   // 68 KB; the partial sums (4*64*64 floats) fit in it too
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(linear_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    GLIS_REQUIRE(e == cudaSuccess, GLIS_E_CUDA, "cudaFuncSetAttribute(linear_fwd_kernel): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  if (ksplit > 1) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * (size_t)M * N, st);
    GLIS_REQUIRE(e == cudaSuccess, GLIS_E_CUDA, "glis_conv_forward(linear): memset failed: %s", cudaGetErrorString(e));
  }
  dim3 grid((N + LN_TN - 1) / LN_TN, (M + LN_TM - 1) / LN_TM, ksplit);
  linear_fwd_kernel<<<grid, LN_NT, smem, st>>>(in, wpack, M, N, K, *ep, out, ksplit);
  GLIS_CHECK_LAUNCH("glis_conv_forward(fp32, linear)");
  return GLIS_OK;
}

}  // namespace glis

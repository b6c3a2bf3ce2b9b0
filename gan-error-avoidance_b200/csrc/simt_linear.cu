// Batch-sized ("skinny") fp32 GEMMs of the G-LIS step: the LIS linears, G's initial linear and
// its data gradient, the discriminator / reverser heads.  out[M, N] = X[M, K] * Wp[K, N] with
// M = batch (64-128 rows).  The work is tiny and latency-bound, so a block stages one whole
// 256-deep K chunk of X (64 rows) and of its 32 weight columns in shared memory with every
// load in flight at once, synchronises once, and runs the FMAs from shared memory.  K longer
// than a chunk is split across blockIdx.z (atomicAdd into a zero-filled output; bias-only
// epilogues), which is also what gives the 12800-deep contractions enough blocks.
#include "common.cuh"
#include "sm100.cuh"

namespace glis {

constexpr int LN_TM = 64, LN_TN = 32, LN_KC = 256, LN_NT = 256;

__global__ void __launch_bounds__(LN_NT)
linear_fwd_kernel(const float* __restrict__ X, const float* __restrict__ Wp, int M, int N, int K,
                  const glis_epilogue_t ep, float* __restrict__ out, int ksplit) {
  extern __shared__ float lsm[];
  float* Xs = lsm;                          // [LN_TM][LN_KC + 1]
  float* Ws = lsm + LN_TM * (LN_KC + 1);    // [LN_KC][LN_TN]
  const int m0 = blockIdx.y * LN_TM, n0 = blockIdx.x * LN_TN;
  const int k0 = blockIdx.z * LN_KC;
  const int kc = min(LN_KC, K - k0);
  const int tid = threadIdx.x;

  // ---- stage X[m0:m0+64, k0:k0+kc] (row-major, padded) and Wp[k0:k0+kc, n0:n0+32]
  const bool x_vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  if (x_vec && kc % 4 == 0) {
    const int k4n = kc >> 2;
    for (int i = tid; i < LN_TM * k4n; i += LN_NT) {
      const int r = i / k4n, k4 = i - r * k4n;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) v = __ldg(reinterpret_cast<const float4*>(X + (size_t)(m0 + r) * K + k0) + k4);
      float* d = Xs + r * (LN_KC + 1) + k4 * 4;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  } else {
    for (int i = tid; i < LN_TM * kc; i += LN_NT) {
      const int r = i / kc, k = i - r * kc;
      Xs[r * (LN_KC + 1) + k] = (m0 + r < M) ? __ldg(X + (size_t)(m0 + r) * K + k0 + k) : 0.f;
    }
  }
  for (int i = tid; i < kc * LN_TN; i += LN_NT) {
    const int k = i / LN_TN, c = i - k * LN_TN;
    Ws[i] = (n0 + c < N) ? __ldg(Wp + (size_t)(k0 + k) * N + n0 + c) : 0.f;
  }
  __syncthreads();

  // ---- 2 rows x 4 columns per thread
  const int r2 = tid >> 3, c4 = tid & 7;
  const float* x0 = Xs + (2 * r2) * (LN_KC + 1);
  const float* x1 = x0 + (LN_KC + 1);
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 8
  for (int k = 0; k < kc; ++k) {
    const float4 w = *reinterpret_cast<const float4*>(Ws + k * LN_TN + c4 * 4);
    const float a0 = x0[k], a1 = x1[k];
    acc[0][0] = fmaf(a0, w.x, acc[0][0]); acc[0][1] = fmaf(a0, w.y, acc[0][1]);
    acc[0][2] = fmaf(a0, w.z, acc[0][2]); acc[0][3] = fmaf(a0, w.w, acc[0][3]);
    acc[1][0] = fmaf(a1, w.x, acc[1][0]); acc[1][1] = fmaf(a1, w.y, acc[1][1]);
    acc[1][2] = fmaf(a1, w.z, acc[1][2]); acc[1][3] = fmaf(a1, w.w, acc[1][3]);
  }

#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + 2 * r2 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + c4 * 4 + j;
      if (n >= N) continue;
      const size_t idx = (size_t)m * N + n;
      float y = acc[i][j];
      if (ksplit > 1) {
        if (ep.bias && blockIdx.z == 0) y += __ldg(ep.bias + n);
        atomicAdd(out + idx, y);
        continue;
      }
      if (ep.bias) y += __ldg(ep.bias + n);
      if (ep.preact) ep.preact[idx] = y;
      float o = y;
      if (ep.act == GLIS_ACT_TPRELU) {
        const float b = __ldg(ep.act_b + n), a = fminf(fmaxf(__ldg(ep.act_a + n), 0.f), 1.f);
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
  const size_t smem = sizeof(float) * (LN_TM * (LN_KC + 1) + LN_KC * LN_TN);
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

// C-ABI dispatch for the convolution-shaped entry points + error plumbing.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>
#include <utility>
#include <vector>

#include <cuda_bf16.h>

#include "common.cuh"

namespace glis {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int simt_conv_forward(const glis_geom_t* g, const float* in, const float* wpack, const glis_epilogue_t* ep,
                      float* out, cudaStream_t st);
int simt_conv_wgrad(const glis_geom_t* g, const float* small, const float* big, float* G, cudaStream_t st);

int tc_conv_supported(const glis_geom_t* g);
int tc_conv_plan_ksplit(const glis_geom_t* g);
int tc_conv_plan_describe(const glis_geom_t* g, int plain_out, int out[15]);
int tc_conv_forward(const glis_geom_t* g, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
                    const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, const glis_epilogue_t* ep, float* out_f32,
                    __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int precision, cudaStream_t st);
int tc_conv_halo_describe(const glis_geom_t* g, int plain_out, int out[20]);
int tc_conv_halo_applies(const glis_geom_t* g, int plain_out);
int tc_conv_pair_describe(const glis_geom_t* g, int plain_out, int out[16]);
int tc_pm_supported(const glis_geom_t* g);
int tc_pm_forward(const glis_geom_t* g, const __nv_bfloat16* x_hi, const __nv_bfloat16* x_lo,
                  const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, const glis_epilogue_t* ep, float* out_f32,
                  __nv_bfloat16* out_hi, __nv_bfloat16* out_lo, int precision, cudaStream_t st);
int tc_wgrad_supported(const glis_geom_t* g);
int tc_wgrad_splits(const glis_geom_t* g);
int tc_wgrad(const glis_geom_t* g, const __nv_bfloat16* s_hi, const __nv_bfloat16* s_lo, const __nv_bfloat16* b_hi,
             const __nv_bfloat16* b_lo, float* G, int n_slabs, int precision, cudaStream_t st);
int split_planes(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int64_t numel, cudaStream_t st);

}  // namespace glis

using namespace glis;

extern "C" const char* glis_last_error(void) { return g_err; }

cudaError_t glis::ensure_max_dynamic_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& d : done)
    if (d.first == kernel && d.second >= bytes) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.emplace_back(kernel, bytes);
  return e;
}

static int g_reserved_sms = -1;   // -1: not decided yet (GLIS_RESERVE_SMS, default 0)
int glis::plan_sms() {
  static int device_sms = 0;
  if (!device_sms) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    (void)cudaGetLastError();
    device_sms = n;
  }
  if (g_reserved_sms < 0) {
    const char* e = getenv("GLIS_RESERVE_SMS");
    g_reserved_sms = e ? atoi(e) : 0;
    if (g_reserved_sms < 0) g_reserved_sms = 0;
  }
  const int n = device_sms - g_reserved_sms;
  return n < 8 ? 8 : n;
}
extern "C" int glis_set_reserved_sms(int n) {
  const int prev = g_reserved_sms < 0 ? 0 : g_reserved_sms;
  g_reserved_sms = n < 0 ? 0 : n;
  return prev;
}

static int g_pdl = -1;   // -1: not decided yet (GLIS_PDL=1 turns it on; default off: measured 2.4 % slower in-step)
int glis::pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("GLIS_PDL");
    g_pdl = e ? atoi(e) : 0;
    if (g_pdl < 0 || g_pdl > 2) g_pdl = 0;
  }
  return g_pdl;
}
extern "C" int glis_set_pdl(int on) {
  const int prev = glis::pdl_enabled();
  g_pdl = on < 0 ? 0 : (on > 2 ? 2 : on);
  return prev;
}
extern "C" int glis_version(void) { return 100; }

extern "C" int glis_conv_forward(const glis_geom_t* g, const float* in, const float* wpack,
                                 const glis_epilogue_t* ep, float* out, int precision, void* stream) {
  int rc = validate_geom(g, "glis_conv_forward");
  if (rc != GLIS_OK) return rc;
  GLIS_REQUIRE(in && wpack && out, GLIS_E_BADARG, "glis_conv_forward: NULL tensor pointer");
  glis_epilogue_t none = {nullptr, GLIS_ACT_NONE, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  if (!ep) ep = &none;
  GLIS_REQUIRE(ep->act == GLIS_ACT_NONE || ep->act == GLIS_ACT_SIGMOID ||
                   (ep->act == GLIS_ACT_TPRELU && ep->act_a && ep->act_b),
               GLIS_E_BADARG, "glis_conv_forward: bad activation descriptor");
  switch (precision) {
    case GLIS_PREC_FP32:
      GLIS_REQUIRE(!(ep->out_hi && g->Co <= 4), GLIS_E_UNSUPPORTED,
                   "glis_conv_forward: bf16 planes are not produced for <= 4 output channels");
      return simt_conv_forward(g, in, wpack, ep, out, (cudaStream_t)stream);
    default:
      set_error("glis_conv_forward: precision %d needs the split-bf16 entry points", precision);
      return GLIS_E_UNSUPPORTED;
  }
}

extern "C" int glis_conv_wgrad(const glis_geom_t* g, const float* small, const float* big, float* G,
                               int precision, void* stream) {
  int rc = validate_geom(g, "glis_conv_wgrad");
  if (rc != GLIS_OK) return rc;
  GLIS_REQUIRE(g->relation == GLIS_CONV, GLIS_E_BADARG, "glis_conv_wgrad: relation must be GLIS_CONV");
  GLIS_REQUIRE(small && big && G, GLIS_E_BADARG, "glis_conv_wgrad: NULL tensor pointer");
  switch (precision) {
    case GLIS_PREC_FP32:
      return simt_conv_wgrad(g, small, big, G, (cudaStream_t)stream);
    default:
      set_error("glis_conv_wgrad: precision %d needs the split-bf16 entry points", precision);
      return GLIS_E_UNSUPPORTED;
  }
}

extern "C" int glis_conv_tc_supported(const glis_geom_t* g) {
  if (validate_geom(g, "glis_conv_tc_supported") != GLIS_OK) return 0;
  return tc_conv_supported(g);
}

extern "C" int glis_conv_tc_ksplit(const glis_geom_t* g) {
  if (validate_geom(g, "glis_conv_tc_ksplit") != GLIS_OK) return 1;
  return tc_conv_plan_ksplit(g);
}

extern "C" int glis_conv_tc_plan(const glis_geom_t* g, int plain_out, int* out15) {
  int rc = validate_geom(g, "glis_conv_tc_plan");
  if (rc != GLIS_OK) return rc;
  GLIS_REQUIRE(out15 != nullptr, GLIS_E_BADARG, "glis_conv_tc_plan: NULL output");
  GLIS_REQUIRE(tc_conv_supported(g), GLIS_E_UNSUPPORTED, "glis_conv_tc_plan: geometry not tileable for tcgen05");
  return tc_conv_plan_describe(g, plain_out, out15);
}

extern "C" int glis_conv_tc_halo_plan(const glis_geom_t* g, int plain_out, int* out20) {
  int rc = validate_geom(g, "glis_conv_tc_halo_plan");
  if (rc != GLIS_OK) return rc;
  GLIS_REQUIRE(out20 != nullptr, GLIS_E_BADARG, "glis_conv_tc_halo_plan: NULL output");
  rc = tc_conv_halo_describe(g, plain_out, out20);
  if (rc != GLIS_OK) set_error("glis_conv_tc_halo_plan: the halo kernel does not apply to this geometry");
  return rc;
}

extern "C" int glis_conv_tc_pair_plan(const glis_geom_t* g, int plain_out, int* out16) {
  int rc = validate_geom(g, "glis_conv_tc_pair_plan");
  if (rc != GLIS_OK) return rc;
  GLIS_REQUIRE(out16 != nullptr, GLIS_E_BADARG, "glis_conv_tc_pair_plan: NULL output");
  rc = tc_conv_halo_applies(g, plain_out) ? GLIS_E_UNSUPPORTED : tc_conv_pair_describe(g, plain_out, out16);
  if (rc != GLIS_OK) set_error("glis_conv_tc_pair_plan: the pair kernel does not take this launch");
  return rc;
}

extern "C" int glis_conv_forward_bf16(const glis_geom_t* g, const void* x_hi, const void* x_lo, const void* w_hi,
                                      const void* w_lo, const glis_epilogue_t* ep, float* out_f32, void* out_hi,
                                      void* out_lo, int precision, void* stream) {
  int rc = validate_geom(g, "glis_conv_forward_bf16");
  if (rc != GLIS_OK) return rc;
  GLIS_REQUIRE(precision == GLIS_PREC_BF16X3 || precision == GLIS_PREC_BF16, GLIS_E_BADARG,
               "glis_conv_forward_bf16: precision must be GLIS_PREC_BF16X3 or GLIS_PREC_BF16");
  GLIS_REQUIRE(out_f32 || out_hi, GLIS_E_BADARG, "glis_conv_forward_bf16: no output tensor");
  glis_epilogue_t none = {nullptr, GLIS_ACT_NONE, nullptr, nullptr, nullptr, nullptr, nullptr, 0};
  if (!ep) ep = &none;
  GLIS_REQUIRE(ep->act == GLIS_ACT_NONE || ep->act == GLIS_ACT_SIGMOID ||
                   (ep->act == GLIS_ACT_TPRELU && ep->act_a && ep->act_b),
               GLIS_E_BADARG, "glis_conv_forward_bf16: bad activation descriptor");
  if (tc_pm_supported(g) && (ep->act_channels == 0 || ep->act_channels == g->Co))   // image-side 1x1 products
    return tc_pm_forward(g, (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, (const __nv_bfloat16*)w_hi,
                         (const __nv_bfloat16*)w_lo, ep, out_f32, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo,
                         precision, (cudaStream_t)stream);
  return tc_conv_forward(g, (const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, (const __nv_bfloat16*)w_hi,
                         (const __nv_bfloat16*)w_lo, ep, out_f32, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo,
                         precision, (cudaStream_t)stream);
}

extern "C" int glis_split_bf16(const float* x, void* hi, void* lo, int64_t numel, void* stream) {
  GLIS_REQUIRE(x && hi && numel >= 0, GLIS_E_BADARG, "glis_split_bf16: bad arguments");
  if (numel == 0) return GLIS_OK;
  return split_planes(x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, numel, (cudaStream_t)stream);
}

extern "C" int glis_wgrad_tc_supported(const glis_geom_t* g) {
  if (validate_geom(g, "glis_wgrad_tc_supported") != GLIS_OK) return 0;
  return tc_wgrad_supported(g);
}

extern "C" int glis_conv_wgrad_bf16(const glis_geom_t* g, const void* small_hi, const void* small_lo,
                                    const void* big_hi, const void* big_lo, float* G, int precision, void* stream) {
  int rc = validate_geom(g, "glis_conv_wgrad_bf16");
  if (rc != GLIS_OK) return rc;
  GLIS_REQUIRE(precision == GLIS_PREC_BF16X3 || precision == GLIS_PREC_BF16, GLIS_E_BADARG,
               "glis_conv_wgrad_bf16: precision must be GLIS_PREC_BF16X3 or GLIS_PREC_BF16");
  GLIS_REQUIRE(G != nullptr, GLIS_E_BADARG, "glis_conv_wgrad_bf16: G is NULL");
  return tc_wgrad(g, (const __nv_bfloat16*)small_hi, (const __nv_bfloat16*)small_lo, (const __nv_bfloat16*)big_hi,
                  (const __nv_bfloat16*)big_lo, G, 0, precision, (cudaStream_t)stream);
}

extern "C" int glis_wgrad_tc_splits(const glis_geom_t* g) {
  if (validate_geom(g, "glis_wgrad_tc_splits") != GLIS_OK) return 0;
  return tc_wgrad_splits(g);
}

extern "C" int glis_conv_wgrad_bf16_slabs(const glis_geom_t* g, const void* small_hi, const void* small_lo,
                                          const void* big_hi, const void* big_lo, float* slabs, int n_slabs, int precision,
                                          void* stream) {
  int rc = validate_geom(g, "glis_conv_wgrad_bf16_slabs");
  if (rc != GLIS_OK) return rc;
  GLIS_REQUIRE(precision == GLIS_PREC_BF16X3 || precision == GLIS_PREC_BF16, GLIS_E_BADARG,
               "glis_conv_wgrad_bf16_slabs: precision must be GLIS_PREC_BF16X3 or GLIS_PREC_BF16");
  GLIS_REQUIRE(slabs != nullptr && n_slabs > 0, GLIS_E_BADARG, "glis_conv_wgrad_bf16_slabs: no slabs");
  return tc_wgrad(g, (const __nv_bfloat16*)small_hi, (const __nv_bfloat16*)small_lo, (const __nv_bfloat16*)big_hi,
                  (const __nv_bfloat16*)big_lo, slabs, n_slabs, precision, (cudaStream_t)stream);
}

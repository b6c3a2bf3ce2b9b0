// C-ABI dispatch for the convolution-shaped entry points + error plumbing.
#include <stdarg.h>

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

}  // namespace glis

using namespace glis;

extern "C" const char* glis_last_error(void) { return g_err; }
extern "C" int glis_version(void) { return 100; }

extern "C" int glis_conv_forward(const glis_geom_t* g, const float* in, const float* wpack,
                                 const glis_epilogue_t* ep, float* out, int precision, void* stream) {
  int rc = validate_geom(g, "glis_conv_forward");
  if (rc != GLIS_OK) return rc;
  GLIS_REQUIRE(in && wpack && out, GLIS_E_BADARG, "glis_conv_forward: NULL tensor pointer");
  glis_epilogue_t none = {nullptr, GLIS_ACT_NONE, nullptr, nullptr, nullptr};
  if (!ep) ep = &none;
  GLIS_REQUIRE(ep->act == GLIS_ACT_NONE || ep->act == GLIS_ACT_SIGMOID ||
                   (ep->act == GLIS_ACT_TPRELU && ep->act_a && ep->act_b),
               GLIS_E_BADARG, "glis_conv_forward: bad activation descriptor");
  switch (precision) {
    case GLIS_PREC_FP32:
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

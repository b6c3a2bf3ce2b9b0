/*
 * glis_b200.h — C ABI of the B200-native G-LIS training-step kernels.
 *
 * The reference (aleju/gan-error-avoidance) has no FFI layer of its own: every device
 * op is a stock PyTorch call made from its Python nn.Modules.  Each entry point below
 * therefore cites the reference *call site* it replaces (paths relative to the
 * upstream repository).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless a name ends in `_host`.  The library
 *     borrows them for the duration of the call: it never allocates, frees or retains
 *     device memory (workspaces are caller-provided).
 *   - Activations are fp32, NHWC ("channels_last"): element (n,y,x,c) at ((n*H+y)*W+x)*C+c.
 *     A (B,F) matrix is the H=W=1 case.
 *   - Parameters keep the reference's shapes ("master" layout): conv (Cout,Cin,KH,KW),
 *     transposed conv (Cin,Cout,KH,KW), linear (out,in), scale/bias one value per
 *     output channel, TPReLU slope/translation one value per channel.
 *   - Every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*); no
 *     host synchronisation.  Entry points are re-entrant.
 *   - Return value: 0 on success, a negative GLIS_E* code otherwise;
 *     glis_last_error() returns a thread-local description.  Nothing throws or exits.
 */
#ifndef GLIS_B200_H
#define GLIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLIS_OK 0
#define GLIS_E_BADARG (-1)      /* null pointer, negative size, inconsistent geometry */
#define GLIS_E_UNSUPPORTED (-2) /* valid request this build has no kernel for */
#define GLIS_E_CUDA (-3)        /* CUDA runtime error at enqueue */
#define GLIS_E_WORKSPACE (-4)   /* caller workspace too small */

/* gather relation between the "out" grid of a kernel and the tensor it reads */
#define GLIS_CONV 0   /* in = out*stride - pad + k*dil   (F.conv2d forward) */
#define GLIS_TCONV 1  /* in = (out + pad - k)/stride     (F.conv_transpose2d forward) */

/* epilogue activation */
#define GLIS_ACT_NONE 0
#define GLIS_ACT_TPRELU 1  /* (t>0 ? t : a*t) + b, t = y-b, a = clamp(act_a, 0, 1) */
#define GLIS_ACT_SIGMOID 2

/* arithmetic of the contraction */
#define GLIS_PREC_FP32 0    /* fp32 FFMA (exact-mode reference kernels, edge layers) */
#define GLIS_PREC_BF16X3 1  /* tcgen05 bf16 hi/lo split, 3 MMAs per k-step, ~2^-16 rel. */
#define GLIS_PREC_BF16 2    /* tcgen05 bf16, 1 MMA per k-step, ~2^-8 rel. */

/* Geometry of one (transposed) convolution.  `in`/`out` refer to the tensors of THIS
 * kernel launch (a dgrad launch swaps the layer's roles).  A linear layer is
 * KH=KW=Hi=Wi=Ho=Wo=1. */
typedef struct glis_geom {
  int32_t relation;          /* GLIS_CONV | GLIS_TCONV */
  int32_t N, Hi, Wi, Ci;     /* tensor that is read */
  int32_t Ho, Wo, Co;        /* tensor that is written */
  int32_t KH, KW, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
} glis_geom_t;

/* Fused epilogue of a forward launch: y = acc + bias[c]; out = act(y). */
typedef struct glis_epilogue {
  const float* bias;   /* per out-channel, or NULL */
  int32_t act;         /* GLIS_ACT_* */
  const float* act_a;  /* TPReLU slope, raw parameter (clamped to [0,1] by the kernel), per out-channel */
  const float* act_b;  /* TPReLU translation (per out-channel) */
  float* preact;       /* if non-NULL, y (before act) is also stored here (NHWC) */
  void* out_hi;        /* fp32 kernels only: if non-NULL, bf16 hi plane of the activated output */
  void* out_lo;        /*   ... and its lo plane (may be NULL) — feeds the next tensor-core layer */
  int32_t act_channels; /* 0: act_a / act_b hold one value per OUTPUT channel; C > 0: the TPReLU has C channels and
                         * output channel co uses act_a[co % C] (a linear layer whose output is a (C,h,w) map
                         * written in NHWC feature order, see glis_wn_prepare_perm) */
  int32_t split_slabs;  /* glis_conv_forward_bf16, plain-output launches only: when > 0 and the launch splits K
                         * (glis_conv_tc_ksplit) the partial sums of share s are STORED at out_f32 + s * N*Ho*Wo*Co
                         * (the caller provides that many slabs, >= the K split) instead of being added atomically
                         * into one zero-filled output: a deterministic split-K (summed by
                         * glis_tprelu_forward_planes_sum in a fixed order) */
} glis_epilogue_t;

const char* glis_last_error(void);
int glis_version(void);

/* Programmatic dependent launch (process-wide mode; env GLIS_PDL): 0 = off, 1 = small launches only (<= 4 blocks per SM:
 * losses, heads, LIS, TPReLU kernels), 2 = every kernel.  When on, the
 * library's stream-ordered kernels are enqueued with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel's
 * CTAs are scheduled while its predecessor in the stream drains, run their prologue and block in
 * `griddepcontrol.wait` until the predecessor has completed — stream semantics are unchanged (no memory access
 * precedes the wait), only the launch latency between dependent kernels is hidden.  No reference counterpart
 * (PyTorch launches every kernel fully serialised).  Returns the previous setting. */
int glis_set_pdl(int on);

/* Sum all-reduce of the slice [offset, offset + count) (floats) of a flat buffer that exists once per rank, over
 * NVLink peer memory: bufs[r] / flags[r] = rank r's buffer / flag array (>= 128 * world int32, zero-initialised) mapped
 * into THIS process (symmetric memory), epochs = this rank's own 128 int32 counters (zero-initialised).  One kernel
 * per rank: barrier, each rank sums its 1/world of the slice from all copies in rank order and stores the result into
 * every copy, barrier (csrc/peer_allreduce.cu).  Every rank must enqueue the same calls in the same order.  offset %
 * 4 == 0, count % (4 * world) == 0, world in {2, 4, 8}.  Replaces the ncclAllReduce of the gradient exchange
 * (no reference counterpart: the reference is single-GPU). */
int glis_peer_allreduce(void* const* bufs, void* const* flags, int rank, int world, int64_t offset, int64_t count,
                        void* epochs, int blocks, void* stream);

/* SMs the launch planners of the persistent tcgen05 kernels leave FREE (process-wide; env GLIS_RESERVE_SMS, default 0):
 * under data parallelism the communication library's kernels need SMs of their own while 148 one-CTA-per-SM compute
 * CTAs are resident.  Only the plans change (tile shapes, K splits, grid sizes), never the results.  No reference
 * counterpart.  Returns the previous setting. */
int glis_set_reserved_sms(int n);

/* ---- weight normalisation ------------------------------------------------------
 * Replaces `_WeightNormalizedConvNd.weight_norm` (common/modules/WeightNormalizedConv.py:29-38)
 * and `WeightNormalizedLinear.weight_norm` (WeightNormalizedLinear.py:30-31), plus the
 * division / scale of `norm_scale_bias` (Conv.py:40-47, Linear.py:33-37), folded into the
 * weights once per parameter update:  w_hat[o] = w[o] * scale[o] / sqrt(c*|w[o]|^2 + 1e-6).
 *
 *   w        master weights; out_axis = 0 for conv/linear (Cout,Cin,T), 1 for transposed (Cin,Cout,T)
 *   c        weight_norm_factor (1, or 1/(stride_h*stride_w) for transposed conv)
 *   norm     [Cout]  out: sqrt(c*sum w^2 + 1e-6)
 *   pack_io  [T][Cin][Cout]  out (may be NULL): w_hat laid out for a launch that reads Cin, writes Cout
 *   pack_oi  [T][Cout][Cin]  out (may be NULL): w_hat laid out for the data-gradient launch
 */
int glis_wn_prepare(const float* w, const float* scale, int out_axis, int Cout, int Cin, int T,
                    float c, float* norm, float* pack_io, float* pack_oi, void* stream);

/* Row-permuted packs for a linear layer feeding View(C, h, w) (common/model.py:206-212): with P = h*w,
 * pack row o' = p*C + c holds master row o = c*P + p, so the layer's output features come out in NHWC
 * order and the View costs nothing on the device.  perm_c = C, perm_p = P (0, 0 = no permutation);
 * requires T == 1, out_axis == 0 and Cout == C*P.  Otherwise as glis_wn_prepare / glis_wn_prepare_bf16. */
int glis_wn_prepare_perm(const float* w, const float* scale, int out_axis, int Cout, int Cin, int T, float c,
                         float* norm, float* pack_io, float* pack_oi, int perm_c, int perm_p, void* stream);
int glis_wn_prepare_bf16_perm(const float* w, const float* scale, int out_axis, int Cout, int Cin, int T, float c,
                              float* norm, void* fwd_hi, void* fwd_lo, void* bwd_hi, void* bwd_lo, int perm_c,
                              int perm_p, void* stream);
/* Everything glis_wn_prepare* / glis_wn_pack_matrix_bf16 do, for a LIST of layers in two launches (all norms, then
 * all packs): one call per network and parameter update.  Per layer: any subset of the pack pointers may be NULL;
 * need_norm = 0 reuses the norm already stored in `norm`.  mat_* / matt_* are the E [A][J] / E^T [J][A] matrix packs
 * of the image-side layers (A = mat_rows, J = T*Cin*Cout / A, in master element order). */
typedef struct glis_wn_layer {
  const float* w; const float* scale; float* norm;
  float* pack_io; float* pack_oi;
  void* fwd_hi; void* fwd_lo; void* bwd_hi; void* bwd_lo;
  void* mat_hi; void* mat_lo; void* matt_hi; void* matt_lo;
  int32_t out_axis, Cout, Cin, T, perm_c, perm_p, mat_rows, need_norm;
  float c;
  int32_t reserved;
} glis_wn_layer_t;
int glis_wn_prepare_multi(const glis_wn_layer_t* layers, int n, void* stream);

/* Weight gradient of a linear layer, G[row(a)][b] += sum_m dy[m][a] * x[m][b] with dy in the permuted
 * feature order of glis_wn_prepare_perm: row(a) = (a % perm_c)*perm_p + a / perm_c (identity if perm_c == 0). */
int glis_linear_wgrad(const float* dy, const float* x, float* G, int M, int Ca, int Cb, int perm_c, int perm_p,
                      void* stream);

/* glis_wn_project (below) for a LIST of layers in one launch: the projections of a whole network once all its raw
 * weight gradients are in. */
typedef struct glis_wn_proj {
  const float* G; const float* w; const float* scale; const float* norm;
  float* dw; float* dscale;
  int32_t out_axis, Cout, Cin, T, accumulate;
  float c;
} glis_wn_proj_t;
int glis_wn_project_multi(const glis_wn_proj_t* items, int n, void* stream);

/* glis_wn_project reading the raw gradient as the sum of n_slabs K-split slabs (slab s at G + s * slab_stride), added
 * in slab order: the deterministic companion of glis_conv_wgrad_bf16_slabs. */
int glis_wn_project_slabs(const float* G, int n_slabs, int64_t slab_stride, const float* w, const float* scale,
                          const float* norm, int out_axis, int Cout, int Cin, int T, float c, float* dw, float* dscale,
                          int accumulate, void* stream);

/* Weight gradient of a weight-normalised LINEAR layer and the projection below in ONE kernel (the raw gradient
 * never exists in memory): with G[row(a)][:] = sum_m dy[m][a] x[m][:] (row(a) as in glis_linear_wgrad),
 * dw[o] (+)= (s/n)(G[o] - w[o] <G[o],w[o]>/n^2), dscale[o] (+)= <G[o],w[o]>/n  (c = 1).  For batch-sized M and rows of
 * at most 1024 inputs (glis_linear_wgrad_project_supported): G's initial linear, the LIS linears.  Only master rows
 * [row_begin, row_begin + row_count) are produced: a large layer's gradient can be made — and, under data
 * parallelism, exchanged — in row chunks. */
int glis_linear_wgrad_project_supported(int M, int Ca, int Cb);
int glis_linear_wgrad_project(const float* dy, const float* x, const float* w, const float* scale, const float* norm,
                              float* dw, float* dscale, int M, int Ca, int Cb, int perm_c, int perm_p, int accumulate,
                              int row_begin, int row_count, void* stream);

/* Backward of the normalisation: given the raw gradient G w.r.t. w_hat
 * (master layout), dw[o] (+)= (s/n)(G[o] - c w[o] <G[o],w[o]>/n^2), dscale[o] (+)= <G[o],w[o]>/n.
 * `accumulate` != 0 adds into dw / dscale instead of overwriting. */
int glis_wn_project(const float* G, const float* w, const float* scale, const float* norm,
                    int out_axis, int Cout, int Cin, int T, float c, float* dw, float* dscale,
                    int accumulate, void* stream);

/* ---- convolution-shaped contractions --------------------------------------------
 * Forward of F.conv2d (WeightNormalizedConv.py:80), F.conv_transpose2d (:98), F.linear
 * (WeightNormalizedLinear.py:42) and, with the roles swapped, their data gradients.
 *   in   [N,Hi,Wi,Ci]   wpack [KH*KW][Ci][Co] (from glis_wn_prepare)   out [N,Ho,Wo,Co]
 */
int glis_conv_forward(const glis_geom_t* g, const float* in, const float* wpack,
                      const glis_epilogue_t* ep, float* out, int precision, void* stream);

/* ---- tensor-core (tcgen05) contractions on split-bf16 operands -----------------------
 * An fp32 value x is carried as two bf16 planes hi = bf16(x), lo = bf16(x - hi) with the same
 * element order (NHWC).  GLIS_PREC_BF16X3 multiplies (hi+lo)*(hi+lo) minus the lo*lo term with
 * three MMAs per k-step (~2^-16 relative, the fp32-faithful mode); GLIS_PREC_BF16 uses hi only. */

/* fp32 -> hi/lo planes (lo may be NULL). */
int glis_split_bf16(const float* x, void* hi, void* lo, int64_t numel, void* stream);

/* As glis_wn_prepare, but emitting K-major bf16 packs: fwd [T][Cout][Cin] for the launch that
 * reads Cin and writes Cout, bwd [T][Cin][Cout] for its data gradient (any may be NULL). */
int glis_wn_prepare_bf16(const float* w, const float* scale, int out_axis, int Cout, int Cin, int T,
                         float c, float* norm, void* fwd_hi, void* fwd_lo, void* bwd_hi, void* bwd_lo,
                         void* stream);

/* 1 if glis_conv_forward_bf16 can tile this geometry (Ci % 8 == 0 and >= 32 or a multiple of 64, no dilation, input
 * divisible by the stride for GLIS_CONV, output rows <= 256 pixels, Cout >= 32), else 0. */
int glis_conv_tc_supported(const glis_geom_t* g);
/* The K split (1 = none) glis_conv_forward_bf16 would use for a PLAIN-output launch (fp32 sums only) of this
 * geometry.  A caller whose fused-epilogue launch cannot fill the machine with its tiles alone (few pixels,
 * deep K: D's last level at batch 64) may then run it as split-K sums + glis_tprelu_forward_planes. */
int glis_conv_tc_ksplit(const glis_geom_t* g);
/* The launch plan glis_conv_forward_bf16 would use (host only, for inspection and tests): out15 = {tile width,
 * tile rows, tile images, UMMA N, TMEM columns, 64-channel blocks, K split, staged weight rows, pipeline stages,
 * row tiles per image, pixel tiles, channel tiles, tiles, work items, dynamic shared memory bytes}. */
int glis_conv_tc_plan(const glis_geom_t* g, int plain_out, int* out15);
/* The plan of the HALO kernel (csrc/tc_conv2.cu: the taps of a stride-parity class share one pixel box with a halo;
 * 4x4 stride-2 and 3x3 stride-1 layers) when a launch of this geometry takes it, GLIS_E_UNSUPPORTED otherwise
 * (1x1 products, GLIS_TC_HALO=0): out20 = {tw, th, tn, n_mma, tmem_cols, kblocks, ksplit, a_rows, weight slots,
 * tiles_h, tiles_x, tiles_co, total_tiles, n_groups, dynamic shared memory bytes, hx, hy, box rows, pixel-slot rows,
 * tap classes + 1000 * channels per stage (64: 128-byte rows / SWIZZLE_128B, 32: 64-byte rows / SWIZZLE_64B)}.  glis_conv_forward_bf16 uses this kernel whenever it applies; glis_conv_tc_plan keeps describing the
 * one-box-per-tap kernel of csrc/tc_conv.cu. */
int glis_conv_tc_halo_plan(const glis_geom_t* g, int plain_out, int* out20);
/* The plan of the CTA-PAIR kernel (csrc/tc_conv_pair.cu: tcgen05.mma.cta_group::2, M = 256 output channels across
 * two CTAs, each staging its 128 weight rows and half of the pixel tile) when a launch of this geometry takes it —
 * GLIS_TC_PAIR=1, Cout a multiple of 256, Cin a multiple of 64, a map too narrow for the halo kernel — and
 * GLIS_E_UNSUPPORTED otherwise: out16 = {tw, th, tn, half rows, half images, pixel rows per half, UMMA N, TMEM columns,
 * 64-channel blocks, K split, pipeline stages, row tiles per image, pixel tiles, channel pairs, tiles, work items}. */
int glis_conv_tc_pair_plan(const glis_geom_t* g, int plain_out, int* out16);

/* Same contraction as glis_conv_forward on tcgen05: TMA-fed implicit GEMM, accumulators in
 * TMEM, epilogue fused.  x planes [N,Hi,Wi,Ci] bf16, w packs [KH*KW][Co][Ci] bf16.  Outputs
 * (each optional, at least one of out_f32 / out_hi): fp32 NHWC, and the hi/lo planes of the
 * activated output for the next tensor-core layer. */
int glis_conv_forward_bf16(const glis_geom_t* g, const void* x_hi, const void* x_lo, const void* w_hi,
                           const void* w_lo, const glis_epilogue_t* ep, float* out_f32, void* out_hi,
                           void* out_lo, int precision, void* stream);

/* Raw weight gradient in master layout: G[a][b][tap] = sum_pix small[pix][a] * big[pix*s-p+k][b].
 * conv layer:  small = dy (Ca=Cout), big = x (Cb=Cin);  transposed: small = x (Ca=Cin), big = dy.
 * `g` describes the gather small(out grid: Ho,Wo,Co=Ca) <- big(Hi,Wi,Ci=Cb), relation GLIS_CONV.
 * G must be zero-filled by the caller when `accumulate` == 0 is intended (the kernel always adds).
 */
int glis_conv_wgrad(const glis_geom_t* g, const float* small, const float* big, float* G,
                    int precision, void* stream);

/* 1 if glis_conv_wgrad_bf16 can tile this geometry (channels multiples of 8, Co >= 64, Ci >= 32, fine
 * grid divisible by the stride, coarse rows <= 64 pixels). */
int glis_wgrad_tc_supported(const glis_geom_t* g);

/* glis_conv_wgrad on tcgen05 (both operands MN-major straight from the NHWC planes); adds into G. */
int glis_conv_wgrad_bf16(const glis_geom_t* g, const void* small_hi, const void* small_lo,
                         const void* big_hi, const void* big_lo, float* G, int precision, void* stream);

/* Deterministic form: the kernel splits the pixel contraction over glis_wgrad_tc_splits(g) CTAs per output tile; with
 * slabs every split STORES its partial sums into its own slab (slab s at slabs + s * Co*Ci*KH*KW floats, n_slabs =
 * that count) instead of adding them atomically into one buffer; glis_wn_project_slabs adds the slabs in a fixed
 * order.  Same inputs -> bit-identical weight gradients; nothing to zero-fill. */
int glis_wgrad_tc_splits(const glis_geom_t* g);
/* out[i] = sum_s slabs[s * slab_stride + i], added in slab order (numel, slab_stride multiples of 4). */
int glis_slab_reduce(const float* slabs, int n_slabs, int64_t slab_stride, float* out, int64_t numel, void* stream);
int glis_conv_wgrad_bf16_slabs(const glis_geom_t* g, const void* small_hi, const void* small_lo, const void* big_hi,
                               const void* big_lo, float* slabs, int n_slabs, int precision, void* stream);

/* ---- image-side layers (C <= 4 colour channels on one side; 4x4 kernel, stride 2, pad 1) ------------
 * Unfolding the image side into J = 16*C columns per coarse pixel, j = c*16 + kh*4 + kw, turns
 * D / R level 0 (common/model.py:31-36), G level 0 (:249-251) and their gradients into 1x1
 * contractions for glis_conv_forward_bf16 / glis_conv_wgrad_bf16 (see csrc/image_side.cu). */

/* x fp32 [N,H,W,C] -> bf16 hi/lo planes [N,H/2,W/2,16*C] (lo may be NULL); zero padding applied. */
int glis_unfold4x4s2_bf16(const float* x, int N, int H, int W, int C, void* hi, void* lo, void* stream);
/* cols fp32 [N,Hi,Wi,16*C] -> out fp32 [N,2Hi,2Wi,C] = fold(cols) + bias[c], act in {NONE, SIGMOID}. */
int glis_fold4x4s2(const float* cols, int N, int Hi, int Wi, int C, const float* bias, int act, float* out,
                   void* stream);
/* Effective weights as the matrix E[a][j] in master memory order (conv: a = Cout, j = ci*T + tap,
 * out_axis 0; transposed: a = Cin, j = co*T + tap, out_axis 1), scaled by scale/norm of the true
 * output channel: E (A x J) and E^T (J x A) as K-major bf16 hi/lo packs (each optional). */
int glis_wn_pack_matrix_bf16(const float* w, const float* scale, const float* norm, int out_axis, int A, int J,
                             int T, void* e_hi, void* e_lo, void* et_hi, void* et_lo, void* stream);

/* ---- the LIS module as one cluster kernel per direction (csrc/lis.cu) ------------------------
 * Reference: the residual blocks of GeneratorLearnedInputSpace (common/model.py:176-192 build them, :281-297 run
 * `x = x + lis(x)`), for norm='weight' (no scale / bias on the two linears, TPReLU between them; biases optional).
 * io1 / io2, oi1 / oi2: the [K][N] fp32 packs of glis_wn_prepare for linear 1 / 2.  code % 32 == 0, <= 256. */
int glis_lis_supported(int code);
/* u' = u + TPReLU(u W1^ + bias1) W2^ + bias2.  h (pre-activations) and act (activated) are what backward and the
 * weight gradients need; both may be NULL (no_grad). */
int glis_lis_forward(const float* u, const float* io1, const float* bias1, const float* a_raw, const float* b_t,
                     const float* io2, const float* bias2, int B, int code, float* h, float* act, float* u_out,
                     void* stream);
/* Given du_out = d(loss)/d(u'):  dh = (du_out W2^) * TPReLU'(h)  (the gradient at linear 1's output, for its weight
 * gradient),  du_in = du_out + dh W1^,  and the TPReLU parameter sums added into da / db (both or neither). */
int glis_lis_backward(const float* du_out, const float* oi2, const float* h, const float* a_raw, const float* b_t,
                      const float* oi1, int B, int code, float* dh, float* du_in, float* da, float* db,
                      void* stream);

/* ---- pointwise / reductions -------------------------------------------------------
 * TPReLU forward (common/modules/TPReLU.py:16-18) on a tensor whose channel of element i is
 * (i / inner) % C  (inner = 1 for NHWC and (B,C); H*W for NCHW-contiguous). a_raw is clamped here. */
int glis_tprelu_forward(const float* x, const float* a_raw, const float* b, float* out,
                        int64_t numel, int C, int inner, void* stream);
/* TPReLU forward of an NHWC / (B,C) tensor of pre-activations x (channel of element i = (i % C) % act_channels,
 * act_channels = 0 meaning C) writing the fp32 result (out may be NULL) and/or its bf16 hi/lo planes (hi may be
 * NULL; lo may be NULL): the epilogue of a split-K tensor-core launch, as one pointwise pass. */
int glis_tprelu_forward_planes(const float* x, const float* a_raw, const float* b, float* out, void* out_hi,
                               void* out_lo, int64_t numel, int C, int act_channels, void* stream);
/* As glis_tprelu_forward_planes on x = slab_0 + slab_1 + ... (nslabs partial sums, slab_stride elements apart,
 * added in that order); preact (may be NULL) receives x. */
int glis_tprelu_forward_planes_sum(const float* slabs, int nslabs, int64_t slab_stride, const float* a_raw,
                                   const float* b, float* preact, float* out, void* out_hi, void* out_lo,
                                   int64_t numel, int C, int act_channels, void* stream);
/* dy = dout * s * (1 - s), s = the sigmoid a contraction's epilogue applied (GLIS_ACT_SIGMOID): backward of the
 * nn.Sigmoid that ends the generators (common/model.py:136, :259), same element order for all three. */
int glis_sigmoid_backward(const float* s, const float* dout, float* dy, int64_t numel, void* stream);
/* dx = dout*(t<=0 ? clamp(a) : 1); da_raw += sum dout*t*[t<=0]*[0<=a_raw<=1]; db += sum dout*[t<=0]*(1-clamp(a)). */
int glis_tprelu_backward(const float* x, const float* a_raw, const float* b, const float* dout,
                         float* dx, float* da, float* db, int64_t numel, int C, int inner,
                         void* stream);
/* As glis_tprelu_backward, writing dx as fp32 (dx may be NULL) and/or as bf16 hi/lo planes
 * (dx_hi may be NULL; dx_lo may be NULL) so the tensor-core dgrad / wgrad can consume it directly.
 * da and db may BOTH be NULL (frozen TPReLU parameters, e.g. D during the G update): no sums. */
int glis_tprelu_backward_planes(const float* x, const float* a_raw, const float* b, const float* dout,
                                float* dx, void* dx_hi, void* dx_lo, float* da, float* db, int64_t numel,
                                int C, int inner, void* stream);
/* out[c] (+)= sum over elements of channel c (bias gradient: WeightNormalizedConv.py:47-48 backward). */
int glis_channel_sum(const float* x, float* out, int64_t numel, int C, int inner, int accumulate,
                     void* stream);

/* Mean binary cross entropy on logits against a constant target (nn.Sigmoid + nn.BCELoss,
 * common/model.py:61, g_lis/main.py:311,555,564,578).  loss[0] = mean; dlogit[i] = gscale*(p_i - t)/B
 * (may be NULL); prob (may be NULL) receives sigmoid(logit). */
int glis_bce_logits(const float* logit, float target, int B, float gscale, float* loss,
                    float* dlogit, float* prob, void* stream);

/* --ls (LSGAN): nn.MSELoss() on D's sigmoid output against a constant target (g_lis/main.py:308-311,
 * r_iterative/main.py:208-211).  loss[0] = mean((p - t)^2), p = sigmoid(logit); dlogit[i] = gscale * 2 (p_i - t)
 * p_i (1 - p_i) / B (may be NULL); prob (may be NULL) receives p. */
int glis_lsq_logits(const float* logit, float target, int B, float gscale, float* loss,
                    float* dlogit, float* prob, void* stream);

/* Training-image augmentation on the device (the reference augments on the host with imgaug, one PIL image at a time:
 * g_lis/main.py:176-231): in NCHW [0,1] (C <= 4), out NHWC [0,1]; params = 12 floats per image
 * [a00 a01 a02 a10 a11 a12 mul alpha sigma flip border _]: (a..) the inverse affine map in pixel coordinates
 * (output pixel -> source position, bilinear), flip != 0 mirrors the output horizontally, border 0 = black /
 * 1 = symmetric; then v*mul, 0.5 + alpha (v - 0.5), + sigma N(0,1) (one draw per pixel, Philox keyed by seed), clamp. */
int glis_augment(const float* in_nchw, float* out_nhwc, const float* params, int N, int C, int H, int W, uint64_t seed,
                 void* stream);

/* nn.Dropout(p) before D's final convolution (common/model.py:52-53) and nn.Dropout2d(p) in R (:344-346):
 * out = keep ? x / (1 - p) : 0.  The mask is a pure function of (seed, *counter + call, element):
 * keep(e) = u01(Philox4x32-10(key = seed, counter = {e / 4, *counter + call})[e % 4]) >= p, u01(r) = (r >> 8) / 2^24.
 * `counter` (device uint64, may be NULL = 0) is advanced once per training iteration with glis_counter_add, `call`
 * numbers the dropout calls inside an iteration: a CUDA-graph replay draws a fresh mask each time, the backward
 * pass is the same call on the gradient (nothing is stored), and a host can reproduce every mask.
 * channel_mode == 0: mask element e = element index in storage order.
 * channel_mode != 0 (Dropout2d): e = (i / per_image) * C + (i / inner) % C, one mask element per (image, channel);
 *   NHWC: inner = 1, per_image = H*W*C;  NCHW: inner = H*W, per_image = C*H*W. */
int glis_dropout(const float* x, float* out, int64_t numel, int C, int inner, int64_t per_image, int channel_mode,
                 float p, uint64_t seed, const void* counter, uint64_t call, void* stream);
/* *counter += inc (one thread; stream-ordered, capturable). */
int glis_counter_add(void* counter, uint64_t inc, void* stream);

/* lambda * mean((u - z)^2) (nn.MSELoss, g_lis/main.py:312,584-585); du (+)= 2*lambda*(u-z)/numel.
 * accumulate: bit 0 = add into du instead of overwriting; bit 1 = report the UNSCALED mean in loss[0] while du keeps
 * the factor lambda (the R-iterative trainer logs MSE(code, first_code) but back-propagates lambda^r times it,
 * r_iterative/main.py:489-497). */
int glis_mse_scaled(const float* u, const float* z, int64_t numel, float lambda, float* loss,
                    float* du, int accumulate, void* stream);

/* RMSprop without momentum (optim.RMSprop, g_lis/main.py:313-314) over one flat buffer:
 * v = alpha v + (1-alpha) g^2 ; p -= lr g / (sqrt(v)+eps).  gscale multiplies g first
 * (1/world_size after a sum all-reduce). */
int glis_rmsprop(float* p, const float* g, float* v, int64_t numel, float lr, float alpha,
                 float eps, float gscale, void* stream);

/* Standard-normal fill (torch.randn at g_lis/main.py:561,576): Philox4x32-10 + Box-Muller. */
int glis_randn(float* out, int64_t numel, uint64_t seed, uint64_t offset, void* stream);
/* U[0,1) fill (synthetic "real" batches of the benchmark). */
int glis_uniform(float* out, int64_t numel, uint64_t seed, uint64_t offset, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GLIS_B200_H */

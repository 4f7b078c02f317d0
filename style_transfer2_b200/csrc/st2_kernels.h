// Internal launch wrappers shared between translation units of libst2.
#pragma once
#include "st2_common.cuh"

// epilogue modes of the convolution kernels
enum { EPI_BIAS_RELU = 0, EPI_MASK = 1, EPI_RAW = 2 };

// ---- st2_layers.cu (layout-agnostic / CUDA-core kernels) ------------------------------------
// first layer: x fp32 NCHW (3 planes) -> NHWC T (64 ch), bias + ReLU     (K1, conv1_1)
template <typename T>
int launch_conv_first_fwd(st2_ctx* ctx, const float* x, const float* w_fwd /*[27][64]*/, const float* bias,
                          T* out, int H, int W, long long x_plane_stride = 0, int lo = 0, int hi = 0);
// first layer data gradient: g NHWC T (64 ch) -> fp32 NCHW (3 planes)    (K2, conv1_1)
template <typename T>
int launch_conv_first_bwd(st2_ctx* ctx, const T* g, const float* w_bwd /*[9][64][3] flipped*/, float* gx,
                          int H, int W, int lo = 0, int hi = 0);
// exact fp32 3x3 conv on NHWC float; w = [tap][cin][cout]; epi per EPI_*; act used by EPI_MASK
int launch_conv_exact(st2_ctx* ctx, const float* in, const float* w, const float* bias, const float* act,
                      float* out, int H, int W, int cin, int cout, int epi, int lo = 0, int hi = 0);
// Row strips: `lo` / `hi` = 1 when the row above the first / below the last of the H rows passed is an
// addressable halo row holding the neighbouring strip's data (0: outside is the zero pad).
template <typename T>
int launch_pool_fwd(st2_ctx* ctx, const T* in, T* out, int C, int H, int W);
// route g_pool back to the window's first maximum of `act`; apply_mask multiplies by act > 0
template <typename T>
int launch_pool_bwd(st2_ctx* ctx, const T* act, const T* g_pool, T* g_out, int C, int H, int W,
                    int apply_mask);
// out = mask(gin) + cc (F - Fc) + sc S + dc F ; coefficients read from scal (device doubles) or host
struct CombineArgs {
  const void* gin;      // nullable
  const void* act;      // F (always given)
  const void* fc;       // nullable
  const void* sraw;     // nullable
  void* out;
  long long n;
  int apply_mask;
  const double* coef;   // device: coef[0]=cc coef[1]=sc coef[2]=dc (nullable -> host values below)
  float h_cc, h_sc, h_dc;
};
template <typename T> int launch_combine(st2_ctx* ctx, const CombineArgs& a);
// sums[0] += sum (F-Fc)^2 (if fc), sums[1] += sum F^2
template <typename T>
int launch_feature_sums(st2_ctx* ctx, const T* act, const T* fc, long long n, double* sum_diff_sq,
                        double* sum_sq);
// Gd (C*C doubles, pre-zeroed) += sum_p F[p,i] F[p,j]; strides in elements
template <typename T>
int launch_gram_generic(st2_ctx* ctx, const T* F, int C, long long HW, long long sp, long long sc, double* Gd);
// D = Gd/(C*HW) - A (A nullable -> D = G); out_f32 = D ; *sum_dsq += sum D^2
int launch_gram_finalize(st2_ctx* ctx, const double* Gd, const float* A, float* D, int C, long long HW,
                         double* sum_dsq);
// row strips: Gd -> fp32 un-normalised sum (all-reduced by the caller), then D = Gs/(C*HW_total) - A
int launch_gram_acc_to_f32(st2_ctx* ctx, const double* Gd, float* out, int C);
int launch_gram_from_sum(st2_ctx* ctx, const float* Gs, const float* A, float* D, int C, double HW_total,
                         double* sum_dsq);
// raw[p,i] = sum_j D[i,j] F[p,j]  (strided, any T); *sum_rawsq += sum raw^2
template <typename T>
int launch_style_grad_generic(st2_ctx* ctx, const T* F, const float* D, T* raw, int C, long long HW,
                              long long sp, long long sc, double* sum_rawsq);
template <typename T>
int launch_export_nchw(st2_ctx* ctx, const T* nhwc, float* nchw, int C, int H, int W);
template <typename T>
int launch_import_nchw(st2_ctx* ctx, const float* nchw, T* nhwc, int C, int H, int W);
int launch_add_inplace(st2_ctx* ctx, float* y, const float* x, float coef_host, const double* coef_dev,
                       long long n);

// ---- st2_pixel.cu: st2_pixel_terms on a row strip (x planes xps floats apart; wrap = 0: rows -1 and H
// of x are addressable halo rows) ------------------------------------------------------------------
int pixel_terms_strip(st2_ctx* ctx, const float* x, long long xps, int wrap, const float* bwd, float* grad_out,
                      int C, int H, int W, float tv, float tv_power, float p, float p_power, float divisor,
                      double* scal, double* part = nullptr, unsigned int* counter = nullptr);

// ---- row strips: halo rows pushed and awaited INSIDE the consuming convolution kernel -------------------------
// (one process per GPU; st2_net.cu halo_args).  At kernel start the first `push_blocks` CTAs store this strip's first /
// last row of the tensor the kernel is about to convolve into the neighbours' halo rows (peer memory) and the last of
// them raises the neighbours' flags; the kernel then works through its tiles with the tile rows that touch a halo row
// scheduled LAST, and its TMA producer polls this strip's own flags only right before the first of those tiles -- by
// then the neighbours' rows have long arrived.  No exchange launch, no wait on the critical path.
struct HaloArgs {
  const unsigned char* src[2];          // my first / last interior row (dir 0: goes to the strip above, 1: below)
  unsigned char* dst[2];                // halo row over there; nullptr: no neighbour on that side
  unsigned long long* flag[2];          // flag to raise over there
  const unsigned long long* wait[2];    // my flags (raised by the strip above / below); nullptr: canvas edge
  long long bytes;                      // one row
  unsigned long long epoch;
  unsigned int* counter;                // in my slab header; 0 between kernels
  int* err;                             // sticky: a wait timed out
  int push_blocks;                      // 0: nothing to do (not a strip / old exchange kernel used)
};

// ---- st2_conv_tc.cu (tcgen05 implicit GEMM, fp16 NHWC) -----------------------------------------
struct TcConvPlan;     // tensor maps + tile geometry for one (layer, direction, canvas)
// halo = 1 (row strips): `in` points at a buffer of H + 2 rows whose first and last row are halo rows
// (neighbouring strip's data, or zeros at the canvas edge); outputs are the H interior rows.
int tc_conv_plan_create(st2_ctx* ctx, const __half* in, const __half* w_packed, int H, int W, int cin,
                        int cout, int taps, TcConvPlan** out, int halo = 0);
void tc_conv_plan_destroy(TcConvPlan* p);
// epi per EPI_*; bias fp32 (EPI_BIAS_RELU); act fp16 NHWC (EPI_MASK); sumsq nullable (sum of fp32 outputs^2)
// inj (EPI_MASK only): out = mask(acc) + coef[0] (act - fc) + coef[1] sraw + coef[2] act -- the loss diffs of
// the blob below enter under its ReLU mask in the same pass (fc / sraw nullable; coef: 3 device doubles)
// pool / pool_wp (EPI_BIAS_RELU only): also write the 2x2/2 ceil-mode max-pool of the output (NHWC, pool_wp pixels
// per row) from the same epilogue; *pooled (tc_conv_launch) tells whether the launched kernel did it
// sfuse (EPI_MASK, CTA-pair kernel with N = 128 only, after tc_conv_set_style_fuse): the style gradient D' F of the blob
// below is contracted by the same kernel into a second accumulator and added with coef[1] -- sraw must then be nullptr
struct TcInject { const __half* fc; const __half* sraw; const double* coef; __half* pool; int pool_wp; int sfuse; };
int tc_conv_launch(st2_ctx* ctx, TcConvPlan* p, const float* bias, const __half* act, __half* out, int epi,
                   float out_scale, double* sumsq, const TcInject* inj = nullptr, bool* pooled = nullptr,
                   const HaloArgs* halo = nullptr);
// does the kernel this plan launches carry the in-kernel halo push / wait?
bool tc_conv_supports_halo(const TcConvPlan* p);
// style fusion into the data-gradient convolution ABOVE a 128-channel style layer: act_below = that layer's activations
// (same geometry as this convolution's output; on strips the padded tensor), d_scaled = C x C fp16 (rewritten per evaluation)
bool tc_conv_supports_style_fuse(const TcConvPlan* p);
int tc_conv_set_style_fuse(st2_ctx* ctx, TcConvPlan* p, const __half* act_below, const __half* d_scaled);
// conv1_1 data gradient on the tensor cores: plan made with cin = 64, cout = 16 (the 3 image planes padded),
// weights [16][tap'][64] fp16; writes fp32 NCHW (3 dense planes of H x W)
// dual plan (tc_conv_dual_plan_create): gx = convT_W(grad) + dual_coef[1] * convT_W'(act) in one launch
int tc_conv_first_bwd_launch(st2_ctx* ctx, TcConvPlan* p, float* gx, const double* dual_coef = nullptr,
                             const HaloArgs* halo = nullptr);
// conv1_1 data gradient with TWO 64-channel sources: the gradient w.r.t. conv1_1 and conv1_1's activations (the style
// gradient folded into the weights, st2_net.cu style_fold_kernel).  w_dual: [16][9][128] fp16 (channels 0..63 for
// `grad`, 64..127 for `act`); separate accumulators, combined in the epilogue with a device scalar.
int tc_conv_dual_plan_create(st2_ctx* ctx, const __half* grad, const __half* act, const __half* w_dual, int H, int W,
                             TcConvPlan** out, int halo = 0);
// The same convolution in "1x1 + stencil" form (one pointwise 64 -> 27 contraction of every patch pixel, then each
// output pixel sums its 9 x 3 neighbours in shared memory): 8 MMAs per tile and source instead of 36.  act = nullptr:
// gradient source only.  w_all: [sources][32 rows = tap' * 3 + plane (27 used)][64 ch] fp16.  Launched through
// tc_conv_first_bwd_launch like the other two plans.
int tc_conv_stencil_plan_create(st2_ctx* ctx, const __half* grad, const __half* act, const __half* w_all, int H, int W,
                                TcConvPlan** out, int halo = 0);
// ---- st2_conv_first_tc.cu: conv1_1 forward on the tensor cores (sliding-window K over pixel pairs) ---------
struct TcFirstPlan;
int tc_first_plan_create(st2_ctx* ctx, int H, int W, int halo_strip, TcFirstPlan** out);
void tc_first_plan_destroy(TcFirstPlan* p);
int tc_first_pack_weights(st2_ctx* ctx, const float* w_oihw, __half* out /* 2 x 6144 halves: hi and lo parts */);
// x: row 0 of plane 0 of the H rows (planes xps floats apart, 0 = dense); lo / hi: halo rows addressable
int tc_first_fwd_launch(st2_ctx* ctx, TcFirstPlan* p, const float* x, long long xps, int lo, int hi, const __half* wpk,
                        const float* bias, __half* out);
// Gram partials with tcgen05 (MN-major operands): Gd += F^T F for fp16 NHWC features
struct TcGramPlan;
int tc_gram_plan_create(st2_ctx* ctx, const __half* F, int C, long long HW, TcGramPlan** out);
void tc_gram_plan_destroy(TcGramPlan* p);
// D = F^T F / (C*HW) - A (A nullable); *sum_dsq += sum D^2 (nullable)
int tc_gram_launch(st2_ctx* ctx, TcGramPlan* p, const float* A, float* D, double* sum_dsq);
// row strips: Gsum (C x C fp32) = F^T F of this strip, un-normalised
int tc_gram_sum_launch(st2_ctx* ctx, TcGramPlan* p, float* Gsum);
// the two halves of the above, so that an evaluation finishes ALL its Grams with one launch: the split-K
// contraction into the plan's partial tiles, then (raw = 0) D_i = sum / (C HW) - A_i and sum D_i^2, or (raw = 1)
// D_i = sum, for n <= 8 plans at once
int tc_gram_mma_launch(st2_ctx* ctx, TcGramPlan* p);
int tc_gram_finalize_all(st2_ctx* ctx, int n, TcGramPlan* const* plans, const float* const* A, float* const* D,
                         double* const* sum_dsq, int raw);

// ---- st2_elementwise.cu: 16-byte-vectorised variants (fall back to the scalar kernels) ---------
template <typename T> int launch_pool_fwd_v(st2_ctx* ctx, const T* in, T* out, int C, int H, int W);
template <typename T>
int launch_pool_bwd_v(st2_ctx* ctx, const T* act, const T* g_pool, T* g_out, int C, int H, int W, int apply_mask);
template <typename T> int launch_combine_v(st2_ctx* ctx, const CombineArgs& a);
template <typename T>
int launch_feature_sums_v(st2_ctx* ctx, const T* act, const T* fc, long long n, double* sum_diff_sq, double* sum_sq);

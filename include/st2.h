/* libst2 -- C ABI of the B200-native style-transfer hot path.
 *
 * The reference (crowsonkb/style_transfer2) has no FFI of its own for this path: its worker
 * reaches native code through pycaffe (worker.py:46-50,61,84-86,96-106), NumPy BLAS
 * (worker.py:114,263), SciPy BLAS (utils.py:34,43) and Pillow (utils.py:130-131).  The entry
 * points below are what a maintainer binds (ctypes) in place of those calls.
 * Each group names the reference interface it replaces.
 *
 * Conventions
 *   - every function returns 0 on success or a negative ST2_ERR_* code; st2_last_error() holds
 *     the message; nothing throws.
 *   - "dev" pointers are caller-owned device memory (e.g. torch allocations) on the context's
 *     device; "host" pointers are host memory.  Image-space vectors are fp32 NCHW with N = 1
 *     (the reference's layout, worker.py:63-66).  The library owns only its workspaces
 *     (packed weights, activations in kernel-native NHWC, masks, scalar blocks).
 *   - all work is enqueued on the stream given by st2_set_stream(); no call synchronises the
 *     host except those documented as returning host values.
 *   - one host thread per context; contexts are not thread-safe.
 *   - there is no CPU fallback: every entry point fails with ST2_ERR_CUDA when no sm_100 device
 *     is usable.
 */
#ifndef ST2_H_
#define ST2_H_

#ifdef __cplusplus
extern "C" {
#endif

#define ST2_NUM_BLOBS 22          /* data, conv1_1 ... pool5 (vgg19.prototxt:3-337) */
#define ST2_NUM_CONVS 16
#define ST2_SCAL_PER_BLOB 24
#define ST2_SCAL_GLOBAL_BASE (ST2_NUM_BLOBS * ST2_SCAL_PER_BLOB)
#define ST2_SCAL_TOTAL (ST2_SCAL_GLOBAL_BASE + 32)

#define ST2_ERR_CUDA (-1)
#define ST2_ERR_ARG (-2)
#define ST2_ERR_STATE (-3)
#define ST2_ERR_UNSUPPORTED (-4)

/* arithmetic of the conv stack */
#define ST2_PREC_FP32 0           /* CUDA-core fp32, exact-order reference path */
#define ST2_PREC_FP16 1           /* tcgen05 kind::f16, fp16 operands (RN), fp32 accumulate in TMEM */

#define ST2_RESAMPLE_LANCZOS 0
#define ST2_RESAMPLE_BILINEAR 1

/* indices into the global part of the scalar block (doubles), base ST2_SCAL_GLOBAL_BASE */
#define ST2_G_SCD_LOSS 0
#define ST2_G_TV_NORM 1           /* raw sum n2^(beta/2) */
#define ST2_G_P_NORM 2            /* raw sum |x|^p (before the 1/p) */
#define ST2_G_SCD_GRAD_SQ 3
#define ST2_G_T_GRAD_SQ 4
#define ST2_G_P_GRAD_SQ 5
#define ST2_G_GRAD_SQ 6
#define ST2_G_T_LOSS 7
#define ST2_G_P_LOSS 8
#define ST2_G_LOSS 9
#define ST2_G_SCD_GRAD 10
#define ST2_G_T_GRAD 11
#define ST2_G_P_GRAD 12
#define ST2_G_GRAD 13
#define ST2_G_HALO_TIMEOUT 14     /* row strips: 1.0 when a halo wait timed out (a neighbour died) */
#define ST2_G_PROTOCOL_ERROR 15   /* row strips: 1.0 when deferred sums were used before every normaliser was frozen */

typedef struct st2_ctx st2_ctx;
typedef struct st2_plan st2_plan;
typedef struct st2_lbfgs st2_lbfgs;

/* ---- context ------------------------------------------------------------------------------
 * replaces CaffeModel.__init__ / reload_net (worker.py:36-61): device selection + net load. */
int st2_ctx_create(int device, st2_ctx** out);
void st2_ctx_destroy(st2_ctx* ctx);
const char* st2_last_error(st2_ctx* ctx);       /* ctx may be NULL: error of the last failed create */
int st2_set_stream(st2_ctx* ctx, void* cuda_stream);
long long st2_launch_count(st2_ctx* ctx);       /* kernels launched so far through this context */
/* per-category device timing with CUDA events on the launch stream (off by default).  Categories:
 * 0 tcgen05 conv 3x3 (fwd+dgrad), 1 conv1_1 fwd/dgrad, 2 pool fwd/bwd, 3 Gram, 4 style gradient,
 * 5 loss reductions/combine, 6 pixel terms, 7 optimizer, 8 exact fp32 conv, 9 halo exchange (row strips:
 * peer stores + the wait for the neighbours' rows), 10 the one launch that finishes all Grams of an evaluation.
 * st2_profile_read SYNCHRONISES, writes elapsed milliseconds and span counts (ST2_PROF_CATS each)
 * accumulated since the last read, and clears them. */
#define ST2_PROF_CATS 12
int st2_profile(st2_ctx* ctx, int enable);
int st2_profile_read(st2_ctx* ctx, double* ms_out, long long* count_out);
/* measurement helpers: st2_debug_flags switches parts of the tcgen05 conv kernel off (timing
 * experiments only: 1 no epilogue stores, 2 no MMA, 4 no A loads, 8 no B loads; results are then
 * wrong by design).  st2_bench_layer times `reps` back-to-back launches of the conv producing
 * `blob` (direction 0) or of the data-gradient conv consuming its gradient (direction 1) with CUDA
 * events on the launch stream and returns the mean milliseconds per launch (SYNCHRONISES). */
int st2_debug_flags(st2_ctx* ctx, int flags);
int st2_bench_layer(st2_plan* plan, int blob, int direction, int reps, float* ms_out);
/* conv weights in Caffe blob layout: w = (Cout, Cin, 3, 3) fp32, b = (Cout) fp32, host memory.
 * conv_index follows prototxt order (0 = conv1_1 ... 15 = conv5_4). */
int st2_set_conv_weights(st2_ctx* ctx, int conv_index, const float* w_host, const float* b_host,
                         int cout, int cin);
/* topology queries (CaffeModel.layers, worker.py:73-75) */
int st2_blob_count(void);
const char* st2_blob_name(int blob);
int st2_blob_channels(int blob);
int st2_blob_kind(int blob);                    /* 0 input, 1 conv (+in-place ReLU), 2 max-pool */

/* ---- plan: one canvas size ------------------------------------------------------------------
 * replaces net.blobs['data'].reshape(...) (worker.py:84) and the blobs it implies. */
int st2_plan_create(st2_ctx* ctx, int height, int width, int precision, st2_plan** out);
void st2_plan_destroy(st2_plan* plan);
int st2_plan_blob_dims(st2_plan* plan, int blob, int* c, int* h, int* w);

/* ---- model seam (CaffeModel.forward / backward, worker.py:77-106) ---------------------------- */
/* run data -> top_blob on x (fp32 NCHW 1x3xHxW, dev). */
int st2_forward(st2_plan* plan, const float* x_dev, int top_blob);
/* copy blob's activation out as fp32 NCHW (dev).  Post-ReLU values for conv blobs. */
int st2_blob_export(st2_plan* plan, int blob, float* out_dev);
/* backward with caller-supplied diffs (fp32 NCHW, dev) added at the named blobs with the
 * reference's segment semantics (diff at convX_Y enters below reluX_Y, unmasked); writes
 * d/d(data) (fp32 NCHW) to grad_dev.  Requires the activations of the preceding st2_forward. */
int st2_backward(st2_plan* plan, int n_diffs, const int* blobs, const float* const* diffs_dev,
                 float* grad_dev);

/* ---- objective (StyleTransfer.set_content/set_style/set_weights/opfunc, worker.py:204-301) -- */
/* freeze the current activation of `blob` as the content target F_c (after st2_forward on the
 * content image). */
int st2_capture_content(st2_plan* plan, int blob);
/* gram_matrix (worker.py:109-114) of the current activation of `blob`: C x C fp32 -> out_dev */
int st2_gram(st2_plan* plan, int blob, float* out_dev);
/* gram_matrix of an arbitrary fp32 NCHW (1 x C x H x W) device array, no plan needed */
int st2_gram_nchw(st2_ctx* ctx, const float* x_dev, int channels, long long hw, float* out_dev);
/* install the style target A (C x C fp32, dev; copied) */
int st2_set_style_gram(st2_plan* plan, int blob, const float* gram_dev);
/* loss weights of one blob (a row of the weights table, worker.py:226-229); NaN counts as 0 */
int st2_set_blob_weights(st2_plan* plan, int blob, float content, float style, float deepdream);
/* evaluation order of the weighted blobs (the DataFrame index order, worker.py:234-235) */
int st2_set_eval_order(st2_plan* plan, int n, const int* blobs);
int st2_set_params(st2_plan* plan, float tv, float tv_power, float p, float p_power);
/* norms (worker.py:172-175 reset; 253-254, 265-266, 274-275 lazily frozen) */
int st2_reset_norms(st2_plan* plan);
int st2_set_norm(st2_plan* plan, int kind /*0 c,1 s,2 d*/, int blob, double value);
/* opfunc: loss and (want_grad) gradient of the objective at x.  grad_dev: fp32 NCHW.  All
 * scalars (loss terms, RMS traces, norms) land in the plan's scalar block. */
int st2_eval(st2_plan* plan, const float* x_dev, float* grad_dev, int want_grad);
/* The same evaluation in four phases; st2_eval == begin, mid, end, final back to back.  On a row strip
 * (below) the caller all-reduces (sum) st2_strip_reduce_block(which) across the strips after
 * begin (which = 0), mid (1) and end (2).  grad_dev may be NULL when want_grad was 0. */
int st2_eval_begin(st2_plan* plan, const float* x_dev, int want_grad);
int st2_eval_mid(st2_plan* plan);
int st2_eval_end(st2_plan* plan, float* grad_dev);
int st2_eval_final(st2_plan* plan);
/* read the scalar block (ST2_SCAL_TOTAL doubles) -- SYNCHRONISES the stream. */
int st2_read_scalars(st2_plan* plan, double* host_out);
/* enqueue a copy of the scalar block into PINNED host memory; no synchronisation (pair it with a
 * stream event).  Lets a caller keep a per-evaluation trace without stalling the pipeline. */
int st2_copy_scalars_async(st2_plan* plan, double* pinned_host_out);
/* device address of the scalar block (for callers that keep everything on the device) */
double* st2_scalars_dev(st2_plan* plan);

/* ---- row strips: one canvas split over several GPUs (new; the reference holds the whole image in one
 * caffe.Net, worker.py:84-86, capped by max_size, app.py:183-185) --------------------------------
 * A strip plan holds rows [row0, row1) of an H_total x W canvas plus one halo row on either side of
 * every activation / gradient tensor.  Strips start at multiples of 32 rows (all five pools stay strip-local).  x / grad / L-BFGS
 * vectors of a strip are dense fp32 (3, row1-row0, W).  Neighbouring strips (circular: the TV term
 * wraps around the canvas) are attached either through a CUDA IPC handle (one process per GPU; halo
 * rows then travel as peer stores over NVLink) or directly when they live in the same process.
 * st2_forward / st2_capture_content / st2_eval_* work on strip plans and must be called by all strips
 * in the same order; st2_gram returns the strip's UN-normalised Gram sum; st2_eval (one shot) works
 * only for world == 1; st2_backward is not available. */
#define ST2_IPC_HANDLE_BYTES 64
int st2_strip_plan_create(st2_ctx* ctx, int height_total, int width, int row0, int row1, int rank,
                          int world, int precision, st2_plan** out);
int st2_strip_ipc_handle(st2_plan* plan, void* handle_out /* ST2_IPC_HANDLE_BYTES */);
/* side 0: the strip above (rank-1, or the last strip for rank 0), side 1: the strip below.  Exactly one of
 * ipc_handle / local_peer is given.  peer_rows: the rows that strip holds. */
int st2_strip_attach(st2_plan* plan, int side, const void* ipc_handle, st2_plan* local_peer, int peer_rows);
/* conv1_1's style gradient can be folded into its data-gradient weights instead of being materialised (fp16 path);
 * grad(conv1_1) then has another meaning, and strips read a row of each other's: enable only when EVERY strip of the
 * canvas holds >= 16 rows and >= 16 columns (the tensor-core conv1_1 kernels' minimum).  Off by default on strips. */
int st2_strip_set_fold(st2_plan* plan, int enable);
/* Steady state: once every active normaliser is frozen (any completed evaluation with the current weights), block 1
 * is only needed for trace values, so its all-reduce can wait: with `enable`, st2_eval_end does not expect block 1 to
 * be reduced yet, and the caller all-reduces blocks 1 and 2 (and whatever else it has, e.g. the L-BFGS dot products)
 * in ONE collective between st2_eval_end and st2_eval_final: two all-reduces per iteration instead of four.  Using it
 * before the normalisers are frozen sets ST2_G_PROTOCOL_ERROR. */
int st2_strip_set_deferred(st2_plan* plan, int enable);
/* which 0: Gram sums (fp32), 1: per-blob partial sums (f64), 2: six pixel-space sums (f64) */
int st2_strip_reduce_block(st2_plan* plan, int which, void** dev_out, long long* count_out);
/* 1 when a halo wait timed out since the plan was created -- SYNCHRONISES */
int st2_strip_halo_error(st2_plan* plan, int* err_out);

/* ---- pixel-space pieces, usable on their own -------------------------------------------------
 * utils.tv_norm / p_norm (utils.py:285-304) + gradient assembly (worker.py:295-297).
 * grad_out = bwd + tv*tv_grad(x/divisor) + p*p_grad(x/divisor); sums -> scal_dev[ST2_G_*]
 * (accumulated).  The objective uses divisor = 255 (worker.py:283,287). */
int st2_pixel_terms(st2_ctx* ctx, const float* x_dev, const float* bwd_dev, float* grad_out_dev,
                    int channels, int height, int width, float tv, float tv_power, float p,
                    float p_power, float divisor, double* scal_dev);
/* CaffeModel.preprocess / deprocess (worker.py:63-71) */
int st2_preprocess_u8(st2_ctx* ctx, const unsigned char* hwc_dev, float* nchw_dev, int h, int w);
int st2_preprocess_f32(st2_ctx* ctx, const float* hwc_dev, float* nchw_dev, int h, int w);
int st2_deprocess(st2_ctx* ctx, const float* nchw_dev, float* hwc_dev, int h, int w);

/* ---- level-1 helpers (utils.dot / utils.axpy, utils.py:29-46) -------------------------------- */
int st2_dot(st2_ctx* ctx, const float* a_dev, const float* b_dev, long long n, double* host_out);
int st2_axpy(st2_ctx* ctx, float alpha, const float* x_dev, float* y_dev, long long n);
int st2_sumsq(st2_ctx* ctx, const float* a_dev, long long n, double* host_out);

/* ---- L-BFGS (optimizers.LBFGSOptimizer, optimizers.py:49-125) -------------------------------- */
int st2_lbfgs_create(st2_ctx* ctx, long long n, int n_corr, st2_lbfgs** out);
void st2_lbfgs_destroy(st2_lbfgs* opt);
int st2_lbfgs_reset(st2_lbfgs* opt);                       /* objective_changed(): drop history */
/* s = -step * inv_hv(g);  x += s   (optimizers.py:67-69, 89-108). */
int st2_lbfgs_advance(st2_lbfgs* opt, float* x_dev, const float* g_dev, float step_size);
/* y = g_new - g_prev; keep (s, y) iff s.y > 1e-10; FIFO cap n_corr (optimizers.py:72-74,79-87) */
int st2_lbfgs_commit(st2_lbfgs* opt, const float* g_new_dev, const float* g_prev_dev);
/* The same two calls split around their cross-vector reductions, for callers that shard the vectors
 * over several GPUs (row strips): after *_begin, all-reduce (sum) the st2_lbfgs_sums_count() doubles
 * at st2_lbfgs_sums_dev(), then call *_end.  advance_begin only produces sums on the first step
 * after a reset/load.  st2_lbfgs_set_global_length gives the un-sharded vector length (p.size in
 * optimizers.py:102). */
int st2_lbfgs_advance_begin(st2_lbfgs* opt, const float* g_dev);
int st2_lbfgs_advance_end(st2_lbfgs* opt, float* x_dev, const float* g_dev, float step_size);
int st2_lbfgs_commit_begin(st2_lbfgs* opt, const float* g_new_dev, const float* g_prev_dev);
int st2_lbfgs_commit_end(st2_lbfgs* opt);
double* st2_lbfgs_sums_dev(st2_lbfgs* opt);
int st2_lbfgs_sums_count(void);
int st2_lbfgs_set_global_length(st2_lbfgs* opt, double n_total);
/* state transfer for teacher-forced tests: count pairs, oldest first; S, Y: count x n fp32 (dev) */
int st2_lbfgs_load(st2_lbfgs* opt, int count, const float* s_dev, const float* y_dev,
                   const double* sy_host);
int st2_lbfgs_export(st2_lbfgs* opt, int* count_out, float* s_dev, float* y_dev, double* sy_host);

/* ---- Adam (optimizers.AdamOptimizer.step, optimizers.py:20-27; utils.DecayingMean) ----------- */
/* m1 <- b1 m1 + (1-b1) g ; m2 <- b2 m2 + (1-b2) g^2 ; x -= step * m1hat / (sqrt(m2hat) + 1e-8)
 * with m1hat = m1 / (1 - b1^items1), m2hat = m2 / (1 - b2^items2); items counted AFTER this update */
int st2_adam_step(st2_ctx* ctx, float* x_dev, const float* g_dev, float* m1_dev, float* m2_dev,
                  long long n, float step_size, double b1, double b2, int items1, int items2);

/* ---- resampling (utils.resample_nchw -> Pillow Image.resize on mode 'F', utils.py:130-160) ---- */
int st2_resample(st2_ctx* ctx, const float* src_dev, int planes, int h_in, int w_in, float* dst_dev,
                 int h_out, int w_out, int method, int clamp_min_zero);

#ifdef __cplusplus
}
#endif
#endif /* ST2_H_ */

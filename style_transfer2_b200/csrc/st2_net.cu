// Context, weights, per-canvas plan and the orchestration of forward / backward / objective.
// Reference: worker.py:32-106 (CaffeModel seam), 109-114 (gram_matrix), 231-301 (opfunc).
#include "st2_kernels.h"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

// --------------------------------------------------------------------------------------------
const BlobInfo g_blobs[ST2_NUM_BLOBS] = {
    {"data", KIND_INPUT, 3, -1},
    {"conv1_1", KIND_CONV, 64, 0},   {"conv1_2", KIND_CONV, 64, 1},   {"pool1", KIND_POOL, 64, -1},
    {"conv2_1", KIND_CONV, 128, 2},  {"conv2_2", KIND_CONV, 128, 3},  {"pool2", KIND_POOL, 128, -1},
    {"conv3_1", KIND_CONV, 256, 4},  {"conv3_2", KIND_CONV, 256, 5},  {"conv3_3", KIND_CONV, 256, 6},
    {"conv3_4", KIND_CONV, 256, 7},  {"pool3", KIND_POOL, 256, -1},
    {"conv4_1", KIND_CONV, 512, 8},  {"conv4_2", KIND_CONV, 512, 9},  {"conv4_3", KIND_CONV, 512, 10},
    {"conv4_4", KIND_CONV, 512, 11}, {"pool4", KIND_POOL, 512, -1},
    {"conv5_1", KIND_CONV, 512, 12}, {"conv5_2", KIND_CONV, 512, 13}, {"conv5_3", KIND_CONV, 512, 14},
    {"conv5_4", KIND_CONV, 512, 15}, {"pool5", KIND_POOL, 512, -1},
};

static std::string g_create_error;

std::vector<const void*>& st2_kernel_registry() {
  static std::vector<const void*> v;
  return v;
}

std::vector<St2SmemOptIn>& st2_smem_registry() {
  static std::vector<St2SmemOptIn> v;
  return v;
}

int st2_fail(st2_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_error = buf;
  return code;
}

static cudaEvent_t prof_event(st2_ctx* ctx) {
  cudaEvent_t e = nullptr;
  if (!ctx->prof_pool.empty()) { e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
  else cudaEventCreate(&e);
  return e;
}

ProfScope::ProfScope(st2_ctx* c, int cat) : ctx(c), idx(-1) {
  if (!c || !c->prof_on) return;
  st2_ctx::ProfSpan sp;
  sp.cat = cat; sp.a = prof_event(c); sp.b = nullptr;
  cudaEventRecord(sp.a, c->stream);
  idx = (int)c->prof_spans.size();
  c->prof_spans.push_back(sp);
}

ProfScope::~ProfScope() {
  if (idx < 0) return;
  cudaEvent_t e = prof_event(ctx);
  cudaEventRecord(e, ctx->stream);
  ctx->prof_spans[idx].b = e;
}

// --------------------------------------------------------------------------------------------
namespace {

// dst index helpers for the four weight packs; src is OIHW (cout, cin, 3, 3)
__global__ void pack_weights_kernel(const float* __restrict__ w, int cout, int cin, float* f_fwd, float* f_bwd,
                                    __half* h_fwd, __half* h_bwd) {
  const long long total = (long long)cout * cin * 9;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i % 3), r = (int)((i / 3) % 3);
    const int ci = (int)((i / 9) % cin), co = (int)(i / (9LL * cin));
    const float v = w[i];
    const int tap = r * 3 + s, tapf = (2 - r) * 3 + (2 - s);
    f_fwd[((long long)tap * cin + ci) * cout + co] = v;              // [tap][ci][co]
    f_bwd[((long long)tapf * cout + co) * cin + ci] = v;             // [tap'][co][ci]
    if (h_fwd) h_fwd[((long long)co * 9 + tap) * cin + ci] = __float2half_rn(v);     // [co][tap][ci]
    if (h_bwd) h_bwd[((long long)ci * 9 + tapf) * cout + co] = __float2half_rn(v);   // [ci][tap'][co] (conv1_1: 16 rows, 3 used)
  }
}

struct EvalSpec {
  int n;
  int order[ST2_NUM_BLOBS];
  float cw[ST2_NUM_BLOBS], sw[ST2_NUM_BLOBS], dw[ST2_NUM_BLOBS];
  int C[ST2_NUM_BLOBS];
  double nelem[ST2_NUM_BLOBS];
  float tv, tv_power, p, p_power;
  double N;
};

__device__ __forceinline__ bool w_on(float w) { return fabsf(w) > 1e-15f; }   // NaN -> false (worker.py:234)

// zero the per-evaluation accumulators, keep norms / valid flags
__global__ void clear_volatile_kernel(double* scal) {
  pdl_trigger();
  pdl_wait();
  for (int i = threadIdx.x; i < ST2_SCAL_TOTAL; i += blockDim.x) {
    if (i >= ST2_SCAL_GLOBAL_BASE) { scal[i] = 0.0; continue; }
    const int f = i % ST2_SCAL_PER_BLOB;
    if (f < SB_C_NORM || f > SB_D_VALID) scal[i] = 0.0;
  }
}

// fp16 copy of D scaled by a power of two chosen from rms(D): |D'| <= ~1, |D' F| <= ~max|F|
__global__ void style_scale_kernel(const float* __restrict__ D, __half* __restrict__ Dh, int C, double* sb) {
  pdl_trigger();
  pdl_wait();
  const double n = (double)C * C;
  const double rms = sqrt(sb[SB_S_GRAMSQ] / n);
  float ds = 1.0f;
  if (rms > 0.0 && isfinite(rms)) ds = exp2f(-ceilf(log2f((float)rms * (float)C)));
  if (blockIdx.x == 0 && threadIdx.x == 0) sb[SB_S_DSCALE] = (double)ds;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)C * C;
       i += (long long)gridDim.x * blockDim.x)
    Dh[i] = __float2half_rn(D[i] * ds);
}

// conv1_1 only.  Its style gradient s = sc (D F) enters blob.diff below the ReLU (worker.py:100) and goes straight
// into conv1_1's own data-gradient convolution (worker.py:104), which is linear:
//     gx += sc * convT_W(D F) = sc * convT_W'(F),     W'[plane][tap'][j] = sum_i W[plane][tap'][i] D[i][j].
// So instead of materialising s (a 134 MB fp16 tensor at 1024^2, written by a 1x1 contraction and read back by
// conv1_2's data-gradient epilogue) the 64 x 64 matrix D is folded into the 64 -> 3 weights once per evaluation and
// the data-gradient kernel makes a second pass over F with W' (accumulating into the fp32 gradient planes).
// The trace value and the normaliser need sum s^2, which needs no s either:
//     sum_p |D' F_p|^2 = <D'^T D', F^T F> = <D'^T D', (D + A) C HW>      (D' = ds D, the scaled copy as before).
// 32 blocks x 128 threads: one (j, k) pair of D'^T D' per thread; D (16 KB) stays in L1 / L2.
// wdual: the data-gradient weights of conv1_1 as [16 rows][9 taps][128]: channels 0..63 = W (gradient channels,
// written once at plan creation), 64..127 = W' (activation channels, rewritten here every evaluation).
__global__ void __launch_bounds__(128)
style_fold_kernel(const float* __restrict__ D, const float* __restrict__ A, const float* __restrict__ gsum_local,
                  const float* __restrict__ w_oihw, __half* __restrict__ wdual, int stencil_layout, double n_total,
                  double* sb, double* raw_sum) {
  // gsum_local (row strips): THIS strip's un-normalised Gram sum F^T F, so that the strip contributes exactly its own
  // sum_p |D' F_p|^2 to the all-reduced total (like a strip that does not fold); nullptr: the whole canvas,
  // F^T F = (D + A) C HW.
  constexpr int C = 64;
  pdl_trigger();
  pdl_wait();
  const double rms = sqrt(sb[SB_S_GRAMSQ] / (double)(C * C));
  float ds = 1.0f;
  if (rms > 0.0 && isfinite(rms)) ds = exp2f(-ceilf(log2f((float)rms * (float)C)));
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // 0 .. 4095
  if (idx == 0) sb[SB_S_DSCALE] = (double)ds;
  if (idx < 3 * 9 * C) {
    const int j = idx % C, tapf = (idx / C) % 9, plane = idx / (9 * C);
    const int r = 2 - tapf / 3, sx = 2 - tapf % 3;              // tap' = flipped tap of the forward weights
    float acc = 0.f;
    // (unroll 32: every unrolled group of rows is one L2 round trip for the whole warp, and the groups are serial)
#pragma unroll 32
    for (int i = 0; i < C; ++i) acc = fmaf(__ldg(w_oihw + ((i * 3 + plane) * 3 + r) * 3 + sx), __ldg(D + i * C + j), acc);
    // stencil_layout: [source 1][row = tap' * 3 + plane][64]; else the dual pack [plane][tap'][128], channels 64..127
    if (stencil_layout) wdual[(32 + tapf * 3 + plane) * C + j] = __float2half_rn(acc * ds);
    else wdual[(plane * 9 + tapf) * 128 + 64 + j] = __float2half_rn(acc * ds);
  }
  double tot = 0.0;
  if (idx < C * C) {
    const int j = idx / C, k = idx % C;
    double e = 0.0;
#pragma unroll 32
    for (int i = 0; i < C; ++i) e += (double)__ldg(D + i * C + j) * (double)__ldg(D + i * C + k);
    tot = gsum_local != nullptr ? e * (double)gsum_local[idx]
                                : e * ((double)D[idx] + (A != nullptr ? (double)A[idx] : 0.0)) * n_total;
  }
  // one atomic per block: same-address fp64 atomics serialise in L2, and the kernel's end waits for all of them
  __shared__ double red[4];
  tot = warp_sum_d(tot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(raw_sum, (red[0] + red[1] + red[2] + red[3]) * (double)ds * (double)ds);
}

// sum_p |D' F_p|^2 = <D'^T D', F^T F> for a style layer whose gradient is never materialised (see above; ds is read
// from sb[SB_S_DSCALE], written by style_scale_kernel just before).  gsum_local: this strip's un-normalised Gram sum,
// or nullptr for the whole canvas (F^T F = (D + A) C HW).  One (j, k) pair per thread.
__global__ void __launch_bounds__(256)
style_rawsq_kernel(const float* __restrict__ D, const float* __restrict__ A, const float* __restrict__ gsum_local, int C,
                   double n_total, const double* sb, double* raw_sum) {
  pdl_trigger();
  pdl_wait();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double tot = 0.0;
  if (idx < (long long)C * C) {
    const int j = (int)(idx / C), k = (int)(idx % C);
    double e = 0.0;
#pragma unroll 32
    for (int i = 0; i < C; ++i) e += (double)__ldg(D + (long long)i * C + j) * (double)__ldg(D + (long long)i * C + k);
    tot = gsum_local != nullptr ? e * (double)gsum_local[idx]
                                : e * ((double)D[idx] + (A != nullptr ? (double)A[idx] : 0.0)) * n_total;
  }
  __shared__ double red[8];
  tot = warp_sum_d(tot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    const double ds = sb[SB_S_DSCALE];
    if (t != 0.0) atomicAdd(raw_sum, t * ds * ds);
  }
}

// conv1_1 data-gradient weights [16][9][64] -> the gradient-channel half of the dual pack [16][9][128]
__global__ void dual_pack_kernel(const __half* __restrict__ wbwd, __half* __restrict__ wdual) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 16 * 9 * 64) return;
  const int j = idx % 64, rt = idx / 64;
  wdual[rt * 128 + j] = wbwd[idx];
}
// ... -> the stencil form's pointwise weights [row = tap' * 3 + plane (27 of 32)][64] (tc_conv_first_stencil_kernel)
__global__ void stencil_pack_kernel(const __half* __restrict__ wbwd, __half* __restrict__ wall) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 27 * 64) return;
  const int j = idx % 64, row = idx / 64, tapf = row / 3, plane = row % 3;
  wall[idx] = wbwd[(plane * 9 + tapf) * 64 + j];
}

// worker.py:249-277: freeze normalisers on first use, derive combine coefficients and trace values
// mode 0: everything.  Row strips with deferred sums (st2_strip_set_deferred): mode 1 before the backward pass -- the
// combine coefficients only, from the FROZEN normalisers (the sums are not reduced yet; a normaliser that is not
// frozen is a protocol error and raises the sticky flag); mode 2 after the merged all-reduce -- the trace values.
__global__ void coef_kernel(EvalSpec es, double* scal, int mode) {
  pdl_trigger();
  pdl_wait();
  const int k = threadIdx.x;
  if (k >= es.n) return;
  const int b = es.order[k];
  double* sb = scal + b * ST2_SCAL_PER_BLOB;
  double* g = scal + ST2_SCAL_GLOBAL_BASE;
  const double n = es.nelem[b];
  const double C = (double)es.C[b];
  const bool coefs = mode != 2, traces = mode != 1;
  if (w_on(es.cw[b])) {
    const double msd = sb[SB_C_SUMSQ] / n;
    if (sb[SB_C_VALID] == 0.0) {
      if (mode == 0) { sb[SB_C_NORM] = (2.0 / n) * sqrt(msd); sb[SB_C_VALID] = 1.0; }
      else g[ST2_G_PROTOCOL_ERROR] = 1.0;
    }
    const double cn = sb[SB_C_NORM];
    if (coefs) sb[SB_C_COEF] = (double)es.cw[b] / cn * (2.0 / n);
    if (traces) {
      sb[SB_C_LOSS] = (double)es.cw[b] * msd / cn;
      sb[SB_C_GRAD] = fabs((double)es.cw[b] / cn) * (2.0 / n) * sqrt(msd);
    }
  }
  if (w_on(es.sw[b])) {
    double ds = sb[SB_S_DSCALE];
    if (ds == 0.0) ds = 1.0;
    const double kk = 2.0 / (C * C * n);                       // 2 / (gram_diff.size * feat.size)
    const double raw2 = sb[SB_S_RAWSQ] / (ds * ds);
    if (sb[SB_S_VALID] == 0.0) {
      if (mode == 0) { sb[SB_S_NORM] = kk * sqrt(raw2 / n); sb[SB_S_VALID] = 1.0; }
      else g[ST2_G_PROTOCOL_ERROR] = 1.0;
    }
    const double sn = sb[SB_S_NORM];
    if (coefs) sb[SB_S_COEF] = (double)es.sw[b] / sn * kk / ds;
    if (traces) {
      sb[SB_S_LOSS] = (double)es.sw[b] * (sb[SB_S_GRAMSQ] / (C * C)) / sn;
      sb[SB_S_GRAD] = fabs((double)es.sw[b] / sn) * kk * sqrt(raw2 / n);
    }
  }
  if (w_on(es.dw[b])) {
    const double msf = sb[SB_D_SUMSQ] / n;
    if (sb[SB_D_VALID] == 0.0) {
      if (mode == 0) { sb[SB_D_NORM] = (2.0 / n) * sqrt(msf); sb[SB_D_VALID] = 1.0; }
      else g[ST2_G_PROTOCOL_ERROR] = 1.0;
    }
    const double dn = sb[SB_D_NORM];
    if (coefs) sb[SB_D_COEF] = (double)es.dw[b] / dn * (-2.0 / n);
    if (traces) {
      sb[SB_D_LOSS] = -(double)es.dw[b] * msf / dn;
      sb[SB_D_GRAD] = fabs((double)es.dw[b] / dn) * (2.0 / n) * sqrt(msf);
    }
  }
}

// worker.py:279-301: totals in the reference's accumulation order
__global__ void final_kernel(EvalSpec es, double* scal, const int* halo_err) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double* g = scal + ST2_SCAL_GLOBAL_BASE;
  g[ST2_G_HALO_TIMEOUT] = (halo_err != nullptr && *halo_err != 0) ? 1.0 : 0.0;
  double scd = 0.0;
  for (int k = 0; k < es.n; ++k) {
    const int b = es.order[k];
    const double* sb = scal + b * ST2_SCAL_PER_BLOB;
    if (w_on(es.cw[b])) scd += sb[SB_C_LOSS];
    if (w_on(es.sw[b])) scd += sb[SB_S_LOSS];
    if (w_on(es.dw[b])) scd += sb[SB_D_LOSS];
  }
  g[ST2_G_SCD_LOSS] = scd;
  g[ST2_G_T_LOSS] = (double)es.tv * g[ST2_G_TV_NORM];
  g[ST2_G_P_LOSS] = (double)es.p * (g[ST2_G_P_NORM] / (double)es.p_power);
  g[ST2_G_LOSS] = scd + g[ST2_G_T_LOSS] + g[ST2_G_P_LOSS];
  g[ST2_G_SCD_GRAD] = sqrt(g[ST2_G_SCD_GRAD_SQ] / es.N);
  g[ST2_G_T_GRAD] = sqrt(g[ST2_G_T_GRAD_SQ] / es.N);
  g[ST2_G_P_GRAD] = sqrt(g[ST2_G_P_GRAD_SQ] / es.N);
  g[ST2_G_GRAD] = sqrt(g[ST2_G_GRAD_SQ] / es.N);
}

struct Blob {
  int C = 0, H = 0, W = 0;    // H: rows held here (the strip's rows when the canvas is row-tiled)
  int Hg = 0;                 // rows of the whole canvas at this level
  void* act = nullptr;        // T NHWC (data: caller's fp32 NCHW x); strips: first interior row
  void* grad = nullptr;       // T NHWC gradient w.r.t. the (post-ReLU) blob; strips: first interior row
  void* act_pad = nullptr;    // strips: start of the top halo row (act - one row)
  void* grad_pad = nullptr;
  void* fc = nullptr;         // content target, same layout as act
  void* sraw = nullptr;       // unscaled style gradient D F
  void* inj = nullptr;        // imported diff for the model-seam backward
  float* gram_target = nullptr;   // A, C x C fp32
  float* D = nullptr;             // gram(F) - A, C x C fp32
  __half* Dh = nullptr;           // scaled fp16 copy for the tcgen05 1x1 contraction
  float cw = 0.f, sw = 0.f, dw = 0.f;
  TcConvPlan* tc_fwd = nullptr;   // producing this blob (conv blobs, idx >= 1)
  TcConvPlan* tc_bwd = nullptr;   // consuming grad of this blob, producing grad of the blob below
  TcConvPlan* tc_style = nullptr;
  TcConvPlan* tc_sfold = nullptr;  // conv1_1: data gradient with two sources (gradient + activations), dual weights
  __half* wfold = nullptr;         // conv1_1: [16][9][128] fp16: W for the gradient channels | W' = W D for the activations
  bool dfuse_set = false;          // the tensor maps of the style fusion are installed in the plan of the layer above
  TcGramPlan* tc_gram = nullptr;
  TcFirstPlan* tc_first = nullptr; // conv1_1 only
  long long n() const { return (long long)C * H * W; }
  double n_total() const { return (double)C * (double)Hg * (double)W; }
};

struct Inject {
  bool on = false;
  bool fold = false;          // conv1_1: the style term is applied by a second data-gradient pass (style_fold_kernel)
  bool dfuse = false;         // the data-gradient kernel of the layer ABOVE contracts D' F itself (TcInject::sfuse)
  const void* fc = nullptr;
  const void* sraw = nullptr;
  const double* coef = nullptr;
  float hcc = 0.f, hsc = 0.f, hdc = 0.f;
};

}  // namespace

// ---- row strips: one canvas split over several GPUs ------------------------------
// Every activation / gradient tensor of a strip carries one halo row above and below its own rows.
// After a layer is produced, a push kernel stores the strip's first / last row straight into the
// neighbouring strips' halo rows (peer memory: CUDA IPC mapping over NVLink, or the same address
// space when several strips share a GPU) and raises a flag there; the consumer's next convolution is
// preceded by a one-thread kernel that spins on its own flags.  No host synchronisation, no NCCL on
// the halo path.  All of a strip's halo-carrying buffers live in ONE allocation (the slab) so a
// neighbour needs a single IPC handle; its layout is a pure function of (rows, W, element size).
constexpr int kSlots = 2 * ST2_NUM_BLOBS;      // slot b: act of blob b (0 = x); ST2_NUM_BLOBS + b: grad of blob b
constexpr size_t kSlabHeader = 4096;
struct SlabHeader {
  unsigned long long flag_from[2][kSlots];     // [0]: raised by the strip above, [1]: by the strip below
  unsigned int counters[2];
  int err;                                     // sticky: a wait timed out
  unsigned int push_counter;                   // in-kernel halo push (HaloArgs::counter)
};
static_assert(sizeof(SlabHeader) <= kSlabHeader, "slab header");

struct StripLayout {
  int rows[ST2_NUM_BLOBS], W[ST2_NUM_BLOBS];
  size_t act_off[ST2_NUM_BLOBS], grad_off[ST2_NUM_BLOBS], row_bytes[ST2_NUM_BLOBS];
  size_t xp_off, total;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static StripLayout strip_layout(int rows0, int W0, size_t esz) {
  StripLayout L;
  size_t off = kSlabHeader;
  L.xp_off = off;
  off = align_up(off + sizeof(float) * 3 * (size_t)(rows0 + 2) * W0, 1024);
  int h = rows0, w = W0;
  for (int i = 0; i < ST2_NUM_BLOBS; ++i) {
    if (g_blobs[i].kind == KIND_POOL) { h = pool_extent(h); w = pool_extent(w); }
    L.rows[i] = h; L.W[i] = w;
    L.row_bytes[i] = (size_t)w * g_blobs[i].channels * esz;
    L.act_off[i] = L.grad_off[i] = 0;
    if (i == 0) continue;
    L.act_off[i] = off;
    off = align_up(off + L.row_bytes[i] * (size_t)(h + 2), 1024);
    L.grad_off[i] = off;
    off = align_up(off + L.row_bytes[i] * (size_t)(h + 2), 1024);
  }
  L.total = off;
  return L;
}

struct HaloPush {
  const unsigned char* src[2];          // my first / last interior row
  unsigned char* dst[2];                // the bottom halo row of the strip above / top halo row of the strip below
  unsigned long long* flag[2];          // the flag to raise over there
  long long src_seg_stride, dst_seg_stride[2];
  long long seg_bytes;
  int nseg;
  unsigned long long epoch;
  unsigned int* counters;
};

// grid (blocks, 2): y = 0 pushes up, y = 1 pushes down; the last block of a direction to finish raises the flag
// over there.  Block (0, 0) then also does this strip's own waiting: thread 0 for the strip above, thread 1 for the
// strip below (one launch per exchange instead of a push + a wait kernel).
__global__ void halo_exchange_kernel(const HaloPush a, const unsigned long long* wait_up,
                                     const unsigned long long* wait_dn, int* err) {
  const int dir = blockIdx.y;
  if (a.dst[dir] != nullptr) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (int sgm = 0; sgm < a.nseg; ++sgm) {
      const unsigned char* sp = a.src[dir] + sgm * a.src_seg_stride;
      unsigned char* dp = a.dst[dir] + sgm * a.dst_seg_stride[dir];
      if ((a.seg_bytes & 15) == 0 && (((uintptr_t)sp | (uintptr_t)dp) & 15) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(sp);
        uint4* d4 = reinterpret_cast<uint4*>(dp);
        for (long long i = tid; i < (a.seg_bytes >> 4); i += nth) d4[i] = s4[i];
      } else {
        const unsigned int* s1 = reinterpret_cast<const unsigned int*>(sp);
        unsigned int* d1 = reinterpret_cast<unsigned int*>(dp);
        for (long long i = tid; i < (a.seg_bytes >> 2); i += nth) d1[i] = s1[i];
      }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned int prev = atomicAdd(&a.counters[dir], 1u);
      if (prev == gridDim.x - 1) {
        a.counters[dir] = 0;
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag[dir]), "l"(a.epoch) : "memory");
      }
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 2) {
    const unsigned long long* f = threadIdx.x == 0 ? wait_up : wait_dn;
    if (f == nullptr || *reinterpret_cast<volatile int*>(err) != 0) return;
    unsigned long long t0, t1, v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= a.epoch) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 10000000000ull) { atomicExch(err, 1); break; }     // 10 s: the neighbour died; fail loudly, don't hang
      __nanosleep(64);
    }
  }
}

// x (3 dense planes of rows x W) -> interior rows of the padded planes
__global__ void pack_x_kernel(const float* __restrict__ x, float* __restrict__ xp, int rows, int W) {
  const long long plane = (long long)rows * W, total = 3 * plane;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / plane);
    const long long r = i - (long long)c * plane;
    xp[(long long)c * (rows + 2) * W + W + r] = x[i];
  }
}

// the all-reduced partial sums -> their scalar-block slots
__global__ void scatter_sums_kernel(const double* __restrict__ red, double* __restrict__ scal) {
  const int b = threadIdx.x;
  if (b >= ST2_NUM_BLOBS) return;
  double* sb = scal + b * ST2_SCAL_PER_BLOB;
  sb[SB_C_SUMSQ] = red[3 * b + 0];
  sb[SB_D_SUMSQ] = red[3 * b + 1];
  sb[SB_S_RAWSQ] = red[3 * b + 2];
}

struct st2_plan {
  st2_ctx* ctx;
  int H, W, prec;             // H: rows held here
  size_t esz;
  // row-strip state (strip == false: the plan holds the whole canvas)
  bool strip = false, edge_top = true, edge_bot = true;
  bool async_halo = false;     // halo rows travel inside the consuming convolution kernels (all neighbours over IPC)
  bool strip_fold = false;     // every strip of the canvas folds conv1_1's style gradient (st2_strip_set_fold)
  bool deferred_sums = false;  // st2_strip_set_deferred: block 1 is all-reduced after st2_eval_end, with block 2
  bool eval_deferred = false;  // ... as decided for the evaluation in flight
  int rank = 0, world = 1, row0 = 0, H_total = 0;
  unsigned char* slab = nullptr;
  StripLayout lay;
  unsigned char* peer[2] = {nullptr, nullptr};      // slab of the strip above / below (circular)
  bool peer_ipc[2] = {false, false};
  StripLayout peer_lay[2];
  unsigned long long epoch[kSlots] = {};
  float* xp = nullptr;                               // padded copy of x: 3 x (H + 2) x W
  float* gram_red = nullptr;                         // strip Gram sums, fp32, all-reduced by the caller
  float* gram_local[ST2_NUM_BLOBS] = {};             // strip Gram sums as they were before the all-reduce (conv1_1 / conv2_1:
                                                     // their style gradient is never materialised, see fold_eligible / dfuse_eligible)
  long long gram_red_off[ST2_NUM_BLOBS] = {};
  long long gram_red_used = 0;
  double* red = nullptr;                             // 3 partial sums per blob, all-reduced by the caller
  // evaluation state carried between the phases
  EvalSpec es;
  Inject inj[ST2_NUM_BLOBS];
  int eval_top = 0, eval_want_grad = 0, eval_phase = 0;
  const float* eval_x = nullptr;
  Blob b[ST2_NUM_BLOBS];
  double* scal = nullptr;
  double* gram_acc = nullptr;     // 512 x 512 doubles
  double* part = nullptr;         // per-block partial sums of the reducing kernels (block_accumulate_last)
  unsigned int* part_counter = nullptr;
  float* bwd = nullptr;           // d(scd)/d(data), fp32 NCHW
  float* data_sraw = nullptr;
  int order_n = 0;
  int order[ST2_NUM_BLOBS];
  float tv = 1.f, tv_power = 1.f, p = 1.f, p_power = 1.f;
  const float* x_cur = nullptr;
  int top_cur = -1;
};

static inline bool host_w_on(float w) { return fabsf(w) > 1e-15f; }

// Push this strip's first / last row of `slot` into the neighbours' halo rows, then wait for theirs.
// slot < ST2_NUM_BLOBS: activation of blob `slot` (0 = the padded x, circular for the TV term);
// otherwise the gradient of blob slot - ST2_NUM_BLOBS.
static int halo_exchange(st2_plan* pl, int slot) {
  st2_ctx* ctx = pl->ctx;
  if (!pl->peer[0] || !pl->peer[1]) return st2_fail(ctx, ST2_ERR_STATE, "strip plan: neighbours not attached");
  const bool is_x = (slot == 0);
  const int b = slot % ST2_NUM_BLOBS;
  const bool is_grad = slot >= ST2_NUM_BLOBS;
  const bool up = is_x || !pl->edge_top, dn = is_x || !pl->edge_bot;
  if (!up && !dn) return 0;
  const unsigned long long epoch = ++pl->epoch[slot];
  SlabHeader* mine = reinterpret_cast<SlabHeader*>(pl->slab);
  HaloPush a;
  memset(&a, 0, sizeof(a));
  a.epoch = epoch;
  a.counters = mine->counters;
  const StripLayout& L = pl->lay;
  for (int side = 0; side < 2; ++side) {
    if (!(side == 0 ? up : dn)) continue;
    const StripLayout& PL = pl->peer_lay[side];
    SlabHeader* theirs = reinterpret_cast<SlabHeader*>(pl->peer[side]);
    a.flag[side] = &theirs->flag_from[side == 0 ? 1 : 0][slot];     // I am its neighbour below (0) / above (1)
    if (is_x) {
      // planar: 3 segments of one row; their padded plane may have a different row count
      const long long W = pl->W;
      a.nseg = 3; a.seg_bytes = W * 4;
      a.src_seg_stride = (long long)(pl->H + 2) * W * 4;
      a.dst_seg_stride[side] = (long long)(PL.rows[0] + 2) * W * 4;
      const unsigned char* base = pl->slab + L.xp_off;
      a.src[side] = base + (side == 0 ? 1 : pl->H) * W * 4;         // first / last interior row of plane 0
      unsigned char* pbase = pl->peer[side] + PL.xp_off;
      a.dst[side] = pbase + (side == 0 ? (long long)(PL.rows[0] + 1) * W * 4 : 0);
    } else {
      const size_t rb = L.row_bytes[b];
      a.nseg = 1; a.seg_bytes = (long long)rb;
      const unsigned char* base = pl->slab + (is_grad ? L.grad_off[b] : L.act_off[b]);
      a.src[side] = base + (side == 0 ? (size_t)1 : (size_t)L.rows[b]) * rb;
      unsigned char* pbase = pl->peer[side] + (is_grad ? PL.grad_off[b] : PL.act_off[b]);
      a.dst[side] = pbase + (side == 0 ? (size_t)(PL.rows[b] + 1) * rb : 0);
    }
  }
  ProfScope ps(ctx, 9);
  long long vecs = a.seg_bytes / 16 * a.nseg;
  int blocks = (int)((vecs + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > 32) blocks = 32;
  halo_exchange_kernel<<<dim3(blocks, 2), 256, 0, ctx->stream>>>(a, up ? &mine->flag_from[0][slot] : nullptr,
                                                                 dn ? &mine->flag_from[1][slot] : nullptr, &mine->err);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

// One process per GPU (every neighbour attached through CUDA IPC): the exchange of `slot` is carried by the
// convolution kernel that consumes the tensor (HaloArgs, st2_kernels.h) instead of a launch of its own.  Several strips
// of one process on one GPU keep the exchange kernel: CTAs spinning inside a convolution would starve the other
// strips' kernels of SMs.
static bool halo_args(st2_plan* pl, int slot, HaloArgs* a) {
  memset(a, 0, sizeof(*a));
  if (!pl->strip || !pl->async_halo || slot == 0 || !pl->peer[0] || !pl->peer[1]) return false;
  const bool up = !pl->edge_top, dn = !pl->edge_bot;
  if (!up && !dn) return false;
  const int b = slot % ST2_NUM_BLOBS;
  const bool is_grad = slot >= ST2_NUM_BLOBS;
  SlabHeader* mine = reinterpret_cast<SlabHeader*>(pl->slab);
  const StripLayout& L = pl->lay;
  const size_t rb = L.row_bytes[b];
  if (rb % 16) return false;
  a->epoch = ++pl->epoch[slot];
  a->bytes = (long long)rb;
  a->counter = &mine->push_counter;
  a->err = &mine->err;
  a->push_blocks = (int)((rb / 16 + 255) / 256);
  if (a->push_blocks > 16) a->push_blocks = 16;
  if (a->push_blocks < 1) a->push_blocks = 1;
  const unsigned char* base = pl->slab + (is_grad ? L.grad_off[b] : L.act_off[b]);
  for (int side = 0; side < 2; ++side) {
    if (!(side == 0 ? up : dn)) continue;
    const StripLayout& PL = pl->peer_lay[side];
    SlabHeader* theirs = reinterpret_cast<SlabHeader*>(pl->peer[side]);
    a->flag[side] = &theirs->flag_from[side == 0 ? 1 : 0][slot];
    a->wait[side] = &mine->flag_from[side][slot];
    a->src[side] = base + (side == 0 ? (size_t)1 : (size_t)L.rows[b]) * rb;
    unsigned char* pbase = pl->peer[side] + (is_grad ? PL.grad_off[b] : PL.act_off[b]);
    a->dst[side] = pbase + (side == 0 ? (size_t)(PL.rows[b] + 1) * rb : 0);
  }
  return true;
}

template <typename T>
static int forward_impl(st2_plan* pl, const float* x, int top) {
  st2_ctx* ctx = pl->ctx;
  pl->b[0].act = const_cast<float*>(x);
  const int lo = (pl->strip && !pl->edge_top) ? 1 : 0, hi = (pl->strip && !pl->edge_bot) ? 1 : 0;
  if (pl->strip) {
    // padded copy of x with the neighbours' boundary rows (circular: the TV term wraps around the canvas)
    long long blocks = ((long long)3 * pl->H * pl->W + 255) / 256;
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    pack_x_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(x, pl->xp, pl->H, pl->W);
    ST2_LAUNCH_CHECK(ctx);
    int rc = halo_exchange(pl, 0);
    if (rc) return rc;
  }
  bool pool_fused = false;
  for (int i = 1; i <= top; ++i) {
    Blob& cur = pl->b[i];
    Blob& below = pl->b[i - 1];
    int rc = 0;
    const int fcat = g_blobs[i].kind != KIND_CONV ? 2 : (g_blobs[i].conv_index == 0 ? 1 : (pl->prec == ST2_PREC_FP32 ? 8 : 0));
    ProfScope ps(ctx, fcat);
    if (g_blobs[i].kind == KIND_CONV) {
      const int ci = g_blobs[i].conv_index;
      if (!ctx->w_oihw[ci]) return st2_fail(ctx, ST2_ERR_STATE, "weights of %s not loaded", g_blobs[i].name);
      if (ci == 0 && cur.tc_first) {
        if (pl->strip)
          rc = tc_first_fwd_launch(ctx, cur.tc_first, pl->xp + pl->W, (long long)(pl->H + 2) * pl->W, lo, hi, ctx->wh_first,
                                   ctx->bias[0], (__half*)cur.act);
        else
          rc = tc_first_fwd_launch(ctx, cur.tc_first, x, 0, 0, 0, ctx->wh_first, ctx->bias[0], (__half*)cur.act);
      } else if (ci == 0) {
        if (pl->strip)
          rc = launch_conv_first_fwd<T>(ctx, pl->xp + pl->W, ctx->wf32_fwd[0], ctx->bias[0], (T*)cur.act, cur.H, cur.W,
                                        (long long)(pl->H + 2) * pl->W, lo, hi);
        else
          rc = launch_conv_first_fwd<T>(ctx, x, ctx->wf32_fwd[0], ctx->bias[0], (T*)cur.act, cur.H, cur.W);
      } else if (pl->prec == ST2_PREC_FP32) {
        rc = launch_conv_exact(ctx, (const float*)below.act, ctx->wf32_fwd[ci], ctx->bias[ci], nullptr,
                               (float*)cur.act, cur.H, cur.W, below.C, cur.C, EPI_BIAS_RELU, lo, hi);
      } else {
        // a max-pool right above this convolution is computed by the same epilogue when the kernel supports it
        TcInject ti;
        ti.fc = nullptr; ti.sraw = nullptr; ti.coef = nullptr; ti.pool = nullptr; ti.pool_wp = 0; ti.sfuse = 0;
        if (i + 1 <= top && g_blobs[i + 1].kind == KIND_POOL) { ti.pool = (__half*)pl->b[i + 1].act; ti.pool_wp = pl->b[i + 1].W; }
        // row strips: this kernel also carries the exchange of its input's boundary rows (or the exchange kernel ran
        // after the producer, see the end of the loop body)
        HaloArgs ha;
        const bool carried = tc_conv_supports_halo(cur.tc_fwd) && halo_args(pl, i - 1, &ha);
        rc = tc_conv_launch(ctx, cur.tc_fwd, ctx->bias[ci], nullptr, (__half*)cur.act, EPI_BIAS_RELU, 1.f, nullptr, &ti,
                            &pool_fused, carried ? &ha : nullptr);
      }
    } else if (pool_fused) {
      pool_fused = false;                       // written by the convolution below
    } else {
      rc = launch_pool_fwd_v<T>(ctx, (const T*)below.act, (T*)cur.act, below.C, below.H, below.W);
    }
    if (rc) return rc;
    // the next convolution reads one row of the neighbouring strips
    if (pl->strip && i < top && g_blobs[i + 1].kind == KIND_CONV) {
      const bool in_kernel = pl->async_halo && pl->prec == ST2_PREC_FP16 && g_blobs[i + 1].conv_index > 0 &&
                             tc_conv_supports_halo(pl->b[i + 1].tc_fwd) && pl->lay.row_bytes[i] % 16 == 0;
      if (!in_kernel && (rc = halo_exchange(pl, i))) return rc;
    }
  }
  pl->x_cur = x;
  pl->top_cur = top;
  return 0;
}

// Gradient of sum_b <inj_b, blob_b> w.r.t. data with the reference's segment semantics
// (worker.py:88-106): gradient from above passes through reluX_Y (mask), the injected diff does not.
template <typename T>
static int backward_impl(st2_plan* pl, int top, const Inject* inj, float* grad_out) {
  st2_ctx* ctx = pl->ctx;
  if (top > pl->top_cur) return st2_fail(ctx, ST2_ERR_STATE, "backward above the last forward's top blob");
  const int lo = (pl->strip && !pl->edge_top) ? 1 : 0, hi = (pl->strip && !pl->edge_bot) ? 1 : 0;
  if (top == 0) {
    ST2_CUDA(ctx, cudaMemsetAsync(grad_out, 0, sizeof(float) * pl->b[0].n(), ctx->stream));
  }
  bool fused_inj = false;      // the injection of blob i was already applied by the data-gradient epilogue above it
  for (int i = top; i >= 1; --i) {
    Blob& cur = pl->b[i];
    Blob& below = pl->b[i - 1];
    const bool have_above = (i < top);
    const bool is_conv = g_blobs[i].kind == KIND_CONV;
    int rc = 0;
    if (inj[i].on && fused_inj) {
      fused_inj = false;
    } else if (inj[i].on) {
      CombineArgs a;
      a.gin = have_above ? cur.grad : nullptr;
      a.act = cur.act; a.fc = inj[i].fc; a.sraw = inj[i].sraw; a.out = cur.grad; a.n = cur.n();
      a.apply_mask = is_conv ? 1 : 0;
      a.coef = inj[i].coef; a.h_cc = inj[i].hcc; a.h_sc = inj[i].hsc; a.h_dc = inj[i].hdc;
      ProfScope ps(ctx, 5);
      rc = launch_combine_v<T>(ctx, a);
      if (rc) return rc;
    } else if (!have_above) {
      return st2_fail(ctx, ST2_ERR_STATE, "top blob carries no diff");
    }
    // does the producer of grad(below) apply below's ReLU mask itself?
    const bool below_conv = g_blobs[i - 1].kind == KIND_CONV;
    const int mask_below = (below_conv && !inj[i - 1].on) ? 1 : 0;
    const int bcat = !is_conv ? 2 : (g_blobs[i].conv_index == 0 ? 1 : (pl->prec == ST2_PREC_FP32 ? 8 : 0));
    // the data-gradient convolution reads one row of the neighbouring strips' gradient: carried by the kernel itself
    // where it can be, else by the exchange kernel
    HaloArgs ha;
    bool carried = false;
    if (pl->strip && is_conv) {
      const bool tc_consumer = pl->prec == ST2_PREC_FP16 && cur.tc_bwd != nullptr && tc_conv_supports_halo(cur.tc_bwd);
      carried = tc_consumer && halo_args(pl, ST2_NUM_BLOBS + i, &ha);
      if (!carried && (rc = halo_exchange(pl, ST2_NUM_BLOBS + i))) return rc;
    }
    const HaloArgs* hap = carried ? &ha : nullptr;
    ProfScope ps(ctx, bcat);
    if (is_conv) {
      const int ci = g_blobs[i].conv_index;
      if (ci == 0 && cur.tc_bwd) {
        // fold: ONE launch reads the gradient and the activations (two patches per tile, two accumulators)
        rc = inj[i].fold ? tc_conv_first_bwd_launch(ctx, cur.tc_sfold, grad_out, inj[i].coef, hap)
                         : tc_conv_first_bwd_launch(ctx, cur.tc_bwd, grad_out, nullptr, hap);
      } else if (ci == 0) {
        rc = launch_conv_first_bwd<T>(ctx, (const T*)cur.grad, ctx->wf32_bwd[0], grad_out, cur.H, cur.W, lo, hi);
      } else if (pl->prec == ST2_PREC_FP32) {
        rc = launch_conv_exact(ctx, (const float*)cur.grad, ctx->wf32_bwd[ci], nullptr, (const float*)below.act,
                               (float*)below.grad, cur.H, cur.W, cur.C, below.C, mask_below ? EPI_MASK : EPI_RAW, lo, hi);
      } else if (below_conv && inj[i - 1].on && inj[i - 1].coef != nullptr && !ctx->knobs.no_fused_inject) {
        TcInject ti;
        ti.fc = (const __half*)inj[i - 1].fc; ti.sraw = (const __half*)inj[i - 1].sraw; ti.coef = inj[i - 1].coef;
        ti.pool = nullptr; ti.pool_wp = 0;
        ti.sfuse = inj[i - 1].dfuse ? 1 : 0;      // the style gradient of the blob below is contracted by this kernel
        rc = tc_conv_launch(ctx, cur.tc_bwd, nullptr, (const __half*)below.act, (__half*)below.grad, EPI_MASK, 1.f,
                            nullptr, &ti, nullptr, hap);
        fused_inj = true;
      } else {
        rc = tc_conv_launch(ctx, cur.tc_bwd, nullptr, (const __half*)below.act, (__half*)below.grad,
                            mask_below ? EPI_MASK : EPI_RAW, 1.f, nullptr, nullptr, nullptr, hap);
      }
    } else {
      rc = launch_pool_bwd_v<T>(ctx, (const T*)below.act, (const T*)cur.grad, (T*)below.grad, below.C, below.H,
                              below.W, mask_below);
    }
    if (rc) return rc;
  }
  if (inj[0].on) {
    CombineArgs a;
    a.gin = grad_out; a.act = pl->b[0].act; a.fc = inj[0].fc; a.sraw = inj[0].sraw; a.out = grad_out;
    a.n = pl->b[0].n(); a.apply_mask = 0;
    a.coef = inj[0].coef; a.h_cc = inj[0].hcc; a.h_sc = inj[0].hsc; a.h_dc = inj[0].hdc;
    int rc = launch_combine_v<float>(ctx, a);
    if (rc) return rc;
  }
  return 0;
}

template <typename T>
static int gram_of_blob(st2_plan* pl, int blob, const float* A, float* D, double* sum_dsq) {
  st2_ctx* ctx = pl->ctx;
  Blob& B = pl->b[blob];
  const long long HW = (long long)B.H * B.W;
  if (blob != 0 && pl->prec == ST2_PREC_FP16 && g_blobs[blob].kind == KIND_CONV && B.C % 64 == 0 &&
      !ctx->knobs.no_tc_gram) {
    if (!B.tc_gram) {
      int rc0 = tc_gram_plan_create(ctx, (const __half*)B.act, B.C, HW, &B.tc_gram);
      if (rc0) return rc0;
    }
    return tc_gram_launch(ctx, B.tc_gram, A, D, sum_dsq);
  }
  ST2_CUDA(ctx, cudaMemsetAsync(pl->gram_acc, 0, sizeof(double) * B.C * B.C, ctx->stream));
  int rc;
  if (blob == 0) rc = launch_gram_generic<float>(ctx, (const float*)B.act, B.C, HW, 1, HW, pl->gram_acc);
  else rc = launch_gram_generic<T>(ctx, (const T*)B.act, B.C, HW, B.C, 1, pl->gram_acc);
  if (rc) return rc;
  return launch_gram_finalize(ctx, pl->gram_acc, A, D, B.C, HW, sum_dsq);
}

// row strips: un-normalised Gram sum of this strip's rows of `blob` -> out (C x C fp32)
template <typename T>
static int gram_sum_of_blob(st2_plan* pl, int blob, float* out) {
  st2_ctx* ctx = pl->ctx;
  Blob& B = pl->b[blob];
  const long long HW = (long long)B.H * B.W;
  if (blob != 0 && pl->prec == ST2_PREC_FP16 && g_blobs[blob].kind == KIND_CONV && B.C % 64 == 0 &&
      !ctx->knobs.no_tc_gram) {
    if (!B.tc_gram) {
      int rc0 = tc_gram_plan_create(ctx, (const __half*)B.act, B.C, HW, &B.tc_gram);
      if (rc0) return rc0;
    }
    return tc_gram_sum_launch(ctx, B.tc_gram, out);
  }
  ST2_CUDA(ctx, cudaMemsetAsync(pl->gram_acc, 0, sizeof(double) * B.C * B.C, ctx->stream));
  int rc;
  if (blob == 0) rc = launch_gram_generic<float>(ctx, (const float*)B.act, B.C, HW, 1, HW, pl->gram_acc);
  else rc = launch_gram_generic<T>(ctx, (const T*)B.act, B.C, HW, B.C, 1, pl->gram_acc);
  if (rc) return rc;
  return launch_gram_acc_to_f32(ctx, pl->gram_acc, out, B.C);
}

static int ensure(st2_ctx* ctx, void** p, size_t bytes) {
  if (*p) return 0;
  ST2_CUDA(ctx, cudaMalloc(p, bytes));
  return 0;
}

// The style layer whose gradient s = sc (D F) is folded into the weights of its own data-gradient convolution instead
// of being materialised (style_fold_kernel): conv1_1 on the fp16 path.  (Tried for conv2_1 as a second, accumulating
// pass of its 128 -> 64 data-gradient kernel: that pass cost 82 us at 1024^2 against the 41 + 41 us it saved.)
// On row strips the choice must be the SAME on every strip: with the fold, grad(conv1_1) does not contain the style
// term, and a strip reads one row of its neighbours' grad(conv1_1) -- so the caller, who knows all strips, says
// whether every strip can fold (st2_strip_set_fold; a strip shorter than 16 rows cannot run the tensor-core conv1_1
// kernels).
static bool fold_eligible(const st2_plan* pl, int b) {
  return b == 1 && pl->prec == ST2_PREC_FP16 && !pl->ctx->knobs.no_style_fuse && pl->b[b].tc_bwd != nullptr &&
         (!pl->strip || pl->strip_fold);
}
// conv2_1 (128 channels): its style gradient D' F is contracted inside the data-gradient kernel of conv2_2, the layer
// above, as one extra pipeline stage per tile into a second accumulator (TcInject::sfuse) -- when that kernel is the
// 128-wide CTA-pair kernel and the backward pass starts above conv2_1.  grad(conv2_1) holds the same values either
// way, so strips may decide differently.
static bool dfuse_eligible(const st2_plan* pl, int b, int want_grad) {
  return b == 4 && want_grad && pl->prec == ST2_PREC_FP16 && !pl->ctx->knobs.no_style_fuse && pl->eval_top >= b + 1 &&
         !pl->ctx->knobs.no_fused_inject && tc_conv_supports_style_fuse(pl->b[b + 1].tc_bwd);
}

// The objective (worker.py:231-301) in four phases.  On a whole-canvas plan st2_eval runs them back to
// back.  On a row strip the caller all-reduces (sum) one small block between consecutive phases:
//   begin: forward; local feature sums; the strip's Gram sums            -> all-reduce gram_red (fp32)
//   mid:   D = G - A from the global Gram; style gradient D F of the strip -> all-reduce red (3 sums / blob)
//   end:   normalisers + coefficients; backward; pixel terms              -> all-reduce 6 pixel-space sums
//   final: totals in the reference's accumulation order
template <typename T>
static int eval_begin_impl(st2_plan* pl, const float* x, int want_grad) {
  st2_ctx* ctx = pl->ctx;
  EvalSpec& es = pl->es;
  memset(&es, 0, sizeof(es));
  int top = 0;
  for (int k = 0; k < pl->order_n; ++k) {
    const int b = pl->order[k];
    const Blob& B = pl->b[b];
    if (!(host_w_on(B.cw) || host_w_on(B.sw) || host_w_on(B.dw))) continue;
    es.order[es.n++] = b;
    if (b > top) top = b;
  }
  for (int b = 0; b < ST2_NUM_BLOBS; ++b) {
    es.cw[b] = pl->b[b].cw; es.sw[b] = pl->b[b].sw; es.dw[b] = pl->b[b].dw;
    es.C[b] = pl->b[b].C; es.nelem[b] = pl->b[b].n_total();
  }
  es.tv = pl->tv; es.tv_power = pl->tv_power; es.p = pl->p; es.p_power = pl->p_power;
  es.N = pl->b[0].n_total();
  pl->eval_top = top; pl->eval_want_grad = want_grad; pl->eval_x = x;

  st2_launch_pdl(ctx, true, clear_volatile_kernel, 1, 256, 0, pl->scal);
  ST2_LAUNCH_CHECK(ctx);
  if (pl->strip) ST2_CUDA(ctx, cudaMemsetAsync(pl->red, 0, sizeof(double) * 3 * ST2_NUM_BLOBS, ctx->stream));
  int rc = forward_impl<T>(pl, x, top);
  if (rc) return rc;

  pl->gram_red_used = 0;
  for (int b = 0; b < ST2_NUM_BLOBS; ++b) pl->inj[b] = Inject();
  // tensor-core Grams: the split-K contractions are launched per layer, ONE launch then finishes them all
  TcGramPlan* gp[8]; const float* gA[8]; float* gD[8]; double* gS[8];
  int n_gram = 0;
  for (int k = 0; k < es.n; ++k) {
    const int b = es.order[k];
    Blob& B = pl->b[b];
    double* sb = pl->scal + b * ST2_SCAL_PER_BLOB;
    double* c_sum = pl->strip ? pl->red + 3 * b + 0 : sb + SB_C_SUMSQ;
    double* d_sum = pl->strip ? pl->red + 3 * b + 1 : sb + SB_D_SUMSQ;
    const bool c_on = host_w_on(B.cw), s_on = host_w_on(B.sw), d_on = host_w_on(B.dw);
    if (c_on && !B.fc) return st2_fail(ctx, ST2_ERR_STATE, "content weight on %s but no content target", g_blobs[b].name);
    if (s_on && !B.gram_target) return st2_fail(ctx, ST2_ERR_STATE, "style weight on %s but no style target", g_blobs[b].name);
    if (c_on || d_on) {
      ProfScope ps(ctx, 5);
      if (b == 0) rc = launch_feature_sums_v<float>(ctx, (const float*)B.act, c_on ? (const float*)B.fc : nullptr, B.n(),
                                                  c_sum, d_sum);
      else rc = launch_feature_sums_v<T>(ctx, (const T*)B.act, c_on ? (const T*)B.fc : nullptr, B.n(), c_sum, d_sum);
      if (rc) return rc;
    }
    if (s_on) {
      const size_t e = (b == 0) ? sizeof(float) : pl->esz;
      if ((rc = ensure(ctx, (void**)&B.D, sizeof(float) * B.C * B.C))) return rc;
      if (!fold_eligible(pl, b) && !dfuse_eligible(pl, b, want_grad) && (rc = ensure(ctx, &B.sraw, e * B.n()))) return rc;
      ProfScope ps(ctx, 3);
      if (pl->strip) {
        pl->gram_red_off[b] = pl->gram_red_used;
        pl->gram_red_used += (long long)B.C * B.C;
      }
      const bool tc = b != 0 && pl->prec == ST2_PREC_FP16 && g_blobs[b].kind == KIND_CONV && B.C % 64 == 0 &&
                      !ctx->knobs.no_tc_gram && n_gram < 8;
      if (tc) {
        if (!B.tc_gram && (rc = tc_gram_plan_create(ctx, (const __half*)B.act, B.C, (long long)B.H * B.W, &B.tc_gram)))
          return rc;
        if ((rc = tc_gram_mma_launch(ctx, B.tc_gram))) return rc;
        gp[n_gram] = B.tc_gram;
        gA[n_gram] = pl->strip ? nullptr : B.gram_target;
        gD[n_gram] = pl->strip ? pl->gram_red + pl->gram_red_off[b] : B.D;
        gS[n_gram] = pl->strip ? nullptr : sb + SB_S_GRAMSQ;
        ++n_gram;
      } else if (pl->strip) {
        if ((rc = gram_sum_of_blob<T>(pl, b, pl->gram_red + pl->gram_red_off[b]))) return rc;
      } else {
        if ((rc = gram_of_blob<T>(pl, b, B.gram_target, B.D, sb + SB_S_GRAMSQ))) return rc;
      }
    }
    pl->inj[b].on = true;
    pl->inj[b].fc = c_on ? B.fc : nullptr;
    pl->inj[b].sraw = s_on ? B.sraw : nullptr;
    pl->inj[b].coef = sb + SB_C_COEF;          // [C_COEF, S_COEF, D_COEF] are consecutive
  }
  if (n_gram) {
    ProfScope ps(ctx, 10);
    if ((rc = tc_gram_finalize_all(ctx, n_gram, gp, gA, gD, gS, pl->strip ? 1 : 0))) return rc;
  }
  if (pl->strip) {
    for (int b = 1; b <= 4; b += 3) {            // conv1_1 and conv2_1
      if (!host_w_on(pl->b[b].sw) || !(b == 1 ? fold_eligible(pl, b) : dfuse_eligible(pl, b, want_grad))) continue;
      const size_t bytes = sizeof(float) * pl->b[b].C * pl->b[b].C;
      if ((rc = ensure(ctx, (void**)&pl->gram_local[b], bytes))) return rc;
      ST2_CUDA(ctx, cudaMemcpyAsync(pl->gram_local[b], pl->gram_red + pl->gram_red_off[b], bytes, cudaMemcpyDeviceToDevice,
                                    ctx->stream));
    }
  }
  pl->eval_phase = 1;
  return 0;
}

template <typename T>
static int eval_mid_impl(st2_plan* pl) {
  st2_ctx* ctx = pl->ctx;
  if (pl->eval_phase != 1) return st2_fail(ctx, ST2_ERR_STATE, "st2_eval_mid: call st2_eval_begin first");
  const EvalSpec& es = pl->es;
  const int want_grad = pl->eval_want_grad;
  int rc = 0;
  for (int k = 0; k < es.n; ++k) {
    const int b = es.order[k];
    Blob& B = pl->b[b];
    if (!host_w_on(B.sw)) continue;
    double* sb = pl->scal + b * ST2_SCAL_PER_BLOB;
    double* raw_sum = pl->strip ? pl->red + 3 * b + 2 : sb + SB_S_RAWSQ;
    const long long HW = (long long)B.H * B.W;
    if (pl->strip) {
      ProfScope ps(ctx, 3);
      if ((rc = launch_gram_from_sum(ctx, pl->gram_red + pl->gram_red_off[b], B.gram_target, B.D, B.C,
                                     (double)B.Hg * B.W, sb + SB_S_GRAMSQ)))
        return rc;
    }
    ProfScope ps(ctx, 4);
    if (b == 1 && fold_eligible(pl, b)) {
      const bool stencil = !ctx->knobs.no_stencil;
      if (!B.wfold) {
        if ((rc = ensure(ctx, (void**)&B.wfold, sizeof(__half) * 16 * 9 * 128))) return rc;
        ST2_CUDA(ctx, cudaMemsetAsync(B.wfold, 0, sizeof(__half) * 16 * 9 * 128, ctx->stream));
        if (stencil) stencil_pack_kernel<<<cdiv(27 * 64, 256), 256, 0, ctx->stream>>>(ctx->wh_bwd[0], B.wfold);
        else dual_pack_kernel<<<cdiv(16 * 9 * 64, 256), 256, 0, ctx->stream>>>(ctx->wh_bwd[0], B.wfold);
        ST2_LAUNCH_CHECK(ctx);
      }
      if (!B.tc_sfold) {
        const __half* gsrc = (const __half*)(pl->strip ? B.grad_pad : B.grad);
        const __half* asrc = (const __half*)(pl->strip ? B.act_pad : B.act);
        rc = stencil ? tc_conv_stencil_plan_create(ctx, gsrc, asrc, B.wfold, B.H, B.W, &B.tc_sfold, pl->strip ? 1 : 0)
                     : tc_conv_dual_plan_create(ctx, gsrc, asrc, B.wfold, B.H, B.W, &B.tc_sfold, pl->strip ? 1 : 0);
        if (rc) return rc;
      }
      st2_launch_pdl(ctx, true, style_fold_kernel, 32, 128, 0, B.D, B.gram_target, pl->strip ? pl->gram_local[1] : nullptr,
                     ctx->w_oihw[0], B.wfold, stencil ? 1 : 0, B.n_total(), sb, raw_sum);
      ST2_LAUNCH_CHECK(ctx);
      pl->inj[b].sraw = nullptr;
      pl->inj[b].fold = true;
      continue;
    }
    if (dfuse_eligible(pl, b, want_grad)) {
      if ((rc = ensure(ctx, (void**)&B.Dh, sizeof(__half) * B.C * B.C))) return rc;
      TcConvPlan* above = pl->b[b + 1].tc_bwd;
      if (!B.dfuse_set) {
        if ((rc = tc_conv_set_style_fuse(ctx, above, (const __half*)(pl->strip ? B.act_pad : B.act), B.Dh))) return rc;
        B.dfuse_set = true;
      }
      st2_launch_pdl(ctx, true, style_scale_kernel, cdiv((long long)B.C * B.C, 256), 256, 0, B.D, B.Dh, B.C, sb);
      ST2_LAUNCH_CHECK(ctx);
      st2_launch_pdl(ctx, true, style_rawsq_kernel, cdiv((long long)B.C * B.C, 256), 256, 0, B.D, B.gram_target,
                     pl->strip ? pl->gram_local[b] : nullptr, B.C, B.n_total(), sb, raw_sum);
      ST2_LAUNCH_CHECK(ctx);
      pl->inj[b].sraw = nullptr;
      pl->inj[b].dfuse = true;
      continue;
    }
    if (want_grad) {
      if (b == 0) {
        rc = launch_style_grad_generic<float>(ctx, (const float*)B.act, B.D, (float*)B.sraw, B.C, HW, 1, HW, raw_sum);
      } else if (pl->prec == ST2_PREC_FP16 && g_blobs[b].kind == KIND_CONV && B.C % 64 == 0) {
        if ((rc = ensure(ctx, (void**)&B.Dh, sizeof(__half) * B.C * B.C))) return rc;
        if (!B.tc_style &&
            (rc = tc_conv_plan_create(ctx, (const __half*)B.act, B.Dh, B.H, B.W, B.C, B.C, 1, &B.tc_style)))
          return rc;
        st2_launch_pdl(ctx, true, style_scale_kernel, cdiv((long long)B.C * B.C, 256), 256, 0, B.D, B.Dh, B.C, sb);
        ST2_LAUNCH_CHECK(ctx);
        rc = tc_conv_launch(ctx, B.tc_style, nullptr, nullptr, (__half*)B.sraw, EPI_RAW, 1.f, raw_sum);
      } else {
        rc = launch_style_grad_generic<T>(ctx, (const T*)B.act, B.D, (T*)B.sraw, B.C, HW, B.C, 1, raw_sum);
      }
    } else {
      // loss-only evaluation still freezes the style normaliser (worker.py:265-266), which needs |D F|
      if (b == 0)
        rc = launch_style_grad_generic<float>(ctx, (const float*)B.act, B.D, (float*)B.sraw, B.C, HW, 1, HW, raw_sum);
      else
        rc = launch_style_grad_generic<T>(ctx, (const T*)B.act, B.D, (T*)B.sraw, B.C, HW, B.C, 1, raw_sum);
    }
    if (rc) return rc;
  }
  pl->eval_phase = 2;
  return 0;
}

template <typename T>
static int eval_end_impl(st2_plan* pl, float* grad_out) {
  st2_ctx* ctx = pl->ctx;
  if (pl->eval_phase != 2) return st2_fail(ctx, ST2_ERR_STATE, "st2_eval_end: call st2_eval_mid first");
  const EvalSpec& es = pl->es;
  const float* x = pl->eval_x;
  int rc = 0;
  // deferred sums (steady state on strips): the per-blob sums are all-reduced AFTER the backward pass, together with
  // the pixel sums (and the caller's L-BFGS dot products); the coefficients need only the frozen normalisers
  pl->eval_deferred = pl->strip && pl->deferred_sums;
  if (pl->strip && !pl->eval_deferred) {
    scatter_sums_kernel<<<1, 32, 0, ctx->stream>>>(pl->red, pl->scal);
    ST2_LAUNCH_CHECK(ctx);
  }
  if (es.n > 0) {
    ProfScope ps(ctx, 5);
    st2_launch_pdl(ctx, true, coef_kernel, 1, 32, 0, es, pl->scal, pl->eval_deferred ? 1 : 0);
    ST2_LAUNCH_CHECK(ctx);
  }
  double* gscal = pl->scal + ST2_SCAL_GLOBAL_BASE;
  float* bwd = nullptr;
  if (pl->eval_want_grad) {
    if (!grad_out) return st2_fail(ctx, ST2_ERR_ARG, "st2_eval: grad_dev is null");
    if (es.n > 0) {
      if ((rc = backward_impl<T>(pl, pl->eval_top, pl->inj, pl->bwd))) return rc;
    } else {
      ST2_CUDA(ctx, cudaMemsetAsync(pl->bwd, 0, sizeof(float) * pl->b[0].n(), ctx->stream));
    }
    bwd = pl->bwd;
  } else {
    grad_out = nullptr;
  }
  {
    ProfScope ps(ctx, 6);
    if (pl->strip)
      rc = pixel_terms_strip(ctx, pl->xp + pl->W, (long long)(pl->H + 2) * pl->W, 0, bwd, grad_out, 3, pl->H, pl->W,
                             pl->tv, pl->tv_power, pl->p, pl->p_power, 255.0f, gscal, pl->part, pl->part_counter);
    else
      rc = pixel_terms_strip(ctx, x, (long long)pl->H * pl->W, 1, bwd, grad_out, 3, pl->H, pl->W, pl->tv, pl->tv_power,
                             pl->p, pl->p_power, 255.0f, gscal, pl->part, pl->part_counter);
  }
  if (rc) return rc;
  pl->eval_phase = 3;
  return 0;
}

static int eval_final_impl(st2_plan* pl) {
  st2_ctx* ctx = pl->ctx;
  if (pl->eval_phase != 3) return st2_fail(ctx, ST2_ERR_STATE, "st2_eval_final: call st2_eval_end first");
  if (pl->eval_deferred) {
    scatter_sums_kernel<<<1, 32, 0, ctx->stream>>>(pl->red, pl->scal);
    ST2_LAUNCH_CHECK(ctx);
    if (pl->es.n > 0) {
      st2_launch_pdl(ctx, true, coef_kernel, 1, 32, 0, pl->es, pl->scal, 2);
      ST2_LAUNCH_CHECK(ctx);
    }
  }
  st2_launch_pdl(ctx, true, final_kernel, 1, 32, 0, pl->es, pl->scal,
                 pl->strip ? &reinterpret_cast<SlabHeader*>(pl->slab)->err : nullptr);
  ST2_LAUNCH_CHECK(ctx);
  pl->eval_phase = 0;
  return 0;
}

// =============================================================================================
extern "C" {

int st2_ctx_create(int device, st2_ctx** out) {
  if (!out) return ST2_ERR_ARG;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return st2_fail(nullptr, ST2_ERR_CUDA, "no CUDA device (%s); libst2 has no CPU fallback",
                    cudaGetErrorString(e));
  if (device < 0) device = 0;                    // config.ini `gpu = -1` meant CPU in the reference
  if (device >= count) return st2_fail(nullptr, ST2_ERR_ARG, "device %d out of range (%d present)", device, count);
  cudaDeviceProp prop;
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return st2_fail(nullptr, ST2_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  if (prop.major != 10)
    return st2_fail(nullptr, ST2_ERR_CUDA, "device %d is sm_%d%d; libst2 is built for sm_100a only", device,
                    prop.major, prop.minor);
  for (const void* fn : st2_kernel_registry()) {          // load every kernel now (see st2_common.cuh)
    cudaFuncAttributes attr;
    if ((e = cudaFuncGetAttributes(&attr, fn)) != cudaSuccess)
      return st2_fail(nullptr, ST2_ERR_CUDA, "loading the sm_100a kernels failed: %s", cudaGetErrorString(e));
  }
  for (const St2SmemOptIn& k : st2_smem_registry()) {
    if ((e = cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, k.bytes)) != cudaSuccess)
      return st2_fail(nullptr, ST2_ERR_CUDA, "shared-memory opt-in (%d bytes) failed on device %d: %s", k.bytes, device,
                      cudaGetErrorString(e));
  }
  st2_ctx* ctx = new st2_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  {
    st2_ctx::Knobs& k = ctx->knobs;
    k.no_fused_inject = getenv("ST2_NO_FUSED_INJECT") != nullptr;
    k.no_tc_gram = getenv("ST2_NO_TC_GRAM") != nullptr;
    k.no_tc_first = getenv("ST2_NO_TC_FIRST") != nullptr;
    k.no_ws = getenv("ST2_NO_WS") != nullptr;
    k.force_pair = getenv("ST2_FORCE_PAIR") != nullptr;
    k.wsp = getenv("ST2_WSP") != nullptr;
    k.no_pair = getenv("ST2_NO_PAIR") != nullptr;
    k.no_pool_fusion = getenv("ST2_NO_POOL_FUSION") != nullptr;
    k.no_style_fuse = getenv("ST2_NO_STYLE_FUSE") != nullptr;
    k.no_graph = getenv("ST2_NO_GRAPH") != nullptr;
    k.no_inkernel_halo = getenv("ST2_NO_INKERNEL_HALO") != nullptr;
    k.no_stencil = getenv("ST2_NO_STENCIL") != nullptr;
    k.no_ws128 = getenv("ST2_NO_WS128") != nullptr;
    k.no_pdl = getenv("ST2_NO_PDL") != nullptr;
    if (const char* v = getenv("ST2_TC_BN")) k.tc_bn = atoi(v);
    if (const char* v = getenv("ST2_PAIR_MIN_TILES")) k.pair_min_tiles = atoll(v);
  }
  if ((e = cudaMalloc(&ctx->dot_scratch, sizeof(double))) != cudaSuccess) {
    delete ctx;
    return st2_fail(nullptr, ST2_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(e));
  }
  *out = ctx;
  return 0;
}

void st2_ctx_destroy(st2_ctx* ctx) {
  if (!ctx) return;
  for (int i = 0; i < ST2_NUM_CONVS; ++i) {
    cudaFree(ctx->w_oihw[i]); cudaFree(ctx->bias[i]); cudaFree(ctx->wf32_fwd[i]); cudaFree(ctx->wf32_bwd[i]);
    cudaFree(ctx->wh_fwd[i]); cudaFree(ctx->wh_bwd[i]);
  }
  cudaFree(ctx->wh_first);
  cudaFree(ctx->wh_bwd_all);
  cudaFree(ctx->dot_scratch);
  delete ctx;
}

const char* st2_last_error(st2_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int st2_set_stream(st2_ctx* ctx, void* s) {
  if (!ctx) return ST2_ERR_ARG;
  ctx->stream = (cudaStream_t)s;
  return 0;
}

long long st2_launch_count(st2_ctx* ctx) { return ctx ? ctx->launches : 0; }

int st2_debug_flags(st2_ctx* ctx, int flags) {
  if (!ctx) return ST2_ERR_ARG;
  ctx->debug_flags = flags;
  return 0;
}

int st2_profile(st2_ctx* ctx, int enable) {
  if (!ctx) return ST2_ERR_ARG;
  ctx->prof_on = enable != 0;
  return 0;
}

int st2_profile_read(st2_ctx* ctx, double* ms_out, long long* count_out) {
  if (!ctx || !ms_out || !count_out) return ST2_ERR_ARG;
  ST2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < ST2_PROF_CATS; ++i) { ms_out[i] = 0.0; count_out[i] = 0; }
  for (auto& sp : ctx->prof_spans) {
    float ms = 0.f;
    if (sp.b && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess && sp.cat >= 0 && sp.cat < ST2_PROF_CATS) {
      ms_out[sp.cat] += ms;
      count_out[sp.cat] += 1;
    }
    ctx->prof_pool.push_back(sp.a);
    if (sp.b) ctx->prof_pool.push_back(sp.b);
  }
  ctx->prof_spans.clear();
  return 0;
}

int st2_blob_count(void) { return ST2_NUM_BLOBS; }
const char* st2_blob_name(int b) { return (b >= 0 && b < ST2_NUM_BLOBS) ? g_blobs[b].name : nullptr; }
int st2_blob_channels(int b) { return (b >= 0 && b < ST2_NUM_BLOBS) ? g_blobs[b].channels : ST2_ERR_ARG; }
int st2_blob_kind(int b) { return (b >= 0 && b < ST2_NUM_BLOBS) ? g_blobs[b].kind : ST2_ERR_ARG; }

int st2_set_conv_weights(st2_ctx* ctx, int ci, const float* w, const float* b, int cout, int cin) {
  if (!ctx || ci < 0 || ci >= ST2_NUM_CONVS || !w || !b) return st2_fail(ctx, ST2_ERR_ARG, "st2_set_conv_weights: bad arguments");
  int blob = -1;
  for (int i = 0; i < ST2_NUM_BLOBS; ++i) if (g_blobs[i].conv_index == ci) blob = i;
  const int want_cout = g_blobs[blob].channels, want_cin = g_blobs[blob - 1].channels;
  if (cout != want_cout || cin != want_cin)
    return st2_fail(ctx, ST2_ERR_ARG, "%s expects (%d,%d,3,3) weights, got (%d,%d,3,3)", g_blobs[blob].name,
                    want_cout, want_cin, cout, cin);
  const size_t nw = (size_t)cout * cin * 9;
  int rc;
  if ((rc = ensure(ctx, (void**)&ctx->w_oihw[ci], nw * 4))) return rc;
  if ((rc = ensure(ctx, (void**)&ctx->bias[ci], (size_t)cout * 4))) return rc;
  if ((rc = ensure(ctx, (void**)&ctx->wf32_fwd[ci], nw * 4))) return rc;
  if ((rc = ensure(ctx, (void**)&ctx->wf32_bwd[ci], nw * 4))) return rc;
  if (ci > 0) {
    if ((rc = ensure(ctx, (void**)&ctx->wh_fwd[ci], nw * 2))) return rc;
    if ((rc = ensure(ctx, (void**)&ctx->wh_bwd[ci], nw * 2))) return rc;
  } else {
    // conv1_1 data gradient on the tensor cores: N = 3 image planes padded to 16 rows of zeros
    if ((rc = ensure(ctx, (void**)&ctx->wh_bwd[0], (size_t)16 * 9 * cout * 2))) return rc;
    ST2_CUDA(ctx, cudaMemsetAsync(ctx->wh_bwd[0], 0, (size_t)16 * 9 * cout * 2, ctx->stream));
    if ((rc = ensure(ctx, (void**)&ctx->wh_first, (size_t)2 * 3 * 2 * 2 * 64 * 8 * 2))) return rc;
  }
  ctx->cin[ci] = cin; ctx->cout[ci] = cout;
  ST2_CUDA(ctx, cudaMemcpyAsync(ctx->w_oihw[ci], w, nw * 4, cudaMemcpyHostToDevice, ctx->stream));
  ST2_CUDA(ctx, cudaMemcpyAsync(ctx->bias[ci], b, (size_t)cout * 4, cudaMemcpyHostToDevice, ctx->stream));
  pack_weights_kernel<<<cdiv((long long)nw, 256), 256, 0, ctx->stream>>>(ctx->w_oihw[ci], cout, cin, ctx->wf32_fwd[ci],
                                                                       ctx->wf32_bwd[ci], ctx->wh_fwd[ci], ctx->wh_bwd[ci]);
  ST2_LAUNCH_CHECK(ctx);
  if (ci == 0 && (rc = tc_first_pack_weights(ctx, ctx->w_oihw[0], ctx->wh_first))) return rc;
  if (ci == 0) {
    if ((rc = ensure(ctx, (void**)&ctx->wh_bwd_all, sizeof(__half) * 32 * 64))) return rc;
    ST2_CUDA(ctx, cudaMemsetAsync(ctx->wh_bwd_all, 0, sizeof(__half) * 32 * 64, ctx->stream));
    stencil_pack_kernel<<<cdiv(27 * 64, 256), 256, 0, ctx->stream>>>(ctx->wh_bwd[0], ctx->wh_bwd_all);
    ST2_LAUNCH_CHECK(ctx);
  }
  ST2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // host buffers may be freed by the caller
  return 0;
}

static int plan_create_common(st2_ctx* ctx, int H, int W, int prec, bool strip, int rank, int world, int row0,
                              int H_total, st2_plan** out) {
  st2_plan* pl = new st2_plan();
  pl->ctx = ctx; pl->H = H; pl->W = W; pl->prec = prec;
  pl->esz = prec == ST2_PREC_FP16 ? 2 : 4;
  pl->strip = strip; pl->rank = rank; pl->world = world; pl->row0 = row0; pl->H_total = H_total;
  pl->edge_top = (rank == 0); pl->edge_bot = (rank == world - 1);
  if (strip) {
    pl->lay = strip_layout(H, W, pl->esz);
    ST2_CUDA(ctx, cudaMalloc(&pl->slab, pl->lay.total));
    // halo rows at the canvas edges are never written: they stay zero = the convolutions' zero pad
    ST2_CUDA(ctx, cudaMemsetAsync(pl->slab, 0, pl->lay.total, ctx->stream));
    pl->xp = reinterpret_cast<float*>(pl->slab + pl->lay.xp_off);
    ST2_CUDA(ctx, cudaMalloc(&pl->red, sizeof(double) * 3 * ST2_NUM_BLOBS));
    long long gtot = 0;
    for (int i = 0; i < ST2_NUM_BLOBS; ++i) gtot += (long long)g_blobs[i].channels * g_blobs[i].channels;
    ST2_CUDA(ctx, cudaMalloc(&pl->gram_red, sizeof(float) * gtot));
  }
  int h = H, w = W, hg = H_total;
  for (int i = 0; i < ST2_NUM_BLOBS; ++i) {
    if (g_blobs[i].kind == KIND_POOL) { h = pool_extent(h); w = pool_extent(w); hg = pool_extent(hg); }
    Blob& B = pl->b[i];
    B.C = g_blobs[i].channels; B.H = h; B.W = w; B.Hg = hg;
    pl->order[i] = i;
    if (i == 0) continue;
    if (strip) {
      B.act_pad = pl->slab + pl->lay.act_off[i];
      B.grad_pad = pl->slab + pl->lay.grad_off[i];
      B.act = (unsigned char*)B.act_pad + pl->lay.row_bytes[i];
      B.grad = (unsigned char*)B.grad_pad + pl->lay.row_bytes[i];
    } else {
      ST2_CUDA(ctx, cudaMalloc(&B.act, pl->esz * B.n()));
      ST2_CUDA(ctx, cudaMalloc(&B.grad, pl->esz * B.n()));
    }
  }
  pl->order_n = ST2_NUM_BLOBS;
  ST2_CUDA(ctx, cudaMalloc(&pl->scal, sizeof(double) * ST2_SCAL_TOTAL));
  ST2_CUDA(ctx, cudaMemsetAsync(pl->scal, 0, sizeof(double) * ST2_SCAL_TOTAL, ctx->stream));
  ST2_CUDA(ctx, cudaMalloc(&pl->gram_acc, sizeof(double) * 512 * 512));
  ST2_CUDA(ctx, cudaMalloc(&pl->part, sizeof(double) * ST2_PART_BLOCKS * 8));
  ST2_CUDA(ctx, cudaMalloc(&pl->part_counter, sizeof(unsigned int) * 4));
  ST2_CUDA(ctx, cudaMemsetAsync(pl->part_counter, 0, sizeof(unsigned int) * 4, ctx->stream));
  ST2_CUDA(ctx, cudaMalloc(&pl->bwd, sizeof(float) * pl->b[0].n()));
  if (prec == ST2_PREC_FP16) {
    const int halo = strip ? 1 : 0;
    if (ctx->wh_bwd[0] && pl->b[1].H >= 16 && pl->b[1].W >= 16 && !ctx->knobs.no_tc_first) {
      Blob& c11 = pl->b[1];
      int rc = ctx->wh_bwd_all && !ctx->knobs.no_stencil
                   ? tc_conv_stencil_plan_create(ctx, (const __half*)(strip ? c11.grad_pad : c11.grad), nullptr,
                                                 ctx->wh_bwd_all, c11.H, c11.W, &c11.tc_bwd, halo)
                   : tc_conv_plan_create(ctx, (const __half*)(strip ? c11.grad_pad : c11.grad), ctx->wh_bwd[0], c11.H,
                                         c11.W, 64, 16, 9, &c11.tc_bwd, halo);
      if (rc) return rc;
      if (ctx->wh_first && (rc = tc_first_plan_create(ctx, c11.H, c11.W, halo, &c11.tc_first))) return rc;
    }
    for (int i = 2; i < ST2_NUM_BLOBS; ++i) {
      if (g_blobs[i].kind != KIND_CONV) continue;
      const int ci = g_blobs[i].conv_index;
      if (!ctx->wh_fwd[ci]) {
        if (ci > 0 && ctx->wh_fwd[1] == nullptr) return st2_fail(ctx, ST2_ERR_STATE, "load weights before creating an fp16 plan");
        continue;        // a network cut below this layer: no weights, never evaluated (st2_forward reports it)
      }
      Blob& cur = pl->b[i];
      Blob& below = pl->b[i - 1];
      int rc = tc_conv_plan_create(ctx, (const __half*)(strip ? below.act_pad : below.act), ctx->wh_fwd[ci], cur.H, cur.W,
                                   below.C, cur.C, 9, &cur.tc_fwd, halo);
      if (rc) return rc;
      rc = tc_conv_plan_create(ctx, (const __half*)(strip ? cur.grad_pad : cur.grad), ctx->wh_bwd[ci], cur.H, cur.W, cur.C,
                               below.C, 9, &cur.tc_bwd, halo);
      if (rc) return rc;
    }
  }
  *out = pl;
  return 0;
}

int st2_plan_create(st2_ctx* ctx, int H, int W, int prec, st2_plan** out) {
  if (!ctx || !out || H < 1 || W < 1 || (prec != ST2_PREC_FP32 && prec != ST2_PREC_FP16))
    return st2_fail(ctx, ST2_ERR_ARG, "st2_plan_create: bad arguments");
  return plan_create_common(ctx, H, W, prec, false, 0, 1, 0, H, out);
}

int st2_strip_plan_create(st2_ctx* ctx, int H_total, int W, int row0, int row1, int rank, int world, int prec,
                          st2_plan** out) {
  if (!ctx || !out || H_total < 1 || W < 1 || world < 1 || rank < 0 || rank >= world || row0 < 0 || row1 <= row0 ||
      row1 > H_total || (prec != ST2_PREC_FP32 && prec != ST2_PREC_FP16))
    return st2_fail(ctx, ST2_ERR_ARG, "st2_strip_plan_create: bad arguments");
  // strips start on multiples of 32 rows so that all FIVE 2x2/2 pools (pool1 .. pool5) keep their windows inside
  // a strip: the pool kernels are strip-local, a window straddling a boundary would be paired wrongly
  if (row0 % 32 || (rank != world - 1 && row1 % 32) || (rank == 0 && row0 != 0) || (rank == world - 1 && row1 != H_total))
    return st2_fail(ctx, ST2_ERR_ARG, "st2_strip_plan_create: strip [%d, %d) of %d rows is not 32-row aligned", row0, row1,
                    H_total);
  return plan_create_common(ctx, row1 - row0, W, prec, true, rank, world, row0, H_total, out);
}

int st2_strip_ipc_handle(st2_plan* pl, void* handle_out) {
  if (!pl || !pl->strip || !handle_out) return st2_fail(pl ? pl->ctx : nullptr, ST2_ERR_ARG, "st2_strip_ipc_handle: not a strip plan");
  static_assert(sizeof(cudaIpcMemHandle_t) == ST2_IPC_HANDLE_BYTES, "ipc handle size");
  cudaIpcMemHandle_t h;
  ST2_CUDA(pl->ctx, cudaIpcGetMemHandle(&h, pl->slab));
  memcpy(handle_out, &h, sizeof(h));
  return 0;
}

int st2_strip_attach(st2_plan* pl, int side, const void* ipc_handle, st2_plan* local_peer, int peer_rows) {
  if (!pl || !pl->strip || side < 0 || side > 1 || (!ipc_handle == !local_peer) || peer_rows < 1)
    return st2_fail(pl ? pl->ctx : nullptr, ST2_ERR_ARG, "st2_strip_attach: bad arguments");
  st2_ctx* ctx = pl->ctx;
  if (pl->peer[side] && pl->peer_ipc[side]) cudaIpcCloseMemHandle(pl->peer[side]);
  pl->peer[side] = nullptr;
  if (local_peer) {
    if (!local_peer->strip || local_peer->H != peer_rows || local_peer->W != pl->W || local_peer->esz != pl->esz)
      return st2_fail(ctx, ST2_ERR_ARG, "st2_strip_attach: local peer does not match");
    pl->peer[side] = local_peer->slab;
    pl->peer_ipc[side] = false;
  } else {
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    void* p = nullptr;
    ST2_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    pl->peer[side] = (unsigned char*)p;
    pl->peer_ipc[side] = true;
  }
  pl->peer_lay[side] = strip_layout(peer_rows, pl->W, pl->esz);
  pl->async_halo = pl->peer[0] && pl->peer[1] && pl->peer_ipc[0] && pl->peer_ipc[1] && !ctx->knobs.no_inkernel_halo;
  return 0;
}

int st2_strip_set_fold(st2_plan* pl, int enable) {
  if (!pl || !pl->strip) return st2_fail(pl ? pl->ctx : nullptr, ST2_ERR_ARG, "st2_strip_set_fold: not a strip plan");
  if (enable && pl->prec == ST2_PREC_FP16 && pl->b[1].tc_bwd == nullptr)
    return st2_fail(pl->ctx, ST2_ERR_STATE, "st2_strip_set_fold: this strip is too small for the tensor-core conv1_1 kernels");
  pl->strip_fold = enable != 0;
  return 0;
}

int st2_strip_set_deferred(st2_plan* pl, int enable) {
  if (!pl || !pl->strip) return st2_fail(pl ? pl->ctx : nullptr, ST2_ERR_ARG, "st2_strip_set_deferred: not a strip plan");
  pl->deferred_sums = enable != 0;
  return 0;
}

int st2_strip_reduce_block(st2_plan* pl, int which, void** dev_out, long long* count_out) {
  if (!pl || !pl->strip || !dev_out || !count_out || which < 0 || which > 2)
    return st2_fail(pl ? pl->ctx : nullptr, ST2_ERR_ARG, "st2_strip_reduce_block: bad arguments");
  if (which == 0) { *dev_out = pl->gram_red; *count_out = pl->gram_red_used; }
  else if (which == 1) { *dev_out = pl->red; *count_out = 3 * ST2_NUM_BLOBS; }
  else { *dev_out = pl->scal + ST2_SCAL_GLOBAL_BASE + ST2_G_TV_NORM; *count_out = 6; }
  return 0;
}

int st2_strip_halo_error(st2_plan* pl, int* err_out) {
  if (!pl || !pl->strip || !err_out) return ST2_ERR_ARG;
  ST2_CUDA(pl->ctx, cudaStreamSynchronize(pl->ctx->stream));
  ST2_CUDA(pl->ctx, cudaMemcpy(err_out, &reinterpret_cast<SlabHeader*>(pl->slab)->err, sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

void st2_plan_destroy(st2_plan* pl) {
  if (!pl) return;
  for (int side = 0; side < 2; ++side)
    if (pl->peer[side] && pl->peer_ipc[side]) cudaIpcCloseMemHandle(pl->peer[side]);
  for (int i = 0; i < ST2_NUM_BLOBS; ++i) {
    Blob& B = pl->b[i];
    if (i > 0 && !pl->strip) { cudaFree(B.act); cudaFree(B.grad); }
    cudaFree(B.fc); cudaFree(B.sraw); cudaFree(B.inj); cudaFree(B.gram_target); cudaFree(B.D); cudaFree(B.Dh);
    tc_conv_plan_destroy(B.tc_fwd); tc_conv_plan_destroy(B.tc_bwd); tc_conv_plan_destroy(B.tc_style);
    tc_conv_plan_destroy(B.tc_sfold); cudaFree(B.wfold);
    tc_gram_plan_destroy(B.tc_gram);
    tc_first_plan_destroy(B.tc_first);
  }
  cudaFree(pl->slab); cudaFree(pl->red); cudaFree(pl->gram_red);
  for (int i = 0; i < ST2_NUM_BLOBS; ++i) cudaFree(pl->gram_local[i]);
  cudaFree(pl->scal); cudaFree(pl->gram_acc); cudaFree(pl->bwd); cudaFree(pl->part); cudaFree(pl->part_counter);
  delete pl;
}

int st2_plan_blob_dims(st2_plan* pl, int blob, int* c, int* h, int* w) {
  if (!pl || blob < 0 || blob >= ST2_NUM_BLOBS) return ST2_ERR_ARG;
  if (c) *c = pl->b[blob].C;
  if (h) *h = pl->b[blob].H;
  if (w) *w = pl->b[blob].W;
  return 0;
}

int st2_forward(st2_plan* pl, const float* x, int top) {
  if (!pl || !x || top < 0 || top >= ST2_NUM_BLOBS) return st2_fail(pl ? pl->ctx : nullptr, ST2_ERR_ARG, "st2_forward: bad arguments");
  return pl->prec == ST2_PREC_FP16 ? forward_impl<__half>(pl, x, top) : forward_impl<float>(pl, x, top);
}

int st2_bench_layer(st2_plan* pl, int blob, int direction, int reps, float* ms_out) {
  if (!pl || !ms_out || blob < 2 || blob >= ST2_NUM_BLOBS || reps < 1 || g_blobs[blob].kind != KIND_CONV)
    return st2_fail(pl ? pl->ctx : nullptr, ST2_ERR_ARG, "st2_bench_layer: needs a conv blob above conv1_1");
  st2_ctx* ctx = pl->ctx;
  Blob& cur = pl->b[blob];
  Blob& below = pl->b[blob - 1];
  const int ci = g_blobs[blob].conv_index;
  cudaEvent_t e0, e1;
  ST2_CUDA(ctx, cudaEventCreate(&e0));
  ST2_CUDA(ctx, cudaEventCreate(&e1));
  int rc = 0;
  for (int it = -2; it < reps && !rc; ++it) {
    if (it == 0) ST2_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    if (pl->prec == ST2_PREC_FP16) {
      rc = direction == 0
               ? tc_conv_launch(ctx, cur.tc_fwd, ctx->bias[ci], nullptr, (__half*)cur.act, EPI_BIAS_RELU, 1.f, nullptr)
               : tc_conv_launch(ctx, cur.tc_bwd, nullptr, (const __half*)below.act, (__half*)below.grad, EPI_MASK, 1.f, nullptr);
    } else {
      rc = direction == 0
               ? launch_conv_exact(ctx, (const float*)below.act, ctx->wf32_fwd[ci], ctx->bias[ci], nullptr,
                                   (float*)cur.act, cur.H, cur.W, below.C, cur.C, EPI_BIAS_RELU)
               : launch_conv_exact(ctx, (const float*)cur.grad, ctx->wf32_bwd[ci], nullptr, (const float*)below.act,
                                   (float*)below.grad, cur.H, cur.W, cur.C, below.C, EPI_MASK);
    }
  }
  if (rc) return rc;
  ST2_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  ST2_CUDA(ctx, cudaEventSynchronize(e1));
  float ms = 0.f;
  ST2_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = ms / reps;
  return 0;
}

int st2_blob_export(st2_plan* pl, int blob, float* out) {
  if (!pl || !out || blob < 0 || blob >= ST2_NUM_BLOBS) return st2_fail(pl ? pl->ctx : nullptr, ST2_ERR_ARG, "st2_blob_export: bad arguments");
  st2_ctx* ctx = pl->ctx;
  if (blob > pl->top_cur) return st2_fail(ctx, ST2_ERR_STATE, "%s not computed by the last forward", g_blobs[blob].name);
  Blob& B = pl->b[blob];
  if (blob == 0) {
    ST2_CUDA(ctx, cudaMemcpyAsync(out, B.act, sizeof(float) * B.n(), cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
  }
  return pl->prec == ST2_PREC_FP16 ? launch_export_nchw<__half>(ctx, (const __half*)B.act, out, B.C, B.H, B.W)
                                   : launch_export_nchw<float>(ctx, (const float*)B.act, out, B.C, B.H, B.W);
}

int st2_backward(st2_plan* pl, int n, const int* blobs, const float* const* diffs, float* grad_out) {
  if (!pl || n < 1 || !blobs || !diffs || !grad_out) return st2_fail(pl ? pl->ctx : nullptr, ST2_ERR_ARG, "st2_backward: bad arguments");
  st2_ctx* ctx = pl->ctx;
  if (pl->strip) return st2_fail(ctx, ST2_ERR_UNSUPPORTED, "st2_backward: not available on row strips (use st2_eval_*)");
  Inject inj[ST2_NUM_BLOBS];
  int top = 0;
  for (int k = 0; k < n; ++k) {
    const int b = blobs[k];
    if (b < 0 || b >= ST2_NUM_BLOBS || !diffs[k]) return st2_fail(ctx, ST2_ERR_ARG, "st2_backward: bad blob %d", b);
    if (b > pl->top_cur) return st2_fail(ctx, ST2_ERR_STATE, "%s not computed by the last forward", g_blobs[b].name);
    Blob& B = pl->b[b];
    inj[b].on = true; inj[b].hsc = 1.f;
    if (b == 0) {
      inj[b].sraw = diffs[k];
    } else {
      int rc = ensure(ctx, &B.inj, pl->esz * B.n());
      if (rc) return rc;
      rc = pl->prec == ST2_PREC_FP16 ? launch_import_nchw<__half>(ctx, diffs[k], (__half*)B.inj, B.C, B.H, B.W)
                                     : launch_import_nchw<float>(ctx, diffs[k], (float*)B.inj, B.C, B.H, B.W);
      if (rc) return rc;
      inj[b].sraw = B.inj;
    }
    if (b > top) top = b;
  }
  return pl->prec == ST2_PREC_FP16 ? backward_impl<__half>(pl, top, inj, grad_out)
                                   : backward_impl<float>(pl, top, inj, grad_out);
}

int st2_capture_content(st2_plan* pl, int blob) {
  if (!pl || blob < 0 || blob >= ST2_NUM_BLOBS) return ST2_ERR_ARG;
  st2_ctx* ctx = pl->ctx;
  if (blob > pl->top_cur) return st2_fail(ctx, ST2_ERR_STATE, "%s not computed by the last forward", g_blobs[blob].name);
  Blob& B = pl->b[blob];
  const size_t bytes = (blob == 0 ? sizeof(float) : pl->esz) * B.n();
  int rc = ensure(ctx, &B.fc, bytes);
  if (rc) return rc;
  ST2_CUDA(ctx, cudaMemcpyAsync(B.fc, B.act, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

int st2_gram(st2_plan* pl, int blob, float* out) {
  if (!pl || !out || blob < 0 || blob >= ST2_NUM_BLOBS) return ST2_ERR_ARG;
  if (blob > pl->top_cur) return st2_fail(pl->ctx, ST2_ERR_STATE, "%s not computed by the last forward", g_blobs[blob].name);
  if (pl->strip)     /* un-normalised Gram sum of this strip: the caller all-reduces and divides by C*H_total*W */
    return pl->prec == ST2_PREC_FP16 ? gram_sum_of_blob<__half>(pl, blob, out) : gram_sum_of_blob<float>(pl, blob, out);
  return pl->prec == ST2_PREC_FP16 ? gram_of_blob<__half>(pl, blob, nullptr, out, nullptr)
                                   : gram_of_blob<float>(pl, blob, nullptr, out, nullptr);
}

int st2_set_style_gram(st2_plan* pl, int blob, const float* gram) {
  if (!pl || !gram || blob < 0 || blob >= ST2_NUM_BLOBS) return ST2_ERR_ARG;
  st2_ctx* ctx = pl->ctx;
  Blob& B = pl->b[blob];
  int rc = ensure(ctx, (void**)&B.gram_target, sizeof(float) * B.C * B.C);
  if (rc) return rc;
  ST2_CUDA(ctx, cudaMemcpyAsync(B.gram_target, gram, sizeof(float) * B.C * B.C, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

int st2_set_blob_weights(st2_plan* pl, int blob, float c, float s, float d) {
  if (!pl || blob < 0 || blob >= ST2_NUM_BLOBS) return ST2_ERR_ARG;
  pl->b[blob].cw = c; pl->b[blob].sw = s; pl->b[blob].dw = d;
  return 0;
}

int st2_set_eval_order(st2_plan* pl, int n, const int* blobs) {
  if (!pl || n < 0 || n > ST2_NUM_BLOBS || (n && !blobs)) return ST2_ERR_ARG;
  for (int i = 0; i < n; ++i) {
    if (blobs[i] < 0 || blobs[i] >= ST2_NUM_BLOBS) return st2_fail(pl->ctx, ST2_ERR_ARG, "st2_set_eval_order: bad blob");
    pl->order[i] = blobs[i];
  }
  pl->order_n = n;
  return 0;
}

int st2_set_params(st2_plan* pl, float tv, float tv_power, float p, float p_power) {
  if (!pl) return ST2_ERR_ARG;
  pl->tv = tv; pl->tv_power = tv_power; pl->p = p; pl->p_power = p_power;
  return 0;
}

int st2_reset_norms(st2_plan* pl) {
  if (!pl) return ST2_ERR_ARG;
  ST2_CUDA(pl->ctx, cudaMemsetAsync(pl->scal, 0, sizeof(double) * ST2_SCAL_TOTAL, pl->ctx->stream));
  return 0;
}

int st2_set_norm(st2_plan* pl, int kind, int blob, double value) {
  if (!pl || kind < 0 || kind > 2 || blob < 0 || blob >= ST2_NUM_BLOBS) return ST2_ERR_ARG;
  st2_ctx* ctx = pl->ctx;
  double* sb = pl->scal + blob * ST2_SCAL_PER_BLOB;
  const double one = 1.0;
  ST2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ST2_CUDA(ctx, cudaMemcpy(sb + SB_C_NORM + kind, &value, sizeof(double), cudaMemcpyHostToDevice));
  ST2_CUDA(ctx, cudaMemcpy(sb + SB_C_VALID + kind, &one, sizeof(double), cudaMemcpyHostToDevice));
  return 0;
}

int st2_eval_begin(st2_plan* pl, const float* x, int want_grad) {
  if (!pl || !x) return st2_fail(pl ? pl->ctx : nullptr, ST2_ERR_ARG, "st2_eval_begin: bad arguments");
  return pl->prec == ST2_PREC_FP16 ? eval_begin_impl<__half>(pl, x, want_grad) : eval_begin_impl<float>(pl, x, want_grad);
}

int st2_eval_mid(st2_plan* pl) {
  if (!pl) return ST2_ERR_ARG;
  return pl->prec == ST2_PREC_FP16 ? eval_mid_impl<__half>(pl) : eval_mid_impl<float>(pl);
}

int st2_eval_end(st2_plan* pl, float* grad) {
  if (!pl) return ST2_ERR_ARG;
  return pl->prec == ST2_PREC_FP16 ? eval_end_impl<__half>(pl, grad) : eval_end_impl<float>(pl, grad);
}

int st2_eval_final(st2_plan* pl) { return pl ? eval_final_impl(pl) : ST2_ERR_ARG; }

int st2_eval(st2_plan* pl, const float* x, float* grad, int want_grad) {
  if (!pl || !x) return st2_fail(pl ? pl->ctx : nullptr, ST2_ERR_ARG, "st2_eval: bad arguments");
  if (pl->strip && pl->world > 1)
    return st2_fail(pl->ctx, ST2_ERR_STATE, "st2_eval on a row strip: use st2_eval_begin/mid/end/final with all-reduces");
  int rc = st2_eval_begin(pl, x, want_grad);
  if (!rc) rc = st2_eval_mid(pl);
  if (!rc) rc = st2_eval_end(pl, grad);
  if (!rc) rc = st2_eval_final(pl);
  return rc;
}

int st2_read_scalars(st2_plan* pl, double* host_out) {
  if (!pl || !host_out) return ST2_ERR_ARG;
  st2_ctx* ctx = pl->ctx;
  ST2_CUDA(ctx, cudaMemcpyAsync(host_out, pl->scal, sizeof(double) * ST2_SCAL_TOTAL, cudaMemcpyDeviceToHost, ctx->stream));
  ST2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int st2_copy_scalars_async(st2_plan* pl, double* pinned_host_out) {
  if (!pl || !pinned_host_out) return ST2_ERR_ARG;
  ST2_CUDA(pl->ctx, cudaMemcpyAsync(pinned_host_out, pl->scal, sizeof(double) * ST2_SCAL_TOTAL,
                                    cudaMemcpyDeviceToHost, pl->ctx->stream));
  return 0;
}

double* st2_scalars_dev(st2_plan* pl) { return pl ? pl->scal : nullptr; }

int st2_gram_nchw(st2_ctx* ctx, const float* x, int C, long long HW, float* out) {
  if (!ctx || !x || !out || C < 1 || HW < 1) return st2_fail(ctx, ST2_ERR_ARG, "st2_gram_nchw: bad arguments");
  double* acc = nullptr;
  ST2_CUDA(ctx, cudaMallocAsync(&acc, sizeof(double) * C * C, ctx->stream));
  ST2_CUDA(ctx, cudaMemsetAsync(acc, 0, sizeof(double) * C * C, ctx->stream));
  int rc = launch_gram_generic<float>(ctx, x, C, HW, 1, HW, acc);
  if (!rc) rc = launch_gram_finalize(ctx, acc, nullptr, out, C, HW, nullptr);
  ST2_CUDA(ctx, cudaFreeAsync(acc, ctx->stream));
  return rc;
}

}  // extern "C"

static St2KernelReg g_reg_net({ST2_KFN(halo_exchange_kernel), ST2_KFN(pack_x_kernel),
                                  ST2_KFN(clear_volatile_kernel), ST2_KFN(style_scale_kernel), ST2_KFN(style_fold_kernel), ST2_KFN(style_rawsq_kernel), ST2_KFN(dual_pack_kernel), ST2_KFN(stencil_pack_kernel), ST2_KFN(scatter_sums_kernel),
                                  ST2_KFN(coef_kernel), ST2_KFN(final_kernel), ST2_KFN(pack_weights_kernel)});

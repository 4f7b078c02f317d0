// L-BFGS (optimizers.LBFGSOptimizer, optimizers.py:49-125) in compact ("vector-free") form.
//
// The reference runs the two-loop recursion as 2m+1 sdot and 2m saxpy calls over N-vectors
// (~464 N bytes of traffic per step at m = 10).  Here the recursion runs in *coefficient space*:
// every intermediate q is a combination  cg*g + sum_j a_j S_j + sum_j b_j Y_j,  and each dot product
// the recursion needs (S_i.q, Y_i.q) follows from a small table of pairwise dot products that is
// maintained incrementally.  One optimizer step is then two passes over the history:
//
//   pass A (after the objective evaluation):  y = g_new - g_prev  -> Y[new];  in the same sweep all
//          dots of {y, g_new} with every stored S_j, Y_j plus s.y, y.y, g.s, g.y, g.g  (4m+5 sums)
//   tiny   accept the pair iff s.y > 1e-10 (optimizers.py:79-87), update the tables, FIFO cap
//   tiny   two-loop recursion on the coefficients (optimizers.py:89-108), single thread, double
//   pass B s = -step * (cg g + sum a_j S_j + sum b_j Y_j) -> S[new];  x += s;  and s.Y_j for the tables
//
// = (2m+4)*4N bytes per pass, 192 N per step, 4 launches, no host synchronisation; and -- the point
// for row-strip tiling across GPUs -- every cross-vector reduction of a step sits in ONE block of
// sums (`st2_lbfgs_sums_dev`), so a multi-GPU step needs a single small all-reduce.
#include "st2_common.cuh"

#include <math.h>
#include <string.h>

#define MAXM 10                       // n_corr <= 10
#define SLOTS (MAXM + 1)              // one staging slot for the pair being formed

namespace {

constexpr int kThreads = 256;

// sums block (doubles): what pass A / pass B / the cold pass accumulate; all-reduced across ranks
enum {
  SUM_SY = 0,                 // [MAXM] S_j . y      (logical j)
  SUM_YY = SUM_SY + MAXM,     // [MAXM] Y_j . y
  SUM_SG = SUM_YY + MAXM,     // [MAXM] S_j . g
  SUM_YG = SUM_SG + MAXM,     // [MAXM] Y_j . g
  SUM_s_y = SUM_YG + MAXM,    // s . y
  SUM_y_y,                    // y . y
  SUM_g_s,                    // g . s
  SUM_g_y,                    // g . y
  SUM_g_g,                    // g . g
  SUM_SNY,                    // [MAXM] s_new . Y_j   (pass B)
  SUM_TOTAL = SUM_SNY + MAXM
};

struct LbfgsDev {
  int count, head, n_corr, pad;
  double sy[SLOTS], yy[SLOTS];          // per physical slot: s.y (fp32-valued, as utils.dot) and y.y
  double SY[SLOTS][SLOTS];              // S_p . Y_q
  double YY[SLOTS][SLOTS];              // Y_p . Y_q
  double gS[SLOTS], gY[SLOTS], gg;      // dots of the current gradient
  double cg, a[SLOTS], b[SLOTS];        // direction coefficients, by physical slot
  double sums[SUM_TOTAL];
};

template <int V> struct Pack { float v[V]; };
template <int V> __device__ __forceinline__ Pack<V> ld(const float* p, long long i);
template <> __device__ __forceinline__ Pack<1> ld<1>(const float* p, long long i) { Pack<1> r; r.v[0] = p[i]; return r; }
template <> __device__ __forceinline__ Pack<4> ld<4>(const float* p, long long i) {
  const float4 u = reinterpret_cast<const float4*>(p)[i];
  Pack<4> r; r.v[0] = u.x; r.v[1] = u.y; r.v[2] = u.z; r.v[3] = u.w; return r;
}
template <int V> __device__ __forceinline__ void st(float* p, long long i, const Pack<V>& a);
template <> __device__ __forceinline__ void st<1>(float* p, long long i, const Pack<1>& a) { p[i] = a.v[0]; }
template <> __device__ __forceinline__ void st<4>(float* p, long long i, const Pack<4>& a) {
  reinterpret_cast<float4*>(p)[i] = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
template <int V> __device__ __forceinline__ float dotp(const Pack<V>& a, const Pack<V>& b) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < V; ++e) s = fmaf(a.v[e], b.v[e], s);
  return s;
}

// block-wide reduction of NV fp32 partials into doubles (one atomicAdd per value per block)
template <int NV>
__device__ __forceinline__ void block_sums(const float (&v)[NV], double* dst) {
  __shared__ float sh[NV][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float w = warp_sum(v[i]);
    if (lane == 0) sh[i][warp] = w;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NV; i += blockDim.x) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += (double)sh[i][w];
    atomicAdd(&dst[i], t);
  }
}

#define PACK_LOOP(k, npk) \
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < (npk); k += (long long)gridDim.x * blockDim.x)

// ---- pass A: y = g_new - g_prev -> Y[new]; 4m + 5 sums ------------------------------------------
template <int V>
__global__ void __launch_bounds__(kThreads, 1)
lbfgs_pass_a(LbfgsDev* st_, const float* __restrict__ S, float* __restrict__ Y, const float* __restrict__ g_new,
             const float* __restrict__ g_prev, long long n) {
  pdl_trigger();
  pdl_wait();
  const int count = st_->count, head = st_->head;
  const int pnew = (head + count) % SLOTS;
  const float* s = S + (long long)pnew * n;
  float* y = Y + (long long)pnew * n;
  float acc[4 * MAXM + 5];
#pragma unroll
  for (int i = 0; i < 4 * MAXM + 5; ++i) acc[i] = 0.f;
  PACK_LOOP(k, n / V) {
    const Pack<V> gn = ld<V>(g_new, k), gp = ld<V>(g_prev, k), sv = ld<V>(s, k);
    Pack<V> yv;
#pragma unroll
    for (int e = 0; e < V; ++e) yv.v[e] = gn.v[e] - gp.v[e];
    st<V>(y, k, yv);
    // issue every history load before the first use: the kernel lives on memory-level parallelism
    Pack<V> sj[MAXM], yj[MAXM];
#pragma unroll
    for (int j = 0; j < MAXM; ++j) {
      if (j < count) {
        const long long off = (long long)((head + j) % SLOTS) * n;
        sj[j] = ld<V>(S + off, k);
        yj[j] = ld<V>(Y + off, k);
      }
    }
#pragma unroll
    for (int j = 0; j < MAXM; ++j) {
      if (j < count) {
        acc[j] += dotp<V>(sj[j], yv);
        acc[MAXM + j] += dotp<V>(yj[j], yv);
        acc[2 * MAXM + j] += dotp<V>(sj[j], gn);
        acc[3 * MAXM + j] += dotp<V>(yj[j], gn);
      }
    }
    acc[4 * MAXM + 0] += dotp<V>(sv, yv);
    acc[4 * MAXM + 1] += dotp<V>(yv, yv);
    acc[4 * MAXM + 2] += dotp<V>(gn, sv);
    acc[4 * MAXM + 3] += dotp<V>(gn, yv);
    acc[4 * MAXM + 4] += dotp<V>(gn, gn);
  }
  block_sums<4 * MAXM + 5>(acc, st_->sums);
}

// ---- cold pass: dots of g with the stored history (first step after reset / load) ---------------
template <int V>
__global__ void __launch_bounds__(kThreads)
lbfgs_pass_g(LbfgsDev* st_, const float* __restrict__ S, const float* __restrict__ Y, const float* __restrict__ g,
             long long n) {
  pdl_trigger();
  pdl_wait();
  const int count = st_->count, head = st_->head;
  float acc[2 * MAXM + 1];
#pragma unroll
  for (int i = 0; i < 2 * MAXM + 1; ++i) acc[i] = 0.f;
  PACK_LOOP(k, n / V) {
    const Pack<V> gv = ld<V>(g, k);
#pragma unroll
    for (int j = 0; j < MAXM; ++j) {
      if (j < count) {
        const long long off = (long long)((head + j) % SLOTS) * n;
        acc[j] += dotp<V>(ld<V>(S + off, k), gv);
        acc[MAXM + j] += dotp<V>(ld<V>(Y + off, k), gv);
      }
    }
    acc[2 * MAXM] += dotp<V>(gv, gv);
  }
  // same slots of the sums block as pass A uses for the g dots
  float sg[MAXM], yg[MAXM], gg[1];
#pragma unroll
  for (int j = 0; j < MAXM; ++j) { sg[j] = acc[j]; yg[j] = acc[MAXM + j]; }
  gg[0] = acc[2 * MAXM];
  block_sums<MAXM>(sg, st_->sums + SUM_SG);
  __syncthreads();
  block_sums<MAXM>(yg, st_->sums + SUM_YG);
  __syncthreads();
  block_sums<1>(gg, st_->sums + SUM_g_g);
}

// ---- load path: dots of Y[phys] with all stored vectors (fills SY[.][phys], YY[phys][.]) ---------
template <int V>
__global__ void __launch_bounds__(kThreads)
lbfgs_pass_y(LbfgsDev* st_, const float* __restrict__ S, const float* __restrict__ Y, long long n, int phys) {
  pdl_trigger();
  pdl_wait();
  const int count = st_->count, head = st_->head;
  const float* y = Y + (long long)phys * n;
  float acc[2 * MAXM];
#pragma unroll
  for (int i = 0; i < 2 * MAXM; ++i) acc[i] = 0.f;
  PACK_LOOP(k, n / V) {
    const Pack<V> yv = ld<V>(y, k);
#pragma unroll
    for (int j = 0; j < MAXM; ++j) {
      if (j < count) {
        const long long off = (long long)((head + j) % SLOTS) * n;
        acc[j] += dotp<V>(ld<V>(S + off, k), yv);
        acc[MAXM + j] += dotp<V>(ld<V>(Y + off, k), yv);
      }
    }
  }
  block_sums<2 * MAXM>(acc, st_->sums);          // SUM_SY / SUM_YY slots
}

__global__ void lbfgs_store_y_dots(LbfgsDev* st_, int phys) {
  if (threadIdx.x != 0) return;
  for (int j = 0; j < st_->count; ++j) {
    const int q = (st_->head + j) % SLOTS;
    st_->SY[q][phys] = st_->sums[SUM_SY + j];
    st_->YY[q][phys] = st_->sums[SUM_YY + j];
    st_->YY[phys][q] = st_->sums[SUM_YY + j];
  }
  st_->yy[phys] = st_->YY[phys][phys];
  for (int i = 0; i < SUM_TOTAL; ++i) st_->sums[i] = 0.0;
}

__global__ void lbfgs_clear_sums(LbfgsDev* st_, int first, int last) {
  for (int i = first + threadIdx.x; i < last; i += blockDim.x) st_->sums[i] = 0.0;
}

// ---- accept (optimizers.py:79-87) + table update -----------------------------------------------
__global__ void lbfgs_accept(LbfgsDev* st_) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x != 0) return;
  const int count = st_->count, head = st_->head;
  const int pnew = (head + count) % SLOTS;
  const double* u = st_->sums;
  // dots of the new gradient with the existing history (valid whether or not the pair is kept)
  for (int j = 0; j < count; ++j) {
    const int q = (head + j) % SLOTS;
    st_->gS[q] = u[SUM_SG + j];
    st_->gY[q] = u[SUM_YG + j];
  }
  st_->gg = u[SUM_g_g];
  const double sy = (double)(float)u[SUM_s_y];           // utils.dot returns an fp32-valued float
  if (sy > 1e-10) {
    for (int j = 0; j < count; ++j) {
      const int q = (head + j) % SLOTS;
      st_->SY[q][pnew] = u[SUM_SY + j];                  // S_q . y
      st_->SY[pnew][q] = u[SUM_SNY + j];                 // s . Y_q   (pass B of the step that made s)
      st_->YY[q][pnew] = u[SUM_YY + j];
      st_->YY[pnew][q] = u[SUM_YY + j];
    }
    st_->SY[pnew][pnew] = u[SUM_s_y];
    st_->YY[pnew][pnew] = u[SUM_y_y];
    st_->sy[pnew] = sy;
    st_->yy[pnew] = (double)(float)u[SUM_y_y];
    st_->gS[pnew] = u[SUM_g_s];
    st_->gY[pnew] = u[SUM_g_y];
    if (count == st_->n_corr) st_->head = (head + 1) % SLOTS;
    else st_->count = count + 1;
  }
}

// cold: take the g dots from the sums block (after pass G)
__global__ void lbfgs_take_g(LbfgsDev* st_) {
  if (threadIdx.x != 0) return;
  for (int j = 0; j < st_->count; ++j) {
    const int q = (st_->head + j) % SLOTS;
    st_->gS[q] = st_->sums[SUM_SG + j];
    st_->gY[q] = st_->sums[SUM_YG + j];
  }
  st_->gg = st_->sums[SUM_g_g];
}

// ---- two-loop recursion on coefficients (optimizers.py:89-108) ------------------------------------
__global__ void lbfgs_coefficients(LbfgsDev* st_, double n_total) {
  // the tables come to shared memory in one coalesced sweep; the recursion itself is ~4 m^2 dependent double
  // operations on one thread (global-memory latency per operand made this kernel 14 us, now ~3)
  __shared__ LbfgsDev sh;
  static_assert(sizeof(LbfgsDev) % 8 == 0, "LbfgsDev is copied as doubles");
  pdl_trigger();
  pdl_wait();
  {
    const double* src = reinterpret_cast<const double*>(st_);
    double* dst = reinterpret_cast<double*>(&sh);
    for (int i = threadIdx.x; i < (int)(sizeof(LbfgsDev) / 8); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int m = sh.count, head = sh.head;
    int ph[MAXM];
    double a[MAXM], b[MAXM], alpha[MAXM];
    for (int j = 0; j < MAXM; ++j) { ph[j] = (head + j) % SLOTS; a[j] = 0.0; b[j] = 0.0; alpha[j] = 0.0; }
    double cg = 1.0;
    for (int i = m - 1; i >= 0; --i) {                    // newest -> oldest; q has g and Y components only
      double sq = cg * sh.gS[ph[i]];
      for (int j = 0; j < m; ++j) sq += b[j] * sh.SY[ph[i]][ph[j]];
      alpha[i] = sq / sh.sy[ph[i]];
      b[i] -= alpha[i];
    }
    if (m > 0) {
      const double gamma = (double)(float)(sh.sy[ph[m - 1]] / sh.yy[ph[m - 1]]);
      cg *= gamma;
      for (int j = 0; j < m; ++j) b[j] *= gamma;
    } else {
      cg = 1.0 / sqrt(sh.gg / n_total);                   // unit-RMS first step
    }
    for (int i = 0; i < m; ++i) {                         // oldest -> newest
      double yq = cg * sh.gY[ph[i]];
      for (int j = 0; j < m; ++j) yq += b[j] * sh.YY[ph[i]][ph[j]] + a[j] * sh.SY[ph[j]][ph[i]];
      const double beta = yq / sh.sy[ph[i]];
      a[i] += alpha[i] - beta;
    }
    sh.cg = cg;
    for (int j = 0; j < SLOTS; ++j) { sh.a[j] = 0.0; sh.b[j] = 0.0; }
    for (int j = 0; j < m; ++j) { sh.a[ph[j]] = a[j]; sh.b[ph[j]] = b[j]; }
  }
  __syncthreads();
  if (threadIdx.x == 0) st_->cg = sh.cg;
  for (int j = threadIdx.x; j < SLOTS; j += blockDim.x) { st_->a[j] = sh.a[j]; st_->b[j] = sh.b[j]; }
  // every sum is consumed by now: zero the block for pass B (s_new . Y_j) and the next pass A
  for (int i = threadIdx.x; i < SUM_TOTAL; i += blockDim.x) st_->sums[i] = 0.0;
}

// ---- pass B: s = -step * direction -> S[new]; x += s; s.Y_j ---------------------------------------
template <int V>
__global__ void __launch_bounds__(kThreads, 1)
lbfgs_pass_b(LbfgsDev* st_, float* __restrict__ S, const float* __restrict__ Y, const float* __restrict__ g,
             float* __restrict__ x, long long n, float step) {
  pdl_trigger();
  pdl_wait();
  const int count = st_->count, head = st_->head;
  const int pnew = (head + count) % SLOTS;
  float* s_new = S + (long long)pnew * n;
  const float cg = (float)st_->cg;
  float ca[MAXM], cb[MAXM];
#pragma unroll
  for (int j = 0; j < MAXM; ++j) {
    const int q = (head + j) % SLOTS;
    ca[j] = (j < count) ? (float)st_->a[q] : 0.f;
    cb[j] = (j < count) ? (float)st_->b[q] : 0.f;
  }
  float acc[MAXM];
#pragma unroll
  for (int j = 0; j < MAXM; ++j) acc[j] = 0.f;
  PACK_LOOP(k, n / V) {
    const Pack<V> gv = ld<V>(g, k);
    Pack<V> q;
#pragma unroll
    for (int e = 0; e < V; ++e) q.v[e] = cg * gv.v[e];
    Pack<V> sj[MAXM], yj[MAXM];
#pragma unroll
    for (int j = 0; j < MAXM; ++j) {
      if (j < count) {
        const long long off = (long long)((head + j) % SLOTS) * n;
        sj[j] = ld<V>(S + off, k);
        yj[j] = ld<V>(Y + off, k);
      }
    }
    Pack<V> xv = ld<V>(x, k);
#pragma unroll
    for (int j = 0; j < MAXM; ++j) {
      if (j < count) {
#pragma unroll
        for (int e = 0; e < V; ++e) q.v[e] = fmaf(ca[j], sj[j].v[e], fmaf(cb[j], yj[j].v[e], q.v[e]));
      }
    }
    Pack<V> sv;
#pragma unroll
    for (int e = 0; e < V; ++e) { sv.v[e] = -step * q.v[e]; xv.v[e] += sv.v[e]; }
    st<V>(s_new, k, sv);
    st<V>(x, k, xv);
#pragma unroll
    for (int j = 0; j < MAXM; ++j)
      if (j < count) acc[j] += dotp<V>(sv, yj[j]);
  }
  block_sums<MAXM>(acc, st_->sums + SUM_SNY);
}

inline int grid_for(long long n, int sm_count) {
  long long blocks = (n / 4 + kThreads - 1) / kThreads;
  const long long cap = (long long)sm_count * 2;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

struct st2_lbfgs {
  st2_ctx* ctx;
  long long n;          // local vector length
  double n_total;       // global length (== n on one GPU; sum over ranks when row-tiled)
  int n_corr;
  bool have_gdots;      // gS/gY/gg describe the gradient that the next advance will use
  bool sums_clean;      // sums[0, SUM_SNY) known to be zero (set by advance_end)
  float *S, *Y;
  LbfgsDev* st;
};

#define LAUNCH_V(kernel, grid, ...)                                                             \
  do {                                                                                          \
    if (o->n % 4 == 0) st2_launch_pdl(ctx, true, kernel<4>, grid, kThreads, 0, __VA_ARGS__);    \
    else st2_launch_pdl(ctx, true, kernel<1>, grid, kThreads, 0, __VA_ARGS__);                  \
    ST2_LAUNCH_CHECK(ctx);                                                                      \
  } while (0)

extern "C" {

int st2_lbfgs_create(st2_ctx* ctx, long long n, int n_corr, st2_lbfgs** out) {
  if (!ctx || !out || n <= 0 || n_corr < 1 || n_corr > MAXM)
    return st2_fail(ctx, ST2_ERR_ARG, "st2_lbfgs_create: bad arguments (n=%lld n_corr=%d)", n, n_corr);
  st2_lbfgs* o = new st2_lbfgs();
  o->ctx = ctx; o->n = n; o->n_total = (double)n; o->n_corr = n_corr; o->have_gdots = false; o->sums_clean = false;
  ST2_CUDA(ctx, cudaMalloc(&o->S, sizeof(float) * n * SLOTS));
  ST2_CUDA(ctx, cudaMalloc(&o->Y, sizeof(float) * n * SLOTS));
  ST2_CUDA(ctx, cudaMalloc(&o->st, sizeof(LbfgsDev)));
  *out = o;
  return st2_lbfgs_reset(o);
}

void st2_lbfgs_destroy(st2_lbfgs* o) {
  if (!o) return;
  cudaFree(o->S); cudaFree(o->Y); cudaFree(o->st);
  delete o;
}

int st2_lbfgs_reset(st2_lbfgs* o) {
  if (!o) return ST2_ERR_ARG;
  LbfgsDev h;
  memset(&h, 0, sizeof(h));
  h.n_corr = o->n_corr;
  ST2_CUDA(o->ctx, cudaMemcpyAsync(o->st, &h, sizeof(h), cudaMemcpyHostToDevice, o->ctx->stream));
  ST2_CUDA(o->ctx, cudaStreamSynchronize(o->ctx->stream));     // h is on this stack frame
  o->have_gdots = false;
  o->sums_clean = false;
  return 0;
}

int st2_lbfgs_set_global_length(st2_lbfgs* o, double n_total) {
  if (!o || n_total < 1) return ST2_ERR_ARG;
  o->n_total = n_total;
  return 0;
}

double* st2_lbfgs_sums_dev(st2_lbfgs* o) { return o ? o->st->sums : nullptr; }
int st2_lbfgs_sums_count(void) { return SUM_TOTAL; }

// s = -step * inv_hv(g); x += s -- split so a multi-GPU caller can all-reduce the sums in between
int st2_lbfgs_advance_begin(st2_lbfgs* o, const float* g) {
  if (!o || !g) return st2_fail(o ? o->ctx : nullptr, ST2_ERR_ARG, "st2_lbfgs_advance_begin: null");
  st2_ctx* ctx = o->ctx;
  cudaStream_t s = ctx->stream;
  if (!o->have_gdots) {
    ProfScope ps(ctx, 7);
    LAUNCH_V(lbfgs_pass_g, grid_for(o->n, ctx->sm_count), o->st, o->S, o->Y, g, o->n);
  }
  return 0;
}

int st2_lbfgs_advance_end(st2_lbfgs* o, float* x, const float* g, float step) {
  if (!o || !x || !g) return st2_fail(o ? o->ctx : nullptr, ST2_ERR_ARG, "st2_lbfgs_advance_end: null");
  st2_ctx* ctx = o->ctx;
  cudaStream_t s = ctx->stream;
  ProfScope ps(ctx, 7);
  if (!o->have_gdots) {
    lbfgs_take_g<<<1, 32, 0, s>>>(o->st);
    ST2_LAUNCH_CHECK(ctx);
  }
  st2_launch_pdl(ctx, true, lbfgs_coefficients, 1, 128, 0, o->st, o->n_total);
  ST2_LAUNCH_CHECK(ctx);
  LAUNCH_V(lbfgs_pass_b, grid_for(o->n, ctx->sm_count), o->st, o->S, o->Y, g, x, o->n, step);
  o->have_gdots = false;
  o->sums_clean = true;
  return 0;
}

int st2_lbfgs_advance(st2_lbfgs* o, float* x, const float* g, float step) {
  int rc = st2_lbfgs_advance_begin(o, g);
  return rc ? rc : st2_lbfgs_advance_end(o, x, g, step);
}

// y = g_new - g_prev; keep (s, y) iff s.y > 1e-10; FIFO cap n_corr
int st2_lbfgs_commit_begin(st2_lbfgs* o, const float* g_new, const float* g_prev) {
  if (!o || !g_new || !g_prev) return st2_fail(o ? o->ctx : nullptr, ST2_ERR_ARG, "st2_lbfgs_commit_begin: null");
  st2_ctx* ctx = o->ctx;
  cudaStream_t s = ctx->stream;
  ProfScope ps(ctx, 7);
  // sums [0, SUM_SNY) are zero here: lbfgs_coefficients cleared the block and only pass B (s_new . Y_j) wrote since
  if (!o->sums_clean) {
    lbfgs_clear_sums<<<1, 64, 0, s>>>(o->st, 0, SUM_SNY);
    ST2_LAUNCH_CHECK(ctx);
  }
  o->sums_clean = false;
  LAUNCH_V(lbfgs_pass_a, grid_for(o->n, ctx->sm_count), o->st, o->S, o->Y, g_new, g_prev, o->n);
  return 0;
}

int st2_lbfgs_commit_end(st2_lbfgs* o) {
  if (!o) return ST2_ERR_ARG;
  st2_ctx* ctx = o->ctx;
  ProfScope ps(ctx, 7);
  st2_launch_pdl(ctx, true, lbfgs_accept, 1, 32, 0, o->st);
  ST2_LAUNCH_CHECK(ctx);
  o->have_gdots = true;
  return 0;
}

int st2_lbfgs_commit(st2_lbfgs* o, const float* g_new, const float* g_prev) {
  int rc = st2_lbfgs_commit_begin(o, g_new, g_prev);
  return rc ? rc : st2_lbfgs_commit_end(o);
}

int st2_lbfgs_load(st2_lbfgs* o, int count, const float* s_dev, const float* y_dev, const double* sy_host) {
  if (!o || count < 0 || count > o->n_corr) return st2_fail(o ? o->ctx : nullptr, ST2_ERR_ARG, "st2_lbfgs_load: bad count");
  st2_ctx* ctx = o->ctx;
  cudaStream_t s = ctx->stream;
  LbfgsDev h;
  memset(&h, 0, sizeof(h));
  h.count = count; h.head = 0; h.n_corr = o->n_corr;
  for (int i = 0; i < count; ++i) h.sy[i] = sy_host[i];
  ST2_CUDA(ctx, cudaStreamSynchronize(s));
  ST2_CUDA(ctx, cudaMemcpy(o->st, &h, sizeof(h), cudaMemcpyHostToDevice));
  if (count) {
    ST2_CUDA(ctx, cudaMemcpyAsync(o->S, s_dev, sizeof(float) * o->n * count, cudaMemcpyDeviceToDevice, s));
    ST2_CUDA(ctx, cudaMemcpyAsync(o->Y, y_dev, sizeof(float) * o->n * count, cudaMemcpyDeviceToDevice, s));
  }
  for (int p = 0; p < count; ++p) {                     // rebuild the dot tables by brute force (cold path)
    LAUNCH_V(lbfgs_pass_y, grid_for(o->n, ctx->sm_count), o->st, o->S, o->Y, o->n, p);
    lbfgs_store_y_dots<<<1, 32, 0, s>>>(o->st, p);
    ST2_LAUNCH_CHECK(ctx);
  }
  o->have_gdots = false;
  return 0;
}

int st2_lbfgs_export(st2_lbfgs* o, int* count_out, float* s_dev, float* y_dev, double* sy_host) {
  if (!o || !count_out) return st2_fail(o ? o->ctx : nullptr, ST2_ERR_ARG, "st2_lbfgs_export: null");
  st2_ctx* ctx = o->ctx;
  cudaStream_t s = ctx->stream;
  LbfgsDev h;
  ST2_CUDA(ctx, cudaStreamSynchronize(s));
  ST2_CUDA(ctx, cudaMemcpy(&h, o->st, sizeof(h), cudaMemcpyDeviceToHost));
  *count_out = h.count;
  for (int i = 0; i < h.count; ++i) {
    const int phys = (h.head + i) % SLOTS;
    if (sy_host) sy_host[i] = h.sy[phys];
    if (s_dev) ST2_CUDA(ctx, cudaMemcpyAsync(s_dev + (long long)i * o->n, o->S + (long long)phys * o->n,
                                             sizeof(float) * o->n, cudaMemcpyDeviceToDevice, s));
    if (y_dev) ST2_CUDA(ctx, cudaMemcpyAsync(y_dev + (long long)i * o->n, o->Y + (long long)phys * o->n,
                                             sizeof(float) * o->n, cudaMemcpyDeviceToDevice, s));
  }
  ST2_CUDA(ctx, cudaStreamSynchronize(s));
  return 0;
}

}  // extern "C"

static St2KernelReg g_reg_lbfgs({
    ST2_KFN(lbfgs_pass_a<1>), ST2_KFN(lbfgs_pass_a<4>), ST2_KFN(lbfgs_pass_b<1>), ST2_KFN(lbfgs_pass_b<4>),
    ST2_KFN(lbfgs_pass_g<1>), ST2_KFN(lbfgs_pass_g<4>), ST2_KFN(lbfgs_pass_y<1>), ST2_KFN(lbfgs_pass_y<4>),
    ST2_KFN(lbfgs_take_g), ST2_KFN(lbfgs_coefficients), ST2_KFN(lbfgs_clear_sums), ST2_KFN(lbfgs_accept),
    ST2_KFN(lbfgs_store_y_dots)});

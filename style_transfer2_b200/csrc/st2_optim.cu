// Optimizer kernels: L-BFGS two-loop recursion and Adam as fused, bandwidth-bound passes.
// Reference: optimizers.py:7-125, utils.py:29-69.  All vectors fp32, length n = 3*H*W.
//
// L-BFGS keeps the whole recursion on the device: history ring (S, Y), s.y, y.y, the alphas and
// every intermediate dot product live in device memory, so one optimizer step is a fixed sequence
// of launches with no host synchronisation (CUDA-graph friendly).  Each launch fuses the axpy of
// one history step with the dot product the next step needs:
//     loop 1 (newest -> oldest):  q <- q - alpha_i Y_i   fused with   S_{i-1} . q
//     loop 2 (oldest -> newest):  q <- q + (alpha_i - beta_i) S_i   fused with   Y_{i+1} . q
// which moves 16 bytes per element per history step instead of the reference's 20 + 8.
#include "st2_common.cuh"

#include <math.h>
#include <string.h>

namespace {

constexpr int kThreads = 256;
constexpr int kVec = 4;

inline int grid_for(long long n, int sm_count) {
  long long blocks = (n + (long long)kThreads * kVec - 1) / ((long long)kThreads * kVec);
  long long cap = (long long)sm_count * 6;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// ------------------------------------------------------------------------------------------------
struct LbfgsDev {
  int count;                       // pairs held (<= n_corr)
  int head;                        // physical slot of the oldest pair
  int pad[2];
  double sy[ST2_MAX_HIST];         // per physical slot
  double yy[ST2_MAX_HIST];
  double alpha[ST2_MAX_HIST];      // per logical index
  double acc1[ST2_MAX_HIST + 1];   // loop-1 dot products: acc1[j] feeds kernel j
  double acc_mid;                  // Y_0 . q after loop 1
  double acc2[ST2_MAX_HIST + 1];   // loop-2 dot products: acc2[i] feeds kernel i (i >= 1)
  double acc_sy, acc_yy;           // commit
};

__device__ __forceinline__ const float* slot_ptr(const float* base, long long n, int slots, int head,
                                                 int logical) {
  return base + (long long)((head + logical) % slots) * n;
}

// V floats per access: 4 (16-byte vectors, when n % 4 == 0) or 1
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

#define ST2_PACK_LOOP(k, npk) \
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < (npk); k += (long long)gridDim.x * blockDim.x)

// acc1[0] = (count > 0) ? S_newest . g : g . g
template <int V>
__global__ void lbfgs_first_dot(LbfgsDev* st_, const float* __restrict__ S, const float* __restrict__ g,
                                long long n, int slots) {
  const int count = st_->count;
  const float* a = count > 0 ? slot_ptr(S, n, slots, st_->head, count - 1) : g;
  float acc = 0.f;
  ST2_PACK_LOOP(k, n / V) {
    const Pack<V> av = ld<V>(a, k), gv = ld<V>(g, k);
#pragma unroll
    for (int e = 0; e < V; ++e) acc = fmaf(av.v[e], gv.v[e], acc);
  }
  float v[1] = {acc};
  double* dst[1] = {&st_->acc1[0]};
  block_accumulate<1>(v, dst);
}

// loop-1 step j (logical i = count-1-j).  q_out = src - alpha_i * Y_i, then the next dot.
template <int V>
__global__ void lbfgs_loop1(LbfgsDev* st_, const float* __restrict__ S, const float* __restrict__ Y,
                            const float* __restrict__ g, float* __restrict__ q, long long n, int slots,
                            int j) {
  const int count = st_->count;
  const int i = count - 1 - j;
  if (i < 0) return;
  const int head = st_->head;
  const int phys = (head + i) % slots;
  const double alpha_d = st_->acc1[j] / st_->sy[phys];
  if (blockIdx.x == 0 && threadIdx.x == 0) st_->alpha[i] = alpha_d;
  const float na = -(float)alpha_d;
  const float* src = (j == 0) ? g : q;
  const float* y = Y + (long long)phys * n;
  const float* nxt = (i > 0) ? slot_ptr(S, n, slots, head, i - 1) : slot_ptr(Y, n, slots, head, 0);
  float acc = 0.f;
  ST2_PACK_LOOP(k, n / V) {
    const Pack<V> yv = ld<V>(y, k), sv = ld<V>(src, k), nv = ld<V>(nxt, k);
    Pack<V> o;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      o.v[e] = fmaf(na, yv.v[e], sv.v[e]);
      acc = fmaf(nv.v[e], o.v[e], acc);
    }
    st<V>(q, k, o);
  }
  float v[1] = {acc};
  double* dst[1] = {(i > 0) ? &st_->acc1[j + 1] : &st_->acc_mid};
  block_accumulate<1>(v, dst);
}

// no history: q = g / sqrt(g.g / n); s = -step q; x += s   (optimizers.py:100-102, 67-69)
template <int V>
__global__ void lbfgs_cold_step(LbfgsDev* st_, float* __restrict__ S, const float* __restrict__ g,
                                float* __restrict__ x, long long n, int slots, float step) {
  if (st_->count != 0) return;
  const float scale = (float)sqrt(st_->acc1[0] / (double)n);
  float* s_new = S + (long long)(st_->head % slots) * n;
  ST2_PACK_LOOP(k, n / V) {
    const Pack<V> gv = ld<V>(g, k);
    Pack<V> xv = ld<V>(x, k), sv;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float qv = gv.v[e] / scale;
      sv.v[e] = -step * qv;
      xv.v[e] += sv.v[e];
    }
    st<V>(s_new, k, sv);
    st<V>(x, k, xv);
  }
}

// loop-2 step i.  i == 0 applies the H0 scaling gamma = s.y / y.y of the newest pair first.
template <int V>
__global__ void lbfgs_loop2(LbfgsDev* st_, float* __restrict__ S, const float* __restrict__ Y,
                            float* __restrict__ q, float* __restrict__ x, long long n, int slots, int i,
                            float step) {
  const int count = st_->count;
  if (i >= count) return;
  const int head = st_->head;
  const int phys = (head + i) % slots;
  const int newest = (head + count - 1) % slots;
  const float gamma = (float)(st_->sy[newest] / st_->yy[newest]);
  const double ydotq = (i == 0) ? st_->acc_mid * (double)gamma : st_->acc2[i];
  const double beta = ydotq / st_->sy[phys];
  const float coef = (float)(st_->alpha[i] - beta);
  const float pre = (i == 0) ? gamma : 1.0f;
  const float* s = S + (long long)phys * n;
  const bool last = (i == count - 1);
  const float* nxt = last ? s : slot_ptr(Y, n, slots, head, i + 1);
  float* s_new = S + (long long)((head + count) % slots) * n;
  float acc = 0.f;
  ST2_PACK_LOOP(k, n / V) {
    Pack<V> qv = ld<V>(q, k);
    const Pack<V> sv = ld<V>(s, k);
#pragma unroll
    for (int e = 0; e < V; ++e) qv.v[e] = fmaf(coef, sv.v[e], i == 0 ? qv.v[e] * pre : qv.v[e]);
    if (last) {
      Pack<V> xv = ld<V>(x, k), dv;
#pragma unroll
      for (int e = 0; e < V; ++e) { dv.v[e] = -step * qv.v[e]; xv.v[e] += dv.v[e]; }
      st<V>(s_new, k, dv);
      st<V>(x, k, xv);
    } else {
      const Pack<V> nv = ld<V>(nxt, k);
#pragma unroll
      for (int e = 0; e < V; ++e) acc = fmaf(nv.v[e], qv.v[e], acc);
      st<V>(q, k, qv);
    }
  }
  if (!last) {
    float v[1] = {acc};
    double* dst[1] = {&st_->acc2[i + 1]};
    block_accumulate<1>(v, dst);
  }
}

// y = g_new - g_prev into the staging slot; s.y and y.y
template <int V>
__global__ void lbfgs_make_pair(LbfgsDev* st_, const float* __restrict__ S, float* __restrict__ Y,
                                const float* __restrict__ g_new, const float* __restrict__ g_prev,
                                long long n, int slots) {
  const int phys = (st_->head + st_->count) % slots;
  const float* s = S + (long long)phys * n;
  float* y = Y + (long long)phys * n;
  float a_sy = 0.f, a_yy = 0.f;
  ST2_PACK_LOOP(k, n / V) {
    const Pack<V> a = ld<V>(g_new, k), b = ld<V>(g_prev, k), sv = ld<V>(s, k);
    Pack<V> yv;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      yv.v[e] = a.v[e] - b.v[e];
      a_sy = fmaf(sv.v[e], yv.v[e], a_sy);
      a_yy = fmaf(yv.v[e], yv.v[e], a_yy);
    }
    st<V>(y, k, yv);
  }
  float v[2] = {a_sy, a_yy};
  double* dst[2] = {&st_->acc_sy, &st_->acc_yy};
  block_accumulate<2>(v, dst);
}

// optimizers.py:79-87: keep the pair iff s.y > 1e-10, drop the oldest beyond n_corr
__global__ void lbfgs_accept(LbfgsDev* st, int slots, int n_corr) {
  const double sy = (double)(float)st->acc_sy;       // utils.dot returns an fp32-valued float
  if (sy > 1e-10) {
    const int phys = (st->head + st->count) % slots;
    st->sy[phys] = sy;
    st->yy[phys] = (double)(float)st->acc_yy;
    if (st->count == n_corr) st->head = (st->head + 1) % slots;
    else st->count += 1;
  }
}

__global__ void lbfgs_clear_acc(LbfgsDev* st) {
  for (int i = threadIdx.x; i <= ST2_MAX_HIST; i += blockDim.x) { st->acc1[i] = 0.0; st->acc2[i] = 0.0; }
  if (threadIdx.x == 0) { st->acc_mid = 0.0; st->acc_sy = 0.0; st->acc_yy = 0.0; }
}

__global__ void lbfgs_yy_of_slot(LbfgsDev* st, const float* __restrict__ Y, long long n, int phys) {
  const float* y = Y + (long long)phys * n;
  float acc = 0.f;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x)
    acc += y[k] * y[k];
  float v[1] = {acc};
  double* dst[1] = {&st->yy[phys]};
  block_accumulate<1>(v, dst);
}

// ------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m1,
                            float* __restrict__ m2, long long n, float step, float b1, float omb1,
                            float b2, float omb2, float c1, float c2) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x) {
    const float gv = g[k];
    const float a = b1 * m1[k] + omb1 * gv;            // utils.py:58-60
    const float b = b2 * m2[k] + omb2 * (gv * gv);
    m1[k] = a;
    m2[k] = b;
    const float num = step * (a / c1);                 // optimizers.py:26
    const float den = sqrtf(b / c2) + 1e-8f;
    x[k] -= num / den;
  }
}

__global__ void dot_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                           double* out) {
  float acc = 0.f;
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x)
    acc += a[k] * b[k];
  float v[1] = {acc};
  double* dst[1] = {out};
  block_accumulate<1>(v, dst);
}

__global__ void axpy_kernel(float alpha, const float* __restrict__ x, float* __restrict__ y,
                            long long n) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n;
       k += (long long)gridDim.x * blockDim.x)
    y[k] = fmaf(alpha, x[k], y[k]);
}

}  // namespace

struct st2_lbfgs {
  st2_ctx* ctx;
  long long n;
  int n_corr, slots;
  int count_ub;                    // host-side upper bound of the device pair count
  float *S, *Y, *q;
  LbfgsDev* st;
};

extern "C" {

int st2_lbfgs_create(st2_ctx* ctx, long long n, int n_corr, st2_lbfgs** out) {
  if (!ctx || !out || n <= 0 || n_corr < 1 || n_corr + 1 > ST2_MAX_HIST)
    return st2_fail(ctx, ST2_ERR_ARG, "st2_lbfgs_create: bad arguments (n=%lld n_corr=%d)", n, n_corr);
  st2_lbfgs* o = new st2_lbfgs();
  o->ctx = ctx; o->n = n; o->n_corr = n_corr; o->slots = n_corr + 1; o->count_ub = 0;
  ST2_CUDA(ctx, cudaMalloc(&o->S, sizeof(float) * n * o->slots));
  ST2_CUDA(ctx, cudaMalloc(&o->Y, sizeof(float) * n * o->slots));
  ST2_CUDA(ctx, cudaMalloc(&o->q, sizeof(float) * n));
  ST2_CUDA(ctx, cudaMalloc(&o->st, sizeof(LbfgsDev)));
  ST2_CUDA(ctx, cudaMemsetAsync(o->st, 0, sizeof(LbfgsDev), ctx->stream));
  *out = o;
  return 0;
}

void st2_lbfgs_destroy(st2_lbfgs* o) {
  if (!o) return;
  cudaFree(o->S); cudaFree(o->Y); cudaFree(o->q); cudaFree(o->st);
  delete o;
}

int st2_lbfgs_reset(st2_lbfgs* o) {
  if (!o) return ST2_ERR_ARG;
  ST2_CUDA(o->ctx, cudaMemsetAsync(o->st, 0, sizeof(LbfgsDev), o->ctx->stream));
  o->count_ub = 0;
  return 0;
}

int st2_lbfgs_advance(st2_lbfgs* o, float* x, const float* g, float step) {
  if (!o || !x || !g) return st2_fail(o ? o->ctx : nullptr, ST2_ERR_ARG, "st2_lbfgs_advance: null");
  st2_ctx* ctx = o->ctx;
  cudaStream_t s = ctx->stream;
  const int grid = grid_for(o->n, ctx->sm_count);
  ProfScope ps(ctx, 7);
  lbfgs_clear_acc<<<1, 32, 0, s>>>(o->st);
  ST2_LAUNCH_CHECK(ctx);
  if (o->n % 4 == 0) lbfgs_first_dot<4><<<grid, kThreads, 0, s>>>(o->st, o->S, g, o->n, o->slots);
  else lbfgs_first_dot<1><<<grid, kThreads, 0, s>>>(o->st, o->S, g, o->n, o->slots);
  ST2_LAUNCH_CHECK(ctx);
  for (int j = 0; j < o->count_ub; ++j) {
    if (o->n % 4 == 0) lbfgs_loop1<4><<<grid, kThreads, 0, s>>>(o->st, o->S, o->Y, g, o->q, o->n, o->slots, j);
    else lbfgs_loop1<1><<<grid, kThreads, 0, s>>>(o->st, o->S, o->Y, g, o->q, o->n, o->slots, j);
    ST2_LAUNCH_CHECK(ctx);
  }
  // count may be 0 on the device even when the host bound is > 0 (all pairs rejected)
  if (o->n % 4 == 0) lbfgs_cold_step<4><<<grid, kThreads, 0, s>>>(o->st, o->S, g, x, o->n, o->slots, step);
  else lbfgs_cold_step<1><<<grid, kThreads, 0, s>>>(o->st, o->S, g, x, o->n, o->slots, step);
  ST2_LAUNCH_CHECK(ctx);
  for (int i = 0; i < o->count_ub; ++i) {
    if (o->n % 4 == 0) lbfgs_loop2<4><<<grid, kThreads, 0, s>>>(o->st, o->S, o->Y, o->q, x, o->n, o->slots, i, step);
    else lbfgs_loop2<1><<<grid, kThreads, 0, s>>>(o->st, o->S, o->Y, o->q, x, o->n, o->slots, i, step);
    ST2_LAUNCH_CHECK(ctx);
  }
  return 0;
}

int st2_lbfgs_commit(st2_lbfgs* o, const float* g_new, const float* g_prev) {
  if (!o || !g_new || !g_prev) return st2_fail(o ? o->ctx : nullptr, ST2_ERR_ARG, "st2_lbfgs_commit: null");
  st2_ctx* ctx = o->ctx;
  cudaStream_t s = ctx->stream;
  const int grid = grid_for(o->n, ctx->sm_count);
  ProfScope ps(ctx, 7);
  if (o->n % 4 == 0) lbfgs_make_pair<4><<<grid, kThreads, 0, s>>>(o->st, o->S, o->Y, g_new, g_prev, o->n, o->slots);
  else lbfgs_make_pair<1><<<grid, kThreads, 0, s>>>(o->st, o->S, o->Y, g_new, g_prev, o->n, o->slots);
  ST2_LAUNCH_CHECK(ctx);
  lbfgs_accept<<<1, 1, 0, s>>>(o->st, o->slots, o->n_corr);
  ST2_LAUNCH_CHECK(ctx);
  if (o->count_ub < o->n_corr) o->count_ub++;
  return 0;
}

int st2_lbfgs_load(st2_lbfgs* o, int count, const float* s_dev, const float* y_dev, const double* sy_host) {
  if (!o || count < 0 || count > o->n_corr) return st2_fail(o ? o->ctx : nullptr, ST2_ERR_ARG, "st2_lbfgs_load: bad count");
  st2_ctx* ctx = o->ctx;
  cudaStream_t s = ctx->stream;
  LbfgsDev h;
  memset(&h, 0, sizeof(h));
  h.count = count; h.head = 0;
  for (int i = 0; i < count; ++i) h.sy[i] = sy_host[i];
  ST2_CUDA(ctx, cudaStreamSynchronize(s));
  ST2_CUDA(ctx, cudaMemcpy(o->st, &h, sizeof(h), cudaMemcpyHostToDevice));
  if (count) {
    ST2_CUDA(ctx, cudaMemcpyAsync(o->S, s_dev, sizeof(float) * o->n * count, cudaMemcpyDeviceToDevice, s));
    ST2_CUDA(ctx, cudaMemcpyAsync(o->Y, y_dev, sizeof(float) * o->n * count, cudaMemcpyDeviceToDevice, s));
  }
  const int grid = grid_for(o->n, ctx->sm_count);
  for (int i = 0; i < count; ++i) {
    lbfgs_yy_of_slot<<<grid, kThreads, 0, s>>>(o->st, o->Y, o->n, i);
    ST2_LAUNCH_CHECK(ctx);
  }
  o->count_ub = count;
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
    const int phys = (h.head + i) % o->slots;
    if (sy_host) sy_host[i] = h.sy[phys];
    if (s_dev) ST2_CUDA(ctx, cudaMemcpyAsync(s_dev + (long long)i * o->n, o->S + (long long)phys * o->n,
                                             sizeof(float) * o->n, cudaMemcpyDeviceToDevice, s));
    if (y_dev) ST2_CUDA(ctx, cudaMemcpyAsync(y_dev + (long long)i * o->n, o->Y + (long long)phys * o->n,
                                             sizeof(float) * o->n, cudaMemcpyDeviceToDevice, s));
  }
  ST2_CUDA(ctx, cudaStreamSynchronize(s));
  return 0;
}

int st2_adam_step(st2_ctx* ctx, float* x, const float* g, float* m1, float* m2, long long n,
                  float step, double b1, double b2, int items1, int items2) {
  if (!ctx || !x || !g || !m1 || !m2 || n <= 0 || items1 < 1 || items2 < 1)
    return st2_fail(ctx, ST2_ERR_ARG, "st2_adam_step: bad arguments");
  // utils.py:58-64: the python-float decay constants act as weak scalars on fp32 arrays, i.e.
  // they are evaluated in double and rounded to fp32 once.
  const float fb1 = (float)b1, fb2 = (float)b2;
  const float omb1 = (float)(1.0 - b1), omb2 = (float)(1.0 - b2);
  const float c1 = (float)(1.0 - pow(b1, (double)items1));
  const float c2 = (float)(1.0 - pow(b2, (double)items2));
  ProfScope ps(ctx, 7);
  adam_kernel<<<grid_for(n, ctx->sm_count), kThreads, 0, ctx->stream>>>(x, g, m1, m2, n, step, fb1, omb1,
                                                                        fb2, omb2, c1, c2);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

static int reduce_to_host(st2_ctx* ctx, const float* a, const float* b, long long n, double* host_out) {
  static double* scratch = nullptr;          // one double per process is enough (single stream)
  if (!scratch) ST2_CUDA(ctx, cudaMalloc(&scratch, sizeof(double)));
  ST2_CUDA(ctx, cudaMemsetAsync(scratch, 0, sizeof(double), ctx->stream));
  dot_kernel<<<grid_for(n, ctx->sm_count), kThreads, 0, ctx->stream>>>(a, b, n, scratch);
  ST2_LAUNCH_CHECK(ctx);
  ST2_CUDA(ctx, cudaMemcpyAsync(host_out, scratch, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ST2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return 0;
}

int st2_dot(st2_ctx* ctx, const float* a, const float* b, long long n, double* host_out) {
  if (!ctx || !a || !b || !host_out || n <= 0) return st2_fail(ctx, ST2_ERR_ARG, "st2_dot: bad arguments");
  return reduce_to_host(ctx, a, b, n, host_out);
}

int st2_sumsq(st2_ctx* ctx, const float* a, long long n, double* host_out) {
  if (!ctx || !a || !host_out || n <= 0) return st2_fail(ctx, ST2_ERR_ARG, "st2_sumsq: bad arguments");
  return reduce_to_host(ctx, a, a, n, host_out);
}

int st2_axpy(st2_ctx* ctx, float alpha, const float* x, float* y, long long n) {
  if (!ctx || !x || !y || n <= 0) return st2_fail(ctx, ST2_ERR_ARG, "st2_axpy: bad arguments");
  axpy_kernel<<<grid_for(n, ctx->sm_count), kThreads, 0, ctx->stream>>>(alpha, x, y, n);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

}  // extern "C"

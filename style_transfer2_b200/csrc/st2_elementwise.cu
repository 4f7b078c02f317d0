// 16-byte-vectorised versions of the bandwidth-bound layer kernels (max-pool fwd/bwd, loss combine,
// feature reductions).  Each falls back to the scalar kernel in st2_layers.cu when the channel
// count / element count / alignment does not allow 16-byte accesses (e.g. the 3-plane data blob at
// odd sizes).  Semantics are identical to the scalar versions.
#include "st2_kernels.h"

namespace {

constexpr int kThreads = 256;

template <typename T> struct Vec16 {};
template <> struct Vec16<__half> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __half* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(hp[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
  }
  static __device__ __forceinline__ void store(__half* p, const float (&v)[8]) {
    uint4 u;
    __half2* hp = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) hp[e] = h2_sat(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 u = *reinterpret_cast<const float4*>(p);
    v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

inline int vgrid(long long items, int sm_count) {
  long long b = (items + kThreads - 1) / kThreads, cap = (long long)sm_count * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename T>
__global__ void pool_fwd_vec_kernel(const T* __restrict__ in, T* __restrict__ out, int C, int H, int W, int Ho,
                                    int Wo) {
  pdl_trigger();
  pdl_wait();
  constexpr int N = Vec16<T>::N;
  const int cv = C / N;
  const long long total = (long long)Ho * Wo * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * N;
    const long long p = i / cv;
    const int wo = (int)(p % Wo), ho = (int)(p / Wo);
    const int h = ho * 2, w = wo * 2;
    const bool has_r = w + 1 < W, has_d = h + 1 < H;
    const T* b00 = in + ((long long)h * W + w) * C + c;
    float m[N], t[N];
    Vec16<T>::load(b00, m);
    if (has_r) {
      Vec16<T>::load(b00 + C, t);
#pragma unroll
      for (int e = 0; e < N; ++e) m[e] = fmaxf(m[e], t[e]);
    }
    if (has_d) {
      Vec16<T>::load(b00 + (long long)W * C, t);
#pragma unroll
      for (int e = 0; e < N; ++e) m[e] = fmaxf(m[e], t[e]);
      if (has_r) {
        Vec16<T>::load(b00 + (long long)W * C + C, t);
#pragma unroll
        for (int e = 0; e < N; ++e) m[e] = fmaxf(m[e], t[e]);
      }
    }
    Vec16<T>::store(out + p * C + c, m);
  }
}

// first maximum in (h, w) scan order wins (strict '>'), optional ReLU mask of the layer below
template <typename T>
__global__ void pool_bwd_vec_kernel(const T* __restrict__ act, const T* __restrict__ gp, T* __restrict__ gout,
                                    int C, int H, int W, int Ho, int Wo, int apply_mask) {
  pdl_trigger();
  pdl_wait();
  constexpr int N = Vec16<T>::N;
  const int cv = C / N;
  const long long total = (long long)Ho * Wo * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv) * N;
    const long long p = i / cv;
    const int wo = (int)(p % Wo), ho = (int)(p / Wo);
    const int h = ho * 2, w = wo * 2;
    const bool has_r = w + 1 < W, has_d = h + 1 < H;
    const long long o00 = ((long long)h * W + w) * C + c;
    const long long o01 = o00 + C, o10 = o00 + (long long)W * C, o11 = o10 + C;
    float best[N], t[N], g[N];
    int arg[N];
    Vec16<T>::load(act + o00, best);
#pragma unroll
    for (int e = 0; e < N; ++e) arg[e] = 0;
    if (has_r) {
      Vec16<T>::load(act + o01, t);
#pragma unroll
      for (int e = 0; e < N; ++e) if (t[e] > best[e]) { best[e] = t[e]; arg[e] = 1; }
    }
    if (has_d) {
      Vec16<T>::load(act + o10, t);
#pragma unroll
      for (int e = 0; e < N; ++e) if (t[e] > best[e]) { best[e] = t[e]; arg[e] = 2; }
      if (has_r) {
        Vec16<T>::load(act + o11, t);
#pragma unroll
        for (int e = 0; e < N; ++e) if (t[e] > best[e]) { best[e] = t[e]; arg[e] = 3; }
      }
    }
    Vec16<T>::load(gp + p * C + c, g);
    if (apply_mask) {
#pragma unroll
      for (int e = 0; e < N; ++e) if (!(best[e] > 0.f)) g[e] = 0.f;
    }
    float o[N];
#pragma unroll
    for (int e = 0; e < N; ++e) o[e] = arg[e] == 0 ? g[e] : 0.f;
    Vec16<T>::store(gout + o00, o);
    if (has_r) {
#pragma unroll
      for (int e = 0; e < N; ++e) o[e] = arg[e] == 1 ? g[e] : 0.f;
      Vec16<T>::store(gout + o01, o);
    }
    if (has_d) {
#pragma unroll
      for (int e = 0; e < N; ++e) o[e] = arg[e] == 2 ? g[e] : 0.f;
      Vec16<T>::store(gout + o10, o);
      if (has_r) {
#pragma unroll
        for (int e = 0; e < N; ++e) o[e] = arg[e] == 3 ? g[e] : 0.f;
        Vec16<T>::store(gout + o11, o);
      }
    }
  }
}

template <typename T>
__global__ void combine_vec_kernel(const T* __restrict__ gin, const T* __restrict__ act, const T* __restrict__ fc,
                                   const T* __restrict__ sraw, T* __restrict__ out, long long nvec, int apply_mask,
                                   const double* __restrict__ coef, float h_cc, float h_sc, float h_dc) {
  pdl_trigger();
  pdl_wait();
  constexpr int N = Vec16<T>::N;
  float cc = h_cc, sc = h_sc, dc = h_dc;
  if (coef != nullptr) { cc = (float)coef[0]; sc = (float)coef[1]; dc = (float)coef[2]; }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const long long o = i * N;
    float f[N], v[N], t[N];
    Vec16<T>::load(act + o, f);
    if (gin != nullptr) {
      Vec16<T>::load(gin + o, v);
      if (apply_mask) {
#pragma unroll
        for (int e = 0; e < N; ++e) if (!(f[e] > 0.f)) v[e] = 0.f;
      }
    } else {
#pragma unroll
      for (int e = 0; e < N; ++e) v[e] = 0.f;
    }
    if (fc != nullptr) {
      Vec16<T>::load(fc + o, t);
#pragma unroll
      for (int e = 0; e < N; ++e) v[e] = fmaf(cc, f[e] - t[e], v[e]);
    }
    if (sraw != nullptr) {
      Vec16<T>::load(sraw + o, t);
#pragma unroll
      for (int e = 0; e < N; ++e) v[e] = fmaf(sc, t[e], v[e]);
    }
    if (dc != 0.f) {
#pragma unroll
      for (int e = 0; e < N; ++e) v[e] = fmaf(dc, f[e], v[e]);
    }
    Vec16<T>::store(out + o, v);
  }
}

template <typename T>
__global__ void feature_sums_vec_kernel(const T* __restrict__ act, const T* __restrict__ fc, long long nvec,
                                        double* sum_diff_sq, double* sum_sq) {
  pdl_trigger();
  pdl_wait();
  constexpr int N = Vec16<T>::N;
  float a = 0.f, b = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    float f[N], t[N];
    Vec16<T>::load(act + i * N, f);
#pragma unroll
    for (int e = 0; e < N; ++e) b = fmaf(f[e], f[e], b);
    if (fc != nullptr) {
      Vec16<T>::load(fc + i * N, t);
#pragma unroll
      for (int e = 0; e < N; ++e) { const float d = f[e] - t[e]; a = fmaf(d, d, a); }
    }
  }
  float v[2] = {a, b};
  double* dst[2] = {fc != nullptr ? sum_diff_sq : nullptr, sum_sq};
  block_accumulate<2>(v, dst);
}

inline bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

template <typename T>
int launch_pool_fwd_v(st2_ctx* ctx, const T* in, T* out, int C, int H, int W) {
  if (C % Vec16<T>::N) return launch_pool_fwd<T>(ctx, in, out, C, H, W);
  const int Ho = pool_extent(H), Wo = pool_extent(W);
  st2_launch_pdl(ctx, true, pool_fwd_vec_kernel<T>, vgrid((long long)Ho * Wo * (C / Vec16<T>::N), ctx->sm_count), kThreads, 0,
                 in, out, C, H, W, Ho, Wo);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_pool_fwd_v<float>(st2_ctx*, const float*, float*, int, int, int);
template int launch_pool_fwd_v<__half>(st2_ctx*, const __half*, __half*, int, int, int);

template <typename T>
int launch_pool_bwd_v(st2_ctx* ctx, const T* act, const T* g_pool, T* g_out, int C, int H, int W, int apply_mask) {
  if (C % Vec16<T>::N) return launch_pool_bwd<T>(ctx, act, g_pool, g_out, C, H, W, apply_mask);
  const int Ho = pool_extent(H), Wo = pool_extent(W);
  st2_launch_pdl(ctx, true, pool_bwd_vec_kernel<T>, vgrid((long long)Ho * Wo * (C / Vec16<T>::N), ctx->sm_count), kThreads, 0,
                 act, g_pool, g_out, C, H, W, Ho, Wo, apply_mask);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_pool_bwd_v<float>(st2_ctx*, const float*, const float*, float*, int, int, int, int);
template int launch_pool_bwd_v<__half>(st2_ctx*, const __half*, const __half*, __half*, int, int, int, int);

template <typename T>
int launch_combine_v(st2_ctx* ctx, const CombineArgs& a) {
  constexpr int N = Vec16<T>::N;
  if (a.n % N || !aligned16(a.gin) || !aligned16(a.act) || !aligned16(a.fc) || !aligned16(a.sraw) || !aligned16(a.out))
    return launch_combine<T>(ctx, a);
  st2_launch_pdl(ctx, true, combine_vec_kernel<T>, vgrid(a.n / N, ctx->sm_count), kThreads, 0, (const T*)a.gin,
                 (const T*)a.act, (const T*)a.fc, (const T*)a.sraw, (T*)a.out, a.n / N, a.apply_mask, a.coef, a.h_cc,
                 a.h_sc, a.h_dc);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_combine_v<float>(st2_ctx*, const CombineArgs&);
template int launch_combine_v<__half>(st2_ctx*, const CombineArgs&);

template <typename T>
int launch_feature_sums_v(st2_ctx* ctx, const T* act, const T* fc, long long n, double* sum_diff_sq, double* sum_sq) {
  constexpr int N = Vec16<T>::N;
  if (n % N || !aligned16(act) || !aligned16(fc)) return launch_feature_sums<T>(ctx, act, fc, n, sum_diff_sq, sum_sq);
  st2_launch_pdl(ctx, true, feature_sums_vec_kernel<T>, vgrid(n / N / 4 + 1, ctx->sm_count), kThreads, 0, act, fc, n / N,
                 sum_diff_sq, sum_sq);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_feature_sums_v<float>(st2_ctx*, const float*, const float*, long long, double*, double*);
template int launch_feature_sums_v<__half>(st2_ctx*, const __half*, const __half*, long long, double*, double*);

static St2KernelReg g_reg_elementwise({
    ST2_KFN(pool_fwd_vec_kernel<float>), ST2_KFN(pool_fwd_vec_kernel<__half>), ST2_KFN(pool_bwd_vec_kernel<float>),
    ST2_KFN(pool_bwd_vec_kernel<__half>), ST2_KFN(combine_vec_kernel<float>), ST2_KFN(combine_vec_kernel<__half>),
    ST2_KFN(feature_sums_vec_kernel<float>), ST2_KFN(feature_sums_vec_kernel<__half>)});

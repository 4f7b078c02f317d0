// CUDA-core kernels of the hot path: first/last-layer convolutions, the exact fp32 convolution
// used by ST2_PREC_FP32, ceil-mode max-pool and its backward, the loss "combine" pass, feature
// reductions, the strided Gram / style-gradient contractions and layout conversions.
// Reference semantics: worker.py:77-106 (model seam), 231-301 (objective); Caffe layers [ext].
#include "st2_kernels.h"

namespace {

constexpr int kThreads = 256;

inline int ew_grid(long long work_items, int sm_count) {
  long long blocks = (work_items + kThreads - 1) / kThreads;
  long long cap = (long long)sm_count * 16;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// =============================================================================== conv1_1 forward
// Work item: 256 consecutive pixels of one row.  Thread = (2 pixels, 32 of the 64 couts): every
// warp-broadcast float4 of weights read from shared memory feeds 8 FMAs (the kernel is bound by the
// shared-memory pipe otherwise), 27 inputs per pixel in registers.  fp16 output is staged through
// an XOR-swizzled shared tile so global stores are whole 128-byte lines; fp32 output is written
// directly (128 contiguous bytes per thread).  1728 FMA per pixel.
template <typename T>
__global__ void __launch_bounds__(256)
conv_first_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                      const float* __restrict__ bias, T* __restrict__ out, int H, int W, long long xps,
                      int lo, int hi) {
  // x: row 0 of plane 0 of the H local rows; planes xps floats apart; rows -lo .. H-1+hi addressable
  // (halo rows of a row strip), anything outside is the zero pad.
  __shared__ __align__(16) float sw[27 * 64];          // [tap][ci][co]
  __shared__ __align__(16) float sb[64];
  __shared__ float sx[3][3][258];
  __shared__ uint4 so[sizeof(T) == 2 ? 128 * 8 : 1];
  for (int i = threadIdx.x; i < 27 * 64; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
  const int tiles_w = (W + 255) / 256;
  const int items = H * tiles_w;
  const int px = threadIdx.x & 127, half = threadIdx.x >> 7;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int h = item / tiles_w;
    const int w0 = (item - h * tiles_w) * 256;
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 3 * 258; i += blockDim.x) {
      const int c = i / (3 * 258), r = (i / 258) % 3, col = i % 258;
      const int hh = h + r - 1, ww = w0 + col - 1;
      float v = 0.f;
      if (hh >= -lo && hh < H + hi && ww >= 0 && ww < W) v = __ldg(&x[(long long)c * xps + (long long)hh * W + ww]);
      sx[c][r][col] = v;
    }
    __syncthreads();
    float acc0[32], acc1[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc0[j] = acc1[j] = sb[half * 32 + j];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float v0 = sx[c][r][px + s], v1 = sx[c][r][px + 128 + s];
          const float4* wr = reinterpret_cast<const float4*>(&sw[((r * 3 + s) * 3 + c) * 64 + half * 32]);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 wv = wr[q];
            acc0[4 * q + 0] = fmaf(v0, wv.x, acc0[4 * q + 0]);
            acc0[4 * q + 1] = fmaf(v0, wv.y, acc0[4 * q + 1]);
            acc0[4 * q + 2] = fmaf(v0, wv.z, acc0[4 * q + 2]);
            acc0[4 * q + 3] = fmaf(v0, wv.w, acc0[4 * q + 3]);
            acc1[4 * q + 0] = fmaf(v1, wv.x, acc1[4 * q + 0]);
            acc1[4 * q + 1] = fmaf(v1, wv.y, acc1[4 * q + 1]);
            acc1[4 * q + 2] = fmaf(v1, wv.z, acc1[4 * q + 2]);
            acc1[4 * q + 3] = fmaf(v1, wv.w, acc1[4 * q + 3]);
          }
        }
    if (sizeof(T) == 4) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int ww = w0 + px + k * 128;
        if (ww < W) {
          const float* a = k ? acc1 : acc0;
          float4* op = reinterpret_cast<float4*>(out + ((long long)h * W + ww) * 64 + half * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            op[q] = make_float4(fmaxf(a[4 * q], 0.f), fmaxf(a[4 * q + 1], 0.f), fmaxf(a[4 * q + 2], 0.f),
                                fmaxf(a[4 * q + 3], 0.f));
        }
      }
    } else {
      // two rounds of 128 pixels through the 16 KB staging tile
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float* a = k ? acc1 : acc0;
        if (k) __syncthreads();
        uint4* row = so + px * 8;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          __half2* hp = reinterpret_cast<__half2*>(&u);
#pragma unroll
          for (int e = 0; e < 4; ++e)
            hp[e] = h2_relu_sat(a[8 * q + 2 * e], a[8 * q + 2 * e + 1]);
          row[(half * 4 + q) ^ (px & 7)] = u;
        }
        __syncthreads();
        const int base_w = w0 + k * 128;
        const int npx = min(128, W - base_w);
        if (npx > 0) {
          uint4* dst = reinterpret_cast<uint4*>(out + ((long long)h * W + base_w) * 64);
          for (int i = threadIdx.x; i < npx * 8; i += blockDim.x) {
            const int pp = i >> 3, c = i & 7;
            dst[i] = so[pp * 8 + (c ^ (pp & 7))];
          }
        }
      }
    }
  }
}

// =============================================================================== conv1_1 dgrad
// gx[ci][h][w] = sum_{tap,co} g[h+1-r][w+1-s][co] * W[co][ci][r][s];  w_bwd = [tap'][co][ci] with
// tap' already flipped so the kernel reads g at (h + r' - 1, w + s' - 1).
// A block owns an 8 x 16 pixel tile, one warp per row.  Eight lanes share a pixel, each owning 8 of
// the 64 channels, so a warp's 16-byte loads cover four whole 128-byte pixel rows (coalesced, and
// the 3 x 3 neighbourhood re-hits L1).  Each lane handles 4 pixels of the row per tap so that the
// 24 weights it fetches from shared memory (bank-conflict-free padding) feed 96 FMAs -- the kernel
// is bound by the L1/shared pipe otherwise.  Outputs are reduced over the 8 lanes with shuffles.
template <typename T> struct Load8 {};
template <> struct Load8<__half> {
  typedef uint4 raw_t;
  static __device__ __forceinline__ raw_t load_raw(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  static __device__ __forceinline__ void cvt(const raw_t& u, float (&v)[8]) {
    const __half2* hp = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(hp[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
  }
};
struct Float8 { float4 a, b; };
template <> struct Load8<float> {
  typedef Float8 raw_t;
  static __device__ __forceinline__ raw_t load_raw(const float* p) {
    raw_t r;
    r.a = __ldg(reinterpret_cast<const float4*>(p));
    r.b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    return r;
  }
  static __device__ __forceinline__ void cvt(const raw_t& r, float (&v)[8]) {
    v[0] = r.a.x; v[1] = r.a.y; v[2] = r.a.z; v[3] = r.a.w; v[4] = r.b.x; v[5] = r.b.y; v[6] = r.b.z; v[7] = r.b.w;
  }
};

template <typename T>
__global__ void __launch_bounds__(256)
conv_first_bwd_kernel(const T* __restrict__ g, const float* __restrict__ w, float* __restrict__ gx, int H,
                      int W, int lo, int hi) {
  __shared__ __align__(16) float sw[9][8][28];          // [tap][channel group][8 co x 3 ci (+4 pad)]
  for (int i = threadIdx.x; i < 9 * 64 * 3; i += blockDim.x) {
    const int ci = i % 3, co = (i / 3) % 64, tap = i / 192;
    sw[tap][co >> 3][(co & 7) * 3 + ci] = w[i];
  }
  __syncthreads();
  const long long HW = (long long)H * W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane >> 3, cgp = lane & 7;
  const int tiles_w = (W + 15) / 16, tiles_h = (H + 7) / 8;
  const long long n_tiles = (long long)tiles_w * tiles_h;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int th = (int)(tile / tiles_w), tw = (int)(tile - (long long)th * tiles_w);
    const int h = th * 8 + warp;                       // warp-uniform
    const int wbase = tw * 16 + sub;                   // this lane's pixels: wbase + 4 j
    float acc[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j][0] = acc[j][1] = acc[j][2] = 0.f;
    if (h < H) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int hh = h + r - 1;
        if (hh < -lo || hh >= H + hi) continue;        // warp-uniform
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          typename Load8<T>::raw_t raw[4];
          bool ok[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int wc = wbase + 4 * j + s - 1;
            ok[j] = wc >= 0 && wc < W;
            const int wcl = min(max(wc, 0), W - 1);
            raw[j] = Load8<T>::load_raw(g + ((long long)hh * W + wcl) * 64 + cgp * 8);
          }
          float wv[24];
          const float4* wr = reinterpret_cast<const float4*>(&sw[r * 3 + s][cgp][0]);
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            const float4 t4 = wr[q];
            wv[4 * q] = t4.x; wv[4 * q + 1] = t4.y; wv[4 * q + 2] = t4.z; wv[4 * q + 3] = t4.w;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v[8];
            Load8<T>::cvt(raw[j], v);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float ve = ok[j] ? v[e] : 0.f;
              acc[j][0] = fmaf(ve, wv[3 * e + 0], acc[j][0]);
              acc[j][1] = fmaf(ve, wv[3 * e + 1], acc[j][1]);
              acc[j][2] = fmaf(ve, wv[3 * e + 2], acc[j][2]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float a = acc[j][c];
        a += __shfl_xor_sync(0xffffffffu, a, 4);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        acc[j][c] = a;
      }
    if (h < H && cgp < 4) {
      // lane cgp of each 8-lane group writes pixel j = cgp (select without dynamic register indexing)
      float o0 = acc[0][0], o1 = acc[0][1], o2 = acc[0][2];
      if (cgp == 1) { o0 = acc[1][0]; o1 = acc[1][1]; o2 = acc[1][2]; }
      if (cgp == 2) { o0 = acc[2][0]; o1 = acc[2][1]; o2 = acc[2][2]; }
      if (cgp == 3) { o0 = acc[3][0]; o1 = acc[3][1]; o2 = acc[3][2]; }
      const int wo = wbase + 4 * cgp;
      if (wo < W) {
        const long long p = (long long)h * W + wo;
        gx[p] = o0;
        gx[HW + p] = o1;
        gx[2 * HW + p] = o2;
      }
    }
  }
}

// =============================================================================== exact fp32 conv
// Implicit GEMM on CUDA cores: block tile = 8x8 pixels x 64 couts, K step = 16 input channels of
// one tap, 4x4 register micro-tile per thread.
template <int EPI>
__global__ void __launch_bounds__(256)
conv_exact_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                  const float* __restrict__ act, float* __restrict__ out, int H, int W, int cin, int cout,
                  int lo, int hi) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64];
  const int tiles_w = (W + 7) / 8;
  const int th = blockIdx.x / tiles_w, tw = blockIdx.x % tiles_w;
  const int h0 = th * 8, w0 = tw * 8;
  const int co0 = blockIdx.y * 64;
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  // A loader: pixel lp, 4 consecutive input channels
  const int lp = t & 63, lk = (t >> 6) * 4;
  const int lph = h0 + (lp >> 3), lpw = w0 + (lp & 7);
  // B loader
  const int bk = t >> 4, bc = (t & 15) * 4;
  float acc[4][4] = {};
  for (int tap = 0; tap < 9; ++tap) {
    const int hh = lph + tap / 3 - 1, wc = lpw + tap % 3 - 1;
    const bool inb = hh >= -lo && hh < H + hi && wc >= 0 && wc < W;
    const float* ip = in + ((long long)hh * W + wc) * cin;
    const float* wp = w + (long long)tap * cin * cout;
    for (int c0 = 0; c0 < cin; c0 += 16) {
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
      if (inb) av = *reinterpret_cast<const float4*>(ip + c0 + lk);
      const float4 bv = *reinterpret_cast<const float4*>(wp + (long long)(c0 + bk) * cout + co0 + bc);
      __syncthreads();
      As[lk + 0][lp] = av.x; As[lk + 1][lp] = av.y; As[lk + 2][lp] = av.z; As[lk + 3][lp] = av.w;
      *reinterpret_cast<float4*>(&Bs[bk][bc]) = bv;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float aa[4] = {a.x, a.y, a.z, a.w};
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pl = ty * 4 + i;
    const int h = h0 + (pl >> 3), ww = w0 + (pl & 7);
    if (h >= H || ww >= W) continue;
    const long long o = ((long long)h * W + ww) * cout + co0 + tx * 4;
    float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    if (EPI == EPI_BIAS_RELU) {
      const float4 b = *reinterpret_cast<const float4*>(bias + co0 + tx * 4);
      v.x = fmaxf(v.x + b.x, 0.f); v.y = fmaxf(v.y + b.y, 0.f);
      v.z = fmaxf(v.z + b.z, 0.f); v.w = fmaxf(v.w + b.w, 0.f);
    } else if (EPI == EPI_MASK) {
      const float4 a = *reinterpret_cast<const float4*>(act + o);
      v.x = a.x > 0.f ? v.x : 0.f; v.y = a.y > 0.f ? v.y : 0.f;
      v.z = a.z > 0.f ? v.z : 0.f; v.w = a.w > 0.f ? v.w : 0.f;
    }
    *reinterpret_cast<float4*>(out + o) = v;
  }
}

// =============================================================================== max-pool 2x2/2 ceil
template <typename T>
__global__ void pool_fwd_kernel(const T* __restrict__ in, T* __restrict__ out, int C, int H, int W, int Ho,
                                int Wo) {
  const long long total = (long long)Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const int wo = (int)(p % Wo), ho = (int)(p / Wo);
    const int h = ho * 2, w = wo * 2;
    float m = to_f<T>(in[((long long)h * W + w) * C + c]);
    if (w + 1 < W) m = fmaxf(m, to_f<T>(in[((long long)h * W + w + 1) * C + c]));
    if (h + 1 < H) {
      m = fmaxf(m, to_f<T>(in[((long long)(h + 1) * W + w) * C + c]));
      if (w + 1 < W) m = fmaxf(m, to_f<T>(in[((long long)(h + 1) * W + w + 1) * C + c]));
    }
    out[i] = from_f<T>(m);
  }
}

// Caffe PoolingLayer backward: the whole diff of a window goes to its first maximum in (h, w) scan
// order (strict '>' update) [ext]; optional ReLU mask of the layer below (act > 0).
template <typename T>
__global__ void pool_bwd_kernel(const T* __restrict__ act, const T* __restrict__ gp, T* __restrict__ gout,
                                int C, int H, int W, int Ho, int Wo, int apply_mask) {
  const long long total = (long long)Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const int wo = (int)(p % Wo), ho = (int)(p / Wo);
    const int h = ho * 2, w = wo * 2;
    const bool has_r = w + 1 < W, has_d = h + 1 < H;
    const long long i00 = ((long long)h * W + w) * C + c;
    const long long i01 = i00 + C, i10 = i00 + (long long)W * C, i11 = i10 + C;
    float best = to_f<T>(act[i00]);
    int arg = 0;
    if (has_r) { const float v = to_f<T>(act[i01]); if (v > best) { best = v; arg = 1; } }
    if (has_d) {
      const float v = to_f<T>(act[i10]); if (v > best) { best = v; arg = 2; }
      if (has_r) { const float v2 = to_f<T>(act[i11]); if (v2 > best) { best = v2; arg = 3; } }
    }
    float g = to_f<T>(gp[i]);
    if (apply_mask && !(best > 0.f)) g = 0.f;
    const T z = from_f<T>(0.f), gv = from_f<T>(g);
    gout[i00] = arg == 0 ? gv : z;
    if (has_r) gout[i01] = arg == 1 ? gv : z;
    if (has_d) {
      gout[i10] = arg == 2 ? gv : z;
      if (has_r) gout[i11] = arg == 3 ? gv : z;
    }
  }
}

// =============================================================================== loss combine
template <typename T>
__global__ void combine_kernel(const T* __restrict__ gin, const T* __restrict__ act, const T* __restrict__ fc,
                               const T* __restrict__ sraw, T* __restrict__ out, long long n, int apply_mask,
                               const double* __restrict__ coef, float h_cc, float h_sc, float h_dc) {
  float cc = h_cc, sc = h_sc, dc = h_dc;
  if (coef != nullptr) { cc = (float)coef[0]; sc = (float)coef[1]; dc = (float)coef[2]; }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float f = to_f<T>(act[i]);
    float v = 0.f;
    if (gin != nullptr) {
      v = to_f<T>(gin[i]);
      if (apply_mask && !(f > 0.f)) v = 0.f;
    }
    if (fc != nullptr) v = fmaf(cc, f - to_f<T>(fc[i]), v);
    if (sraw != nullptr) v = fmaf(sc, to_f<T>(sraw[i]), v);
    if (dc != 0.f) v = fmaf(dc, f, v);
    out[i] = from_f<T>(v);
  }
}

template <typename T>
__global__ void feature_sums_kernel(const T* __restrict__ act, const T* __restrict__ fc, long long n,
                                    double* sum_diff_sq, double* sum_sq) {
  float a = 0.f, b = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float f = to_f<T>(act[i]);
    b = fmaf(f, f, b);
    if (fc != nullptr) { const float d = f - to_f<T>(fc[i]); a = fmaf(d, d, a); }
  }
  float v[2] = {a, b};
  double* dst[2] = {fc != nullptr ? sum_diff_sq : nullptr, sum_sq};
  block_accumulate<2>(v, dst);
}

// =============================================================================== strided Gram
// Block = 64 x 64 tile of G over one chunk of pixels; thread micro-tile 4 x 4.
template <typename T>
__global__ void __launch_bounds__(256)
gram_generic_kernel(const T* __restrict__ F, int C, long long HW, long long sp, long long sc,
                    double* __restrict__ Gd, long long chunk) {
  __shared__ __align__(16) float As[16][64 + 4];
  __shared__ __align__(16) float Bs[16][64 + 4];
  const int i0 = blockIdx.x * 64, j0 = blockIdx.y * 64;
  const long long p_begin = (long long)blockIdx.z * chunk;
  const long long p_end = min(HW, p_begin + chunk);
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  float acc[4][4] = {};
  for (long long p0 = p_begin; p0 < p_end; p0 += 16) {
    __syncthreads();
    for (int e = t; e < 16 * 64; e += 256) {
      const int ch = e & 63, pp = e >> 6;
      const long long p = p0 + pp;
      float a = 0.f, b = 0.f;
      if (p < p_end) {
        if (i0 + ch < C) a = to_f<T>(F[p * sp + (long long)(i0 + ch) * sc]);
        if (j0 + ch < C) b = to_f<T>(F[p * sp + (long long)(j0 + ch) * sc]);
      }
      As[pp][ch] = a;
      Bs[pp][ch] = b;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gi = i0 + ty * 4 + i, gj = j0 + tx * 4 + j;
      if (gi < C && gj < C) atomicAdd(&Gd[(long long)gi * C + gj], (double)acc[i][j]);
    }
}

__global__ void gram_finalize_kernel(const double* __restrict__ Gd, const float* __restrict__ A,
                                     float* __restrict__ D, int C, long long HW, double* sum_dsq) {
  const long long n = (long long)C * C;
  const float denom = (float)((double)C * (double)HW);         // np.float32(x.size), worker.py:114
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float g = (float)Gd[i] / denom;
    if (A != nullptr) g -= A[i];
    D[i] = g;
    acc = fmaf(g, g, acc);
  }
  float v[1] = {acc};
  double* dst[1] = {sum_dsq};
  block_accumulate<1>(v, dst);
}

// Row strips: the strip's un-normalised Gram sum as fp32 (what gets all-reduced), and the finalize from
// the all-reduced sum with the whole canvas' pixel count.
__global__ void gram_acc_to_f32_kernel(const double* __restrict__ Gd, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = (float)Gd[i];
}

__global__ void gram_from_sum_kernel(const float* __restrict__ Gs, const float* __restrict__ A,
                                     float* __restrict__ D, int C, double HW, double* sum_dsq) {
  const long long n = (long long)C * C;
  const float denom = (float)((double)C * HW);
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float g = Gs[i] / denom;
    if (A != nullptr) g -= A[i];
    D[i] = g;
    acc = fmaf(g, g, acc);
  }
  float v[1] = {acc};
  double* dst[1] = {sum_dsq};
  block_accumulate<1>(v, dst);
}

// raw[p, i] = sum_j D[i, j] F[p, j]; block = 64 pixels x 64 channels i, K step 16 channels j
template <typename T>
__global__ void __launch_bounds__(256)
style_grad_generic_kernel(const T* __restrict__ F, const float* __restrict__ D, T* __restrict__ raw, int C,
                          long long HW, long long sp, long long sc, double* sum_rawsq) {
  __shared__ __align__(16) float As[16][64 + 4];     // [j][pixel]
  __shared__ __align__(16) float Bs[16][64 + 4];     // [j][i]
  const long long p0 = (long long)blockIdx.x * 64;
  const int i0 = blockIdx.y * 64;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  float acc[4][4] = {};
  for (int j0 = 0; j0 < C; j0 += 16) {
    __syncthreads();
    for (int e = t; e < 16 * 64; e += 256) {
      const int jj = e & 15, q = e >> 4;              // q: pixel (A) or channel i (B)
      float a = 0.f, b = 0.f;
      if (j0 + jj < C) {
        if (p0 + q < HW) a = to_f<T>(F[(p0 + q) * sp + (long long)(j0 + jj) * sc]);
        if (i0 + q < C) b = D[(long long)(i0 + q) * C + j0 + jj];
      }
      As[jj][q] = a;
      Bs[jj][q] = b;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long p = p0 + ty * 4 + i;
      const int ci = i0 + tx * 4 + j;
      if (p < HW && ci < C) {
        raw[p * sp + (long long)ci * sc] = from_f<T>(acc[i][j]);
        ss = fmaf(acc[i][j], acc[i][j], ss);
      }
    }
  float v[1] = {ss};
  double* dst[1] = {sum_rawsq};
  block_accumulate<1>(v, dst);
}

// =============================================================================== layout conversion
template <typename T>
__global__ void export_nchw_kernel(const T* __restrict__ nhwc, float* __restrict__ nchw, int C, long long HW) {
  __shared__ float tile[32][33];
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long p = p0 + r;
    const int c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < HW && c < C) ? to_f<T>(nhwc[p * C + c]) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r;
    const long long p = p0 + threadIdx.x;
    if (p < HW && c < C) nchw[(long long)c * HW + p] = tile[threadIdx.x][r];
  }
}

template <typename T>
__global__ void import_nchw_kernel(const float* __restrict__ nchw, T* __restrict__ nhwc, int C, long long HW) {
  __shared__ float tile[32][33];
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r;
    const long long p = p0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < HW && c < C) ? nchw[(long long)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long p = p0 + r;
    const int c = c0 + threadIdx.x;
    if (p < HW && c < C) nhwc[p * C + c] = from_f<T>(tile[threadIdx.x][r]);
  }
}

__global__ void add_inplace_kernel(float* __restrict__ y, const float* __restrict__ x, float coef_host,
                                   const double* coef_dev, long long n) {
  const float c = coef_dev ? (float)coef_dev[0] : coef_host;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = fmaf(c, x[i], y[i]);
}

}  // namespace

// =============================================================================== launch wrappers
template <typename T>
int launch_conv_first_fwd(st2_ctx* ctx, const float* x, const float* w, const float* bias, T* out, int H,
                          int W, long long xps, int lo, int hi) {
  const int items = H * ((W + 255) / 256);
  const int blocks = items < ctx->sm_count * 2 ? items : ctx->sm_count * 2;
  conv_first_fwd_kernel<T><<<blocks, 256, 0, ctx->stream>>>(x, w, bias, out, H, W, xps ? xps : (long long)H * W, lo, hi);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_conv_first_fwd<float>(st2_ctx*, const float*, const float*, const float*, float*, int, int, long long, int, int);
template int launch_conv_first_fwd<__half>(st2_ctx*, const float*, const float*, const float*, __half*, int, int, long long, int, int);

template <typename T>
int launch_conv_first_bwd(st2_ctx* ctx, const T* g, const float* w, float* gx, int H, int W, int lo, int hi) {
  const long long hw = (long long)H * W;
  long long blocks = (long long)((W + 15) / 16) * ((H + 7) / 8);
  if (blocks > (long long)ctx->sm_count * 8) blocks = (long long)ctx->sm_count * 8;
  conv_first_bwd_kernel<T><<<(int)blocks, 256, 0, ctx->stream>>>(g, w, gx, H, W, lo, hi);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_conv_first_bwd<float>(st2_ctx*, const float*, const float*, float*, int, int, int, int);
template int launch_conv_first_bwd<__half>(st2_ctx*, const __half*, const float*, float*, int, int, int, int);

int launch_conv_exact(st2_ctx* ctx, const float* in, const float* w, const float* bias, const float* act,
                      float* out, int H, int W, int cin, int cout, int epi, int lo, int hi) {
  if (cin % 16 || cout % 64) return st2_fail(ctx, ST2_ERR_ARG, "conv_exact: cin %% 16 / cout %% 64");
  dim3 grid(((H + 7) / 8) * ((W + 7) / 8), cout / 64);
  if (epi == EPI_BIAS_RELU)
    conv_exact_kernel<EPI_BIAS_RELU><<<grid, 256, 0, ctx->stream>>>(in, w, bias, act, out, H, W, cin, cout, lo, hi);
  else if (epi == EPI_MASK)
    conv_exact_kernel<EPI_MASK><<<grid, 256, 0, ctx->stream>>>(in, w, bias, act, out, H, W, cin, cout, lo, hi);
  else
    conv_exact_kernel<EPI_RAW><<<grid, 256, 0, ctx->stream>>>(in, w, bias, act, out, H, W, cin, cout, lo, hi);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

template <typename T>
int launch_pool_fwd(st2_ctx* ctx, const T* in, T* out, int C, int H, int W) {
  const int Ho = pool_extent(H), Wo = pool_extent(W);
  pool_fwd_kernel<T><<<ew_grid((long long)Ho * Wo * C, ctx->sm_count), kThreads, 0, ctx->stream>>>(
      in, out, C, H, W, Ho, Wo);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_pool_fwd<float>(st2_ctx*, const float*, float*, int, int, int);
template int launch_pool_fwd<__half>(st2_ctx*, const __half*, __half*, int, int, int);

template <typename T>
int launch_pool_bwd(st2_ctx* ctx, const T* act, const T* g_pool, T* g_out, int C, int H, int W, int apply_mask) {
  const int Ho = pool_extent(H), Wo = pool_extent(W);
  pool_bwd_kernel<T><<<ew_grid((long long)Ho * Wo * C, ctx->sm_count), kThreads, 0, ctx->stream>>>(
      act, g_pool, g_out, C, H, W, Ho, Wo, apply_mask);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_pool_bwd<float>(st2_ctx*, const float*, const float*, float*, int, int, int, int);
template int launch_pool_bwd<__half>(st2_ctx*, const __half*, const __half*, __half*, int, int, int, int);

template <typename T>
int launch_combine(st2_ctx* ctx, const CombineArgs& a) {
  combine_kernel<T><<<ew_grid(a.n, ctx->sm_count), kThreads, 0, ctx->stream>>>(
      (const T*)a.gin, (const T*)a.act, (const T*)a.fc, (const T*)a.sraw, (T*)a.out, a.n, a.apply_mask, a.coef,
      a.h_cc, a.h_sc, a.h_dc);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_combine<float>(st2_ctx*, const CombineArgs&);
template int launch_combine<__half>(st2_ctx*, const CombineArgs&);

template <typename T>
int launch_feature_sums(st2_ctx* ctx, const T* act, const T* fc, long long n, double* sum_diff_sq,
                        double* sum_sq) {
  feature_sums_kernel<T><<<ew_grid(n / 4 + 1, ctx->sm_count), kThreads, 0, ctx->stream>>>(act, fc, n, sum_diff_sq,
                                                                                          sum_sq);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_feature_sums<float>(st2_ctx*, const float*, const float*, long long, double*, double*);
template int launch_feature_sums<__half>(st2_ctx*, const __half*, const __half*, long long, double*, double*);

template <typename T>
int launch_gram_generic(st2_ctx* ctx, const T* F, int C, long long HW, long long sp, long long sc, double* Gd) {
  const int tiles = (C + 63) / 64;
  long long want_splits = ((long long)ctx->sm_count * 4) / ((long long)tiles * tiles);
  if (want_splits < 1) want_splits = 1;
  long long chunk = (HW + want_splits - 1) / want_splits;
  if (chunk < 256) chunk = 256;
  chunk = (chunk + 15) / 16 * 16;
  const int splits = (int)((HW + chunk - 1) / chunk);
  dim3 grid(tiles, tiles, splits);
  gram_generic_kernel<T><<<grid, 256, 0, ctx->stream>>>(F, C, HW, sp, sc, Gd, chunk);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_gram_generic<float>(st2_ctx*, const float*, int, long long, long long, long long, double*);
template int launch_gram_generic<__half>(st2_ctx*, const __half*, int, long long, long long, long long, double*);

int launch_gram_finalize(st2_ctx* ctx, const double* Gd, const float* A, float* D, int C, long long HW,
                         double* sum_dsq) {
  gram_finalize_kernel<<<ew_grid((long long)C * C, ctx->sm_count), kThreads, 0, ctx->stream>>>(Gd, A, D, C, HW,
                                                                                               sum_dsq);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

int launch_gram_acc_to_f32(st2_ctx* ctx, const double* Gd, float* out, int C) {
  gram_acc_to_f32_kernel<<<ew_grid((long long)C * C, ctx->sm_count), kThreads, 0, ctx->stream>>>(Gd, out, (long long)C * C);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

int launch_gram_from_sum(st2_ctx* ctx, const float* Gs, const float* A, float* D, int C, double HW_total,
                         double* sum_dsq) {
  gram_from_sum_kernel<<<ew_grid((long long)C * C, ctx->sm_count), kThreads, 0, ctx->stream>>>(Gs, A, D, C, HW_total,
                                                                                               sum_dsq);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

template <typename T>
int launch_style_grad_generic(st2_ctx* ctx, const T* F, const float* D, T* raw, int C, long long HW,
                              long long sp, long long sc, double* sum_rawsq) {
  dim3 grid((unsigned)((HW + 63) / 64), (C + 63) / 64);
  style_grad_generic_kernel<T><<<grid, 256, 0, ctx->stream>>>(F, D, raw, C, HW, sp, sc, sum_rawsq);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_style_grad_generic<float>(st2_ctx*, const float*, const float*, float*, int, long long,
                                              long long, long long, double*);
template int launch_style_grad_generic<__half>(st2_ctx*, const __half*, const float*, __half*, int, long long,
                                               long long, long long, double*);

template <typename T>
int launch_export_nchw(st2_ctx* ctx, const T* nhwc, float* nchw, int C, int H, int W) {
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32), block(32, 8);
  export_nchw_kernel<T><<<grid, block, 0, ctx->stream>>>(nhwc, nchw, C, HW);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_export_nchw<float>(st2_ctx*, const float*, float*, int, int, int);
template int launch_export_nchw<__half>(st2_ctx*, const __half*, float*, int, int, int);

template <typename T>
int launch_import_nchw(st2_ctx* ctx, const float* nchw, T* nhwc, int C, int H, int W) {
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32), block(32, 8);
  import_nchw_kernel<T><<<grid, block, 0, ctx->stream>>>(nchw, nhwc, C, HW);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}
template int launch_import_nchw<float>(st2_ctx*, const float*, float*, int, int, int);
template int launch_import_nchw<__half>(st2_ctx*, const float*, __half*, int, int, int);

int launch_add_inplace(st2_ctx* ctx, float* y, const float* x, float coef_host, const double* coef_dev,
                       long long n) {
  add_inplace_kernel<<<ew_grid(n, ctx->sm_count), kThreads, 0, ctx->stream>>>(y, x, coef_host, coef_dev, n);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

static St2KernelReg g_reg_layers({
    ST2_KFN(conv_first_fwd_kernel<float>), ST2_KFN(conv_first_fwd_kernel<__half>), ST2_KFN(conv_first_bwd_kernel<float>),
    ST2_KFN(conv_first_bwd_kernel<__half>), ST2_KFN(conv_exact_kernel<EPI_BIAS_RELU>), ST2_KFN(conv_exact_kernel<EPI_MASK>),
    ST2_KFN(conv_exact_kernel<EPI_RAW>), ST2_KFN(pool_fwd_kernel<float>), ST2_KFN(pool_fwd_kernel<__half>),
    ST2_KFN(pool_bwd_kernel<float>), ST2_KFN(pool_bwd_kernel<__half>), ST2_KFN(combine_kernel<float>),
    ST2_KFN(combine_kernel<__half>), ST2_KFN(feature_sums_kernel<float>), ST2_KFN(feature_sums_kernel<__half>),
    ST2_KFN(gram_generic_kernel<float>), ST2_KFN(gram_generic_kernel<__half>), ST2_KFN(gram_finalize_kernel),
    ST2_KFN(gram_acc_to_f32_kernel), ST2_KFN(gram_from_sum_kernel), ST2_KFN(style_grad_generic_kernel<float>),
    ST2_KFN(style_grad_generic_kernel<__half>), ST2_KFN(export_nchw_kernel<float>), ST2_KFN(export_nchw_kernel<__half>),
    ST2_KFN(import_nchw_kernel<float>), ST2_KFN(import_nchw_kernel<__half>), ST2_KFN(add_inplace_kernel)});

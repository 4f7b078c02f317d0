// Pixel-space kernels: pre/deprocess (worker.py:63-71), the fused TV + p-norm + gradient-assembly
// regulariser (utils.py:285-304, worker.py:283-297) and the separable Pillow-compatible resampler
// (utils.py:130-160; libImaging/Resample.c [ext]).
#include "st2_kernels.h"

#include <math.h>

namespace {

constexpr int kThreads = 256;
__constant__ float c_mean[3] = {123.68f, 116.779f, 103.939f};       // worker.py:34, RGB order

template <typename Tin>
__global__ void preprocess_kernel(const Tin* __restrict__ hwc, float* __restrict__ nchw, long long HW) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW;
       p += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) nchw[(long long)c * HW + p] = (float)hwc[p * 3 + c] - c_mean[c];
  }
}

__global__ void deprocess_kernel(const float* __restrict__ nchw, float* __restrict__ hwc, long long HW) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < HW;
       p += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) hwc[p * 3 + c] = nchw[(long long)c * HW + p] + c_mean[c];
  }
}

__device__ __forceinline__ float pow_beta(float v, float e, int mode) {
  // mode 0: e == 0 -> 1;  mode 1: e == 1 -> v;  mode 2: general
  return mode == 0 ? 1.0f : (mode == 1 ? v : powf(v, e));
}

__device__ __forceinline__ float ipow(float m, int k) {
  float r = 1.0f;
  for (int i = 0; i < k; ++i) r *= m;
  return r;
}

// One thread per element of x (C planes of H x W).  Circular forward differences.
//   v = x/divisor (255 in the objective, worker.py:283,287); dw = v - v(h, w+1); dh = v - v(h+1, w); n2 = dw^2 + dh^2 + 1e-8
//   tv_norm = sum n2^(beta/2); k = (beta/2) n2^(beta/2-1); gw = 2 dw k; gh = 2 dh k
//   tv_grad = gw + gh - gw(h, w-1) - gh(h-1, w)
//   p_norm = sum |v|^p (the 1/p is applied by the caller); p_grad = sign(v)|v|^(p-1)
//   grad = bwd + tv*tv_grad + p*p_grad
constexpr int kPxRows = 8, kPxCols = 256;      // outputs per work item: 8 rows x 256 columns of one plane

// TVM = 0: beta == 2 (k == 1, the reference's default tv_power: no pow at all); 1: general beta.
// PI > 0: integer p known at compile time (p_power 6 is the stock value, 2 the other common one); 0: run-time p.
// (ncu, round 2: the generic kernel issued ~156 instructions per output and stalled on instruction fetch; the
// specialisations cut the arithmetic to what the stock parameters need.)
template <int TVM, int PI>
__global__ void __launch_bounds__(256)
pixel_terms_kernel(const float* __restrict__ x, const float* __restrict__ bwd, float* __restrict__ grad, int C, int H,
                   int W, long long xps, int wrap, float tv, float beta, float pw, float pp, float divisor,
                   double* scal, double* __restrict__ part, unsigned int* counter) {
  // x: row 0 of plane 0, planes xps floats apart.  wrap = 1: rows wrap around inside the tensor (whole
  // canvas); wrap = 0: rows -1 and H are addressable halo rows (a row strip; the strips at the canvas
  // edges hold the circular neighbours there).  bwd / grad are dense C x H x W.
  // Work item = 8 rows x 256 columns of one plane.  The (8 + 2) x (256 + 2) values v = x / 255 it needs are
  // divided ONCE (IEEE division, as the reference's x / 255) into shared memory; every output then reads its 7
  // neighbours from there.  The staging loads are row-wise (thread = column, ten independent loads in flight per
  // thread: the kernel is bound by load latency, ncu r1v: long-scoreboard stalls, 6.7 % DRAM), and so are the
  // eight bwd loads of the compute phase, issued before the arithmetic.
  // Sums: per-block partials -> `part` -> summed by the last block to finish (one writer per scalar instead of
  // six same-line fp64 atomics per block); part == nullptr: plain atomics (stand-alone st2_pixel_terms).
  __shared__ float sv[kPxRows + 2][kPxCols + 2];
  pdl_trigger();
  pdl_wait();
  const long long HW = (long long)H * W;
  const float half_beta = beta * 0.5f;
  const int m_norm = (TVM == 0 || half_beta == 1.0f) ? 1 : 2;
  const float e_k = half_beta - 1.0f;
  const int m_k = (TVM == 0 || e_k == 0.0f) ? 0 : (e_k == 1.0f ? 1 : 2);
  const int pi = PI > 0 ? PI : (int)pp;
  const bool p_int = PI > 0 || (((float)pi == pp) && pi >= 1 && pi <= 16);
  float s_tv = 0.f, s_p = 0.f, s_b = 0.f, s_t = 0.f, s_pg = 0.f, s_g = 0.f;
  const int bands = (H + kPxRows - 1) / kPxRows, chunks = (W + kPxCols - 1) / kPxCols;
  const int items = C * bands * chunks;
  const int t = threadIdx.x;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int ch = item % chunks, band = (item / chunks) % bands, c = item / (chunks * bands);
    const int h0 = band * kPxRows, w0 = ch * kPxCols;
    const float* xp = x + (long long)c * xps;
    const int w = w0 + t;
    // bwd of this thread's eight outputs: independent of the staging, so in flight across the barrier
    float bw[kPxRows];
#pragma unroll
    for (int r = 0; r < kPxRows; ++r) {
      const int h = h0 + r;
      bw[r] = (bwd != nullptr && grad != nullptr && h < H && w < W) ? __ldg(bwd + (long long)c * HW + (long long)h * W + w) : 0.f;
    }
    __syncthreads();                            // the previous item's readers are done with sv
    {
      // column q = t + 1 of every staged row (w0 + t), plus the two edge columns q = 0 and q = 257 by threads 0..19
      int wc = w;                               // <= W: w == W is the wrap column of a ragged last chunk
      if (wc == W) wc = 0;
      float raw[kPxRows + 2];
#pragma unroll
      for (int r = 0; r < kPxRows + 2; ++r) {
        int hh = h0 + r - 1;
        float v = 0.f;
        if (hh <= H && w <= W) {
          if (wrap) { if (hh < 0) hh = H - 1; else if (hh == H) hh = 0; }
          v = __ldg(xp + (long long)hh * W + wc);
        }
        raw[r] = v;
      }
      float edge = 0.f;
      int er = 0, eq = 0;
      if (t < 2 * (kPxRows + 2)) {
        er = t >> 1;
        eq = (t & 1) ? (kPxCols + 1) : 0;
        int hh = h0 + er - 1, ww = w0 + eq - 1;
        if (hh <= H && ww <= W) {
          if (ww < 0) ww = W - 1; else if (ww == W) ww = 0;            // columns always wrap (utils.py:232-254)
          if (wrap) { if (hh < 0) hh = H - 1; else if (hh == H) hh = 0; }
          edge = __ldg(xp + (long long)hh * W + ww);
        }
      }
#pragma unroll
      for (int r = 0; r < kPxRows + 2; ++r) sv[r][t + 1] = raw[r] / divisor;
      if (t < 2 * (kPxRows + 2)) sv[er][eq] = edge / divisor;
    }
    __syncthreads();
    if (w < W) {
#pragma unroll
      for (int r = 0; r < kPxRows; ++r) {
        const int h = h0 + r;
        if (h >= H) break;
        const int q = t + 1, rr = r + 1;
        const float v = sv[rr][q], vr = sv[rr][q + 1], vl = sv[rr][q - 1], vd = sv[rr + 1][q], vu = sv[rr - 1][q];
        const float vld = sv[rr + 1][q - 1], vur = sv[rr - 1][q + 1];
        // this pixel
        const float dw0 = v - vr, dh0 = v - vd;
        const float n0 = dw0 * dw0 + dh0 * dh0 + 1e-8f;
        const float k0 = half_beta * pow_beta(n0, e_k, m_k);
        // left neighbour (h, w-1)
        const float dw1 = vl - v, dh1 = vl - vld;
        const float n1 = dw1 * dw1 + dh1 * dh1 + 1e-8f;
        const float k1 = half_beta * pow_beta(n1, e_k, m_k);
        // upper neighbour (h-1, w)
        const float dw2 = vu - vur, dh2 = vu - v;
        const float n2 = dw2 * dw2 + dh2 * dh2 + 1e-8f;
        const float k2 = half_beta * pow_beta(n2, e_k, m_k);
        float tg = 2.0f * dw0 * k0 + 2.0f * dh0 * k0;
        tg -= 2.0f * dw1 * k1;
        tg -= 2.0f * dh2 * k2;
        s_tv += pow_beta(n0, half_beta, m_norm);
        const float mag = fabsf(v);
        float mp1;                                           // |v|^(p-1)
        if (PI > 0) {
          mp1 = 1.0f;
#pragma unroll
          for (int e = 0; e < PI - 1; ++e) mp1 *= mag;
        } else if (p_int) mp1 = ipow(mag, pi - 1); else mp1 = powf(mag, pp - 1.0f);
        s_p += p_int ? mp1 * mag : powf(mag, pp);
        const float sgn = (v > 0.f) ? 1.f : (v < 0.f ? -1.f : 0.f);
        const float tgw = tv * tg;
        const float pgw = pw * (sgn * mp1);
        s_t = fmaf(tgw, tgw, s_t);
        s_pg = fmaf(pgw, pgw, s_pg);
        if (grad != nullptr) {
          const long long o = (long long)c * HW + (long long)h * W + w;
          const float b = bw[r];
          float g = b + tgw;                                 // worker.py:295-297 order
          g += pgw;
          grad[o] = g;
          s_b = fmaf(b, b, s_b);
          s_g = fmaf(g, g, s_g);
        }
      }
    }
  }
  float vals[6] = {s_tv, s_p, s_b, s_t, s_pg, s_g};
  double* dst[6] = {scal + ST2_G_TV_NORM, scal + ST2_G_P_NORM, scal + ST2_G_SCD_GRAD_SQ,
                    scal + ST2_G_T_GRAD_SQ, scal + ST2_G_P_GRAD_SQ, scal + ST2_G_GRAD_SQ};
  if (part == nullptr) {
    block_accumulate<6>(vals, dst);
    return;
  }
  block_accumulate_last<6>(vals, dst, part, counter);
}

// ---------------------------------------------------------------------------------- resampler
__device__ __forceinline__ double filt_eval(double t, int method) {
  if (method == ST2_RESAMPLE_BILINEAR) {
    t = fabs(t);
    return t < 1.0 ? 1.0 - t : 0.0;
  }
  if (t < -3.0 || t >= 3.0) return 0.0;                   // lanczos a = 3
  if (t == 0.0) return 1.0;
  const double pt = M_PI * t;
  return (sin(pt) / pt) * (sin(pt / 3.0) / (pt / 3.0));
}

// One pass along one axis.  in: planes x n_other x n_in (axis contiguous when stride_axis == 1).
// Each thread produces one output sample, recomputing its window weights in fp64 (cold path).
__global__ void resample_axis_kernel(const float* __restrict__ src, float* __restrict__ dst, int planes,
                                     int n_in, int n_out, int n_other, long long in_axis_stride,
                                     long long in_other_stride, long long in_plane_stride,
                                     long long out_axis_stride, long long out_other_stride,
                                     long long out_plane_stride, int method, int clamp_min_zero) {
  const double scale = (double)n_in / (double)n_out;
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double support = (method == ST2_RESAMPLE_BILINEAR ? 1.0 : 3.0) * fscale;
  const long long total = (long long)planes * n_other * n_out;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % n_out);
    const long long r = i / n_out;
    const int o = (int)(r % n_other);
    const int pl = (int)(r / n_other);
    const double centre = (xx + 0.5) * scale;
    int xmin = (int)(centre - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(centre + support + 0.5);
    if (xmax > n_in) xmax = n_in;
    const float* sp = src + (long long)pl * in_plane_stride + (long long)o * in_other_stride;
    double wsum = 0.0, acc = 0.0;
    for (int k = xmin; k < xmax; ++k) {
      const double w = filt_eval((k - centre + 0.5) / fscale, method);
      wsum += w;
      acc += w * (double)sp[(long long)k * in_axis_stride];
    }
    // Pillow normalises the weights first (w /= sum) and then accumulates sum in[x] * w[x]
    double val = 0.0;
    if (wsum != 0.0) {
      val = 0.0;
      for (int k = xmin; k < xmax; ++k) {
        const double w = filt_eval((k - centre + 0.5) / fscale, method) / wsum;
        val += (double)sp[(long long)k * in_axis_stride] * w;
      }
    } else {
      val = acc;
    }
    float out = (float)val;
    if (clamp_min_zero && out < 0.f) out = 0.f;
    dst[(long long)pl * out_plane_stride + (long long)o * out_other_stride + (long long)xx * out_axis_stride] = out;
  }
}

inline int ew_grid(long long n, int sm_count) {
  long long b = (n + kThreads - 1) / kThreads, cap = (long long)sm_count * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

int pixel_terms_strip(st2_ctx* ctx, const float* x, long long xps, int wrap, const float* bwd, float* grad_out,
                      int C, int H, int W, float tv, float tv_power, float p, float p_power, float divisor,
                      double* scal, double* part, unsigned int* counter) {
  if (!ctx || !x || !scal || C < 1 || H < 1 || W < 1) return st2_fail(ctx, ST2_ERR_ARG, "st2_pixel_terms: bad arguments");
  // whole waves of resident blocks (registers allow 5 per SM)
  int blocks = C * ((H + kPxRows - 1) / kPxRows) * ((W + kPxCols - 1) / kPxCols);
  if (blocks > ctx->sm_count * 5) blocks = ctx->sm_count * 5;
  if (blocks > ST2_PART_BLOCKS) blocks = ST2_PART_BLOCKS;
#define ST2_PIXEL_LAUNCH(TVM, PI)                                                                                    \
  st2_launch_pdl(ctx, true, pixel_terms_kernel<TVM, PI>, blocks, kThreads, 0, x, bwd, grad_out, C, H, W, xps, wrap, tv,    \
                 tv_power, p, p_power, divisor, scal, part, counter)
  const bool b2 = tv_power == 2.0f;
  if (p_power == 6.0f) { if (b2) ST2_PIXEL_LAUNCH(0, 6); else ST2_PIXEL_LAUNCH(1, 6); }
  else if (p_power == 2.0f) { if (b2) ST2_PIXEL_LAUNCH(0, 2); else ST2_PIXEL_LAUNCH(1, 2); }
  else { if (b2) ST2_PIXEL_LAUNCH(0, 0); else ST2_PIXEL_LAUNCH(1, 0); }
#undef ST2_PIXEL_LAUNCH
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

extern "C" {

int st2_pixel_terms(st2_ctx* ctx, const float* x, const float* bwd, float* grad_out, int C, int H, int W,
                    float tv, float tv_power, float p, float p_power, float divisor, double* scal) {
  return pixel_terms_strip(ctx, x, (long long)H * W, 1, bwd, grad_out, C, H, W, tv, tv_power, p, p_power, divisor, scal,
                           nullptr, nullptr);
}

int st2_preprocess_u8(st2_ctx* ctx, const unsigned char* hwc, float* nchw, int h, int w) {
  if (!ctx || !hwc || !nchw) return st2_fail(ctx, ST2_ERR_ARG, "st2_preprocess_u8: null");
  preprocess_kernel<unsigned char><<<ew_grid((long long)h * w, ctx->sm_count), kThreads, 0, ctx->stream>>>(
      hwc, nchw, (long long)h * w);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

int st2_preprocess_f32(st2_ctx* ctx, const float* hwc, float* nchw, int h, int w) {
  if (!ctx || !hwc || !nchw) return st2_fail(ctx, ST2_ERR_ARG, "st2_preprocess_f32: null");
  preprocess_kernel<float><<<ew_grid((long long)h * w, ctx->sm_count), kThreads, 0, ctx->stream>>>(
      hwc, nchw, (long long)h * w);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

int st2_deprocess(st2_ctx* ctx, const float* nchw, float* hwc, int h, int w) {
  if (!ctx || !hwc || !nchw) return st2_fail(ctx, ST2_ERR_ARG, "st2_deprocess: null");
  deprocess_kernel<<<ew_grid((long long)h * w, ctx->sm_count), kThreads, 0, ctx->stream>>>(nchw, hwc,
                                                                                       (long long)h * w);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

int st2_resample(st2_ctx* ctx, const float* src, int planes, int h_in, int w_in, float* dst, int h_out,
                 int w_out, int method, int clamp_min_zero) {
  if (!ctx || !src || !dst || planes < 1 || h_in < 1 || w_in < 1 || h_out < 1 || w_out < 1 ||
      (method != ST2_RESAMPLE_LANCZOS && method != ST2_RESAMPLE_BILINEAR))
    return st2_fail(ctx, ST2_ERR_ARG, "st2_resample: bad arguments");
  cudaStream_t s = ctx->stream;
  // horizontal pass first into an fp32 temporary (skipped when the width is unchanged), then vertical
  const float* cur = src;
  float* tmp = nullptr;
  if (w_out != w_in) {
    float* hdst = dst;
    if (h_out != h_in) {
      ST2_CUDA(ctx, cudaMallocAsync(&tmp, sizeof(float) * (size_t)planes * h_in * w_out, s));
      hdst = tmp;
    }
    const long long total = (long long)planes * h_in * w_out;
    resample_axis_kernel<<<ew_grid(total, ctx->sm_count), kThreads, 0, s>>>(
        cur, hdst, planes, w_in, w_out, h_in, 1, w_in, (long long)h_in * w_in, 1, w_out,
        (long long)h_in * w_out, method, (h_out == h_in) ? clamp_min_zero : 0);
    ST2_LAUNCH_CHECK(ctx);
    cur = hdst;
  }
  if (h_out != h_in) {
    const long long total = (long long)planes * w_out * h_out;
    resample_axis_kernel<<<ew_grid(total, ctx->sm_count), kThreads, 0, s>>>(
        cur, dst, planes, h_in, h_out, w_out, w_out, 1, (long long)h_in * w_out, w_out, 1,
        (long long)h_out * w_out, method, clamp_min_zero);
    ST2_LAUNCH_CHECK(ctx);
  } else if (w_out == w_in) {
    ST2_CUDA(ctx, cudaMemcpyAsync(dst, src, sizeof(float) * (size_t)planes * h_in * w_in,
                                  cudaMemcpyDeviceToDevice, s));
  }
  if (tmp) ST2_CUDA(ctx, cudaFreeAsync(tmp, s));
  return 0;
}

}  // extern "C"

static St2KernelReg g_reg_pixel({ST2_KFN(pixel_terms_kernel<0, 6>), ST2_KFN(pixel_terms_kernel<1, 6>),
                                    ST2_KFN(pixel_terms_kernel<0, 2>), ST2_KFN(pixel_terms_kernel<1, 2>),
                                    ST2_KFN(pixel_terms_kernel<0, 0>), ST2_KFN(pixel_terms_kernel<1, 0>), ST2_KFN(preprocess_kernel<unsigned char>),
                                    ST2_KFN(preprocess_kernel<float>), ST2_KFN(deprocess_kernel), ST2_KFN(resample_axis_kernel)});

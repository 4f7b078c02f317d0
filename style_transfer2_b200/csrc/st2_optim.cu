// Adam and the level-1 helpers as fused, bandwidth-bound passes (optimizers.py:7-46, utils.py:29-69).
// The L-BFGS kernels live in st2_lbfgs.cu.  All vectors fp32, length n = 3*H*W.
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

extern "C" {

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
  double* scratch = ctx->dot_scratch;        // per context (= per device)
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

static St2KernelReg g_reg_optim({ST2_KFN(adam_kernel), ST2_KFN(dot_kernel), ST2_KFN(axpy_kernel)});

// Shared declarations for libst2 (sm_100a only).  See include/st2.h for the C ABI.
#pragma once
#include <utility>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <initializer_list>
#include <vector>

#include "../../include/st2.h"

#define ST2_MAX_HIST 11          // L-BFGS: n_corr (10) + one staging slot

// ---- network topology (models/vgg19.prototxt:3-337 of the reference) ------------------------
enum BlobKind { KIND_INPUT = 0, KIND_CONV = 1, KIND_POOL = 2 };
struct BlobInfo { const char* name; int kind; int channels; int conv_index; };
extern const BlobInfo g_blobs[ST2_NUM_BLOBS];

// ---- scalar block layout (doubles, device) ---------------------------------------------------
// per blob b: base = b * ST2_SCAL_PER_BLOB
enum {
  SB_C_SUMSQ = 0,     // sum (F - Fc)^2
  SB_S_GRAMSQ,        // sum D^2          (D = gram(F) - A)
  SB_S_RAWSQ,         // sum (D F)^2      (unscaled style gradient)
  SB_D_SUMSQ,         // sum F^2
  SB_C_NORM, SB_S_NORM, SB_D_NORM,          // frozen normalisers (worker.py:253-254,265-266,274-275)
  SB_C_VALID, SB_S_VALID, SB_D_VALID,       // 1.0 once the normaliser is frozen
  SB_C_COEF, SB_S_COEF, SB_D_COEF,          // multipliers used by the combine kernel
  SB_C_LOSS, SB_C_GRAD, SB_S_LOSS, SB_S_GRAD, SB_D_LOSS, SB_D_GRAD,   // trace values
  SB_S_DSCALE,        // power-of-two scale applied to the fp16 copy of D
  ST2_SCAL_PER_BLOB_USED
};
static_assert(ST2_SCAL_PER_BLOB_USED <= ST2_SCAL_PER_BLOB, "scalar block too small");

struct st2_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = 0;
  std::string err;
  // master fp32 weights (device): OIHW + bias, per conv
  float* w_oihw[ST2_NUM_CONVS] = {};
  float* bias[ST2_NUM_CONVS] = {};
  int cin[ST2_NUM_CONVS] = {};
  int cout[ST2_NUM_CONVS] = {};
  // exact-mode packs: forward [tap][ci][co], backward-data [tap'][co][ci]
  float* wf32_fwd[ST2_NUM_CONVS] = {};
  float* wf32_bwd[ST2_NUM_CONVS] = {};
  // tensor-core packs (fp16, K-major): forward [co][tap][ci], backward [ci][tap'][co]
  __half* wh_fwd[ST2_NUM_CONVS] = {};
  __half* wh_bwd[ST2_NUM_CONVS] = {};
  __half* wh_first = nullptr;      // conv1_1 forward pack for the sliding-window tensor-core kernel
  __half* wh_bwd_all = nullptr;    // conv1_1 data-gradient pack of the stencil form: [tap' * 3 + plane (27 of 32)][64]
  void* tmap_encode = nullptr;     // cuTensorMapEncodeTiled entry point
  long long launches = 0;          // kernels launched through this context
  int debug_flags = 0;             // timing experiments only (st2_debug_flags)
  // ST2_* environment knobs (kernel-selection experiments and the small-canvas tests that force the
  // production kernels): read ONCE in st2_ctx_create, never on a launch path
  struct Knobs {
    bool no_fused_inject = false, no_tc_gram = false, no_tc_first = false, no_ws = false, force_pair = false,
         wsp = false, no_pair = false, no_pool_fusion = false, no_style_fuse = false, no_graph = false, no_inkernel_halo = false, no_stencil = false, no_ws128 = false, no_pdl = false;
    int tc_bn = 0;
    long long pair_min_tiles = -1;
  } knobs;
  double* dot_scratch = nullptr;   // st2_dot / st2_sumsq result slot (device)
  // optional per-category device timing (CUDA events on the launch stream), see st2_profile()
  bool prof_on = false;
  struct ProfSpan { int cat; cudaEvent_t a, b; };
  std::vector<ProfSpan> prof_spans;
  std::vector<cudaEvent_t> prof_pool;
};

// RAII span: records an event pair around the launches issued in its scope when profiling is on
struct ProfScope {
  st2_ctx* ctx; int idx;
  ProfScope(st2_ctx* c, int cat);
  ~ProfScope();
};

int st2_fail(st2_ctx* ctx, int code, const char* fmt, ...);

// Every kernel of the library is registered here and loaded when a context is created.  CUDA loads
// kernels lazily on first launch and that load may wait for the device to go idle; strips that spin on a
// neighbour's flag while the host is about to launch a not-yet-loaded kernel would deadlock
// (CUDA programming guide, "Lazy Loading -> Concurrent execution").
std::vector<const void*>& st2_kernel_registry();
struct St2KernelReg {
  St2KernelReg(std::initializer_list<const void*> fns) { for (const void* f : fns) st2_kernel_registry().push_back(f); }
};
#define ST2_KFN(...) reinterpret_cast<const void*>(&__VA_ARGS__)
// Kernels that need more than 48 KB of dynamic shared memory: cudaFuncAttributeMaxDynamicSharedMemorySize is
// per (function, device), so st2_ctx_create sets it for the context's device from this list (a process-wide
// "already set" flag would leave a second device without the opt-in).
struct St2SmemOptIn { const void* fn; int bytes; };
std::vector<St2SmemOptIn>& st2_smem_registry();
struct St2SmemReg {
  St2SmemReg(std::initializer_list<St2SmemOptIn> fns) { for (const St2SmemOptIn& f : fns) st2_smem_registry().push_back(f); }
};

#define ST2_CUDA(ctx, expr)                                                             \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      return st2_fail((ctx), ST2_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,              \
                      cudaGetErrorString(_e), __FILE__, __LINE__);                      \
  } while (0)

#define ST2_LAUNCH_CHECK(ctx)                                                           \
  do {                                                                                  \
    (ctx)->launches++;                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess)                                                              \
      return st2_fail((ctx), ST2_ERR_CUDA, "kernel launch failed: %s (%s:%d)",          \
                      cudaGetErrorString(_e), __FILE__, __LINE__);                      \
  } while (0)

// ---- programmatic dependent launch ---------------------------------------------------------------
// Every kernel of the steady-state iteration is launched with programmatic stream serialisation: it may be scheduled
// while its predecessor in the stream is still draining, runs its prologue (barrier init, TMEM allocation, tensor-map
// prefetch) in that shadow, and only pdl_wait() -- before its first access to global memory -- waits for the
// predecessor to have completed and flushed.  pdl_trigger() at the top of a kernel lets ITS successor be scheduled as
// soon as SM resources free up.  A kernel launched this way MUST execute pdl_wait() (all threads, before any global
// load / store / atomic): the chain is transitive only because every link waits.  ST2_NO_PDL=1 launches plainly, for
// which both instructions are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline void st2_launch_pdl(st2_ctx* ctx, bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                           Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && !ctx->knobs.no_pdl) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);      // the error is picked up by ST2_LAUNCH_CHECK
}

// ---- device helpers --------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of up to NV per-thread fp32 partials; one double atomicAdd per value per block.
template <int NV>
__device__ __forceinline__ void block_accumulate(const float (&v)[NV], double* const (&dst)[NV]) {
  __shared__ double sh[NV][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double w = warp_sum_d((double)v[i]);
    if (lane == 0) sh[i][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double w = lane < nwarps ? sh[i][lane] : 0.0;
      w = warp_sum_d(w);
      if (lane == 0 && dst[i] != nullptr) atomicAdd(dst[i], w);
    }
  }
}

// The same without contended atomics: every block stores its NV partial sums to part[blockIdx.x * NV ..], the last
// block to finish (counter) adds them up in block order -- deterministic -- and does ONE add per destination.
// `counter` must be 0 at launch and is left 0.  At most ST2_PART_BLOCKS blocks.
#define ST2_PART_BLOCKS 1024
template <int NV>
__device__ __forceinline__ void block_accumulate_last(const float (&v)[NV], double* const (&dst)[NV], double* part,
                                                      unsigned int* counter) {
  __shared__ double sh[NV][32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double w = warp_sum_d((double)v[i]);
    if (lane == 0) sh[i][warp] = w;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double w = lane < nwarps ? sh[i][lane] : 0.0;
      w = warp_sum_d(w);
      if (lane == 0) part[(size_t)blockIdx.x * NV + i] = w;
    }
    if (lane == 0) {
      __threadfence();
      last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int i = warp; i < NV; i += nwarps) {
    double w = 0.0;
    for (unsigned b = lane; b < gridDim.x; b += 32) w += __ldcg(part + (size_t)b * NV + i);
    w = warp_sum_d(w);
    if (lane == 0 && dst[i] != nullptr) *dst[i] += w;
  }
  if (threadIdx.x == 0) *counter = 0;
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
// fp16 stores saturate: the reference's line-search-free L-BFGS overshoots by orders of magnitude for a
// few steps (pixels at +-11 000) and recovers; an inf in an activation would instead
// poison every later iterate with NaN.
__device__ __forceinline__ float sat_h(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }
// (a, b) -> half2 {lo: a, hi: b}, finite-saturating, in ONE instruction (F2FP.SATFINITE...PACK_AB): the epilogues of the
// convolution kernels are bound by their instruction count (one warp per scheduler), and clamping in fp32 first cost
// four FMNMX per pair.  |v| > 65504 (and +-inf) -> +-65504 as before; a NaN stays a NaN.
__device__ __forceinline__ __half2 h2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return *reinterpret_cast<__half2*>(&r);
}
// the same with max(v, 0) folded in (ReLU epilogues)
__device__ __forceinline__ __half2 h2_relu_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return *reinterpret_cast<__half2*>(&r);
}
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(sat_h(v)); }

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline int pool_extent(int n) { return n > 1 ? (n - 2 + 1) / 2 + 1 : 1; }   // ceil((n-2)/2)+1

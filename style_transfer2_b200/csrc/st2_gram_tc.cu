// tcgen05 Gram matrix of fp16 NHWC features (K5 of SURVEY 2.4):  G[i, j] = sum_p F[p, i] F[p, j].
//
// F is stored pixel-major ([HW][C], channels contiguous), so both operands of F^T F are "MN-major":
// the TMA box {64 channels, 64 pixels} lands in shared memory as 64 K-rows (pixels) of 128 bytes
// (64 channels), SWIZZLE_128B -- the UMMA canonical MN-major layout with 8-row atoms 1024 B apart
// (SBO) and 64-channel groups 8 KB apart (LBO).  No transposed copy of F is ever made.
// One CTA = one 128 x GN tile of G over one chunk of pixels (split-K across the whole GPU); fp32
// accumulators live in TMEM; partial tiles go to a workspace with plain 128-byte row stores and are
// summed (in double) by ONE finalize launch for all style layers of an evaluation -- deterministic, no atomics.
// (Measured, round 2: adding the partial tiles into a single accumulator with red.global.add.v4.f32 instead is 4x
// SLOWER -- 16.8 M fp32 reductions per iteration run at ~20 per clock GPU-wide in the L2 atomic units.)
#include "st2_kernels.h"
#include "st2_tc.cuh"

int st2_encode_tmap(st2_ctx* ctx, CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                    const cuuint64_t* strides_bytes, const cuuint32_t* box);

namespace {

constexpr int GM = 128;
constexpr int KP = 64;              // pixels per pipeline stage
constexpr int kBoxBytes = 64 * 64 * 2;   // one {64 ch, 64 px} box = 8 KB
constexpr int kThreadsG = 256;

template <int GN> struct GCfg {
  static constexpr int kABytes = 2 * kBoxBytes;
  static constexpr int kBBytes = (GN / 64) * kBoxBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (GN == 256) ? 4 : (GN == 128 ? 6 : 8);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

struct GramGeom {
  int C;
  long long HW;
  long long chunk;     // pixels per split (multiple of 64)
  int m_tiles, n_tiles, splits;
};

template <int GN>
__global__ void __launch_bounds__(kThreadsG, 1)
tc_gram_kernel(const __grid_constant__ CUtensorMap tmap, const GramGeom g, float* __restrict__ partials) {
  using C = GCfg<GN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* done_bar = bars + 2 * C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int mt = tile / g.n_tiles, nt = tile % g.n_tiles;
  const int m0 = mt * GM, n0 = nt * GN;
  const long long p_begin = (long long)split * g.chunk;
  const long long p_end = (p_begin + g.chunk < g.HW) ? p_begin + g.chunk : g.HW;
  const int k_iters = (int)((p_end - p_begin + KP - 1) / KP);

  pdl_trigger();
  if (warp == 0 && lane == 0) tc::prefetch_tmap(&tmap);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(done_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, GN < 32 ? 32 : GN);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < k_iters; ++it) {
      const int p = (int)(p_begin + (long long)it * KP);
      tc::mbar_wait(&empty_bar[stage], phase ^ 1);
      if (tc::elect_one()) {
        tc::mbar_expect_tx(&full_bar[stage], C::kStageBytes);
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          int ca = m0 + a * 64;
          if (ca > g.C - 64) ca = g.C - 64;               // C == 64: rows 64..127 duplicate rows 0..63 (discarded)
          tc::tma_load_2d(smem_a + stage * C::kABytes + a * kBoxBytes, &tmap, &full_bar[stage], ca, p);
        }
#pragma unroll
        for (int b = 0; b < GN / 64; ++b)
          tc::tma_load_2d(smem_b + stage * C::kBBytes + b * kBoxBytes, &tmap, &full_bar[stage], n0 + b * 64, p);
      }
      __syncwarp();
      if (++stage == C::kStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = tc::idesc_f16(GM, GN, 1, 1);      // both operands MN-major
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < k_iters; ++it) {
      tc::mbar_wait(&full_bar[stage], phase);
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t a_addr = tc::smem_u32(smem_a + stage * C::kABytes);
        const uint32_t b_addr = tc::smem_u32(smem_b + stage * C::kBBytes);
#pragma unroll
        for (int k = 0; k < KP / 16; ++k) {                        // 16 pixels = 2 atoms of 8 K-rows = 2048 B
          const uint64_t a_desc = tc::smem_desc_mn_sw128(a_addr + k * 2048, kBoxBytes);
          const uint64_t b_desc = tc::smem_desc_mn_sw128(b_addr + k * 2048, kBoxBytes);
          tc::umma_f16(tmem_base, a_desc, b_desc, idesc, (it | k) != 0);
        }
        tc::umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == C::kStages) { stage = 0; phase ^= 1; }
    }
    if (tc::elect_one()) tc::umma_commit(done_bar);
    __syncwarp();
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int i = m0 + ew * 32 + lane;
    if (lane == 0) tc::mbar_wait(done_bar, 0);
    __syncwarp();
    tc::fence_after_sync();
    const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16);
    // Partial tiles are stored GROUP-major: the C*C outputs are cut into groups of 128 consecutive floats and all
    // splits of a group sit next to each other ([group][split][128]), so the finalize kernel streams contiguous
    // memory (split-major tiles made it gather 512-byte pieces 16 KB .. 1 MB apart: 1.1 TB/s, ncu r2d).
    const long long lin0 = (long long)i * g.C + n0;
#pragma unroll 1
    for (int c = 0; c < GN / 32; ++c) {
      uint32_t r[32];
      tc::tmem_ld_32x32(t_row + c * 32, r);
      tc::tmem_ld_wait();
      if (i < g.C && k_iters > 0) {
        const long long lin = lin0 + c * 32;
        float4* o = reinterpret_cast<float4*>(partials + (((lin >> 7) * gridDim.y + split) << 7) + (lin & 127));
#pragma unroll
        for (int q = 0; q < 8; ++q)
          o[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                             __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3]));
      }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, GN < 32 ? 32 : GN);
  }
}

// One launch finishes every Gram of an evaluation.  blockIdx.y = layer; a block = 8 split groups x 32 outputs:
// thread (kg, o) sums the partials kg, kg + 8, ... of output o (two chains), the eight groups meet in shared memory,
// so an output costs ~nsplit / 16 dependent L2 round trips instead of nsplit / 4 (the per-layer finalize of round 1
// was pure latency: 10-13 us each, five launches).
//   raw == 0:  D = (sum_s partial[s]) / (C*HW) - A ; *sum_dsq += sum D^2        raw == 1: D = sum_s partial[s]
struct GramFinLayer {
  const float* partials; const float* A; float* D; double* sum_dsq;
  int nsplit, C; long long HW;
};
struct GramFinArgs { GramFinLayer l[8]; int raw; };

__global__ void __launch_bounds__(256) gram_finalize_all_kernel(const GramFinArgs a) {
  // 8 split groups x 32 lanes; a lane owns FOUR consecutive outputs (one 16-byte load per partial tile), and all of a
  // thread's loads (<= 19 for 148 splits) are issued before the first add: the kernel streams 67 MB of partial tiles
  // per iteration at 1024^2 and must keep that many bytes in flight, not chase one L2 round trip per split.
  const GramFinLayer L = a.l[blockIdx.y];
  __shared__ double sh[8][32][4];
  pdl_trigger();
  pdl_wait();
  const long long n = (long long)L.C * L.C;                  // multiple of 4096
  const float denom = (float)((double)L.C * (double)L.HW);
  const int kg = threadIdx.x >> 5, o = threadIdx.x & 31;
  constexpr int kMaxPer = 19;                                // ceil(148 / 8)
  float acc = 0.f;
  for (long long base = (long long)blockIdx.x * 128; base < n; base += (long long)gridDim.x * 128) {
    const long long i = base + 4 * o;
    const float4* src = reinterpret_cast<const float4*>(L.partials + (base >> 7) * (long long)L.nsplit * 128 + 4 * o);
    const long long stride4 = 32;                            // [group][split][128 floats]
    // a thread adds its <= 19 partial tiles in fp32 (pairwise where it matters little: each is itself an fp32 sum over
    // thousands of pixels), the eight groups are then combined in double
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = kg; k0 < L.nsplit; k0 += 8 * kMaxPer) {
      float4 v[kMaxPer];
#pragma unroll
      for (int u = 0; u < kMaxPer; ++u) {                    // unconditional loads (clamped index): all issued up front
        const int k = k0 + 8 * u;
        v[u] = __ldcg(src + (long long)(k < L.nsplit ? k : L.nsplit - 1) * stride4);
      }
#pragma unroll
      for (int u = 0; u < kMaxPer; ++u) {
        if (k0 + 8 * u < L.nsplit) { s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w; }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) sh[kg][o][e] = (double)s[e];
    __syncthreads();
    if (kg < 4) {                                            // warp e finishes element e of every lane's quad
      double t = 0.0;
#pragma unroll
      for (int g8 = 0; g8 < 8; ++g8) t += sh[g8][o][kg];
      float gv = (float)t;
      const long long ie = i + kg;
      if (!a.raw) {
        gv = gv / denom;
        if (L.A != nullptr) gv -= L.A[ie];
        acc = fmaf(gv, gv, acc);
      }
      L.D[ie] = gv;
    }
    __syncthreads();
  }
  if (!a.raw && L.sum_dsq != nullptr) {
    __shared__ double wsum[8];
    const double tot = warp_sum_d((double)acc);              // zero in warps 4..7
    if (o == 0) wsum[kg] = tot;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(L.sum_dsq, (wsum[0] + wsum[1]) + (wsum[2] + wsum[3]));
  }
}

}  // namespace

struct TcGramPlan {
  CUtensorMap tmap;
  GramGeom g;
  int gn;
  float* partials;
};

int tc_gram_plan_create(st2_ctx* ctx, const __half* F, int C, long long HW, TcGramPlan** out) {
  if (C % 64 || C < 64 || HW < 1) return st2_fail(ctx, ST2_ERR_ARG, "tc_gram: C must be a multiple of 64");
  TcGramPlan* p = new TcGramPlan();
  p->gn = (C % 256 == 0) ? 256 : (C % 128 == 0 ? 128 : 64);
  GramGeom& g = p->g;
  g.C = C; g.HW = HW;
  g.m_tiles = (C + GM - 1) / GM;
  g.n_tiles = C / p->gn;
  const int tiles = g.m_tiles * g.n_tiles;
  // split-K over all SMs -- but a partial tile is 128 x GN fp32: for C = 512 (8 tiles) a split per SM would write and
  // re-read 19 MB of partial tiles for a 4-17 MB input; half the splits cost the contraction nothing (it is latency-
  // bound there) and halve that traffic
  long long want = ctx->sm_count / tiles / (C >= 512 ? 2 : 1);
  if (want < 1) want = 1;
  long long chunk = (HW + want - 1) / want;
  chunk = (chunk + KP - 1) / KP * KP;
  g.chunk = chunk;
  g.splits = (int)((HW + chunk - 1) / chunk);
  cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)HW};
  cuuint64_t strides[1] = {(cuuint64_t)C * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)KP};
  int rc = st2_encode_tmap(ctx, &p->tmap, F, 2, dims, strides, box);
  if (rc) { delete p; return rc; }
  cudaError_t e = cudaMalloc(&p->partials, sizeof(float) * (size_t)g.splits * C * C);
  if (e != cudaSuccess) { delete p; return st2_fail(ctx, ST2_ERR_CUDA, "tc_gram: workspace: %s", cudaGetErrorString(e)); }
  *out = p;
  return 0;
}

void tc_gram_plan_destroy(TcGramPlan* p) {
  if (!p) return;
  cudaFree(p->partials);
  delete p;
}

template <int GN>
static int launch_gn(st2_ctx* ctx, TcGramPlan* p) {
  dim3 grid(p->g.m_tiles * p->g.n_tiles, p->g.splits);
  st2_launch_pdl(ctx, true, tc_gram_kernel<GN>, grid, kThreadsG, GCfg<GN>::kSmemBytes, p->tmap, p->g, p->partials);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

static int launch_mma(st2_ctx* ctx, TcGramPlan* p) {
  switch (p->gn) {
    case 256: return launch_gn<256>(ctx, p);
    case 128: return launch_gn<128>(ctx, p);
    default:  return launch_gn<64>(ctx, p);
  }
}

int tc_gram_mma_launch(st2_ctx* ctx, TcGramPlan* p) { return launch_mma(ctx, p); }

int tc_gram_finalize_all(st2_ctx* ctx, int n, TcGramPlan* const* plans, const float* const* A, float* const* D,
                         double* const* sum_dsq, int raw) {
  if (n < 1 || n > 8) return st2_fail(ctx, ST2_ERR_ARG, "tc_gram_finalize_all: 1..8 layers");
  GramFinArgs a;
  a.raw = raw;
  for (int i = 0; i < n; ++i) {
    a.l[i].partials = plans[i]->partials; a.l[i].A = A ? A[i] : nullptr; a.l[i].D = D[i];
    a.l[i].sum_dsq = sum_dsq ? sum_dsq[i] : nullptr;
    a.l[i].nsplit = plans[i]->g.splits; a.l[i].C = plans[i]->g.C; a.l[i].HW = plans[i]->g.HW;
  }
  st2_launch_pdl(ctx, true, gram_finalize_all_kernel, dim3(ctx->sm_count * 2, n), 256, 0, a);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

int tc_gram_sum_launch(st2_ctx* ctx, TcGramPlan* p, float* Gsum) {
  int rc = launch_mma(ctx, p);
  if (rc) return rc;
  return tc_gram_finalize_all(ctx, 1, &p, nullptr, &Gsum, nullptr, 1);
}

int tc_gram_launch(st2_ctx* ctx, TcGramPlan* p, const float* A, float* D, double* sum_dsq) {
  int rc = launch_mma(ctx, p);
  if (rc) return rc;
  return tc_gram_finalize_all(ctx, 1, &p, &A, &D, &sum_dsq, 0);
}

static St2SmemReg g_smem_gram_tc({{ST2_KFN(tc_gram_kernel<256>), GCfg<256>::kSmemBytes},
                                   {ST2_KFN(tc_gram_kernel<128>), GCfg<128>::kSmemBytes},
                                   {ST2_KFN(tc_gram_kernel<64>), GCfg<64>::kSmemBytes}});
static St2KernelReg g_reg_gram_tc({ST2_KFN(tc_gram_kernel<256>), ST2_KFN(tc_gram_kernel<128>), ST2_KFN(tc_gram_kernel<64>),
                                      ST2_KFN(gram_finalize_all_kernel)});

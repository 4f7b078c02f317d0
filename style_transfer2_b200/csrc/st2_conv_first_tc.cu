// conv1_1 forward on the tensor cores (K1 of SURVEY 2.4 for the 3-channel input layer).
//
// The CUDA-core kernel (st2_layers.cu) is bound by its shared-memory weight broadcasts (150 us at 1024^2
// for 3.6 GFLOP).  Here the layer becomes an implicit GEMM with a *sliding-window K*:
//   * x (fp32, 3 planes) is first repacked to NHWC with 8 fp16 "channels" per pixel (16 bytes):
//     [hi r, hi g, hi b, lo r, lo g, lo b, 0, 0] with hi = fp16(x), lo = fp16(x - hi) -- the optimisation
//     variable is NOT rounded: hi + lo carries 22 bits of it;
//   * per 16 x 8 pixel tile ONE TMA box {8 ch, 12 px, 18 rows} (3.4 KB, no swizzle) lands in shared memory;
//   * a UMMA K step (16 fp16 = 32 bytes) spans TWO horizontally adjacent pixels.  In the no-swizzle
//     K-major layout a core matrix is 8 rows x 16 bytes with a 16-byte row pitch -- exactly the pitch of
//     consecutive pixels -- so tile row m = pixel w reads its K chunk 0 at pixel w + c and chunk 1 at
//     pixel w + c + 1 (LBO = 16 bytes: the core matrices overlap), the next 8-pixel row group is one patch
//     row further (SBO = 12 * 16 bytes).  Two K steps per filter row (columns w-1,w | w+1,w+2, the last
//     with zero weights), three filter rows: 6 UMMAs of 128 x 64 x 16 per tile, no im2col anywhere;
//   * weights: [hi|lo][filter row][K step][K chunk][cout][8] fp16 (24 KB), resident in shared memory; w = hi + lo
//     like x, so the layer keeps fp32-grade operands (12 UMMAs per tile -- the tensor pipe idles anyway, the
//     layer is bound by its 134 MB of output);
//   * epilogue: bias + ReLU -> fp16 -> swizzled staging tile -> TMA tile store (whole 128-byte lines).
#include "st2_kernels.h"
#include "st2_tc.cuh"

int st2_encode_tmap_ex(st2_ctx* ctx, CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                       const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle128);

namespace {

constexpr int kThreadsF = 256;
constexpr int kTW = 8, kTH = 16, kPW = 12, kPH = 18;
constexpr int kPatchBytesF = 3584;                     // 12 * 18 * 16 = 3456, padded to a multiple of 128
constexpr int kStagesF = 8;
constexpr int kWBytesF = 2 * 3 * 2 * 2 * 64 * 16;      // 24 576: fp16 hi and lo parts of the weights
constexpr int kOutTile = 128 * 128;                    // 16 KB staging tile
constexpr int kSmemF = kWBytesF + kStagesF * kPatchBytesF + 2 * kOutTile + 1024 + 256;

struct FirstGeom { int H, W, tiles_h, tiles_w, hoff; };

// x (3 planes, xps floats apart, rows -lo .. H-1+hi addressable) -> x8: (H + 2) x W pixels of 8 halves
__global__ void pack_x8_kernel(const float* __restrict__ x, long long xps, uint4* __restrict__ x8, int H, int W, int lo,
                               int hi) {
  const long long total = (long long)(H + 2) * W;
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / W) - 1, w = (int)(i % W);
    uint4 o = make_uint4(0, 0, 0, 0);
    if (r >= -lo && r < H + hi) {
      float v[3];
      __half h[8];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        v[c] = __ldg(x + (long long)c * xps + (long long)r * W + w);
        h[c] = __float2half_rn(sat_h(v[c]));
        h[3 + c] = __float2half_rn(v[c] - __half2float(h[c]));
      }
      h[6] = h[7] = __float2half_rn(0.f);
      o = *reinterpret_cast<const uint4*>(h);
    }
    x8[i] = o;
  }
}

// OIHW fp32 (64, 3, 3, 3) -> [part][r][j][chunk][co][8] fp16: chunk c of K step j is filter column s = 2j + c;
// part 0 = fp16(w), part 1 = fp16(w - part 0)
__global__ void pack_w_first_kernel(const float* __restrict__ w, __half* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * 2 * 2 * 64 * 8) return;
  const int e = i & 7, co = (i >> 3) & 63, chunk = (i >> 9) & 1, j = (i >> 10) & 1, r = i >> 11;
  const int s = 2 * j + chunk;
  float v = 0.f;
  if (s < 3 && e < 6) v = w[((co * 3 + (e % 3)) * 3 + r) * 3 + s];      // hi and lo halves of x see the same weight
  const __half hi = __float2half_rn(v);
  out[i] = hi;
  out[3 * 2 * 2 * 64 * 8 + i] = __float2half_rn(v - __half2float(hi));
}

// accumulator chunk (32 couts of one pixel) -> bias + ReLU -> fp16 -> four 16-byte pieces of the pixel's row of
// the SWIZZLE_128B staging tile (piece q at ((chunk0 + q) ^ row%8); sw = 64 | chunk0 << 3 | row%8)
__device__ __forceinline__ void epi_chunk_first(const uint32_t (&r)[32], const float* __restrict__ bias_c,
                                                __half* __restrict__ srow, const int sw) {
  uint4* op = reinterpret_cast<uint4*>(srow);
  const float4* bp = reinterpret_cast<const float4*>(bias_c);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 b0 = __ldg(bp + 2 * q), b1 = __ldg(bp + 2 * q + 1);
    uint4 o;
    __half2* hp = reinterpret_cast<__half2*>(&o);
    hp[0] = h2_relu_sat(__uint_as_float(r[8 * q + 0]) + b0.x, __uint_as_float(r[8 * q + 1]) + b0.y);
    hp[1] = h2_relu_sat(__uint_as_float(r[8 * q + 2]) + b0.z, __uint_as_float(r[8 * q + 3]) + b0.w);
    hp[2] = h2_relu_sat(__uint_as_float(r[8 * q + 4]) + b1.x, __uint_as_float(r[8 * q + 5]) + b1.y);
    hp[3] = h2_relu_sat(__uint_as_float(r[8 * q + 6]) + b1.z, __uint_as_float(r[8 * q + 7]) + b1.w);
    op[(((sw >> 3) & 7) + q) ^ (sw & 7)] = o;
  }
}

// K-major, no swizzle: core matrix = 8 rows x 16 bytes, rows 16 bytes apart
__device__ __forceinline__ uint64_t desc_nosw(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;                                            // layout_type 0 = no swizzle
}

__global__ void __launch_bounds__(kThreadsF, 1)
tc_conv_first_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_o,
                         const FirstGeom g, const __half* __restrict__ wpk, const float* __restrict__ bias) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_o = smem;                                            // 2 x 16 KB, 1024-aligned (SWIZZLE_128B)
  uint8_t* smem_w = smem + 2 * kOutTile;
  uint8_t* smem_p = smem_w + kWBytesF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + kStagesF * kPatchBytesF);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStagesF;
  uint64_t* tmem_full = bars + 2 * kStagesF;
  uint64_t* tmem_empty = bars + 2 * kStagesF + 2;
  uint64_t* w_full = bars + 2 * kStagesF + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStagesF + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pt = g.tiles_h * g.tiles_w;
  pdl_trigger();
  if (warp == 0 && lane == 0) { tc::prefetch_tmap(&tmap_x); tc::prefetch_tmap(&tmap_o); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStagesF; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full[a], 1); tc::mbar_init(&tmem_empty[a], 4); }
    tc::mbar_init(w_full, 1);
    tc::fence_mbar_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, 128);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ producer ====================================
    if (tc::elect_one()) {
      tc::mbar_expect_tx(w_full, kWBytesF);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       tc::smem_u32(smem_w)),
                   "l"(wpk), "r"(kWBytesF), "r"(tc::smem_u32(w_full))
                   : "memory");
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int pt = blockIdx.x; pt < n_pt; pt += gridDim.x) {
      const int th = pt / g.tiles_w, tw = pt - th * g.tiles_w;
      tc::mbar_wait(&empty_bar[stage], phase ^ 1);
      if (tc::elect_one()) {
        tc::mbar_expect_tx(&full_bar[stage], kPW * kPH * 16);
        tc::tma_load_3d(smem_p + stage * kPatchBytesF, &tmap_x, &full_bar[stage], 0, tw * kTW - 1, th * kTH - 1 + g.hoff);
      }
      __syncwarp();
      if (++stage == kStagesF) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc = tc::idesc_f16(128, 64, 0, 0);
    const uint32_t w_addr = tc::smem_u32(smem_w), p_addr0 = tc::smem_u32(smem_p);
    tc::mbar_wait(w_full, 0);
    tc::fence_after_sync();
    int stage = 0; uint32_t phase = 0;
    int local = 0;
    for (int pt = blockIdx.x; pt < n_pt; pt += gridDim.x, ++local) {
      const int acc = local & 1;
      tc::mbar_wait(&tmem_empty[acc], ((local >> 1) & 1) ^ 1);
      tc::mbar_wait(&full_bar[stage], phase);
      tc::fence_after_sync();
      if (tc::elect_one()) {
        const uint32_t d_tmem = tmem_base + acc * 64;
        const uint32_t p_addr = p_addr0 + stage * kPatchBytesF;
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              // A: tile pixel (lh, lw) -> patch pixel (lh + r, lw + 2j [+1 for K chunk 1])
              const uint64_t a_desc = desc_nosw(p_addr + (r * kPW + 2 * j) * 16, 16, kPW * 16);
              // B: [part][r][j][chunk][co][8]: chunks 1024 B apart, 8-row groups 128 B apart
              const uint64_t b_desc = desc_nosw(w_addr + ((part * 3 + r) * 2 + j) * 2048, 1024, 128);
              tc::umma_f16(d_tmem, a_desc, b_desc, idesc, (part | r | j) != 0);
            }
        tc::umma_commit(&empty_bar[stage]);
        tc::umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
      if (++stage == kStagesF) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ====================================
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    int local = 0;
    for (int pt = blockIdx.x; pt < n_pt; pt += gridDim.x, ++local) {
      const int acc = local & 1;
      const int th = pt / g.tiles_w, tw = pt - th * g.tiles_w;
      if (lane == 0) tc::mbar_wait(&tmem_full[acc], (local >> 1) & 1);
      __syncwarp();
      tc::fence_after_sync();
      uint8_t* stg = smem_o + acc * kOutTile;
      if (warp == 4 && lane == 0) tc::bulk_wait_read<1>();
      tc::named_bar_sync(1, 128);
      __half* srow = reinterpret_cast<__half*>(stg + row * 128);
      const uint32_t t_row = tmem_base + acc * 64 + ((uint32_t)(ew * 32) << 16);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(t_row + c * 32, r);
        tc::tmem_ld_wait();
        epi_chunk_first(r, bias + c * 32, srow, 64 | ((c * 4) << 3) | (row & 7));
      }
      tc::fence_before_sync();
      tc::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
      tc::named_bar_sync(1, 128);
      if (warp == 4 && lane == 0) {
        tc::tma_store_3d(&tmap_o, stg, 0, tw * kTW, th * kTH);
        tc::bulk_commit();
      }
    }
    if (warp == 4 && lane == 0) tc::bulk_wait<0>();
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace

struct TcFirstPlan {
  CUtensorMap tmap_x, tmap_o;
  FirstGeom g;
  uint4* x8;
  const void* out_base;
};

int tc_first_plan_create(st2_ctx* ctx, int H, int W, int halo_strip, TcFirstPlan** out) {
  if (H < 16 || W < 16) return st2_fail(ctx, ST2_ERR_ARG, "tc_first: canvas too small");
  TcFirstPlan* p = new TcFirstPlan();
  p->g.H = H; p->g.W = W; p->g.hoff = 1;
  p->g.tiles_h = (H + kTH - 1) / kTH;
  p->g.tiles_w = (W + kTW - 1) / kTW;
  p->out_base = nullptr;
  (void)halo_strip;
  cudaError_t e = cudaMalloc(&p->x8, sizeof(uint4) * (size_t)(H + 2) * W);
  if (e != cudaSuccess) { delete p; return st2_fail(ctx, ST2_ERR_CUDA, "tc_first: x8 buffer: %s", cudaGetErrorString(e)); }
  cuuint64_t dims[3] = {8, (cuuint64_t)W, (cuuint64_t)(H + 2)};
  cuuint64_t strides[2] = {16, (cuuint64_t)W * 16};
  cuuint32_t box[3] = {8, (cuuint32_t)kPW, (cuuint32_t)kPH};
  int rc = st2_encode_tmap_ex(ctx, &p->tmap_x, p->x8, 3, dims, strides, box, 0);
  if (rc) { cudaFree(p->x8); delete p; return rc; }
  *out = p;
  return 0;
}

void tc_first_plan_destroy(TcFirstPlan* p) {
  if (!p) return;
  cudaFree(p->x8);
  delete p;
}

int tc_first_pack_weights(st2_ctx* ctx, const float* w_oihw, __half* out) {
  pack_w_first_kernel<<<(3 * 2 * 2 * 64 * 8 + 255) / 256, 256, 0, ctx->stream>>>(w_oihw, out);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

int tc_first_fwd_launch(st2_ctx* ctx, TcFirstPlan* p, const float* x, long long xps, int lo, int hi, const __half* wpk,
                        const float* bias, __half* out) {
  const int H = p->g.H, W = p->g.W;
  if (p->out_base != (const void*)out) {
    cuuint64_t dims[3] = {64, (cuuint64_t)W, (cuuint64_t)H};
    cuuint64_t strides[2] = {128, (cuuint64_t)W * 128};
    cuuint32_t box[3] = {64, (cuuint32_t)kTW, (cuuint32_t)kTH};
    int rc = st2_encode_tmap_ex(ctx, &p->tmap_o, out, 3, dims, strides, box, 1);
    if (rc) return rc;
    p->out_base = out;
  }
  long long blocks = ((long long)(H + 2) * W + 255) / 256;
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  st2_launch_pdl(ctx, true, pack_x8_kernel, (int)blocks, 256, 0, x, xps ? xps : (long long)H * W, p->x8, H, W, lo, hi);
  ST2_LAUNCH_CHECK(ctx);
  const int n_pt = p->g.tiles_h * p->g.tiles_w;
  const int grid = n_pt < ctx->sm_count ? n_pt : ctx->sm_count;
  st2_launch_pdl(ctx, true, tc_conv_first_fwd_kernel, grid, kThreadsF, kSmemF, p->tmap_x, p->tmap_o, p->g, wpk, bias);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

static St2SmemReg g_smem_first_tc({{ST2_KFN(tc_conv_first_fwd_kernel), kSmemF}});
static St2KernelReg g_reg_first_tc({ST2_KFN(pack_x8_kernel), ST2_KFN(pack_w_first_kernel), ST2_KFN(tc_conv_first_fwd_kernel)});

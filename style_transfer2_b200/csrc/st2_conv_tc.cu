// tcgen05 implicit-GEMM 3x3 / 1x1 convolution on fp16 NHWC activations (K1, K2, K6 of SURVEY 2.4).
//
//   out[p, n] = epi( sum_{tap, c} in[p + off(tap), c] * w[n, tap, c] )        M = pixels, N = cout, K = taps*cin
//
// * A operand: for every (tap, 64-channel block) one TMA *tiled* 3-D box {64 ch, TW, TH} of the NHWC
//   tensor at pixel offset (h0 + dh, w0 + dw).  Out-of-bounds rows/columns are zero-filled by the
//   TMA unit, which IS the pad-1 border; the box lands in shared memory as a 128-row x 128-byte
//   K-major SWIZZLE_128B tile, exactly the UMMA canonical layout -- no im2col buffer exists anywhere.
// * B operand: packed weights [n][tap][c] (K-major), 2-D box {64, BN}.
// * MMA: tcgen05.mma cta_group::1 kind::f16, 128 x BN x 16 per instruction, fp32 accumulators in
//   TMEM, double-buffered (2 x BN columns) so the epilogue of tile i overlaps the main loop of i+1.
// * Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-7 epilogue
//   (tcgen05.ld 32x32b -> bias/ReLU | ReLU-mask | raw -> fp16 -> 64-byte vector stores).
// * Persistent: grid = min(tiles, SMs); static round-robin tile order with the cout block fastest
//   so CTAs working on the same pixels at the same time share the A tile through L2.
// The same kernel runs forward (weights [co][tap][ci]), data-gradient (weights [ci][tap'][co],
// taps flipped) and the style-gradient 1x1 contraction (weights = scaled Gram difference).
#include "st2_kernels.h"
#include "st2_tc.cuh"

#include <stdlib.h>
#include <string.h>

namespace {

constexpr int BM = 128;          // pixels per tile
constexpr int BK = 64;           // fp16 elements per K block = 128 bytes = one swizzle row
constexpr int kNumThreads = 256;
constexpr int kEpiWarp0 = 4;

template <int BN> struct Cfg {
  // One pipeline stage carries KB consecutive 64-wide K blocks: the barrier hand-shake of a stage
  // costs ~450 cycles of issue latency, so a stage must hold >= that much MMA work (2*BN cycles
  // per K block at M = 128).
  static constexpr int KB = (BN == 256) ? 1 : (BN == 128 ? 2 : 3);
  static constexpr int kABlock = BM * BK * 2;                 // 16 KB per K block
  static constexpr int kBBlock = BN * BK * 2;
  static constexpr int kABytes = KB * kABlock;                // per stage
  static constexpr int kBBytes = KB * kBBlock;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : 3;         // 192 KB / 192 KB / 216 KB
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct ConvGeom {
  int H, W, cin, cout, taps;
  int TH, TW, tiles_h, tiles_w, n_blocks, total_tiles, k_iters, cblocks;
  int step_nb, step_tw, step_th;   // decomposition of gridDim.x in (n block, tile column, tile row) digits
  int hoff;  // row strips: the tensor map starts `hoff` halo rows above the first output row
  int dbg;   // st2_debug_flags(): 1 no epilogue stores, 2 no MMA, 4 no A loads, 8 no B loads (timing experiments)
  int rot;   // 1: tile rows are visited in the order 1, 2, .., tiles_h - 1, 0 (the two that touch halo rows last)
  HaloArgs halo;
};

// The kernels are instantiated with and without the in-kernel halo exchange (template parameter HALO): measured on
// one box, the few per-tile instructions of the exchange cost the whole-canvas launches ~2 % when merely compiled in.
// tile row visited at position `th` of the schedule
template <bool HALO>
__device__ __forceinline__ int rot_row(const ConvGeom& g, int th) {
  return HALO ? (th + 1 == g.tiles_h ? 0 : th + 1) : th;
}
// does the tile row at schedule position `th` read a halo row?
template <bool HALO>
__device__ __forceinline__ bool halo_row(const ConvGeom& g, int th) { return HALO && th >= g.tiles_h - 2; }

// All threads of the first push_blocks CTAs, at kernel start.
__device__ __forceinline__ void halo_push_prologue(const HaloArgs& h) {
  if (h.push_blocks == 0) return;
  // the LAST CTAs of the grid push: with a partial last wave of tiles they are the ones with a tile less to do
  const int npush = h.push_blocks < (int)gridDim.x ? h.push_blocks : (int)gridDim.x;
  const int pb = (int)blockIdx.x - ((int)gridDim.x - npush);
  if (pb < 0) return;
  const long long n16 = h.bytes >> 4;
#pragma unroll
  for (int dir = 0; dir < 2; ++dir) {
    if (h.dst[dir] == nullptr) continue;
    const uint4* s4 = reinterpret_cast<const uint4*>(h.src[dir]);
    uint4* d4 = reinterpret_cast<uint4*>(h.dst[dir]);
    for (long long i = (long long)pb * blockDim.x + threadIdx.x; i < n16; i += (long long)npush * blockDim.x)
      d4[i] = s4[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(h.counter, 1u);
    if (prev == (unsigned int)npush - 1) {
      atomicExch(h.counter, 0u);
      __threadfence_system();
#pragma unroll
      for (int dir = 0; dir < 2; ++dir)
        if (h.flag[dir] != nullptr)
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(h.flag[dir]), "l"(h.epoch) : "memory");
    }
  }
}

// One lane of the TMA producer warp, right before the first tile that reads a halo row.
__device__ __forceinline__ void halo_wait_flags(const HaloArgs& h) {
#pragma unroll
  for (int dir = 0; dir < 2; ++dir) {
    const unsigned long long* f = h.wait[dir];
    if (f == nullptr || *reinterpret_cast<volatile int*>(h.err) != 0) continue;
    unsigned long long t0, t1, v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= h.epoch) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 10000000000ull) { atomicExch(h.err, 1); break; }     // 10 s: the neighbour died; fail loudly
      __nanosleep(32);
    }
  }
}
// whole producer warp: returns true once the halo rows are known to be in place
template <bool HALO>
__device__ __forceinline__ bool halo_ready(const ConvGeom& g, bool waited, int th) {
  if (!HALO || waited || !halo_row<HALO>(g, th)) return waited;
  if ((threadIdx.x & 31) == 0) halo_wait_flags(g.halo);
  __syncwarp();
  asm volatile("fence.proxy.async;" ::: "memory");       // the rows are read by the TMA unit (async proxy)
  return true;
}

// Tile coordinates advanced by gridDim.x per step without divisions (mixed-radix add with carry).
struct TileWalk {
  int nb, tw, th;
  __device__ __forceinline__ void init(const ConvGeom& g, int tile) {
    nb = tile % g.n_blocks;
    const int pt = tile / g.n_blocks;
    tw = pt % g.tiles_w;
    th = pt / g.tiles_w;
  }
  __device__ __forceinline__ void next(const ConvGeom& g) {
    nb += g.step_nb;
    if (nb >= g.n_blocks) { nb -= g.n_blocks; ++tw; }
    tw += g.step_tw;
    if (tw >= g.tiles_w) { tw -= g.tiles_w; ++th; }
    if (tw >= g.tiles_w) { tw -= g.tiles_w; ++th; }
    th += g.step_th;
  }
};

// fused 2x2/2 max-pool of a forward epilogue: p = this lane's pooled pixel (32 channels of it), vert = lane distance
// of the pixel one row below (the tile width), writer = this lane is the top-left pixel of its window
struct PoolOut { __half* p; int vert; bool writer; };

__device__ __forceinline__ uint4 hmax_u4(const uint4& a, const uint4& b) {
  uint4 r;
  const __half2* x = reinterpret_cast<const __half2*>(&a);
  const __half2* y = reinterpret_cast<const __half2*>(&b);
  __half2* z = reinterpret_cast<__half2*>(&r);
#pragma unroll
  for (int e = 0; e < 4; ++e) z[e] = __hmax2(x[e], y[e]);
  return r;
}

__device__ __forceinline__ uint4 shfl_xor_u4(const uint4& v, const int m) {
  uint4 r;
  r.x = __shfl_xor_sync(0xffffffffu, v.x, m);
  r.y = __shfl_xor_sync(0xffffffffu, v.y, m);
  r.z = __shfl_xor_sync(0xffffffffu, v.z, m);
  r.w = __shfl_xor_sync(0xffffffffu, v.w, m);
  return r;
}

// One 32-channel chunk of one output pixel: accumulator -> bias+ReLU | ReLU mask (+ loss injection) | raw
// -> fp16 (saturating) -> four 16-byte stores.  a4 / s4: the pixel's 32 channels of the mask source and
// of the style-gradient injection, already in registers.
__device__ __forceinline__ void epi_chunk(const uint32_t (&r)[32], const int epi, const float* __restrict__ bias_c,
                                          const uint4 (&a4)[4], const uint4 (&s4)[4], const __half* __restrict__ fc_c,
                                          const bool have_inj, const bool have_s, const float cc, const float sc,
                                          const float dc, const float out_scale, const bool want_ss, float& ss,
                                          __half* __restrict__ out_c, const int sw = 0, const bool valid = true,
                                          const long long pix_stride = 0, const PoolOut po = PoolOut{nullptr, 0, false},
                                          const uint32_t* r2 = nullptr) {
  // Called by ALL lanes of the warp (shuffles inside); `valid` = this lane's pixel exists.  pix_stride = elements
  // between horizontally adjacent pixels of the output (cout).
  // out_c: where this chunk's four 16-byte pieces go.  sw = 0: consecutive (global memory).  sw != 0: out_c is
  // a 128-byte row of a SWIZZLE_128B staging tile in shared memory, pieces q land at ((chunk0 + q) ^ row%8)
  // with chunk0 = sw >> 3 and row%8 = sw & 7 (bit 6 set marks the mode).
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  uint4 o4[4];
  // The two common forms first, in packed fp16 arithmetic (the epilogue warps are bound by their instruction count):
  // bias + ReLU + saturation = one add and half a convert per value; a bare ReLU mask = half a convert, half a
  // compare and half an AND.  Both give the bits of the fp32 forms below.
  const bool fast_relu = epi == EPI_BIAS_RELU;
  // an injection with nothing to inject is a bare mask too: the launch above a style layer whose gradient is folded
  // (conv1_2's data gradient, the most expensive launch of the iteration) carries the coefficient block but no content
  // target, no style tensor and a zero deep-dream coefficient
  const bool inj_active = have_inj && (fc_c != nullptr || have_s || dc != 0.f);
  const bool fast_mask = epi == EPI_MASK && !inj_active && r2 == nullptr;
  if (fast_relu) {
    const float4* bp = reinterpret_cast<const float4*>(bias_c);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b0 = __ldg(bp + 2 * q), b1 = __ldg(bp + 2 * q + 1);
      __half2* hp = reinterpret_cast<__half2*>(&o4[q]);
      hp[0] = h2_relu_sat(v[8 * q + 0] + b0.x, v[8 * q + 1] + b0.y);
      hp[1] = h2_relu_sat(v[8 * q + 2] + b0.z, v[8 * q + 3] + b0.w);
      hp[2] = h2_relu_sat(v[8 * q + 4] + b1.x, v[8 * q + 5] + b1.y);
      hp[3] = h2_relu_sat(v[8 * q + 6] + b1.z, v[8 * q + 7] + b1.w);
    }
  } else if (fast_mask) {
    const __half2 zero = __float2half2_rn(0.f);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __half2* hp = reinterpret_cast<const __half2*>(&a4[q]);
      uint32_t* op = reinterpret_cast<uint32_t*>(&o4[q]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __half2 h = h2_sat(v[8 * q + 2 * e], v[8 * q + 2 * e + 1]);
        op[e] = *reinterpret_cast<const uint32_t*>(&h) & __hgt2_mask(hp[e], zero);
      }
    }
  } else if (epi == EPI_MASK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __half2* hp = reinterpret_cast<const __half2*>(&a4[q]);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(hp[e]);
        if (!(f.x > 0.f)) v[8 * q + 2 * e] = 0.f;
        if (!(f.y > 0.f)) v[8 * q + 2 * e + 1] = 0.f;
      }
      if (have_inj) {
        // loss diffs of the blob below enter under its ReLU mask (worker.py:100-102)
        if (fc_c != nullptr) {
          const uint4 c4 = __ldg(reinterpret_cast<const uint4*>(fc_c) + q);
          const __half2* cp = reinterpret_cast<const __half2*>(&c4);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(hp[e]), t = __half22float2(cp[e]);
            v[8 * q + 2 * e] = fmaf(cc, f.x - t.x, v[8 * q + 2 * e]);
            v[8 * q + 2 * e + 1] = fmaf(cc, f.y - t.y, v[8 * q + 2 * e + 1]);
          }
        }
        if (have_s) {
          const __half2* sp = reinterpret_cast<const __half2*>(&s4[q]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 t = __half22float2(sp[e]);
            v[8 * q + 2 * e] = fmaf(sc, t.x, v[8 * q + 2 * e]);
            v[8 * q + 2 * e + 1] = fmaf(sc, t.y, v[8 * q + 2 * e + 1]);
          }
        }
        if (dc != 0.f) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(hp[e]);
            v[8 * q + 2 * e] = fmaf(dc, f.x, v[8 * q + 2 * e]);
            v[8 * q + 2 * e + 1] = fmaf(dc, f.y, v[8 * q + 2 * e + 1]);
          }
        }
      }
    }
    if (r2 != nullptr) {
      // the style gradient D' F of the same pixels, contracted by this kernel into a second accumulator (fp32):
      // it enters below the ReLU mask like the other loss diffs
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaf(sc, __uint_as_float(r2[j]), v[j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= out_scale;
    if (want_ss) {
#pragma unroll
      for (int j = 0; j < 32; ++j) ss = fmaf(v[j], v[j], ss);
    }
  }
  if (!fast_relu && !fast_mask) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      __half2* hp = reinterpret_cast<__half2*>(&o4[q]);
#pragma unroll
      for (int e = 0; e < 4; ++e) hp[e] = h2_sat(v[8 * q + 2 * e], v[8 * q + 2 * e + 1]);
    }
  }
  if (po.p != nullptr || po.vert != 0) {
    // Caffe's ceil-mode 2x2/2 max pool of the post-ReLU output, first in registers: the window's other pixels sit in
    // lanes l^1 (next column) and l^vert (next row); pixels beyond the canvas contribute 0, the neutral element
    // of a max over post-ReLU values.  (po.vert != 0 is warp-uniform: every lane takes part in the shuffles.)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 m = valid ? o4[q] : make_uint4(0, 0, 0, 0);
      m = hmax_u4(m, shfl_xor_u4(m, 1));
      m = hmax_u4(m, shfl_xor_u4(m, po.vert));
      if (po.writer && valid) reinterpret_cast<uint4*>(po.p)[q] = m;
    }
  }
  if (sw != 0) {
    uint4* op = reinterpret_cast<uint4*>(out_c);
#pragma unroll
    for (int q = 0; q < 4; ++q) op[(((sw >> 3) & 7) + q) ^ (sw & 7)] = o4[q];
    return;
  }
  // Direct stores.  A lane holds 64 contiguous bytes of ITS pixel; storing them as they are makes every store
  // instruction touch 32 different lines with 16 bytes each (measured: the epilogue then tops out at ~1.8 TB/s and
  // bounds every layer with <= 128 output channels).  4 x 4 transpose of the 16-byte pieces inside each quad of
  // lanes (quads = 4 horizontally adjacent pixels): afterwards slot j holds piece (lane & 3) of pixel quad + j, and
  // store j writes 64 contiguous bytes per quad.
  const int lane = threadIdx.x & 31, me = lane & 3;
  {
    const bool b0 = (lane & 1) != 0;
    uint4 snd = b0 ? o4[0] : o4[1], rcv = shfl_xor_u4(snd, 1);
    if (b0) o4[0] = rcv; else o4[1] = rcv;
    snd = b0 ? o4[2] : o4[3]; rcv = shfl_xor_u4(snd, 1);
    if (b0) o4[2] = rcv; else o4[3] = rcv;
    const bool b1 = (lane & 2) != 0;
    snd = b1 ? o4[0] : o4[2]; rcv = shfl_xor_u4(snd, 2);
    if (b1) o4[0] = rcv; else o4[2] = rcv;
    snd = b1 ? o4[1] : o4[3]; rcv = shfl_xor_u4(snd, 2);
    if (b1) o4[1] = rcv; else o4[3] = rcv;
  }
  const unsigned vmask = __ballot_sync(0xffffffffu, valid);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if ((vmask >> ((lane & ~3) + j)) & 1u)
      *reinterpret_cast<uint4*>(out_c + (long long)(j - me) * pix_stride + me * 8) = o4[j];
}

template <int BN, bool HALO>
__global__ void __launch_bounds__(kNumThreads, 1)
tc_conv_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const ConvGeom g, const float* __restrict__ bias, const __half* __restrict__ act,
               __half* __restrict__ out, const int epi, const float out_scale, double* sumsq,
               const TcInject inj) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by SWIZZLE_128B operand tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;                         // [kStages]
  uint64_t* empty_bar = bars + C::kStages;           // [kStages]
  uint64_t* tmem_full = bars + 2 * C::kStages;       // [2]
  uint64_t* tmem_empty = bars + 2 * C::kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (HALO) halo_push_prologue(g.halo);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a);
    tc::prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full[a], 1); tc::mbar_init(&tmem_empty[a], 4); }
    tc::fence_mbar_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, C::kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  pdl_wait();                      // everything above ran in the predecessor's shadow; global memory from here on
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // whole warp runs the (uniform) loop; one elected lane issues expect_tx + the TMA loads of a stage
    int stage = 0; uint32_t phase = 0;
    const bool three = (g.taps == 9);
    bool waited = false;
    TileWalk tk;
    tk.init(g, blockIdx.x);
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, tk.next(g)) {
      const int nb = tk.nb;
      waited = halo_ready<HALO>(g, waited, tk.th);
      const int h0 = rot_row<HALO>(g, tk.th) * g.TH, w0 = tk.tw * g.TW;
      int tap = 0, cb = 0;                               // K block index it = tap * cblocks + cb
      for (int it = 0; it < g.k_iters; it += C::KB) {
        const int nkb = (g.k_iters - it < C::KB) ? g.k_iters - it : C::KB;
        tc::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (tc::elect_one()) {
          uint32_t bytes = (uint32_t)nkb * (C::kABlock + C::kBBlock);
          if (g.dbg & 4) bytes -= (uint32_t)nkb * C::kABlock;
          if (g.dbg & 8) bytes -= (uint32_t)nkb * C::kBBlock;
          if (bytes) tc::mbar_expect_tx(&full_bar[stage], bytes); else tc::mbar_arrive(&full_bar[stage]);
        }
#pragma unroll
        for (int j = 0; j < C::KB; ++j) {
          if (j < nkb) {
            const int dh = three ? tap / 3 - 1 : 0;
            const int dw = three ? tap % 3 - 1 : 0;
            if (tc::elect_one()) {
              if (!(g.dbg & 4))
                tc::tma_load_3d(smem_a + stage * C::kABytes + j * C::kABlock, &tmap_a, &full_bar[stage], cb * BK,
                                w0 + dw, h0 + dh + g.hoff);
              if (!(g.dbg & 8))
                tc::tma_load_2d(smem_b + stage * C::kBBytes + j * C::kBBlock, &tmap_b, &full_bar[stage],
                                tap * g.cin + cb * BK, nb * BN);
            }
            if (++cb == g.cblocks) { cb = 0; ++tap; }
          }
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc = tc::idesc_f16(BM, BN, 0, 0);
    const uint64_t a_desc0 = tc::smem_desc_k_sw128(tc::smem_u32(smem_a));
    const uint64_t b_desc0 = tc::smem_desc_k_sw128(tc::smem_u32(smem_b));
    int stage = 0; uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      tc::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int it = 0; it < g.k_iters; it += C::KB) {
        const int nkb = (g.k_iters - it < C::KB) ? g.k_iters - it : C::KB;
        tc::mbar_wait(&full_bar[stage], phase);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint64_t a_desc = a_desc0 + (uint64_t)(stage * (C::kABytes >> 4));
          const uint64_t b_desc = b_desc0 + (uint64_t)(stage * (C::kBBytes >> 4));
          if (!(g.dbg & 2)) {
#pragma unroll
            for (int j = 0; j < C::KB; ++j) {
              if (j < nkb) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k)        // +32 bytes (>>4 = 2) per 16-element K step
                  tc::umma_f16(d_tmem, a_desc + (uint64_t)(j * (C::kABlock >> 4) + 2 * k),
                               b_desc + (uint64_t)(j * (C::kBBlock >> 4) + 2 * k), idesc, (it | j | k) != 0);
              }
            }
          }
          if (g.dbg & 16) tc::mbar_arrive(&empty_bar[stage]);   // experiment: plain arrive instead of commit
          else tc::umma_commit(&empty_bar[stage]);     // frees the smem slot when the MMAs retire
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(&tmem_full[acc]);    // accumulator complete -> epilogue
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // ================================ epilogue ====================================
    const int ew = warp - kEpiWarp0;                   // == warp % 4: TMEM lane quarter of this warp
    const int row = ew * 32 + lane;                    // pixel index inside the tile
    float ss = 0.f;
    int local = 0;
    const int row_h = row / g.TW, row_w = row % g.TW;
    // loss injection fused into the data-gradient epilogue (worker.py:249-277 diffs added below the ReLU mask)
    float cc = 0.f, sc = 0.f, dc = 0.f;
    if (inj.coef != nullptr) { cc = (float)inj.coef[0]; sc = (float)inj.coef[1]; dc = (float)inj.coef[2]; }
    constexpr int NCH = BN / 32;
    constexpr bool PF = (BN <= 128);
    const bool masked = (epi == EPI_MASK);
    const bool have_inj = masked && inj.coef != nullptr;
    const bool have_s = have_inj && inj.sraw != nullptr;
    TileWalk tk;
    tk.init(g, blockIdx.x);
    for (int tile = blockIdx.x; tile < g.total_tiles; tile += gridDim.x, ++local, tk.next(g)) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int nb = tk.nb;
      const int h = rot_row<HALO>(g, tk.th) * g.TH + row_h;
      const int w = tk.tw * g.TW + row_w;
      const bool valid = (h < g.H) && (w < g.W) && !(g.dbg & 1);
      const long long obase = ((long long)h * g.W + w) * g.cout + (long long)nb * BN;
      const bool do_pool = (epi == EPI_BIAS_RELU) && inj.pool != nullptr;
      __half* const pool_px = do_pool ? inj.pool + ((long long)(h >> 1) * inj.pool_wp + (w >> 1)) * g.cout + (long long)nb * BN
                                      : nullptr;
      const bool pool_writer = do_pool && !(lane & 1) && !(lane & g.TW);
      const int pool_vert = do_pool ? g.TW : 0;
      // Operands of the epilogue that do not depend on the accumulator (ReLU-mask source, style-gradient
      // injection) are fetched BEFORE waiting for the MMAs of this tile on the narrow tiles: those layers
      // are memory-bound and a DRAM round trip per 32-channel chunk would otherwise serialise behind the wait.
      uint4 pa[PF ? NCH : 1][4], ps[PF ? NCH : 1][4];
      if (PF) {
        if (valid && masked) {
          const uint4* ap = reinterpret_cast<const uint4*>(act + obase);
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) pa[c][q] = __ldg(ap + c * 4 + q);
        }
        if (valid && have_s) {
          const uint4* sp = reinterpret_cast<const uint4*>(inj.sraw + obase);
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) ps[c][q] = __ldg(sp + c * 4 + q);
        }
      }
      if (lane == 0) tc::mbar_wait(&tmem_full[acc], acc_phase);     // one poller per warp
      __syncwarp();
      tc::fence_after_sync();
      const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16);
      if (PF) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t r[32];
          tc::tmem_ld_32x32(t_row + c * 32, r);
          tc::tmem_ld_wait();
          epi_chunk(r, epi, bias + nb * BN + c * 32, pa[PF ? c : 0], ps[PF ? c : 0],
                    (valid && have_inj && inj.fc != nullptr) ? inj.fc + obase + c * 32 : nullptr, have_inj, have_s, cc, sc, dc,
                    out_scale, valid && sumsq != nullptr, ss, out + obase + c * 32, 0, valid, g.cout,
                    PoolOut{do_pool ? pool_px + c * 32 : nullptr, pool_vert, pool_writer});
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          uint32_t r[32];
          tc::tmem_ld_32x32(t_row + c * 32, r);
          tc::tmem_ld_wait();
          uint4 a4[4], s4[4];
          if (valid && masked) {
            const uint4* ap = reinterpret_cast<const uint4*>(act + obase + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) a4[q] = __ldg(ap + q);
          }
          if (valid && have_s) {
            const uint4* sp = reinterpret_cast<const uint4*>(inj.sraw + obase + c * 32);
#pragma unroll
            for (int q = 0; q < 4; ++q) s4[q] = __ldg(sp + q);
          }
          epi_chunk(r, epi, bias + nb * BN + c * 32, a4, s4,
                    (valid && have_inj && inj.fc != nullptr) ? inj.fc + obase + c * 32 : nullptr, have_inj, have_s, cc, sc, dc,
                    out_scale, valid && sumsq != nullptr, ss, out + obase + c * 32, 0, valid, g.cout,
                    PoolOut{do_pool ? pool_px + c * 32 : nullptr, pool_vert, pool_writer});
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
    }
    if (sumsq != nullptr) {
      const double tot = warp_sum_d((double)ss);
      if (lane == 0) atomicAdd(sumsq, tot);
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ================================================================================================
// CTA-pair variant (cta_group::2) for the wide layers: two SMs of a TPC run one M = 256 x BN tile.
// Each CTA loads its own 128 pixels of A and HALF of the weight tile (BN / 2 rows); the MMA reads B from
// both shared memories, so per SM the operand traffic from shared memory drops from (128 + BN) to
// (128 + BN / 2) rows per K step -- the single-CTA kernel at N = 256 is at ~96 B/clk of the 128 B/clk shared
// memory port.  The pair tile is 16 rows x 16 pixels; CTA rank r owns rows 8r .. 8r+7 of it.
// SFUSE (BN = 128 data gradient above a 128-channel style layer, i.e. conv2_2's above conv2_1): after the 9 x 2 K blocks
// of the convolution every tile gets ONE more pipeline stage -- the tile's own 256 x 128 activations F of the blob below
// (tmap_f, centre tap) against the scaled Gram difference D' (tmap_d, 128 x 128) -- into a second accumulator:
// the style gradient D' F, which the epilogue adds under the mask with the device coefficient.  That replaces a 1 x 1
// contraction launch, its 67 MB fp16 output at 1024^2 and the read of that output in this epilogue.
template <int BN, bool HALO, bool SFUSE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
tc_conv2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_f, const __grid_constant__ CUtensorMap tmap_d,
                const ConvGeom g, const float* __restrict__ bias, const __half* __restrict__ act,
                __half* __restrict__ out, const int epi, const TcInject inj) {
  static_assert(!SFUSE || BN == 128, "style fusion: 128-channel layers only (TMEM holds 4 x 128 columns)");
  constexpr int KB = (BN == 256) ? 1 : 2;                      // K blocks per stage
  constexpr int kABlock = BM * BK * 2;                         // 16 KB
  constexpr int kBBlock = (BN / 2) * BK * 2;                   // this CTA's half of the weight tile
  constexpr int kABytes = KB * kABlock, kBBytes = KB * kBBlock;
  constexpr int kStages = (BN == 256) ? 6 : 4;
  constexpr int kTmemCols = SFUSE ? 4 * BN : 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kABytes + kBBytes));
  uint64_t* full_bar = bars;                         // used on the leader: both CTAs' TMA bytes land here
  uint64_t* empty_bar = bars + kStages;              // local copy in each CTA (multicast commit)
  uint64_t* tmem_full = bars + 2 * kStages;          // local copy in each CTA (multicast commit)
  uint64_t* tmem_empty = bars + 2 * kStages + 2;     // used on the leader: 4 local + 4 remote epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  pdl_trigger();
  if (HALO) halo_push_prologue(g.halo);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a); tc::prefetch_tmap(&tmap_b);
    if (SFUSE) { tc::prefetch_tmap(&tmap_f); tc::prefetch_tmap(&tmap_d); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full[a], 1); tc::mbar_init(&tmem_empty[a], 8); }
    tc::fence_mbar_init();
  }
  if (warp == 2) tc::tmem_alloc_2sm(tmem_slot, kTmemCols);
  tc::fence_before_sync();
  tc::cluster_sync_all();
  tc::fence_after_sync();
  pdl_wait();                      // everything above ran in the predecessor's shadow; global memory from here on
  const uint32_t tmem_base = *tmem_slot;

  // pair tiles: (n block fastest, then tile column, then pair-tile row), strided by the number of pairs
  const int n_tiles = g.tiles_h * g.tiles_w * g.n_blocks;      // tiles_h counts 16-row pair tiles here

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    int stage = 0; uint32_t phase = 0;
    bool waited = false;
    for (int tile = pair; tile < n_tiles; tile += n_pairs) {
      const int nb = tile % g.n_blocks;
      const int pt = tile / g.n_blocks;
      waited = halo_ready<HALO>(g, waited, pt / g.tiles_w);
      const int w0 = (pt % g.tiles_w) * g.TW, h0 = rot_row<HALO>(g, pt / g.tiles_w) * (2 * g.TH) + (int)rank * g.TH;
      int tap = 0, cb = 0;
      for (int it = 0; it < g.k_iters; it += KB) {
        const int nkb = (g.k_iters - it < KB) ? g.k_iters - it : KB;
        tc::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (rank == 0 && tc::elect_one())
          tc::mbar_expect_tx(&full_bar[stage], 2u * (uint32_t)nkb * (kABlock + kBBlock));
#pragma unroll
        for (int j = 0; j < KB; ++j) {
          if (j < nkb) {
            const int dh = tap / 3 - 1, dw = tap % 3 - 1;
            if (tc::elect_one()) {
              tc::tma_load_3d_2sm(smem_a + stage * kABytes + j * kABlock, &tmap_a, &full_bar[stage], cb * BK, w0 + dw,
                                  h0 + dh + g.hoff);
              tc::tma_load_2d_2sm(smem_b + stage * kBBytes + j * kBBlock, &tmap_b, &full_bar[stage],
                                  tap * g.cin + cb * BK, nb * BN + (int)rank * (BN / 2));
            }
            if (++cb == g.cblocks) { cb = 0; ++tap; }
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (SFUSE) {
        // the style stage: this CTA's 128 pixels of F (both 64-channel blocks) and its half of D'
        tc::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (rank == 0 && tc::elect_one())
          tc::mbar_expect_tx(&full_bar[stage], 2u * (uint32_t)KB * (kABlock + kBBlock));
#pragma unroll
        for (int j = 0; j < KB; ++j) {
          if (tc::elect_one()) {
            tc::tma_load_3d_2sm(smem_a + stage * kABytes + j * kABlock, &tmap_f, &full_bar[stage], j * BK, w0, h0 + g.hoff);
            tc::tma_load_2d_2sm(smem_b + stage * kBBytes + j * kBBlock, &tmap_d, &full_bar[stage], j * BK,
                                nb * BN + (int)rank * (BN / 2));
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================================ MMA issuer (leader only) ================================
    constexpr uint32_t idesc = tc::idesc_f16(2 * BM, BN, 0, 0);
    const uint64_t a_desc0 = tc::smem_desc_k_sw128(tc::smem_u32(smem_a));
    const uint64_t b_desc0 = tc::smem_desc_k_sw128(tc::smem_u32(smem_b));
    int stage = 0; uint32_t phase = 0;
    int local = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs, ++local) {
      const int acc = local & 1;
      tc::mbar_wait(&tmem_empty[acc], ((local >> 1) & 1) ^ 1);
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int it = 0; it < g.k_iters; it += KB) {
        const int nkb = (g.k_iters - it < KB) ? g.k_iters - it : KB;
        tc::mbar_wait(&full_bar[stage], phase);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint64_t a_desc = a_desc0 + (uint64_t)(stage * (kABytes >> 4));
          const uint64_t b_desc = b_desc0 + (uint64_t)(stage * (kBBytes >> 4));
#pragma unroll
          for (int j = 0; j < KB; ++j) {
            if (j < nkb) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                tc::umma_f16_2sm(d_tmem, a_desc + (uint64_t)(j * (kABlock >> 4) + 2 * k),
                                 b_desc + (uint64_t)(j * (kBBlock >> 4) + 2 * k), idesc, (it | j | k) != 0);
            }
          }
          tc::umma_commit_2sm(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (SFUSE) {
        tc::mbar_wait(&full_bar[stage], phase);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint64_t a_desc = a_desc0 + (uint64_t)(stage * (kABytes >> 4));
          const uint64_t b_desc = b_desc0 + (uint64_t)(stage * (kBBytes >> 4));
#pragma unroll
          for (int j = 0; j < KB; ++j)
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc::umma_f16_2sm(d_tmem + 2 * BN, a_desc + (uint64_t)(j * (kABlock >> 4) + 2 * k),
                               b_desc + (uint64_t)(j * (kBBlock >> 4) + 2 * k), idesc, (j | k) != 0);
          tc::umma_commit_2sm(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit_2sm(&tmem_full[acc]);
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // ================================ epilogue (both CTAs, own 128 pixels) ====================
    const int ew = warp - kEpiWarp0;
    const int row = ew * 32 + lane;
    const int row_h = row / g.TW, row_w = row % g.TW;
    float ss = 0.f;
    float cc = 0.f, sc = 0.f, dc = 0.f;
    if (inj.coef != nullptr) { cc = (float)inj.coef[0]; sc = (float)inj.coef[1]; dc = (float)inj.coef[2]; }
    constexpr int NCH = BN / 32;
    const bool masked = (epi == EPI_MASK);
    const bool have_inj = masked && inj.coef != nullptr;
    const bool have_s = have_inj && inj.sraw != nullptr;
    int local = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs, ++local) {
      const int acc = local & 1;
      const int nb = tile % g.n_blocks;
      const int pt = tile / g.n_blocks;
      const int h = rot_row<HALO>(g, pt / g.tiles_w) * (2 * g.TH) + (int)rank * g.TH + row_h;
      const int w = (pt % g.tiles_w) * g.TW + row_w;
      const bool valid = (h < g.H) && (w < g.W) && !(g.dbg & 1);
      const long long obase = ((long long)h * g.W + w) * g.cout + (long long)nb * BN;
      const bool do_pool = (epi == EPI_BIAS_RELU) && inj.pool != nullptr;
      __half* const pool_px = do_pool ? inj.pool + ((long long)(h >> 1) * inj.pool_wp + (w >> 1)) * g.cout + (long long)nb * BN
                                      : nullptr;
      const bool pool_writer = do_pool && !(lane & 1) && !(lane & g.TW);
      const int pool_vert = do_pool ? g.TW : 0;
      // narrow tiles: epilogue operands fetched before the accumulator wait (see tc_conv_kernel)
      constexpr bool PF = (BN <= 128);
      uint4 pa[PF ? NCH : 1][4], ps[PF ? NCH : 1][4];
      if (PF) {
        if (valid && masked) {
          const uint4* ap = reinterpret_cast<const uint4*>(act + obase);
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) pa[c][q] = __ldg(ap + c * 4 + q);
        }
        if (valid && have_s) {
          const uint4* sp = reinterpret_cast<const uint4*>(inj.sraw + obase);
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) ps[c][q] = __ldg(sp + c * 4 + q);
        }
      }
      if (lane == 0) tc::mbar_wait(&tmem_full[acc], (local >> 1) & 1);
      __syncwarp();
      tc::fence_after_sync();
      const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16);
      if (PF) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t r[32];
          uint32_t r2[SFUSE ? 32 : 1];
          tc::tmem_ld_32x32(t_row + c * 32, r);
          if (SFUSE) tc::tmem_ld_32x32(t_row + 2 * BN + c * 32, reinterpret_cast<uint32_t(&)[32]>(r2));
          tc::tmem_ld_wait();
          epi_chunk(r, epi, bias + nb * BN + c * 32, pa[PF ? c : 0], ps[PF ? c : 0],
                    (valid && have_inj && inj.fc != nullptr) ? inj.fc + obase + c * 32 : nullptr, have_inj, have_s, cc, sc, dc,
                    1.f, false, ss, out + obase + c * 32, 0, valid, g.cout,
                    PoolOut{do_pool ? pool_px + c * 32 : nullptr, pool_vert, pool_writer}, SFUSE ? r2 : nullptr);
        }
      }
#pragma unroll 1
      for (int c = 0; c < (PF ? 0 : NCH); ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(t_row + c * 32, r);
        tc::tmem_ld_wait();
        uint4 a4[4], s4[4];
        if (valid && masked) {
          const uint4* ap = reinterpret_cast<const uint4*>(act + obase + c * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) a4[q] = __ldg(ap + q);
        }
        if (valid && have_s) {
          const uint4* sp = reinterpret_cast<const uint4*>(inj.sraw + obase + c * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) s4[q] = __ldg(sp + q);
        }
        epi_chunk(r, epi, bias + nb * BN + c * 32, a4, s4,
                  (valid && have_inj && inj.fc != nullptr) ? inj.fc + obase + c * 32 : nullptr, have_inj, have_s, cc, sc, dc,
                  1.f, false, ss, out + obase + c * 32, 0, valid, g.cout,
                  PoolOut{do_pool ? pool_px + c * 32 : nullptr, pool_vert, pool_writer});
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_leader(&tmem_empty[acc]);
    }
  }

  tc::fence_before_sync();
  tc::cluster_sync_all();
  if (warp == 2) {
    tc::fence_after_sync();
    tc::tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

// ================================================================================================
// Weight-stationary, halo-reuse variant for the full-resolution layers (cin <= 128, narrow N).
//
// With N = 64 one 64-wide K block is only 128 MMA cycles, but the generic kernel fetches a fresh
// 16 KB A tile per tap for it: 9 x 16 KB per 128 pixels per channel block, more than the L2 -> shared
// memory path delivers (ncu: conv1_2 at 28 % tensor pipe).  Here
//   * the CTA's weights (all 9 taps x KB channel blocks of its N block) are loaded ONCE and stay in
//     shared memory for the whole persistent kernel;
//   * per 16 x 8 pixel tile and channel block ONE TMA box {64 ch, 16 px, 18 rows} = 36 KB (the tile plus
//     its halo, rows 16 px = 2 KB apart) lands in shared memory, and the nine taps are nine UMMA
//     descriptors into that patch: start = patch + (r*16 + s) * 128 B, stride between 8-pixel row groups
//     (SBO) = 2048 B (base_offset stays 0: the swizzle follows absolute address bits, see smem_desc_patch).
// A traffic drops 4x (36 KB instead of 144 KB); the layer becomes MMA / HBM bound.
// Patch = the 16 x 8 pixel tile plus its one-pixel halo: 18 rows x (8 + 2) pixels x 128 B.  The rows of a patch are
// 10 pixels = 1280 B apart (SBO of the UMMA descriptors); round 1 fetched 16 pixels per row (2048 B, a power of two)
// of which 6 were never read -- 37 % of the L2 -> shared-memory fill and of the shared-memory write bandwidth the
// N = 64 layers are bound by.
#ifndef ST2_PATCH_W
#define ST2_PATCH_W 10
#endif
constexpr int kPatchW = ST2_PATCH_W, kPatchH = 18;
constexpr int kPatchTx = kPatchW * kPatchH * 128;             // bytes one TMA box delivers (23 040)
constexpr int kPatchBytes = (kPatchTx + 1023) / 1024 * 1024;  // stage stride: SWIZZLE_128B tiles start 1024-byte aligned
constexpr int kWsTW = 8, kWsTH = 16;

template <int BN, int KB> struct WsCfg {
  static constexpr int kWTile = BN * 128;                      // one tap, one channel block
  static constexpr int kWBytes = KB * 9 * kWTile;
  // Output through shared memory + TMA tile stores where it fits (KB = 1): per-thread 16-byte stores at a
  // 128-byte stride cap the epilogue at ~1.8 TB/s (measured 76 us for conv1_2's 134 MB with everything else
  // switched off), whole-line bulk stores do not.
  static constexpr bool kTmaStore = (KB == 1 && BN == 64);
  static constexpr int kStages = kTmaStore ? 4 : (kWBytes <= 73728 ? 4 : (kWBytes + 3 * kPatchBytes + 1280 <= 232448 ? 3 : 2));
  // three 16 KB tiles: in the masked (data-gradient) epilogue a tile is first the landing zone of the ReLU-mask
  // source (a TMA load issued one tile ahead), then -- in place, thread = pixel row -- the output staging tile of
  // the TMA store; the forward epilogue uses two of them for staging only
  static constexpr int kOutBytes = kTmaStore ? 3 * BM * 128 : 0;
  static constexpr int kAccStride = BN < 32 ? 32 : BN;                 // TMEM columns per accumulator
  static constexpr int kTmemCols = 2 * kAccStride;
  static constexpr int kSmemBytes = kWBytes + kStages * kPatchBytes + kOutBytes + 1024 + 256;
};

// Start addresses are 128 B (one pixel) granular, not 1024 B aligned.  Measured on B200: the swizzle XOR
// is taken from the absolute shared-memory address bits (the same bits the TMA unit used when it wrote
// the patch), so the descriptor's base_offset field must stay 0 -- setting it to (addr >> 7) & 7 reads
// the wrong 16-byte chunks for every tap with a column shift.
__device__ __forceinline__ uint64_t smem_desc_patch(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((kPatchW * 128) >> 4) << 32;                 // SBO: next 8-pixel group = next patch row
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int BN, int KB, bool HALO>
__global__ void __launch_bounds__(kNumThreads, 1)
tc_conv_ws_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_a2,
                  const __grid_constant__ CUtensorMap tmap_m, const ConvGeom g, const float* __restrict__ bias,
                  const __half* __restrict__ act, __half* __restrict__ out, const int epi, const TcInject inj) {
  using C = WsCfg<BN, KB>;
  // <16, 2>: conv1_1 data gradient from TWO 64-channel tensors (the gradient through tmap_a, conv1_1's activations
  // through tmap_a2) with their own weight halves and their own accumulators (TMEM columns 0..15 / 16..31 of the
  // slot): gx = acc1 + coef[1] * acc2 -- the style gradient of conv1_1 never exists as a tensor (st2_net.cu).
  constexpr bool DUAL = (BN == 16 && KB == 2);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_p = smem + C::kWBytes;
  uint8_t* smem_o = smem + C::kWBytes + C::kStages * kPatchBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kWBytes + C::kStages * kPatchBytes + C::kOutBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = bars + 2 * C::kStages + 2;
  uint64_t* w_full = bars + 2 * C::kStages + 4;
  uint64_t* m_full = bars + 2 * C::kStages + 5;      // [3] mask-source tiles (TMA-store variant, masked epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a CTA keeps one N block for its whole life (its weights are resident); pixel tiles are strided
  const int nb = blockIdx.x % g.n_blocks;
  const int pt0 = blockIdx.x / g.n_blocks, pt_step = gridDim.x / g.n_blocks;
  const int n_pt = g.tiles_h * g.tiles_w;
  pdl_trigger();
  if (HALO) halo_push_prologue(g.halo);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmap_a); tc::prefetch_tmap(&tmap_b); tc::prefetch_tmap(&tmap_o);
    if (DUAL) tc::prefetch_tmap(&tmap_a2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full[a], 1); tc::mbar_init(&tmem_empty[a], 4); }
    tc::mbar_init(w_full, 1);
    for (int a = 0; a < 3; ++a) tc::mbar_init(&m_full[a], 1);
    tc::fence_mbar_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, C::kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  pdl_wait();                      // everything above ran in the predecessor's shadow; global memory from here on
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (tc::elect_one()) {
      tc::mbar_expect_tx(w_full, C::kWBytes);
      for (int kb = 0; kb < KB; ++kb)
        for (int tap = 0; tap < 9; ++tap)
          tc::tma_load_2d(smem_w + (kb * 9 + tap) * C::kWTile, &tmap_b, w_full, tap * g.cin + kb * BK, nb * BN);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    bool waited = false;
    for (int pt = pt0; pt < n_pt; pt += pt_step) {
      const int ths = pt / g.tiles_w, tw = pt - ths * g.tiles_w;
      waited = halo_ready<HALO>(g, waited, ths);
      const int th = rot_row<HALO>(g, ths);
      const int h0 = th * kWsTH, w0 = tw * kWsTW;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        tc::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (tc::elect_one()) {
          if (g.dbg & 4) {
            tc::mbar_arrive(&full_bar[stage]);
          } else {
            tc::mbar_expect_tx(&full_bar[stage], kPatchTx);
            tc::tma_load_3d(smem_p + stage * kPatchBytes, (DUAL && kb == 1) ? &tmap_a2 : &tmap_a, &full_bar[stage],
                            DUAL ? 0 : kb * BK, w0 - 1, h0 - 1 + g.hoff);
          }
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc = tc::idesc_f16(BM, BN, 0, 0);
    const uint32_t p_addr0 = tc::smem_u32(smem_p);
    const uint64_t b_desc0 = tc::smem_desc_k_sw128(tc::smem_u32(smem_w));
    tc::mbar_wait(w_full, 0);
    tc::fence_after_sync();
    int stage = 0; uint32_t phase = 0;
    int local = 0;
    for (int pt = pt0; pt < n_pt; pt += pt_step, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      tc::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * C::kAccStride;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        tc::mbar_wait(&full_bar[stage], phase);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t p_addr = p_addr0 + stage * kPatchBytes;
#pragma unroll
          for (int tap = 0; tap < ((g.dbg & 2) ? 0 : 9); ++tap) {
            const uint64_t a_desc = smem_desc_patch(p_addr + ((tap / 3) * kPatchW + (tap % 3)) * 128);
            const uint64_t b_desc = b_desc0 + (uint64_t)(((kb * 9 + tap) * C::kWTile) >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc::umma_f16(d_tmem + (DUAL ? kb * 16 : 0), a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc,
                           DUAL ? (tap | k) != 0 : (kb | tap | k) != 0);
          }
          tc::umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(&tmem_full[acc]);
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // ================================ epilogue ====================================
    const int ew = warp - kEpiWarp0;
    const int row = ew * 32 + lane;
    const int row_h = row / kWsTW, row_w = row % kWsTW;
    float ss = 0.f;
    float cc = 0.f, sc = 0.f, dc = 0.f;
    if (inj.coef != nullptr) { cc = (float)inj.coef[0]; sc = (float)inj.coef[1]; dc = (float)inj.coef[2]; }
    constexpr int NCH = BN >= 32 ? BN / 32 : 1;
    const bool masked = (epi == EPI_MASK);
    const bool have_inj = masked && inj.coef != nullptr;
    const bool have_s = have_inj && inj.sraw != nullptr;
    int local = 0;
    // TMA-store variant, masked epilogue: the tile of the ReLU-mask source (the 16 x 8 pixels x 64 channels this
    // tile's outputs belong to) is fetched by TMA into staging buffer (tile % 3) ONE TILE AHEAD, by the same thread
    // that issues the tile stores -- it knows when the store that last used the buffer has finished reading it.
    const bool mtma = C::kTmaStore && BN == 64 && masked && !(g.dbg & 1);
    auto load_mask_tile = [&](const int ptq, const int buf) {
      const int thq = rot_row<HALO>(g, ptq / g.tiles_w), twq = ptq - (ptq / g.tiles_w) * g.tiles_w;
      tc::mbar_expect_tx(&m_full[buf], BM * 128);
      tc::tma_load_3d(smem_o + buf * (BM * 128), &tmap_m, &m_full[buf], nb * BN, twq * kWsTW, thq * kWsTH);
    };
    if (mtma && warp == kEpiWarp0 && lane == 0 && pt0 < n_pt) load_mask_tile(pt0, 0);
    for (int pt = pt0; pt < n_pt; pt += pt_step, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int ths = pt / g.tiles_w, tw = pt - ths * g.tiles_w;
      const int th = rot_row<HALO>(g, ths);
      const int h = th * kWsTH + row_h, w = tw * kWsTW + row_w;
      const bool valid = (h < g.H) && (w < g.W) && !(g.dbg & 1);
      const long long obase = ((long long)h * g.W + w) * g.cout + (long long)nb * BN;
      if (BN == 16) {
        // conv1_1 data gradient: 64 gradient channels -> the 3 image planes (N padded to 16), fp32 NCHW out
        if (lane == 0) tc::mbar_wait(&tmem_full[acc], acc_phase);
        __syncwarp();
        tc::fence_after_sync();
        uint32_t r[32];
        tc::tmem_ld_32x32(tmem_base + acc * C::kAccStride + ((uint32_t)(ew * 32) << 16), r);
        tc::tmem_ld_wait();
        if (valid) {
          float* gx = reinterpret_cast<float*>(out);
          const long long plane = (long long)g.H * g.W, p = (long long)h * g.W + w;
          if (DUAL) {
            gx[p] = fmaf(sc, __uint_as_float(r[16]), __uint_as_float(r[0]));
            gx[plane + p] = fmaf(sc, __uint_as_float(r[17]), __uint_as_float(r[1]));
            gx[2 * plane + p] = fmaf(sc, __uint_as_float(r[18]), __uint_as_float(r[2]));
          } else {
            gx[p] = __uint_as_float(r[0]);
            gx[plane + p] = __uint_as_float(r[1]);
            gx[2 * plane + p] = __uint_as_float(r[2]);
          }
        }
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
        continue;
      }
      uint4 pa[NCH][4], ps[NCH][4];
      if (masked && (!mtma || have_s)) {
        // pull the NEXT tile's epilogue operands towards L2 now: per tile the loads below are issued and then
        // immediately needed, so their latency (not their bandwidth) is what the epilogue pays (a TMA-fetched mask
        // tile is a tile ahead anyway)
        const int ptn = pt + pt_step;
        if (ptn < n_pt) {
          const int thns = ptn / g.tiles_w, twn = ptn - thns * g.tiles_w;
          const int thn = rot_row<HALO>(g, thns);
          const int hn = thn * kWsTH + row_h, wn = twn * kWsTW + row_w;
          if (hn < g.H && wn < g.W) {
            const long long on = ((long long)hn * g.W + wn) * g.cout + (long long)nb * BN;
            if (!mtma) asm volatile("prefetch.global.L2 [%0];" ::"l"(act + on));
            if (have_s) asm volatile("prefetch.global.L2 [%0];" ::"l"(inj.sraw + on));
          }
        }
      }
      if (C::kTmaStore) {
        // Coalesced fetch: the warp's 32 pixels are 4 tile rows of 8 pixels = 4 x 1 KB contiguous in global
        // memory (cout = 64).  Load j takes 512 contiguous bytes; the pieces are re-sorted to "thread = pixel"
        // through the warp's own 4 KB slice of the staging tile further down.
        if (masked) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int hh = th * kWsTH + ew * 4 + (j >> 1), ww = tw * kWsTW + (j & 1) * 4 + (lane >> 3);
            const long long o = ((long long)hh * g.W + ww) * g.cout + (lane & 7) * 8;
            const bool ok = hh < g.H && ww < g.W && !(g.dbg & 1);
            if (!mtma) pa[j >> 2][j & 3] = ok ? __ldg(reinterpret_cast<const uint4*>(act + o)) : make_uint4(0, 0, 0, 0);
            if (have_s) ps[j >> 2][j & 3] = ok ? __ldg(reinterpret_cast<const uint4*>(inj.sraw + o)) : make_uint4(0, 0, 0, 0);
          }
        }
      } else {
        if (valid && masked) {
          const uint4* ap = reinterpret_cast<const uint4*>(act + obase);
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) pa[c][q] = __ldg(ap + c * 4 + q);
        }
        if (valid && have_s) {
          const uint4* sp = reinterpret_cast<const uint4*>(inj.sraw + obase);
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int q = 0; q < 4; ++q) ps[c][q] = __ldg(sp + c * 4 + q);
        }
      }
      if (lane == 0) tc::mbar_wait(&tmem_full[acc], acc_phase);
      __syncwarp();
      tc::fence_after_sync();
      const uint32_t t_row = tmem_base + acc * C::kAccStride + ((uint32_t)(ew * 32) << 16);
      if (C::kTmaStore) {
        // staging tile: the bulk store issued from it two (forward: buffers 0 / 1) or three (masked: 0 / 1 / 2) tiles ago
        // must have finished reading it
        const int buf = mtma ? local % 3 : acc;
        uint8_t* stg = smem_o + buf * (BM * 128);
        if (warp == kEpiWarp0 && lane == 0) {
          tc::bulk_wait_read<1>();
          // ... which also frees the buffer of the tile after this one: fetch its mask source now
          if (mtma && pt + pt_step < n_pt) load_mask_tile(pt + pt_step, (local + 1) % 3);
        }
        if (mtma) {
          if (lane == 0) tc::mbar_wait(&m_full[buf], (local / 3) & 1);
          __syncwarp();
        } else {
          tc::named_bar_sync(1, 128);
        }
        __half* srow = reinterpret_cast<__half*>(stg + row * 128);
        const bool do_pool = (epi == EPI_BIAS_RELU) && inj.pool != nullptr;
        __half* const pool_px = do_pool ? inj.pool + ((long long)(h >> 1) * inj.pool_wp + (w >> 1)) * g.cout + (long long)nb * BN
                                        : nullptr;
        if (masked) {
          // re-sort the coalesced pieces: piece j of lane l belongs to staging row (ew*4 + j/2)*8 + (j%2)*4 + l/8,
          // chunk l%8; afterwards every thread reads the eight chunks of its own row
          uint4* sl = reinterpret_cast<uint4*>(stg);
          if (mtma) {
            // the mask source of this thread's pixel row, straight from the TMA-written (swizzled) tile
#pragma unroll
            for (int j = 0; j < 8; ++j) pa[j >> 2][j & 3] = sl[row * 8 + (j ^ (row & 7))];
            if (have_s) tc::named_bar_sync(1, 128);            // the tile doubles as scratch for the re-sort below
          }
#pragma unroll
          for (int pass = mtma ? 1 : 0; pass < 2; ++pass) {
            if (pass == 1 && !have_s) break;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int rr = (ew * 4 + (j >> 1)) * 8 + (j & 1) * 4 + (lane >> 3);
              sl[rr * 8 + ((lane & 7) ^ (rr & 7))] = pass ? ps[j >> 2][j & 3] : pa[j >> 2][j & 3];
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 t = sl[row * 8 + (j ^ (row & 7))];
              if (pass) ps[j >> 2][j & 3] = t; else pa[j >> 2][j & 3] = t;
            }
            __syncwarp();
          }
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t r[32];
          tc::tmem_ld_32x32(t_row + c * 32, r);
          tc::tmem_ld_wait();
          epi_chunk(r, epi, bias + nb * BN + c * 32, pa[c], ps[c],
                    (valid && have_inj && inj.fc != nullptr) ? inj.fc + obase + c * 32 : nullptr, have_inj, have_s, cc, sc,
                    dc, 1.f, false, ss, srow, 64 | ((c * 4) << 3) | (row & 7), valid, g.cout,
                    PoolOut{do_pool ? pool_px + c * 32 : nullptr, do_pool ? kWsTW : 0,
                            do_pool && !(lane & 1) && !(lane & kWsTW)});
        }
        tc::fence_before_sync();
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
        tc::named_bar_sync(1, 128);
        if (warp == kEpiWarp0 && lane == 0 && !(g.dbg & 1)) {
          tc::tma_store_3d(&tmap_o, stg, nb * BN, tw * kWsTW, th * kWsTH);
          tc::bulk_commit();
        }
        continue;
      }
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(t_row + c * 32, r);
        tc::tmem_ld_wait();
        epi_chunk(r, epi, bias + nb * BN + c * 32, pa[c], ps[c],
                  (valid && have_inj && inj.fc != nullptr) ? inj.fc + obase + c * 32 : nullptr, have_inj, have_s, cc, sc, dc,
                  1.f, false, ss, out + obase + c * 32, 0, valid, g.cout);
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
    }
    if (C::kTmaStore && warp == kEpiWarp0 && lane == 0) tc::bulk_wait<0>();
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}


// ================================================================================================
// conv1_1 data gradient in "1x1 + stencil" form.  The convolution maps 64 channels to only 3 image planes, so
//     gx[q, plane] = sum_tap sum_ch g[q + off(tap), ch] W[plane][tap][ch] = sum_tap T[q + off(tap)][tap * 3 + plane]
// with T = g W_all a POINTWISE contraction 64 -> 27 (9 taps x 3 planes, N padded to 32): per 16 x 8 pixel tile the
// 18 x 10 pixel patch (tile + halo) is multiplied ONCE by the 27 x 64 weight matrix -- two M = 128 chunks of patch
// pixels x 4 K steps = 8 MMAs -- instead of nine shifted 128 x 16 x 64 contractions = 36 MMAs that each re-read their
// A operand from shared memory (the N = 16 kernel was bound by exactly that: 63 us for 134 MB).  The 27 values of
// every patch pixel go TMEM -> registers -> shared memory, and each output pixel sums its 9 x 3 neighbours there.
// DUAL: a second source (conv1_1's activations with the folded style weights W' = W D', st2_net.cu style_fold_kernel)
// gets its own accumulators; the two are combined with the device coefficient before the stencil.
constexpr int kStPx = kPatchW * kPatchH;                     // 180 patch pixels
constexpr int kStN = 32;                                     // 27 outputs per patch pixel, padded
template <bool DUAL> struct StCfg {
  static constexpr int kSrc = DUAL ? 2 : 1;
  static constexpr int kWBytes = kSrc * kStN * 128;          // [source][32 rows][64 ch] fp16
  static constexpr int kStages = 6;
  static constexpr int kStageBytes = 2 * BM * 128;           // two M = 128 chunks of 128-byte rows (patch + slack)
  static constexpr int kUBytes = kStPx * 27 * 4;             // fp32 [180][27]
  static constexpr int kAccStride = kSrc * 2 * kStN;         // TMEM columns per accumulator set
  static constexpr int kTmemCols = 2 * kAccStride < 32 ? 32 : 2 * kAccStride;
  static constexpr int kSmemBytes = ((kWBytes + 1023) / 1024) * 1024 + kStages * kStageBytes +
                                    ((kUBytes + 1023) / 1024) * 1024 + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(kSmemBytes <= 232448, "stencil kernel: shared memory");
};

template <bool DUAL, bool HALO>
__global__ void __launch_bounds__(kNumThreads, 1)
tc_conv_first_stencil_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                             const __grid_constant__ CUtensorMap tmap_b, const ConvGeom g, float* __restrict__ gx,
                             const double* __restrict__ coef) {
  using C = StCfg<DUAL>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;                                   // [source][32 rows][128 B], SWIZZLE_128B
  uint8_t* smem_p = smem + ((C::kWBytes + 1023) / 1024) * 1024;
  float* smem_u = reinterpret_cast<float*>(smem_p + C::kStages * C::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_u) + ((C::kUBytes + 1023) / 1024) * 1024);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tmem_full = bars + 2 * C::kStages;
  uint64_t* tmem_empty = bars + 2 * C::kStages + 2;
  uint64_t* w_full = bars + 2 * C::kStages + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pt = g.tiles_h * g.tiles_w;
  pdl_trigger();
  if (HALO) halo_push_prologue(g.halo);

  if (warp == 0 && lane == 0) { tc::prefetch_tmap(&tmap_a); tc::prefetch_tmap(&tmap_b); if (DUAL) tc::prefetch_tmap(&tmap_a2); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full[a], 1); tc::mbar_init(&tmem_empty[a], 4); }
    tc::mbar_init(w_full, 1);
    tc::fence_mbar_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, C::kTmemCols);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  pdl_wait();                      // everything above ran in the predecessor's shadow; global memory from here on
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (tc::elect_one()) {
      tc::mbar_expect_tx(w_full, C::kWBytes);
      for (int sidx = 0; sidx < C::kSrc; ++sidx)
        tc::tma_load_2d(smem_w + sidx * (kStN * 128), &tmap_b, w_full, 0, sidx * kStN);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    bool waited = false;
    for (int pt = blockIdx.x; pt < n_pt; pt += gridDim.x) {
      const int ths = pt / g.tiles_w, tw = pt - ths * g.tiles_w;
      waited = halo_ready<HALO>(g, waited, ths);
      const int th = rot_row<HALO>(g, ths);
      const int h0 = th * kWsTH, w0 = tw * kWsTW;
#pragma unroll
      for (int sidx = 0; sidx < C::kSrc; ++sidx) {
        tc::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (tc::elect_one()) {
          tc::mbar_expect_tx(&full_bar[stage], kPatchTx);
          tc::tma_load_3d(smem_p + stage * C::kStageBytes, sidx ? &tmap_a2 : &tmap_a, &full_bar[stage], 0, w0 - 1,
                          h0 - 1 + g.hoff);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    constexpr uint32_t idesc = tc::idesc_f16(BM, kStN, 0, 0);
    const uint64_t a_desc0 = tc::smem_desc_k_sw128(tc::smem_u32(smem_p));
    const uint64_t b_desc0 = tc::smem_desc_k_sw128(tc::smem_u32(smem_w));
    tc::mbar_wait(w_full, 0);
    tc::fence_after_sync();
    int stage = 0; uint32_t phase = 0;
    int local = 0;
    for (int pt = blockIdx.x; pt < n_pt; pt += gridDim.x, ++local) {
      const int acc = local & 1;
      tc::mbar_wait(&tmem_empty[acc], ((local >> 1) & 1) ^ 1);
      tc::fence_after_sync();
#pragma unroll
      for (int sidx = 0; sidx < C::kSrc; ++sidx) {
        tc::mbar_wait(&full_bar[stage], phase);
        tc::fence_after_sync();
        if (tc::elect_one()) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {                  // patch pixels 0..127 and 128..255 (180.. : slack, unused)
            const uint32_t d_tmem = tmem_base + acc * C::kAccStride + (sidx * 2 + c) * kStN;
            const uint64_t a_desc = a_desc0 + (uint64_t)((stage * C::kStageBytes + c * (BM * 128)) >> 4);
            const uint64_t b_desc = b_desc0 + (uint64_t)((sidx * (kStN * 128)) >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc::umma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, k != 0);
          }
          tc::umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(&tmem_full[acc]);
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // ================================ epilogue ====================================
    const int ew = warp - kEpiWarp0;
    const int row = ew * 32 + lane;                    // TMEM lane = patch pixel (chunk 0) / patch pixel - 128 (chunk 1)
    const int out_r = row / kWsTW, out_c = row % kWsTW; // this thread's output pixel inside the tile
    const float sc = (DUAL && coef != nullptr) ? (float)coef[1] : 0.f;
    const long long plane = (long long)g.H * g.W;
    int local = 0;
    for (int pt = blockIdx.x; pt < n_pt; pt += gridDim.x, ++local) {
      const int acc = local & 1;
      const int ths = pt / g.tiles_w, tw = pt - ths * g.tiles_w;
      const int th = rot_row<HALO>(g, ths);
      if (lane == 0) tc::mbar_wait(&tmem_full[acc], (local >> 1) & 1);
      __syncwarp();
      tc::fence_after_sync();
      const uint32_t t_row = tmem_base + acc * C::kAccStride + ((uint32_t)(ew * 32) << 16);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        uint32_t r2[DUAL ? 32 : 1];
        tc::tmem_ld_32x32(t_row + c * kStN, r);
        if (DUAL) tc::tmem_ld_32x32(t_row + (2 + c) * kStN, reinterpret_cast<uint32_t(&)[32]>(r2));
        tc::tmem_ld_wait();
        const int p = row + c * BM;
        if (p < kStPx) {
          float* u = smem_u + p * 27;
#pragma unroll
          for (int n = 0; n < 27; ++n)
            u[n] = DUAL ? fmaf(sc, __uint_as_float(r2[n]), __uint_as_float(r[n])) : __uint_as_float(r[n]);
        }
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
      tc::named_bar_sync(1, 128);                      // the 180 x 27 table of this tile is complete
      const int h = th * kWsTH + out_r, w = tw * kWsTW + out_c;
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float* u = smem_u + ((out_r + t / 3) * kPatchW + out_c + t % 3) * 27 + t * 3;
        o0 += u[0]; o1 += u[1]; o2 += u[2];
      }
      if (h < g.H && w < g.W && !(g.dbg & 1)) {
        const long long q = (long long)h * g.W + w;
        gx[q] = o0; gx[plane + q] = o1; gx[2 * plane + q] = o2;
      }
      tc::named_bar_sync(1, 128);                      // everybody has read the table: the next tile may overwrite it
    }
  }

  tc::fence_before_sync();
  __syncthreads();
  if (warp == 2) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ================================================================================================
// Weight-stationary halo reuse AND CTA pair, for the 128-channel layers at half resolution (conv2_1 forward,
// conv2_2 both directions): N = 128 split over the two CTAs of a pair (64 weight rows each, all 9 taps x KB
// channel blocks resident: 72 / 144 KB), M = 256 = two 16 x 8 pixel tiles stacked vertically, each CTA
// loading ONE 36 KB patch per tile and channel block.  The generic pair kernel moves 24 KB per CTA per
// 256 MMA cycles from L2 for these layers (94 B/clk/SM -- the fill path, not the tensor pipe, was the limit:
// ncu 38-52 % tensor active); here it is 36 KB per 2304 cycles.
template <int KB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
tc_conv_wsp_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const ConvGeom g, const float* __restrict__ bias, const __half* __restrict__ act,
                   __half* __restrict__ out, const int epi, const TcInject inj) {
  constexpr int BN = 128;
  constexpr int kWTile = (BN / 2) * 128;                       // this CTA's 64 rows of one tap, one channel block
  constexpr int kWBytes = KB * 9 * kWTile;
  constexpr int kStages = (KB == 1) ? 4 : 2;
  constexpr int kTmemCols = 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_p = smem + kWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWBytes + kStages * kPatchBytes);
  uint64_t* full_bar = bars;                         // leader's copy collects both CTAs' bytes
  uint64_t* empty_bar = bars + kStages;              // local (multicast commit)
  uint64_t* tmem_full = bars + 2 * kStages;          // local (multicast commit)
  uint64_t* tmem_empty = bars + 2 * kStages + 2;     // leader's copy: 4 local + 4 remote epilogue warps
  uint64_t* w_full = bars + 2 * kStages + 4;         // leader's copy
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int n_tiles = g.tiles_h * g.tiles_w * g.n_blocks;      // tiles_h counts 32-row pair tiles
  // a pair keeps one N block for its whole life (its weights are resident)
  const int nb = pair % g.n_blocks;
  const int pt0 = pair / g.n_blocks, pt_step = n_pairs / g.n_blocks;
  const int n_pt = g.tiles_h * g.tiles_w;
  (void)n_tiles;

  if (warp == 0 && lane == 0) { tc::prefetch_tmap(&tmap_a); tc::prefetch_tmap(&tmap_b); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { tc::mbar_init(&tmem_full[a], 1); tc::mbar_init(&tmem_empty[a], 8); }
    tc::mbar_init(w_full, 1);
    tc::fence_mbar_init();
  }
  if (warp == 2) tc::tmem_alloc_2sm(tmem_slot, kTmemCols);
  tc::fence_before_sync();
  tc::cluster_sync_all();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (rank == 0 && tc::elect_one()) tc::mbar_expect_tx(w_full, 2u * kWBytes);
    if (tc::elect_one()) {
      for (int kb = 0; kb < KB; ++kb)
        for (int tap = 0; tap < 9; ++tap)
          tc::tma_load_2d_2sm(smem_w + (kb * 9 + tap) * kWTile, &tmap_b, w_full, tap * g.cin + kb * BK,
                              nb * BN + (int)rank * (BN / 2));
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int pt = pt0; pt < n_pt; pt += pt_step) {
      const int th = pt / g.tiles_w, tw = pt - th * g.tiles_w;
      const int h0 = th * (2 * kWsTH) + (int)rank * kWsTH, w0 = tw * kWsTW;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        tc::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (rank == 0 && tc::elect_one()) tc::mbar_expect_tx(&full_bar[stage], 2u * kPatchTx);
        if (tc::elect_one())
          tc::tma_load_3d_2sm(smem_p + stage * kPatchBytes, &tmap_a, &full_bar[stage], kb * BK, w0 - 1, h0 - 1 + g.hoff);
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================================ MMA issuer (leader only) ================================
    constexpr uint32_t idesc = tc::idesc_f16(2 * BM, BN, 0, 0);
    const uint32_t p_addr0 = tc::smem_u32(smem_p);
    const uint64_t b_desc0 = tc::smem_desc_k_sw128(tc::smem_u32(smem_w));
    tc::mbar_wait(w_full, 0);
    tc::fence_after_sync();
    int stage = 0; uint32_t phase = 0;
    int local = 0;
    for (int pt = pt0; pt < n_pt; pt += pt_step, ++local) {
      const int acc = local & 1;
      tc::mbar_wait(&tmem_empty[acc], ((local >> 1) & 1) ^ 1);
      tc::fence_after_sync();
      const uint32_t d_tmem = tmem_base + acc * BN;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        tc::mbar_wait(&full_bar[stage], phase);
        tc::fence_after_sync();
        if (tc::elect_one()) {
          const uint32_t p_addr = p_addr0 + stage * kPatchBytes;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint64_t a_desc = smem_desc_patch(p_addr + ((tap / 3) * kPatchW + (tap % 3)) * 128);
            const uint64_t b_desc = b_desc0 + (uint64_t)(((kb * 9 + tap) * kWTile) >> 4);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc::umma_f16_2sm(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (kb | tap | k) != 0);
          }
          tc::umma_commit_2sm(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit_2sm(&tmem_full[acc]);
      __syncwarp();
    }
  } else if (warp >= kEpiWarp0) {
    // ================================ epilogue (both CTAs, own 16 x 8 pixels) =================
    const int ew = warp - kEpiWarp0;
    const int row = ew * 32 + lane;
    const int row_h = row / kWsTW, row_w = row % kWsTW;
    float ss = 0.f;
    float cc = 0.f, sc = 0.f, dc = 0.f;
    if (inj.coef != nullptr) { cc = (float)inj.coef[0]; sc = (float)inj.coef[1]; dc = (float)inj.coef[2]; }
    constexpr int NCH = BN / 32;
    const bool masked = (epi == EPI_MASK);
    const bool have_inj = masked && inj.coef != nullptr;
    const bool have_s = have_inj && inj.sraw != nullptr;
    int local = 0;
    for (int pt = pt0; pt < n_pt; pt += pt_step, ++local) {
      const int acc = local & 1;
      const int th = pt / g.tiles_w, tw = pt - th * g.tiles_w;
      const int h = th * (2 * kWsTH) + (int)rank * kWsTH + row_h, w = tw * kWsTW + row_w;
      const bool valid = (h < g.H) && (w < g.W) && !(g.dbg & 1);
      const long long obase = ((long long)h * g.W + w) * g.cout + (long long)nb * BN;
      if (masked) {
        // next tile's mask / injection operands towards L2 (a tile is only ~2.4 us of MMA here, and the loads
        // below are issued and then immediately needed)
        const int ptn = pt + pt_step;
        if (ptn < n_pt) {
          const int thn = ptn / g.tiles_w, twn = ptn - thn * g.tiles_w;
          const int hn = thn * (2 * kWsTH) + (int)rank * kWsTH + row_h, wn = twn * kWsTW + row_w;
          if (hn < g.H && wn < g.W) {
            const __half* an = act + ((long long)hn * g.W + wn) * g.cout + (long long)nb * BN;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(an));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(an + 64));
            if (have_s) {
              const __half* sn = inj.sraw + ((long long)hn * g.W + wn) * g.cout + (long long)nb * BN;
              asm volatile("prefetch.global.L2 [%0];" ::"l"(sn));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(sn + 64));
            }
          }
        }
      }
      uint4 pa[NCH][4], ps[NCH][4];
      if (valid && masked) {
        const uint4* ap = reinterpret_cast<const uint4*>(act + obase);
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int q = 0; q < 4; ++q) pa[c][q] = __ldg(ap + c * 4 + q);
      }
      if (valid && have_s) {
        const uint4* sp = reinterpret_cast<const uint4*>(inj.sraw + obase);
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
          for (int q = 0; q < 4; ++q) ps[c][q] = __ldg(sp + c * 4 + q);
      }
      if (lane == 0) tc::mbar_wait(&tmem_full[acc], (local >> 1) & 1);
      __syncwarp();
      tc::fence_after_sync();
      const uint32_t t_row = tmem_base + acc * BN + ((uint32_t)(ew * 32) << 16);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32(t_row + c * 32, r);
        tc::tmem_ld_wait();
        epi_chunk(r, epi, bias + nb * BN + c * 32, pa[c], ps[c],
                  (valid && have_inj && inj.fc != nullptr) ? inj.fc + obase + c * 32 : nullptr, have_inj, have_s, cc, sc, dc,
                  1.f, false, ss, out + obase + c * 32, 0, valid, g.cout);
      }
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_leader(&tmem_empty[acc]);
    }
  }

  tc::fence_before_sync();
  tc::cluster_sync_all();
  if (warp == 2) {
    tc::fence_after_sync();
    tc::tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

}  // namespace

namespace {
template <int BN> struct PairCfg {
  static constexpr int KB = (BN == 256) ? 1 : 2;
  static constexpr int kStages = (BN == 256) ? 6 : 4;
  static constexpr int kSmemBytes = kStages * KB * (BM * BK * 2 + (BN / 2) * BK * 2) + 1024 + 256;
};
template <int KB> struct WspCfg {
  static constexpr int kStages = (KB == 1) ? 4 : 2;
  static constexpr int kSmemBytes = KB * 9 * 64 * 128 + kStages * kPatchBytes + 1024 + 256;
};
}  // namespace

struct TcConvPlan {
  CUtensorMap tmap_a, tmap_b;
  CUtensorMap tmap_a2;          // dual-source plans only (tc_conv_dual_plan_create)
  bool dual = false;
  bool stencil = false;         // conv1_1 data gradient in 1x1 + stencil form (tc_conv_stencil_plan_create)
  CUtensorMap tmap_f, tmap_d;   // style fusion (tc_conv_set_style_fuse): activations of the blob below, scaled Gram difference
  bool sfuse = false;
  ConvGeom g;
  int bn;
  int ws_kb;        // > 0: weight-stationary halo-reuse kernel with this many 64-channel K blocks
  bool pair;        // CTA-pair kernel (cta_group::2): tiles_h counts 16-row pair tiles, tmap_b box is BN / 2 rows
  CUtensorMap tmap_o;           // output tile stores of the weight-stationary kernel, encoded on first use
  const void* tmap_o_base = nullptr;
  CUtensorMap tmap_m;           // mask-source tile loads of its masked epilogue, encoded on first use
  const void* tmap_m_base = nullptr;
};

static int get_encoder(st2_ctx* ctx, EncodeTiledFn* fn) {
  if (!ctx->tmap_encode) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    ST2_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (!p || qres != cudaDriverEntryPointSuccess)
      return st2_fail(ctx, ST2_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    ctx->tmap_encode = p;
  }
  *fn = reinterpret_cast<EncodeTiledFn>(ctx->tmap_encode);
  return 0;
}

int st2_encode_tmap(st2_ctx* ctx, CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                    const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  EncodeTiledFn enc;
  int rc = get_encoder(ctx, &enc);
  if (rc) return rc;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                   strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return st2_fail(ctx, ST2_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// same with a choice of shared-memory layout: swizzle128 = 0 -> no swizzle (dense boxes)
int st2_encode_tmap_ex(st2_ctx* ctx, CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                       const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle128) {
  EncodeTiledFn enc;
  int rc = get_encoder(ctx, &enc);
  if (rc) return rc;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                   strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return st2_fail(ctx, ST2_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

int tc_conv_plan_create(st2_ctx* ctx, const __half* in, const __half* w_packed, int H, int W, int cin, int cout,
                        int taps, TcConvPlan** out, int halo) {
  const bool first_bwd = (cin == 64 && cout == 16 && taps == 9);     // conv1_1 data gradient, 3 planes padded to 16
  if (first_bwd && (W < 16 || H < 16)) return st2_fail(ctx, ST2_ERR_ARG, "tc_conv: canvas too small for the conv1_1 gradient kernel");
  if (!first_bwd && (cin % 64 || cout % 64 || (taps != 9 && taps != 1)))
    return st2_fail(ctx, ST2_ERR_ARG, "tc_conv: cin/cout must be multiples of 64 (got %d/%d)", cin, cout);
  TcConvPlan* p = new TcConvPlan();
  ConvGeom& g = p->g;
  g.H = H; g.W = W; g.cin = cin; g.cout = cout; g.taps = taps;
  g.hoff = halo;                 // `in` then points at the first halo row; H counts the output rows only
  g.rot = 0;
  memset(&g.halo, 0, sizeof(g.halo));
  g.TW = (W <= 4) ? 4 : (W <= 8 ? 8 : 16);
  g.TH = BM / g.TW;
  g.tiles_h = (H + g.TH - 1) / g.TH;
  g.tiles_w = (W + g.TW - 1) / g.TW;
  // Pick the N tile: time ~ waves * K blocks * max(MMA cycles, hand-shake cycles / K blocks per stage).
  // Narrower tiles cost more operand traffic per flop but quantise better over the 148 SMs.
  {
    const int cand[3] = {256, 128, 64};
    const int kb[3] = {1, 2, 3};
    double best = 1e30;
    p->bn = 64;
    const int force = ctx->knobs.tc_bn;
    for (int c = 0; c < 3; ++c) {
      const int bn = cand[c];
      if (cout % bn || first_bwd) continue;
      if (force == bn) { p->bn = bn; best = -1; break; }
      const long long tiles = (long long)g.tiles_h * g.tiles_w * (cout / bn);
      const long long waves = (tiles + ctx->sm_count - 1) / ctx->sm_count;
      const double eff = bn == 256 ? 1.0 : (bn == 128 ? 1.30 : 1.8);   // shared-memory operand traffic per flop
      const double per_kblock = 2.0 * bn * eff > 450.0 / kb[c] ? 2.0 * bn * eff : 450.0 / kb[c];
      const double epi = 300.0 + 1.5 * bn;                       // drain of one tile when it is not hidden
      const double cost = (double)waves * ((double)taps * (cin / BK) * per_kblock + epi);
      if (cost < best * 0.97) { best = cost; p->bn = bn; }       // prefer the wider tile on near-ties
    }
  }
  p->ws_kb = 0;
  // Only where N = 64: there the generic kernel starves on A-tile fill.  (Measured: with N = 128 the two
  // kernels tie -- conv2_1 fwd 54 vs 57 us, conv2_2 80 vs 78 us -- so those keep the generic path.)
  // conv2_1 forward (64 -> 128): measured 54.0 vs 57.4 us on the generic kernel once the patches were 10 pixels wide
  const bool ws128 = !ctx->knobs.no_ws128 && !ctx->knobs.no_ws && taps == 9 && cin == 64 && cout == 128 && W >= 16 && H >= 16;
  if (first_bwd || ws128 || (taps == 9 && cin <= 128 && cout == 64 && W >= 16 && H >= 16 && !ctx->knobs.no_ws)) {
    p->ws_kb = cin / 64;
    p->bn = first_bwd ? 16 : (ws128 ? 128 : 64);    // (64,1) (64,2) (16,1) [(128,1)]: weights + 2..4 patches fit in 227 KB
    g.TW = kWsTW; g.TH = kWsTH;
    g.tiles_h = (H + g.TH - 1) / g.TH;
    g.tiles_w = (W + g.TW - 1) / g.TW;
  }
  // ST2_FORCE_PAIR lets small test canvases reach the CTA-pair kernels (normally chosen by size)
  const bool force_pair = ctx->knobs.force_pair;
  // weight-stationary + CTA pair for the 128-channel layers (conv2_1 forward, conv2_2): N = 128, cin <= 128
  bool wsp = false;
  // OPT-IN (ST2_WSP=1): measured at 1024^2 after the epilogue stores were fixed, it wins only conv2_2 forward
  // (65.6 vs 71.3 us) and loses conv2_1 forward (55.8 vs 47.3) and conv2_2 data gradient (89.7 vs 82.4) to the
  // generic kernels -- with 2.4 us of MMA per tile the epilogue (one tile of look-ahead) is what bounds it.
  if (!p->ws_kb && taps == 9 && cout == 128 && cin <= 128 && W >= 16 && H >= 32 && ctx->knobs.wsp &&
      !ctx->knobs.no_pair) {
    const long long pair_tiles = (long long)((H + 31) / 32) * ((W + kWsTW - 1) / kWsTW);
    if (pair_tiles >= ctx->sm_count / 2 || force_pair) {
      wsp = true;
      p->ws_kb = cin / 64;
      p->bn = 128;
      g.TW = kWsTW; g.TH = kWsTH;
      g.tiles_h = (H + 31) / 32;
      g.tiles_w = (W + kWsTW - 1) / kWsTW;
    }
  }
  if (force_pair && !p->ws_kb && taps == 9 && cout >= 128) p->bn = (cout % 256 == 0) ? 256 : 128;
  g.n_blocks = cout / p->bn;
  // CTA pairs for the wide layers when there is at least one wave of 16 x 16 pixel pair tiles
  p->pair = false;
  // (measured: N = 128 with a short K loop -- conv2_1 forward, 9 K blocks -- is faster on the single-CTA kernel)
  if (!p->ws_kb && taps == 9 && p->bn >= 128 && g.TW == 16 && (p->bn == 256 || taps * (cin / BK) >= 18) &&
      !ctx->knobs.no_pair) {
    const long long pair_tiles = (long long)((H + 15) / 16) * g.tiles_w * g.n_blocks;
    const long long min_env = ctx->knobs.pair_min_tiles;          // tuning experiments
    // one full wave of pair tiles; N = 128 pays off from ~0.8 waves on (conv5_1 at 1024^2: 64 pair tiles on 74
    // pairs, 26.8 -> 22.7 us), N = 256 does not (31.6 vs 26.8 us there)
    const long long min_tiles = min_env >= 0 ? min_env : (p->bn == 128 ? (ctx->sm_count * 2) / 5 : ctx->sm_count / 2);
    if (pair_tiles >= min_tiles || force_pair) {
      p->pair = true;
      g.tiles_h = (H + 15) / 16;
    }
  }
  if (wsp) p->pair = true;
  g.total_tiles = g.tiles_h * g.tiles_w * g.n_blocks;
  g.cblocks = cin / BK;
  g.k_iters = taps * g.cblocks;
  {
    cuuint64_t dims[3] = {(cuuint64_t)cin, (cuuint64_t)W, (cuuint64_t)(H + 2 * halo)};
    cuuint64_t strides[2] = {(cuuint64_t)cin * 2, (cuuint64_t)W * cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)g.TW, (cuuint32_t)g.TH};
    if (p->ws_kb) { box[1] = kPatchW; box[2] = kPatchH; }
    int rc = st2_encode_tmap(ctx, &p->tmap_a, in, 3, dims, strides, box);
    if (rc) { delete p; return rc; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)taps * cin, (cuuint64_t)cout};
    cuuint64_t strides[1] = {(cuuint64_t)taps * cin * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(p->pair ? p->bn / 2 : p->bn)};
    int rc = st2_encode_tmap(ctx, &p->tmap_b, w_packed, 2, dims, strides, box);
    if (rc) { delete p; return rc; }
  }
  *out = p;
  return 0;
}

void tc_conv_plan_destroy(TcConvPlan* p) { delete p; }

template <int BN>
static int launch_bn(st2_ctx* ctx, TcConvPlan* p, const float* bias, const __half* act, __half* out, int epi,
                     float out_scale, double* sumsq, const TcInject& inj) {
  const int grid = p->g.total_tiles < ctx->sm_count ? p->g.total_tiles : ctx->sm_count;
  p->g.dbg = ctx->debug_flags;
  p->g.step_nb = grid % p->g.n_blocks;
  p->g.step_tw = (grid / p->g.n_blocks) % p->g.tiles_w;
  p->g.step_th = (grid / p->g.n_blocks) / p->g.tiles_w;
  if (p->g.rot)
    st2_launch_pdl(ctx, false, tc_conv_kernel<BN, true>, grid, kNumThreads, Cfg<BN>::kSmemBytes, p->tmap_a, p->tmap_b, p->g,
                   bias, act, out, epi, out_scale, sumsq, inj);
  else
    st2_launch_pdl(ctx, true, tc_conv_kernel<BN, false>, grid, kNumThreads, Cfg<BN>::kSmemBytes, p->tmap_a, p->tmap_b, p->g,
                   bias, act, out, epi, out_scale, sumsq, inj);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

template <int BN, int KB>
static int launch_ws(st2_ctx* ctx, TcConvPlan* p, const float* bias, const __half* act, __half* out, int epi,
                     const TcInject& inj) {
  if (WsCfg<BN, KB>::kTmaStore && p->tmap_o_base != (const void*)out) {
    cuuint64_t dims[3] = {(cuuint64_t)p->g.cout, (cuuint64_t)p->g.W, (cuuint64_t)p->g.H};
    cuuint64_t strides[2] = {(cuuint64_t)p->g.cout * 2, (cuuint64_t)p->g.W * p->g.cout * 2};
    cuuint32_t box[3] = {(cuuint32_t)BN, (cuuint32_t)kWsTW, (cuuint32_t)kWsTH};
    int rc = st2_encode_tmap(ctx, &p->tmap_o, out, 3, dims, strides, box);
    if (rc) return rc;
    p->tmap_o_base = out;
  }
  if (WsCfg<BN, KB>::kTmaStore && epi == EPI_MASK && p->tmap_m_base != (const void*)act) {
    cuuint64_t dims[3] = {(cuuint64_t)p->g.cout, (cuuint64_t)p->g.W, (cuuint64_t)p->g.H};
    cuuint64_t strides[2] = {(cuuint64_t)p->g.cout * 2, (cuuint64_t)p->g.W * p->g.cout * 2};
    cuuint32_t box[3] = {(cuuint32_t)BN, (cuuint32_t)kWsTW, (cuuint32_t)kWsTH};
    int rc = st2_encode_tmap(ctx, &p->tmap_m, act, 3, dims, strides, box);
    if (rc) return rc;
    p->tmap_m_base = act;
  }
  const CUtensorMap& tm = (WsCfg<BN, KB>::kTmaStore && epi == EPI_MASK) ? p->tmap_m : p->tmap_a;
  const int nbk = p->g.n_blocks;
  const int n_pt = p->g.tiles_h * p->g.tiles_w;
  int per_nb = ctx->sm_count / nbk;
  if (per_nb > n_pt) per_nb = n_pt;
  if (per_nb < 1) per_nb = 1;
  p->g.dbg = ctx->debug_flags;
  if (p->g.rot)
    st2_launch_pdl(ctx, false, tc_conv_ws_kernel<BN, KB, true>, per_nb * nbk, kNumThreads, WsCfg<BN, KB>::kSmemBytes,
                   p->tmap_a, p->tmap_b, p->tmap_o, p->dual ? p->tmap_a2 : p->tmap_a, tm, p->g, bias, act, out, epi, inj);
  else
    st2_launch_pdl(ctx, true, tc_conv_ws_kernel<BN, KB, false>, per_nb * nbk, kNumThreads, WsCfg<BN, KB>::kSmemBytes,
                   p->tmap_a, p->tmap_b, p->tmap_o, p->dual ? p->tmap_a2 : p->tmap_a, tm, p->g, bias, act, out, epi, inj);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

template <int BN, bool SFUSE = false>
static int launch_pair(st2_ctx* ctx, TcConvPlan* p, const float* bias, const __half* act, __half* out, int epi,
                       const TcInject& inj) {
  constexpr int smem = PairCfg<BN>::kSmemBytes;
  int pairs = ctx->sm_count / 2;
  if (pairs > p->g.total_tiles) pairs = p->g.total_tiles;
  p->g.dbg = ctx->debug_flags;
  if (SFUSE) {
    if (p->g.rot)
      st2_launch_pdl(ctx, false, tc_conv2_kernel<128, true, true>, 2 * pairs, kNumThreads, smem, p->tmap_a, p->tmap_b,
                     p->tmap_f, p->tmap_d, p->g, bias, act, out, epi, inj);
    else
      st2_launch_pdl(ctx, true, tc_conv2_kernel<128, false, true>, 2 * pairs, kNumThreads, smem, p->tmap_a, p->tmap_b,
                     p->tmap_f, p->tmap_d, p->g, bias, act, out, epi, inj);
  } else if (p->g.rot) {
    st2_launch_pdl(ctx, false, tc_conv2_kernel<BN, true, false>, 2 * pairs, kNumThreads, smem, p->tmap_a, p->tmap_b,
                   p->tmap_a, p->tmap_b, p->g, bias, act, out, epi, inj);
  } else {
    st2_launch_pdl(ctx, true, tc_conv2_kernel<BN, false, false>, 2 * pairs, kNumThreads, smem, p->tmap_a, p->tmap_b,
                   p->tmap_a, p->tmap_b, p->g, bias, act, out, epi, inj);
  }
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

template <int KB>
static int launch_wsp(st2_ctx* ctx, TcConvPlan* p, const float* bias, const __half* act, __half* out, int epi,
                      const TcInject& inj) {
  constexpr int smem = WspCfg<KB>::kSmemBytes;
  const int nbk = p->g.n_blocks;
  const int n_pt = p->g.tiles_h * p->g.tiles_w;
  int per_nb = (ctx->sm_count / 2) / nbk;
  if (per_nb > n_pt) per_nb = n_pt;
  if (per_nb < 1) per_nb = 1;
  p->g.dbg = ctx->debug_flags;
  tc_conv_wsp_kernel<KB><<<2 * per_nb * nbk, kNumThreads, smem, ctx->stream>>>(p->tmap_a, p->tmap_b, p->g, bias, act, out,
                                                                               epi, inj);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

// ---- conv1_1 data gradient, stencil form (tc_conv_first_stencil_kernel) -----------------------------------------
int tc_conv_stencil_plan_create(st2_ctx* ctx, const __half* grad, const __half* act, const __half* w_all, int H, int W,
                                TcConvPlan** out, int halo) {
  if (W < 16 || H < 16) return st2_fail(ctx, ST2_ERR_ARG, "tc_conv: canvas too small for the conv1_1 gradient kernel");
  TcConvPlan* p = new TcConvPlan();
  ConvGeom& g = p->g;
  memset(&g, 0, sizeof(g));
  g.H = H; g.W = W; g.cin = 64; g.cout = 16; g.taps = 9;
  g.hoff = halo;
  g.TW = kWsTW; g.TH = kWsTH;
  g.tiles_h = (H + g.TH - 1) / g.TH;
  g.tiles_w = (W + g.TW - 1) / g.TW;
  p->bn = 16; p->ws_kb = 1; p->pair = false; p->dual = (act != nullptr); p->stencil = true;
  g.n_blocks = 1;
  g.total_tiles = g.tiles_h * g.tiles_w;
  g.cblocks = 1;
  g.k_iters = 9;
  cuuint64_t dims[3] = {64, (cuuint64_t)W, (cuuint64_t)(H + 2 * halo)};
  cuuint64_t strides[2] = {128, (cuuint64_t)W * 128};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)kPatchW, (cuuint32_t)kPatchH};
  int rc = st2_encode_tmap(ctx, &p->tmap_a, grad, 3, dims, strides, box);
  if (!rc && act) rc = st2_encode_tmap(ctx, &p->tmap_a2, act, 3, dims, strides, box);
  if (!rc) {
    cuuint64_t wd[2] = {64, (cuuint64_t)(act ? 2 : 1) * kStN};
    cuuint64_t ws[1] = {128};
    cuuint32_t wb[2] = {(cuuint32_t)BK, (cuuint32_t)kStN};
    rc = st2_encode_tmap(ctx, &p->tmap_b, w_all, 2, wd, ws, wb);
  }
  if (rc) { delete p; return rc; }
  *out = p;
  return 0;
}

template <bool DUAL>
static int launch_stencil(st2_ctx* ctx, TcConvPlan* p, float* gx, const double* coef) {
  const int n_pt = p->g.tiles_h * p->g.tiles_w;
  const int grid = n_pt < ctx->sm_count ? n_pt : ctx->sm_count;
  p->g.dbg = ctx->debug_flags;
  if (p->g.rot)
    st2_launch_pdl(ctx, false, tc_conv_first_stencil_kernel<DUAL, true>, grid, kNumThreads, StCfg<DUAL>::kSmemBytes,
                   p->tmap_a, DUAL ? p->tmap_a2 : p->tmap_a, p->tmap_b, p->g, gx, coef);
  else
    st2_launch_pdl(ctx, true, tc_conv_first_stencil_kernel<DUAL, false>, grid, kNumThreads, StCfg<DUAL>::kSmemBytes,
                   p->tmap_a, DUAL ? p->tmap_a2 : p->tmap_a, p->tmap_b, p->g, gx, coef);
  ST2_LAUNCH_CHECK(ctx);
  return 0;
}

static void set_halo(TcConvPlan* p, const HaloArgs* halo) {
  if (halo != nullptr && halo->push_blocks > 0) { p->g.halo = *halo; p->g.rot = 1; }
  else { memset(&p->g.halo, 0, sizeof(p->g.halo)); p->g.rot = 0; }
}

bool tc_conv_supports_halo(const TcConvPlan* p) { return p != nullptr && !(p->ws_kb && p->pair); }

bool tc_conv_supports_style_fuse(const TcConvPlan* p) {
  return p != nullptr && p->pair && !p->ws_kb && p->bn == 128 && p->g.cout == 128 && p->g.taps == 9;
}

int tc_conv_set_style_fuse(st2_ctx* ctx, TcConvPlan* p, const __half* act_below, const __half* d_scaled) {
  if (!tc_conv_supports_style_fuse(p)) return st2_fail(ctx, ST2_ERR_ARG, "tc_conv_set_style_fuse: wrong kernel");
  const ConvGeom& g = p->g;
  cuuint64_t dims[3] = {(cuuint64_t)g.cout, (cuuint64_t)g.W, (cuuint64_t)(g.H + 2 * g.hoff)};
  cuuint64_t strides[2] = {(cuuint64_t)g.cout * 2, (cuuint64_t)g.W * g.cout * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)g.TW, (cuuint32_t)g.TH};
  int rc = st2_encode_tmap(ctx, &p->tmap_f, act_below, 3, dims, strides, box);
  if (rc) return rc;
  cuuint64_t dd[2] = {(cuuint64_t)g.cout, (cuuint64_t)g.cout};
  cuuint64_t ds[1] = {(cuuint64_t)g.cout * 2};
  cuuint32_t db[2] = {(cuuint32_t)BK, (cuuint32_t)(p->bn / 2)};
  if ((rc = st2_encode_tmap(ctx, &p->tmap_d, d_scaled, 2, dd, ds, db))) return rc;
  p->sfuse = true;
  return 0;
}


int tc_conv_first_bwd_launch(st2_ctx* ctx, TcConvPlan* p, float* gx, const double* dual_coef, const HaloArgs* halo) {
  if (!p || p->bn != 16 || (p->dual != (dual_coef != nullptr)))
    return st2_fail(ctx, ST2_ERR_ARG, "tc_conv_first_bwd: wrong plan");
  set_halo(p, halo);
  if (p->stencil) return p->dual ? launch_stencil<true>(ctx, p, gx, dual_coef) : launch_stencil<false>(ctx, p, gx, nullptr);
  TcInject inj;
  inj.fc = nullptr; inj.sraw = nullptr; inj.coef = dual_coef; inj.pool = nullptr; inj.pool_wp = 0; inj.sfuse = 0;
  if (p->dual) return launch_ws<16, 2>(ctx, p, nullptr, nullptr, reinterpret_cast<__half*>(gx), EPI_RAW, inj);
  return launch_ws<16, 1>(ctx, p, nullptr, nullptr, reinterpret_cast<__half*>(gx), EPI_RAW, inj);
}

int tc_conv_dual_plan_create(st2_ctx* ctx, const __half* grad, const __half* act, const __half* w_dual, int H, int W,
                             TcConvPlan** out, int halo) {
  if (W < 16 || H < 16) return st2_fail(ctx, ST2_ERR_ARG, "tc_conv: canvas too small for the conv1_1 gradient kernel");
  TcConvPlan* p = new TcConvPlan();
  ConvGeom& g = p->g;
  g.H = H; g.W = W; g.cin = 128; g.cout = 16; g.taps = 9;       // cin = 128: 64 gradient + 64 activation channels
  g.hoff = halo;
  g.TW = kWsTW; g.TH = kWsTH;
  g.tiles_h = (H + g.TH - 1) / g.TH;
  g.tiles_w = (W + g.TW - 1) / g.TW;
  p->bn = 16; p->ws_kb = 2; p->pair = false; p->dual = true;
  g.n_blocks = 1;
  g.total_tiles = g.tiles_h * g.tiles_w;
  g.cblocks = 2;
  g.k_iters = 18;
  g.step_nb = g.step_tw = g.step_th = 0; g.dbg = 0;
  g.rot = 0;
  memset(&g.halo, 0, sizeof(g.halo));
  cuuint64_t dims[3] = {64, (cuuint64_t)W, (cuuint64_t)(H + 2 * halo)};
  cuuint64_t strides[2] = {128, (cuuint64_t)W * 128};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)kPatchW, (cuuint32_t)kPatchH};
  int rc = st2_encode_tmap(ctx, &p->tmap_a, grad, 3, dims, strides, box);
  if (!rc) rc = st2_encode_tmap(ctx, &p->tmap_a2, act, 3, dims, strides, box);
  if (!rc) {
    cuuint64_t wd[2] = {9 * 128, 16};
    cuuint64_t ws[1] = {9 * 128 * 2};
    cuuint32_t wb[2] = {(cuuint32_t)BK, 16};
    rc = st2_encode_tmap(ctx, &p->tmap_b, w_dual, 2, wd, ws, wb);
  }
  if (rc) { delete p; return rc; }
  *out = p;
  return 0;
}

int tc_conv_launch(st2_ctx* ctx, TcConvPlan* p, const float* bias, const __half* act, __half* out, int epi,
                   float out_scale, double* sumsq, const TcInject* inj_in, bool* pooled, const HaloArgs* halo) {
  if (epi == EPI_BIAS_RELU && !bias) return st2_fail(ctx, ST2_ERR_ARG, "tc_conv: bias required");
  if (halo != nullptr && halo->push_blocks > 0 && !tc_conv_supports_halo(p))
    return st2_fail(ctx, ST2_ERR_STATE, "tc_conv: this kernel has no in-kernel halo exchange");
  set_halo(p, halo);
  if (epi == EPI_MASK && !act) return st2_fail(ctx, ST2_ERR_ARG, "tc_conv: act required");
  TcInject inj;
  inj.fc = nullptr; inj.sraw = nullptr; inj.coef = nullptr; inj.pool = nullptr; inj.pool_wp = 0; inj.sfuse = 0;
  if (pooled) *pooled = false;
  if (inj_in) {
    if (epi != EPI_MASK && (inj_in->fc || inj_in->sraw || inj_in->coef))
      return st2_fail(ctx, ST2_ERR_ARG, "tc_conv: injection needs the mask epilogue");
    inj = *inj_in;
    // the fused pool lives in the generic, CTA-pair and TMA-store weight-stationary epilogues
    const bool can_pool = epi == EPI_BIAS_RELU && inj.pool != nullptr && pooled != nullptr && out_scale == 1.f &&
                          sumsq == nullptr && !(p->ws_kb && p->pair) && !(p->ws_kb && !(p->ws_kb == 1 && p->bn == 64)) &&
                          !ctx->knobs.no_pool_fusion;
    if (!can_pool) { inj.pool = nullptr; inj.pool_wp = 0; }
    else *pooled = true;
  }
  if (p->ws_kb && p->pair && out_scale == 1.f && sumsq == nullptr)
    return p->ws_kb == 1 ? launch_wsp<1>(ctx, p, bias, act, out, epi, inj) : launch_wsp<2>(ctx, p, bias, act, out, epi, inj);
  if (p->ws_kb && out_scale == 1.f && sumsq == nullptr) {
    if (p->ws_kb == 1 && p->bn == 64) return launch_ws<64, 1>(ctx, p, bias, act, out, epi, inj);
    if (p->ws_kb == 1 && p->bn == 128) return launch_ws<128, 1>(ctx, p, bias, act, out, epi, inj);
    if (p->ws_kb == 2 && p->bn == 64) return launch_ws<64, 2>(ctx, p, bias, act, out, epi, inj);
    if (p->ws_kb == 1 && p->bn == 16) return launch_ws<16, 1>(ctx, p, bias, act, out, EPI_RAW, inj);
    return st2_fail(ctx, ST2_ERR_STATE, "tc_conv: no weight-stationary kernel for this shape");
  }
  if (inj.sfuse && !(p->pair && !p->ws_kb && p->bn == 128 && p->sfuse && epi == EPI_MASK && inj.coef != nullptr))
    return st2_fail(ctx, ST2_ERR_STATE, "tc_conv: style fusion needs the 128-wide CTA-pair kernel with its maps set");
  if (p->pair && out_scale == 1.f && sumsq == nullptr) {
    if (inj.sfuse) return launch_pair<128, true>(ctx, p, bias, act, out, epi, inj);
    return p->bn == 256 ? launch_pair<256>(ctx, p, bias, act, out, epi, inj) : launch_pair<128>(ctx, p, bias, act, out, epi, inj);
  }
  switch (p->bn) {
    case 256: return launch_bn<256>(ctx, p, bias, act, out, epi, out_scale, sumsq, inj);
    case 128: return launch_bn<128>(ctx, p, bias, act, out, epi, out_scale, sumsq, inj);
    default:  return launch_bn<64>(ctx, p, bias, act, out, epi, out_scale, sumsq, inj);
  }
}

#define ST2_BOTH(K, ...) {ST2_KFN(K<__VA_ARGS__, false>), SM}, {ST2_KFN(K<__VA_ARGS__, true>), SM}
static St2SmemReg g_smem_conv_tc({
    {ST2_KFN(tc_conv_first_stencil_kernel<false, false>), StCfg<false>::kSmemBytes},
    {ST2_KFN(tc_conv_first_stencil_kernel<false, true>), StCfg<false>::kSmemBytes},
    {ST2_KFN(tc_conv_first_stencil_kernel<true, false>), StCfg<true>::kSmemBytes},
    {ST2_KFN(tc_conv_first_stencil_kernel<true, true>), StCfg<true>::kSmemBytes},
#define SM Cfg<256>::kSmemBytes
    ST2_BOTH(tc_conv_kernel, 256),
#undef SM
#define SM Cfg<128>::kSmemBytes
    ST2_BOTH(tc_conv_kernel, 128),
#undef SM
#define SM Cfg<64>::kSmemBytes
    ST2_BOTH(tc_conv_kernel, 64),
#undef SM
#define SM WsCfg<64, 1>::kSmemBytes
    ST2_BOTH(tc_conv_ws_kernel, 64, 1),
#undef SM
#define SM WsCfg<128, 1>::kSmemBytes
    ST2_BOTH(tc_conv_ws_kernel, 128, 1),
#undef SM
#define SM WsCfg<64, 2>::kSmemBytes
    ST2_BOTH(tc_conv_ws_kernel, 64, 2),
#undef SM
#define SM WsCfg<16, 1>::kSmemBytes
    ST2_BOTH(tc_conv_ws_kernel, 16, 1),
#undef SM
#define SM WsCfg<16, 2>::kSmemBytes
    ST2_BOTH(tc_conv_ws_kernel, 16, 2),
#undef SM
    {ST2_KFN(tc_conv2_kernel<256, false, false>), PairCfg<256>::kSmemBytes}, {ST2_KFN(tc_conv2_kernel<256, true, false>), PairCfg<256>::kSmemBytes},
    {ST2_KFN(tc_conv2_kernel<128, false, false>), PairCfg<128>::kSmemBytes}, {ST2_KFN(tc_conv2_kernel<128, true, false>), PairCfg<128>::kSmemBytes},
    {ST2_KFN(tc_conv2_kernel<128, false, true>), PairCfg<128>::kSmemBytes}, {ST2_KFN(tc_conv2_kernel<128, true, true>), PairCfg<128>::kSmemBytes},
    {ST2_KFN(tc_conv_wsp_kernel<1>), WspCfg<1>::kSmemBytes}, {ST2_KFN(tc_conv_wsp_kernel<2>), WspCfg<2>::kSmemBytes}});
#undef ST2_BOTH
#define ST2_BOTH(K, ...) ST2_KFN(K<__VA_ARGS__, false>), ST2_KFN(K<__VA_ARGS__, true>)
static St2KernelReg g_reg_conv_tc({ST2_KFN(tc_conv_first_stencil_kernel<false, false>),
                                      ST2_KFN(tc_conv_first_stencil_kernel<false, true>),
                                      ST2_KFN(tc_conv_first_stencil_kernel<true, false>),
                                      ST2_KFN(tc_conv_first_stencil_kernel<true, true>), ST2_BOTH(tc_conv_kernel, 256), ST2_BOTH(tc_conv_kernel, 128), ST2_BOTH(tc_conv_kernel, 64),
                                      ST2_BOTH(tc_conv_ws_kernel, 64, 1), ST2_BOTH(tc_conv_ws_kernel, 128, 1),
                                      ST2_BOTH(tc_conv_ws_kernel, 64, 2), ST2_BOTH(tc_conv_ws_kernel, 16, 1),
                                      ST2_BOTH(tc_conv_ws_kernel, 16, 2),
                                      ST2_KFN(tc_conv2_kernel<256, false, false>), ST2_KFN(tc_conv2_kernel<256, true, false>),
                                      ST2_KFN(tc_conv2_kernel<128, false, false>), ST2_KFN(tc_conv2_kernel<128, true, false>),
                                      ST2_KFN(tc_conv2_kernel<128, false, true>), ST2_KFN(tc_conv2_kernel<128, true, true>),
                                      ST2_KFN(tc_conv_wsp_kernel<1>), ST2_KFN(tc_conv_wsp_kernel<2>)});
#undef ST2_BOTH

// Column-strip form of the 3x3 64->64 convolution between flat-padded activations (FPA) for sm_100a: the layer that is 93 % of a
// VDSR 4K frame (vdsr/vdsr/model_vdsr.py:64-83, 18 of the 20 layers) -- NO lane shift, no shuffles in the epilogue.
//
// conv_tc.cu stacks the HORIZONTAL taps along N and pays for it in the epilogue: the three accumulator blocks are one lane apart
// and have to be added with warp shuffles + a cross-quadrant exchange, which share the shared-memory crossbar with the tensor
// core's operand fetch (975 TFLOP/s, 1 124 without the lane shift).  This kernel turns the layer around, as the
// fused ESPCN kernel does (espcn_fused.cuh):
//   * a CTA walks COLUMN STRIPS of 126 output pixels from top to bottom, one image row per step; TMEM lane i of a step is pixel
//     x = 126*strip - 1 + i of the row (128 lanes = 126 stored + one apron lane per side);
//   * the HORIZONTAL tap dx is a ROW-SHIFTED A descriptor into the row's 128 x 128-byte input tile (start address -/+ one row;
//     the two rows beyond the tile are neighbouring memory and only reach the two discarded apron lanes);
//   * the VERTICAL taps are stacked along N: input row u feeds the output rows v = u, u+1, u+2 (tap dy = u + 2 - v), whose
//     accumulators are three neighbouring 64-column slots of a ring of SIX (row v in slot v % 6), so ONE N = 192 instruction per
//     (dx, K-step) -- the full-rate shape, 96 cycles -- serves all three; where the three slots wrap around the ring (two rows
//     in six) the instruction is split in an N = 128 and an N = 64 one.  The weights keep one fixed order [dy2 dy1 dy0] per dx.
//     A slot is re-opened four rows after it was read out, so the MMAs never wait for the epilogue (the single three-slot window
//     of the ESPCN kernel does);
//   * row v+2's slot is opened by an instruction against a block of ZERO weights (accumulate off), or by the split
//     instruction itself when the row sits alone in it;
//   * epilogue = tcgen05.ld -> bias -> ReLU -> [ReLU' mask of a saved activation: the data-gradient form] -> bf16 -> swizzled
//     st.shared -> one TMA store per row (clipped at the row end, so the strip that ends a row needs no special case; the FPA's
//     zero column x = W and the zero row above every image are written too: the output is a complete FPA).
// Tiles are 4-D TMA boxes of the FPA ([image][image row][x][channel], out-of-range coordinates = zero fill = the SAME padding):
//   * wide images: box {64, 128 x, 1 row, 1 image} at x = 126*strip - 1;
//   * narrow images (2*(W+1) <= 126): K = 126 / (W+1) images SIDE BY SIDE in one tile, box {64, W+1, 1, K} at x = -1 -- each
//     image's out-of-range column -1 is the zero separator its neighbour's right-hand tap reads (3 x 42 lanes for VDSR's 41-px
//     training patches); the tile rows past K*(W+1) are never written and stay zero.
// A launch may run a CHAIN of layers of one geometry (srk_conv_tc_chain): the (layer, part) phases form one continuous row
// sequence -- ring / slot / phase counters simply continue -- with a grid barrier between layers (epilogue sets count their
// completed TMA stores into a global counter, the producer polls it before the next layer's first load) and the weights
// re-loaded while the barrier drains; small batches can be split into parts whose layers alternate, hiding one part's barrier
// behind the other's work.  (Measured: not faster than PDL launches at the training shape; kept as an entry point.)
// Short images: the host cuts the CTAs' unit ranges at image-group boundaries (`bounds`), since a CTA that straddles two
// column walks pays the two apron rows twice.
//
// Warp roles (11 warps): 0..7 epilogue (two sets of four TMEM-quadrant warps taking rows alternately), 8 TMA producer,
// 9 MMA issuer, 10 set-up (TMEM allocation, weights of the current layer).
#include <algorithm>
#include <cstdlib>

#include "sm100_ptx.cuh"
#include "srk_common.cuh"
#include "strip_walk.cuh"

namespace srk {

constexpr int kCsStripW = 126;
constexpr int kCsStages = 6;                  // input row tiles in flight
constexpr int kCsSlots = 6;                   // accumulator slots (64 TMEM columns each)
constexpr int kCsTileBytes = 128 * 128;
constexpr int kCsOffW = 0;                    // 9 weight blocks [64][64] bf16: dx-major, [dy2 dy1 dy0] inside
constexpr int kCsOffZ = kCsOffW + 9 * 8192;   // 64 rows of zero weights
constexpr int kCsOffIn = kCsOffZ + 8192;
constexpr int kCsOffStage = kCsOffIn + kCsStages * kCsTileBytes;  // one output staging tile per epilogue set
constexpr int kCsOffBias = kCsOffStage + 2 * kCsTileBytes;
constexpr int kCsOffTab = kCsOffBias + 512;
constexpr int kCsOffBars = (kCsOffTab + 3 * kEfTabInts * 4 + 7) / 8 * 8;
constexpr int kCsNumBars = 2 + 2 * kCsStages + 2 * kCsSlots;
constexpr int kCsOffTmemSlot = kCsOffBars + kCsNumBars * 8;
constexpr int kCsSmem = kCsOffTmemSlot + 16 + 1024;
constexpr int kCsThreads = 11 * 32;
static_assert(kCsSmem <= 227 * 1024, "shared-memory plan does not fit");

constexpr int kCsMaxLayers = 20;
constexpr int kCsMaxParts = 3;
constexpr int kCsMaxGrid = 160;

struct alignas(64) CsLayer {
  CUtensorMap map_in;   // FPA 4-D (make_tensor_map_fpa4), box {64, box_x_in, 1, K}
  CUtensorMap map_out;  // FPA 4-D, box {64, rows_out / K, 1, K}
  CUtensorMap map_w;    // [9*64][64] box {64, 64}: SRK_PACK_FWD / SRK_PACK_DGRAD blocks (tap-major, t = dy*3 + dx)
  const float* bias;    // [64] or null
  const __nv_bfloat16* mask;  // FPA of the layer's OUTPUT geometry or null: y *= (mask > 0)  (ReLU' of the saved activation)
  int act;
  int pad_;
};

struct alignas(64) ConvStripParams {
  CsLayer layer[kCsMaxLayers];  // a chain: layer l reads what layer l-1 wrote (same geometry); one grid barrier between layers
  int n_layers;
  int n_total;          // images in the FPA
  int H, W, strips;
  int K;                // images side by side in one tile (1: wide mode, strips of 126 pixels)
  int box_x_in;         // 128 (wide) | Wp: tile rows per image
  int rows_out;         // 126 (wide) | K * Wp: staging rows stored per step
  int parts;            // the image groups are split into `parts` independent parts whose layers alternate (A0 B0 A1 B1 ...):
                        // while part A's layer l drains, is stored and passes its grid barrier, the CTAs compute part B's layer l
  int part_g0[kCsMaxParts + 1];  // image groups [part_g0[q], part_g0[q+1]) belong to part q
  unsigned int* sync;   // grid-barrier counters (one per part), zeroed before the launch (n_layers > 1)
  int use_bounds;       // single-part launches of short images: CTA b owns the units [bounds[b], bounds[b+1]) -- cut so that no CTA
  int bounds[kCsMaxGrid + 1];  // straddles two images (every straddle costs two more apron rows on the critical path)
};

// Grid barrier between the layers of a chain: every epilogue set of every CTA adds one when its stores of the layer are
// complete; the TMA producer waits for 2 * gridDim.x * layer before it reads the next layer's input.  All CTAs are co-resident
// (grid <= number of SMs, one CTA per SM), the spin is bounded (a trap, not a hang).
__device__ __forceinline__ void cs_grid_wait(const unsigned int* ctr, unsigned int target) {
  unsigned int v, n = 0;
  do {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    if (v >= target) break;
    __nanosleep(64);
    if (++n > (1u << 24)) __trap();
  } while (true);
  asm volatile("fence.proxy.async;" ::: "memory");  // the async proxy (TMA loads) must see what the other CTAs' TMA stores wrote
}

__global__ void __launch_bounds__(kCsThreads, 1) conv_strip_kernel(const __grid_constant__ ConvStripParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sb = smem_u32(smem);
  const uint32_t s_w = sb + kCsOffW, s_z = sb + kCsOffZ, s_in = sb + kCsOffIn, s_stage = sb + kCsOffStage, s_bars = sb + kCsOffBars;
  int* const s_tab = reinterpret_cast<int*>(smem + kCsOffTab);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + kCsOffTmemSlot);
  const uint32_t bar_w = s_bars, bar_wfree = s_bars + 8;
  auto FULL = [&](int i) { return s_bars + 8u * (2 + i); };                                // input tile landed (TMA bytes)
  auto EMPTY = [&](int i) { return s_bars + 8u * (2 + kCsStages + i); };                   // input tile consumed (tensor-pipe commit)
  auto ACC_FULL = [&](int i) { return s_bars + 8u * (2 + 2 * kCsStages + i); };            // output row accumulated (commit)
  auto ACC_FREE = [&](int i) { return s_bars + 8u * (2 + 2 * kCsStages + kCsSlots + i); };  // slot read out (4 quadrant warps)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = p.n_layers, P = p.parts;

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_wfree, 1);
    for (int i = 0; i < kCsStages; ++i) {
      mbar_init(FULL(i), 1);
      mbar_init(EMPTY(i), 1);
    }
    for (int i = 0; i < kCsSlots; ++i) {
      mbar_init(ACC_FULL(i), 1);
      mbar_init(ACC_FREE(i), 4);
    }
    fence_mbar_init();
  }
  if (warp == 10) {
    tmem_alloc<512>(smem_u32(tmem_slot));
    if (lane == 0) {
      tma_prefetch_desc(&p.layer[0].map_in);
      tma_prefetch_desc(&p.layer[0].map_out);
      tma_prefetch_desc(&p.layer[0].map_w);
    }
  }
  // zero weights block; input stages: several images side by side fill only K * Wp rows of a tile -- the rest stays zero
  for (int i = threadIdx.x; i < (8192 + kCsStages * kCsTileBytes) / 16; i += kCsThreads) *reinterpret_cast<uint4*>(smem + kCsOffZ + i * 16) = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x >= 64 && threadIdx.x < 64 + P) {  // this CTA's share of every part: an even cut of the part's unit range
    const int q = threadIdx.x - 64;
    const long long units = (long long)(p.part_g0[q + 1] - p.part_g0[q]) * p.strips * p.H;
    const uint32_t u0 = p.use_bounds ? uint32_t(p.bounds[blockIdx.x]) : uint32_t((units * blockIdx.x) / gridDim.x);
    const uint32_t u1 = p.use_bounds ? uint32_t(p.bounds[blockIdx.x + 1]) : uint32_t((units * (blockIdx.x + 1)) / gridDim.x);
    ef_build_segments(s_tab + q * kEfTabInts, u0, u1, p.H, 0, p.strips, 2);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  int Vq[kCsMaxParts], Vsum = 0;  // virtual rows of this CTA per layer and part: every segment = its output rows + 2 (the input rows above / below)
#pragma unroll
  for (int q = 0; q < kCsMaxParts; ++q) {
    Vq[q] = q < P ? s_tab[q * kEfTabInts + s_tab[q * kEfTabInts + 71]] : 0;
    Vsum += Vq[q];
  }
  const bool wide = p.K == 1;

  // The (layer, part) phases of a chain run as ONE sequence of rows: stage, slot and barrier-phase counters simply continue (the
  // two rows a phase's last steps open beyond its end are the first two rows of the next phase, whose results are discarded anyway).
  if (Vsum > 0) {
    if (warp == 10) {
      // ---------------------------------------------------------------- weights: one layer resident at a time
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        if (l > 0) mbar_wait(bar_wfree, (l - 1) & 1);  // the previous layer's MMAs have retired
        if (lane == 0) {
          mbar_arrive_expect_tx(bar_w, 9 * 8192);
          for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) tma_load_2d(s_w + dx * 24576 + (2 - dy) * 8192, &p.layer[l].map_w, 0, (dy * 3 + dx) * 64, bar_w);
        }
        __syncwarp();
      }
    } else if (warp == 8) {
      // ---------------------------------------------------------------- TMA producer: one input row tile per step
      EfSeg w;
      int st = 0, lap = 0;
      const uint32_t tile_bytes = uint32_t(p.box_x_in) * uint32_t(p.K) * 128u;
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
#pragma unroll 1
        for (int q = 0; q < P; ++q) {
          const int V = Vq[q];
          const int* tab = s_tab + q * kEfTabInts;
          if (l > 0 && V > 0) {
            if (lane == 0) cs_grid_wait(p.sync + q, 2u * gridDim.x * uint32_t(l));
            __syncwarp();
          }
          ef_seg_load(w, tab, 0);
#pragma unroll 1
          for (int v = 0; v < V; ++v) {
            ef_seg_seek(w, tab, v);
            if (lap > 0) mbar_wait(EMPTY(st), (lap - 1) & 1);
            if (lane == 0) {
              mbar_arrive_expect_tx(FULL(st), tile_bytes);
              // virtual row j of a segment is the input row ya - 1 + j = row index ya + j of the image (index 0 = the zero row)
              tma_load_4d(s_in + st * kCsTileBytes, &p.layer[l].map_in, 0, wide ? w.s * kCsStripW - 1 : -1, w.ya + (v - w.v0),
                          (p.part_g0[q] + w.n) * p.K, FULL(st));
            }
            __syncwarp();
            if (++st == kCsStages) st = 0, ++lap;
          }
        }
      }
    } else if (warp == 9) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t id192 = umma_idesc_bf16(128, 192, 0, 0), id128 = umma_idesc_bf16(128, 128, 0, 0), id64 = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint64_t hi = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
      int st = 0, lap = 0, s0 = 0, g = 0;  // g % stages, g / stages, g % 6, global row
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        mbar_wait(bar_w, l & 1);
#pragma unroll 1
        for (int v = 0; v < Vsum; ++v, ++g) {
          const int so = (s0 + 2) % kCsSlots;  // slot opened for row g + 2; its previous row (g - 4) must have been read out
          if (g + 2 >= kCsSlots) mbar_wait(ACC_FREE(so), (((g + 2) / kCsSlots) - 1) & 1);
          mbar_wait(FULL(st), lap & 1);
          tc_fence_after();
          const uint32_t a0 = s_in + st * kCsTileBytes - 128;
          if (elect_one()) {
            if (s0 <= 3) {
              umma_bf16(tmem + so * 64, umma_desc(hi, a0 + 128), umma_desc(hi, s_z), id64, 0);
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(tmem + s0 * 64, umma_desc(hi, a0 + dx * 128 + k * 32), umma_desc(hi, s_w + dx * 24576 + k * 32), id192, 1);
            } else if (s0 == 4) {  // rows g, g+1 in slots 4, 5; row g+2 alone in slot 0: its first instruction opens it
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16(tmem + 4 * 64, umma_desc(hi, a0 + dx * 128 + k * 32), umma_desc(hi, s_w + dx * 24576 + k * 32), id128, 1);
                  umma_bf16(tmem, umma_desc(hi, a0 + dx * 128 + k * 32), umma_desc(hi, s_w + dx * 24576 + 16384 + k * 32), id64, (dx | k) != 0);
                }
            } else {  // row g in slot 5; rows g+1, g+2 in slots 0, 1
              umma_bf16(tmem + 64, umma_desc(hi, a0 + 128), umma_desc(hi, s_z), id64, 0);
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  umma_bf16(tmem + 5 * 64, umma_desc(hi, a0 + dx * 128 + k * 32), umma_desc(hi, s_w + dx * 24576 + k * 32), id64, 1);
                  umma_bf16(tmem, umma_desc(hi, a0 + dx * 128 + k * 32), umma_desc(hi, s_w + dx * 24576 + 8192 + k * 32), id128, 1);
                }
            }
            umma_commit(EMPTY(st));
            umma_commit(ACC_FULL(s0));
            if (v == Vsum - 1 && l + 1 < L) umma_commit(bar_wfree);
          }
          __syncwarp();
          if (++st == kCsStages) st = 0, ++lap;
          if (++s0 == kCsSlots) s0 = 0;
        }
      }
    } else if (warp < 8) {
      // ---------------------------------------------------------------- epilogue: set `set` takes the global rows g = set, set + 2, ...
      const int quad = warp & 3, set = warp >> 2, gl = quad * 32 + lane;
      const uint32_t lane_addr = uint32_t(quad * 32) << 16;
      const uint32_t stage = s_stage + set * kCsTileBytes;
      const bool leader = (quad == 0 && lane == 0);
      const int Wp = p.W + 1;
      const int lane_k = wide ? 0 : gl / Wp;               // image of the tile this lane belongs to
      const int lane_x = wide ? gl - 1 : gl - lane_k * Wp - 1;  // pixel within the strip / image (-1: the zero lane left of an image)
      const bool writer = gl >= 1 && gl <= p.rows_out;     // lane i holds staging row i - 1
      const uint32_t srow = stage + uint32_t(gl - 1) * 128u;
      const int sw = (gl - 1) & 7;
      EfSeg w;
      int g = set;
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const CsLayer& ly = p.layer[l];
        const bool relu = ly.act == SRK_ACT_RELU;
        const float* const biasg = ly.bias;
        const __nv_bfloat16* const maskg = ly.mask;
#pragma unroll 1
        for (int q = 0; q < P; ++q) {
        const int* tab = s_tab + q * kEfTabInts;
        ef_seg_load(w, tab, 0);
        const int g_begin = l * Vsum + (q > 0 ? Vq[0] : 0) + (q > 1 ? Vq[1] : 0);
        const int g_end = g_begin + Vq[q];
        const int img0 = p.part_g0[q] * p.K;
#pragma unroll 1
        for (; g < g_end; g += 2) {
          const int v = g - g_begin;
          ef_seg_seek(w, tab, v);
          const int j = v - w.v0;
          const int x = (wide ? w.s * kCsStripW : 0) + lane_x;
          const int img = img0 + w.n * p.K + lane_k;
          const int y = w.ya + j - 2;
          const bool keep = j >= 2 && x >= 0 && x < p.W && img < p.n_total;  // x == W is the FPA's zero column; beyond it the store clips
          uint4 mk[8];
          if (maskg != nullptr && keep) {  // ReLU' mask of this pixel (issued before the wait: the latency hides behind it)
            const uint4* mp = reinterpret_cast<const uint4*>(maskg + ((int64_t(img) * (p.H + 1) + y + 1) * Wp + x) * 64);
#pragma unroll
            for (int c = 0; c < 8; ++c) mk[c] = __ldg(mp + c);
          }
          const int slot = g % kCsSlots;
          mbar_wait(ACC_FULL(slot), (g / kCsSlots) & 1);
          tc_fence_after();
          uint32_t a[64];
          {
            uint32_t lo[32], hi2[32];
            tmem_ld_32x32b_x32(tmem + slot * 64 + lane_addr, lo);
            tmem_ld_32x32b_x32(tmem + slot * 64 + 32 + lane_addr, hi2);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) a[c] = lo[c], a[32 + c] = hi2[c];
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(ACC_FREE(slot));
          if (j >= 2) {  // (uniform over the set: j depends on the row only)
            uint32_t pk[32];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (biasg != nullptr) b = __ldg(reinterpret_cast<const float4*>(biasg) + c);
              float v0 = __uint_as_float(a[4 * c]) + b.x, v1 = __uint_as_float(a[4 * c + 1]) + b.y;
              float v2 = __uint_as_float(a[4 * c + 2]) + b.z, v3 = __uint_as_float(a[4 * c + 3]) + b.w;
              if (relu) v0 = fmaxf(v0, 0.f), v1 = fmaxf(v1, 0.f), v2 = fmaxf(v2, 0.f), v3 = fmaxf(v3, 0.f);
              __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
              pk[2 * c] = keep ? *reinterpret_cast<uint32_t*>(&h0) : 0u;
              pk[2 * c + 1] = keep ? *reinterpret_cast<uint32_t*>(&h1) : 0u;
            }
            if (maskg != nullptr && keep) {
              const __nv_bfloat162 z2 = __floats2bfloat162_rn(0.f, 0.f);
              const uint32_t* mw = reinterpret_cast<const uint32_t*>(mk);
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                const __nv_bfloat162 gt = __hgt2(*reinterpret_cast<const __nv_bfloat162*>(&mw[c]), z2);  // 1.0 where the activation was positive
                const __nv_bfloat162 r = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&pk[c]), gt);
                pk[c] = *reinterpret_cast<const uint32_t*>(&r);
              }
            }
            if (j == 2 && w.ya == 0) {
              // first output row of an image: this strip's part of the FPA's zero row above it is written too, so that a
              // freshly allocated output buffer is a complete FPA
              if (leader) tma_store_wait_read<0>();
              asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
              if (writer) {
#pragma unroll
                for (int c = 0; c < 8; ++c) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(srow + c * 16), "r"(0u) : "memory");
              }
              fence_proxy_async_smem();
              asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
              if (leader) {
                tma_store_4d(&ly.map_out, 0, wide ? w.s * kCsStripW : 0, 0, img0 + w.n * p.K, stage);
                tma_store_commit();
              }
            }
            if (leader) tma_store_wait_read<0>();  // the previous store of this set has finished reading the staging tile
            asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
            if (writer) {
#pragma unroll
              for (int c = 0; c < 8; ++c)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((c ^ sw) << 4)), "r"(pk[4 * c]), "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]),
                             "r"(pk[4 * c + 3])
                             : "memory");
            }
            fence_proxy_async_smem();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
            if (leader) {
              tma_store_4d(&ly.map_out, 0, wide ? w.s * kCsStripW : 0, y + 1, img0 + w.n * p.K, stage);
              tma_store_commit();
            }
          }
        }
        if (l + 1 < L && leader) {  // this set's share of the (layer, part) is in global memory: one arrival at the part's grid barrier
          tma_store_wait_all<0>();
          asm volatile("fence.proxy.async;" ::: "memory");
          __threadfence();
          atomicAdd(p.sync + q, 1u);
        }
        }
      }
      if (leader) tma_store_wait_all<0>();
    }
  } else if (L > 1 && warp < 8 && (warp & 3) == 0 && lane == 0) {
    for (int q = 0; q < P; ++q) atomicAdd(p.sync + q, uint32_t(L - 1));  // a CTA without work still owes the grid barriers its arrivals
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc<512>(tmem);
}

// Geometry of the strip form: wide images are cut into strips of 126 pixels (K = 1); narrow ones sit K side by side in one
// tile, each with its zero column (K * (W+1) <= 126 lanes).  `lanes` = the share of a tile's 126 lanes that carry pixels.
struct CsGeom {
  int K, strips, box_x_in, rows_out;
  double lanes;
};
static CsGeom cs_geom(int W) {
  const int Wp = W + 1;
  CsGeom g;
  if (2 * Wp <= kCsStripW) {
    g.K = kCsStripW / Wp;
    g.strips = 1;
    g.box_x_in = Wp;
    g.rows_out = g.K * Wp;
    g.lanes = double(g.K * Wp) / kCsStripW;
  } else {
    g.K = 1;
    g.strips = (Wp + kCsStripW - 1) / kCsStripW;
    g.box_x_in = 128;
    g.rows_out = kCsStripW;
    g.lanes = double(Wp) / (g.strips * kCsStripW);
  }
  return g;
}

// AUTO form: strips wherever at least 80 % of the lanes carry pixels (3 x 42 lanes for 41-pixel patches, 242-pixel panels, whole
// 4K rows); otherwise the flat stream, which has no such quantisation.
bool conv_strip_applicable(srk_ctx*, int n_img, int /*H*/, int W) {
  const CsGeom g = cs_geom(W);
  return (g.lanes >= 0.8 && n_img >= g.K) || W > 254;  // (rows wider than 254 pixels are beyond the flat stream's shared-memory ring)
}

int launch_conv_strip_chain(srk_ctx* h, int n_layers, const void* const* x_fpa, const void* const* w_packed, const float* const* bias, const int* act,
                            void* const* y_fpa, const void* const* mask_src, int n_img, int H, int W, unsigned int* sync, cudaStream_t stream) {
  SRK_REQUIRE(n_layers >= 1 && n_layers <= kCsMaxLayers, "conv_strip: %d layers (at most %d per chain)", n_layers, kCsMaxLayers);
  SRK_REQUIRE(n_layers == 1 || sync != nullptr, "conv_strip: a chain needs a grid-barrier word");
  SRK_REQUIRE(kCsSmem <= h->smem_optin, "conv_strip: needs %d B smem, device allows %d", kCsSmem, h->smem_optin);
  if (first_use(h, reinterpret_cast<const void*>(&conv_strip_kernel)))
    SRK_CHECK_CUDA(cudaFuncSetAttribute(conv_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCsSmem));
  const CsGeom g = cs_geom(W);
  ConvStripParams p{};
  for (int l = 0; l < n_layers; ++l) {
    SRK_REQUIRE(act[l] == SRK_ACT_NONE || act[l] == SRK_ACT_RELU, "conv_strip: activation %d", act[l]);
    SRK_REQUIRE(x_fpa[l] && w_packed[l] && y_fpa[l], "conv_strip: null buffer in layer %d", l);
    CsLayer& ly = p.layer[l];
    if (int rc = make_tensor_map_fpa4(h, &ly.map_in, x_fpa[l], n_img, H, W, uint32_t(g.box_x_in), uint32_t(g.K))) return rc;
    if (int rc = make_tensor_map_fpa4(h, &ly.map_out, y_fpa[l], n_img, H, W, uint32_t(g.rows_out / g.K), uint32_t(g.K))) return rc;
    if (int rc = make_tensor_map_2d(h, &ly.map_w, w_packed[l], 9 * 64, 64, 64)) return rc;
    ly.bias = bias ? bias[l] : nullptr;
    ly.mask = static_cast<const __nv_bfloat16*>(mask_src ? mask_src[l] : nullptr);
    ly.act = act[l];
  }
  p.n_layers = n_layers;
  p.n_total = n_img;
  p.H = H;
  p.W = W;
  p.strips = g.strips;
  p.K = g.K;
  p.box_x_in = g.box_x_in;
  p.rows_out = g.rows_out;
  p.sync = sync;
  // A CTA follows at most kEfMaxSegs strip segments (one per strip it touches): image groups are processed in chunks small enough
  // for that (one launch for anything but thousands of tiny images; a chain must fit one launch).
  const int groups = (n_img + g.K - 1) / g.K;
  SRK_REQUIRE(g.strips <= (kEfMaxSegs - 3) * h->num_sms, "conv_strip: image width %d too large", W);
  const int chunk = std::max(1, (kEfMaxSegs - 3) * h->num_sms / g.strips);
  SRK_REQUIRE(n_layers == 1 || groups <= chunk, "conv_strip: %d image groups do not fit one launch of a chain", groups);
  for (int g0 = 0; g0 < groups; g0 += chunk) {
    const int ng = std::min(chunk, groups - g0);
    const long long units = (long long)ng * g.strips * H;
    SRK_REQUIRE(units < (1ll << 31), "conv_strip: too many strip rows for one launch");
    const int grid = int(std::min<long long>(h->num_sms, std::max<long long>(1, units / 4)));
    // short chains of small layers (a few rows per CTA and layer) are latency-bound at the grid barrier: two parts alternate
    int parts = 1;
    if (n_layers > 1 && ng >= 2 && units / grid < 24) parts = 2;
    p.parts = parts;
    for (int q = 0; q <= parts; ++q) p.part_g0[q] = g0 + int((long long)ng * q / parts);
    // Short column walks (a few rows per CTA): give every walk a whole number of CTAs and cut it evenly, so that no CTA straddles
    // two walks -- a straddling CTA pays the two apron rows twice and sets the launch's critical path (VDSR training: 83.0 k ->
    // 89.4 k patches/s).
    const int walks = ng * g.strips;
    p.use_bounds = 0;
    if (parts == 1 && grid <= kCsMaxGrid && walks <= grid && units / grid < 32) {
      p.use_bounds = 1;
      int b = 0;
      p.bounds[0] = 0;
      for (int wk = 0; wk < walks; ++wk) {
        const int ctas = (grid * (wk + 1)) / walks - (grid * wk) / walks;  // >= 1
        for (int c = 1; c <= ctas; ++c) p.bounds[++b] = wk * H + int((long long)H * c / ctas);
      }
    }
    if (n_layers > 1) SRK_CHECK_CUDA(cudaMemsetAsync(sync, 0, kCsMaxParts * sizeof(unsigned int), stream));
    SRK_CHECK_CUDA(launch_pdl(conv_strip_kernel, dim3(grid), dim3(kCsThreads), size_t(kCsSmem), stream, p));
  }
  return 0;
}

int launch_conv_strip(srk_ctx* h, const void* x_fpa, const void* w_packed, const float* bias, int act, int n_img, int H, int W, void* y_fpa,
                      const void* mask_src, cudaStream_t stream) {
  return launch_conv_strip_chain(h, 1, &x_fpa, &w_packed, &bias, &act, &y_fpa, &mask_src, n_img, H, W, nullptr, stream);
}

}  // namespace srk

using namespace srk;

extern "C" int srk_conv_tc_chain(srk_handle_t h, int n_layers, const void* const* x_fpa, const void* const* w_packed, const float* const* bias,
                                 const int* act, void* const* y_fpa, const void* const* mask_src, int n_img, int H, int W, void* sync_word,
                                 srk_stream_t stream) {
  SRK_REQUIRE(h && x_fpa && w_packed && act && y_fpa, "srk_conv_tc_chain: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  return launch_conv_strip_chain(h, n_layers, x_fpa, w_packed, bias, act, y_fpa, mask_src, n_img, H, W, static_cast<unsigned int*>(sync_word),
                                 as_stream(stream));
}

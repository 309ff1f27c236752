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
//   * epilogue = tcgen05.ld -> bias -> ReLU -> bf16 -> swizzled st.shared -> one 3-D TMA store per row (clipped at the row end,
//     so the strip that ends a row needs no special case; the zero pad column x = W is written as zero).
// Input rows arrive by 3-D TMA ([image row][x][channel], out-of-range x = zero fill = the SAME padding; the row above the first
// and below the last image row are the FPA's zero rows).
//
// Warp roles (11 warps): 0..7 epilogue (two sets of four TMEM-quadrant warps taking rows alternately), 8 TMA producer,
// 9 MMA issuer, 10 set-up (TMEM allocation, weights, zero block).
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
constexpr int kCsOffTab = kCsOffBias + 256;
constexpr int kCsOffBars = (kCsOffTab + kEfTabInts * 4 + 7) / 8 * 8;
constexpr int kCsNumBars = 1 + 2 * kCsStages + 2 * kCsSlots;
constexpr int kCsOffTmemSlot = kCsOffBars + kCsNumBars * 8;
constexpr int kCsSmem = kCsOffTmemSlot + 16 + 1024;
constexpr int kCsThreads = 11 * 32;
static_assert(kCsSmem <= 227 * 1024, "shared-memory plan does not fit");

struct alignas(64) ConvStripParams {
  CUtensorMap map_in;   // [n_img*(H+1)][Wp][64]  box {64, 128, 1}
  CUtensorMap map_out;  // same tensor shape        box {64, 126, 1}
  CUtensorMap map_w;    // [9*64][64]               box {64, 64}      SRK_PACK_FWD (tap-major, t = dy*3 + dx)
  const float* bias;    // [64] or null
  int n_base, n_img, H, W, strips;  // images [n_base, n_base + n_img) of the FPA are processed by this launch
  long long units;      // n_img * strips * H
  int act;
};

__global__ void __launch_bounds__(kCsThreads, 1) conv_strip_kernel(const __grid_constant__ ConvStripParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sb = smem_u32(smem);
  const uint32_t s_w = sb + kCsOffW, s_z = sb + kCsOffZ, s_in = sb + kCsOffIn, s_stage = sb + kCsOffStage, s_bars = sb + kCsOffBars;
  int* const s_tab = reinterpret_cast<int*>(smem + kCsOffTab);
  float* const s_bias = reinterpret_cast<float*>(smem + kCsOffBias);
  uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(smem + kCsOffTmemSlot);
  const uint32_t bar_w = s_bars;
  auto FULL = [&](int i) { return s_bars + 8u * (1 + i); };                                // input tile landed (TMA bytes)
  auto EMPTY = [&](int i) { return s_bars + 8u * (1 + kCsStages + i); };                   // input tile consumed (tensor-pipe commit)
  auto ACC_FULL = [&](int i) { return s_bars + 8u * (1 + 2 * kCsStages + i); };            // output row accumulated (commit)
  auto ACC_FREE = [&](int i) { return s_bars + 8u * (1 + 2 * kCsStages + kCsSlots + i); };  // slot read out (4 quadrant warps)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t u0 = uint32_t((p.units * blockIdx.x) / gridDim.x), u1 = uint32_t((p.units * (blockIdx.x + 1)) / gridDim.x);

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
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
      tma_prefetch_desc(&p.map_in);
      tma_prefetch_desc(&p.map_out);
      tma_prefetch_desc(&p.map_w);
    }
    for (int i = lane; i < 8192 / 16; i += 32) *reinterpret_cast<uint4*>(smem + kCsOffZ + i * 16) = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  if (threadIdx.x == 64) ef_build_segments(s_tab, u0, u1, p.H, 0, p.strips, 2);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int V = s_tab[s_tab[71]];  // virtual rows of this CTA: every segment = its output rows + 2 (the input rows above / below)
  const int rows_per_img = p.H + 1;

  if (V > 0) {
    if (warp == 10) {
      // ---------------------------------------------------------------- weights: resident for the whole kernel
      if (lane == 0) {
        mbar_arrive_expect_tx(bar_w, 9 * 8192);
        for (int dy = 0; dy < 3; ++dy)
          for (int dx = 0; dx < 3; ++dx) tma_load_2d(s_w + dx * 24576 + (2 - dy) * 8192, &p.map_w, 0, (dy * 3 + dx) * 64, bar_w);
      }
    } else if (warp == 8) {
      // ---------------------------------------------------------------- TMA producer: one 128-pixel input row tile per step
      EfSeg w;
      ef_seg_load(w, s_tab, 0);
      int st = 0, lap = 0;
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        ef_seg_seek(w, s_tab, v);
        if (lap > 0) mbar_wait(EMPTY(st), (lap - 1) & 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(FULL(st), kCsTileBytes);
          // virtual row j of a segment is the input row ya - 1 + j = FPA row index n*(H+1) + ya + j (index 0 = the zero row)
          tma_load_3d(s_in + st * kCsTileBytes, &p.map_in, 0, w.s * kCsStripW - 1, (p.n_base + w.n) * rows_per_img + w.ya + (v - w.v0), FULL(st));
        }
        __syncwarp();
        if (++st == kCsStages) st = 0, ++lap;
      }
    } else if (warp == 9) {
      // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t id192 = umma_idesc_bf16(128, 192, 0, 0), id128 = umma_idesc_bf16(128, 128, 0, 0), id64 = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint64_t hi = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
      mbar_wait(bar_w, 0);
      int st = 0, lap = 0, s0 = 0;  // v % stages, v / stages, v % 6
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        const int so = (s0 + 2) % kCsSlots;  // slot opened for row v + 2; its previous row (v - 4) must have been read out
        if (v + 2 >= kCsSlots) mbar_wait(ACC_FREE(so), (((v + 2) / kCsSlots) - 1) & 1);
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
          } else if (s0 == 4) {  // rows v, v+1 in slots 4, 5; row v+2 alone in slot 0: its first instruction opens it
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16(tmem + 4 * 64, umma_desc(hi, a0 + dx * 128 + k * 32), umma_desc(hi, s_w + dx * 24576 + k * 32), id128, 1);
                umma_bf16(tmem, umma_desc(hi, a0 + dx * 128 + k * 32), umma_desc(hi, s_w + dx * 24576 + 16384 + k * 32), id64, (dx | k) != 0);
              }
          } else {  // row v in slot 5; rows v+1, v+2 in slots 0, 1
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
        }
        __syncwarp();
        if (++st == kCsStages) st = 0, ++lap;
        if (++s0 == kCsSlots) s0 = 0;
      }
    } else if (warp < 8) {
      // ---------------------------------------------------------------- epilogue: set `set` takes the rows v = set, set + 2, ...
      const int quad = warp & 3, set = warp >> 2, gl = quad * 32 + lane;
      const uint32_t lane_addr = uint32_t(quad * 32) << 16;
      const uint32_t stage = s_stage + set * kCsTileBytes;
      const bool leader = (quad == 0 && lane == 0);
      const bool relu = p.act == SRK_ACT_RELU;
      const float4* const bias4 = reinterpret_cast<const float4*>(s_bias);
      EfSeg w;
      ef_seg_load(w, s_tab, 0);
#pragma unroll 1
      for (int v = set; v < V; v += 2) {
        const int slot = v % kCsSlots;
        mbar_wait(ACC_FULL(slot), (v / kCsSlots) & 1);
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
        ef_seg_seek(w, s_tab, v);
        const int j = v - w.v0;
        if (j >= 2) {  // (uniform over the set: j depends on the row only)
          const int x = w.s * kCsStripW + gl - 1;
          const bool keep = x < p.W;  // x == W is the FPA's zero column; x > W is clipped by the store
          uint32_t pk[32];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float4 b = bias4[c];
            float v0 = __uint_as_float(a[4 * c]) + b.x, v1 = __uint_as_float(a[4 * c + 1]) + b.y;
            float v2 = __uint_as_float(a[4 * c + 2]) + b.z, v3 = __uint_as_float(a[4 * c + 3]) + b.w;
            if (relu) v0 = fmaxf(v0, 0.f), v1 = fmaxf(v1, 0.f), v2 = fmaxf(v2, 0.f), v3 = fmaxf(v3, 0.f);
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v0, v1), h1 = __floats2bfloat162_rn(v2, v3);
            pk[2 * c] = keep ? *reinterpret_cast<uint32_t*>(&h0) : 0u;
            pk[2 * c + 1] = keep ? *reinterpret_cast<uint32_t*>(&h1) : 0u;
          }
          if (j == 2 && w.ya == 0) {
            // first output row of an image: this strip's part of the FPA's zero row above it (index n*(H+1)) is written too, so
            // that a freshly allocated output buffer is a complete FPA
            if (leader) tma_store_wait_read<0>();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
            if (gl < kCsStripW) {
#pragma unroll
              for (int c = 0; c < 8; ++c) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(stage + gl * 128 + c * 16), "r"(0u) : "memory");
            }
            fence_proxy_async_smem();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
            if (leader) {
              tma_store_3d(&p.map_out, 0, w.s * kCsStripW, (p.n_base + w.n) * rows_per_img, stage);
              tma_store_commit();
            }
          }
          if (leader) tma_store_wait_read<0>();  // the previous store of this set has finished reading the staging tile
          asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
          if (gl >= 1 && gl <= kCsStripW) {
            const int r = gl - 1;  // lane i holds pixel 126*s + i - 1: staging row i - 1
            const uint32_t dst = stage + r * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((c ^ (r & 7)) << 4)), "r"(pk[4 * c]), "r"(pk[4 * c + 1]), "r"(pk[4 * c + 2]),
                           "r"(pk[4 * c + 3])
                           : "memory");
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
          if (leader) {
            tma_store_3d(&p.map_out, 0, w.s * kCsStripW, (p.n_base + w.n) * rows_per_img + w.ya + j - 1, stage);  // output row ya + j - 2 -> FPA index + 1
            tma_store_commit();
          }
        }
      }
      if (leader) tma_store_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc<512>(tmem);
}

// AUTO form: wide rows only (a strip is 126 pixels: narrow patches would leave most lanes idle).
bool conv_strip_applicable(srk_ctx*, int /*n_img*/, int /*H*/, int W) {
  return std::getenv("SRK_NO_STRIP") == nullptr && W >= 112;  // (the environment switch is for A/B measurements)
}

int launch_conv_strip(srk_ctx* h, const void* x_fpa, const void* w_packed, const float* bias, int act, int n_img, int H, int W, void* y_fpa,
                      cudaStream_t stream) {
  SRK_REQUIRE(act == SRK_ACT_NONE || act == SRK_ACT_RELU, "conv_strip: activation %d", act);
  SRK_REQUIRE(kCsSmem <= h->smem_optin, "conv_strip: needs %d B smem, device allows %d", kCsSmem, h->smem_optin);
  if (first_use(h, reinterpret_cast<const void*>(&conv_strip_kernel)))
    SRK_CHECK_CUDA(cudaFuncSetAttribute(conv_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCsSmem));
  ConvStripParams p{};
  const int Wp = W + 1;
  const uint64_t img_rows = uint64_t(n_img) * (H + 1);
  if (int rc = make_tensor_map_3d(h, &p.map_in, x_fpa, 2, uint64_t(Wp), 64, img_rows, 128, uint64_t(Wp) * 128, 128)) return rc;
  if (int rc = make_tensor_map_3d(h, &p.map_out, y_fpa, 2, uint64_t(Wp), 64, img_rows, 128, uint64_t(Wp) * 128, kCsStripW)) return rc;
  if (int rc = make_tensor_map_2d(h, &p.map_w, w_packed, 9 * 64, 64, 64)) return rc;
  p.bias = bias;
  p.H = H;
  p.W = W;
  p.strips = (Wp + kCsStripW - 1) / kCsStripW;
  p.act = act;
  // A CTA follows at most kEfMaxSegs strip segments (one per strip it touches): images are processed in chunks small enough for
  // that (one launch for anything but thousands of tiny images).
  SRK_REQUIRE(p.strips <= (kEfMaxSegs - 3) * h->num_sms, "conv_strip: image width %d too large", W);
  const int chunk = std::max(1, (kEfMaxSegs - 3) * h->num_sms / p.strips);
  for (int n0 = 0; n0 < n_img; n0 += chunk) {
    p.n_base = n0;
    p.n_img = std::min(chunk, n_img - n0);
    p.units = (long long)p.n_img * p.strips * H;
    SRK_REQUIRE(p.units < (1ll << 31), "conv_strip: too many strip rows for one launch");
    const int grid = int(std::min<long long>(h->num_sms, std::max<long long>(1, p.units / 8)));
    SRK_CHECK_CUDA(launch_pdl(conv_strip_kernel, dim3(grid), dim3(kCsThreads), size_t(kCsSmem), stream, p));
  }
  return 0;
}

}  // namespace srk

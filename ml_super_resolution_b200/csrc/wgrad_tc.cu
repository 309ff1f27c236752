// Tensor-core weight gradient of a 3x3 64->64 conv layer over FPAs (include/srk.h) for sm_100a.
//
//   dW[tap][ci][co] = sum_p X[p + off(tap)][ci] * dY[p][co],   off(u,v) = (u-1)*Wp + (v-1)
//
// is a GEMM whose contraction index is the pixel row p.  Both operands are "MN-major" in smem
// exactly as TMA lands them ([pixel][channel] rows, SW128), so no transpose is ever materialised:
// tcgen05.mma reads A = X^T and B = dY through MN-major descriptors (a_major = b_major = 1).
// One M=128 x N=192 instruction covers SIX taps.  With q = p + (v-1) the sum reads
//   dW[u][v] = sum_q X[q + (u-1)*Wp][ci] * dY[q - (v-1)][co]:
// the two 64-row M atoms are two row-shifted windows of the X ring (u = 0, 1), LBO = Wp rows apart, and the three
// 64-column N atoms are three row-shifted windows of the dY tile (v = 2, 1, 0), LBO = ONE row apart (probe-verified:
// an MN-major atom stride may be any positive multiple of 16 B; it must be positive, so the X ring keeps a mirror of
// its first chunks behind its last slot, and every dY tile is loaded with one extra row on either side).  A second
// instruction pairs the u = 2 window with a block of ones, which makes rows 64..127 of that accumulator the bias
// gradient sum_q dY[q][co].  N = 192 runs at the tensor pipe's full rate (96 cycles), where the earlier five
// M=128 x N=64 instructions per K-step (60 cycles each, operand-fetch bound) cost 300.
// Two fp32 accumulators (2 x 192 TMEM columns) persist over the CTA's whole pixel range (split-K across CTAs); each
// CTA stores its partial block to a workspace and a second, deterministic kernel sums the partials into dW / dbias
// (fixed order: bit-reproducible gradients).
//
//   warp 0: TMA producer for X chunks (+ mirrors)      warp 6: TMA producer for dY chunks
//   warp 1: TMEM allocator + single-thread MMA issuer  warps 2..5: epilogue
#include <algorithm>

#include "sm100_ptx.cuh"
#include "srk_common.cuh"

namespace srk {

constexpr int kXSlots = 10;  // ring + mirror slots (16 KB each)
constexpr int kYRing = 3;
constexpr int kWgThreads = 224;
constexpr int kChunk = 128 * 128;  // bytes
constexpr int kYRows = 130;        // a dY tile: 128 rows + one neighbour row on either side
constexpr int kYSlot = 17 * 1024;  // bytes per dY slot (130 rows, 1024-byte aligned for the swizzle phase)
constexpr int kPartialFloats = 9 * 64 * 64 + 64;  // dW[9][64][64] then dbias[64]

struct alignas(64) WgradParams {
  CUtensorMap map_x;   // [rows_valid][64] box {64,128}
  CUtensorMap map_dy;  // [rows_valid][64] box {64,130}
  float* partial;      // workspace: [gridDim.x][kPartialFloats] per-CTA partial sums
  int Wp;
  int num_chunks;
  int nb;      // look-behind/ahead chunks: ceil((Wp+1)/128)
  int ring;    // X ring slots
  int mirror;  // mirrored leading slots: ceil((Wp+16)/128)
};

struct WgSmem {
  static constexpr int kOffX = 0;
  static constexpr int kOffY = kXSlots * kChunk;
  static constexpr int kOffOnes = kOffY + kYRing * kYSlot;
  static constexpr int kOffBars = kOffOnes + 2048;
  static constexpr int kNumBars = 2 * kXSlots + 2 * kYRing + 1;
  static constexpr int kOffTmemSlot = kOffBars + kNumBars * 8;
  static constexpr int kTotal = kOffTmemSlot + 16 + 1024;
};

// Body shared by the single-layer and the batched launch.  `mx` / `mdy` point into kernel-parameter (constant) space.
__device__ __forceinline__ void wgrad_body(const WgradParams& p, const CUtensorMap* mx, const CUtensorMap* mdy, float* part) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_x = s_base + WgSmem::kOffX;
  const uint32_t s_y = s_base + WgSmem::kOffY;
  const uint32_t s_ones = s_base + WgSmem::kOffOnes;
  const uint32_t s_bars = s_base + WgSmem::kOffBars;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + WgSmem::kOffTmemSlot);
  auto bar_xfull = [&](int s) { return s_bars + 8u * s; };
  auto bar_xempty = [&](int s) { return s_bars + 8u * (kXSlots + s); };
  auto bar_yfull = [&](int s) { return s_bars + 8u * (2 * kXSlots + s); };
  auto bar_yempty = [&](int s) { return s_bars + 8u * (2 * kXSlots + kYRing + s); };
  const uint32_t bar_done = s_bars + 8u * (2 * kXSlots + 2 * kYRing);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_begin = int((int64_t(blockIdx.x) * p.num_chunks) / gridDim.x);
  const int k_end = int((int64_t(blockIdx.x + 1) * p.num_chunks) / gridDim.x);
  const int nb = p.nb, R = p.ring;
  const int c0 = k_begin - nb, c_last = k_end - 1 + nb;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kXSlots; ++i) {
      mbar_init(bar_xfull(i), 1);
      mbar_init(bar_xempty(i), 1);
    }
    for (int i = 0; i < kYRing; ++i) {
      mbar_init(bar_yfull(i), 1);
      mbar_init(bar_yempty(i), 1);
    }
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  pdl_wait();
  pdl_launch_dependents();
  // 16 rows x 64 channels of bf16 1.0 (0x3F80): the swizzle of a constant block is the block itself
  for (int i = threadIdx.x; i < 2048 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + WgSmem::kOffOnes)[i] = 0x3F803F80u;
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc<512>(smem_u32(tmem_slot));
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(mx);
    tma_prefetch_desc(mdy);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (k_begin < k_end) {
    if (warp == 0) {
      if (lane == 0) {
        for (int c = c0; c <= c_last; ++c) {
          const int i = c - c0, slot = i % R, gen = i / R;
          mbar_wait(bar_xempty(slot), (gen & 1) ^ 1);
          const bool mir = slot < p.mirror;
          mbar_arrive_expect_tx(bar_xfull(slot), kChunk * (mir ? 2 : 1));
          tma_load_2d(s_x + slot * kChunk, mx, 0, c * 128, bar_xfull(slot));
          if (mir) tma_load_2d(s_x + (R + slot) * kChunk, mx, 0, c * 128, bar_xfull(slot));
        }
      }
    } else if (warp == 6) {
      if (lane == 0) {
        for (int k = k_begin; k < k_end; ++k) {
          const int j = k - k_begin, slot = j % kYRing, gen = j / kYRing;
          mbar_wait(bar_yempty(slot), (gen & 1) ^ 1);
          mbar_arrive_expect_tx(bar_yfull(slot), kYRows * 128);
          tma_load_2d(s_y + slot * kYSlot, mdy, 0, k * 128 - 1, bar_yfull(slot));  // rows 128k-1 .. 128k+128 (OOB rows: zero)
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, 192, 1, 1);
        constexpr uint64_t b_hi = umma_desc_hi(128 /* N atoms one row apart */, 1024, UMMA_LAYOUT_SW128);
        const int Wp = p.Wp;
        int loaded = c0 - 1;
        const int ring_rows = R * 128;
        // The single issuing thread is the critical path (every instruction pays its full dependent latency), so the
        // loop carries running window rows and pre-built descriptor halves instead of recomputing them with
        // divisions: a_row[] advances 16 rows per K-step and wraps by subtraction.
        //   a_row[0]: X window of tap row u=0 (its second M atom, u=1, sits Wp rows behind it);  a_row[1]: u=2
        int a_row[2] = {(nb * 128 - Wp) % ring_rows, (nb * 128 + Wp) % ring_rows};
        const uint64_t a_hi = umma_desc_hi(uint32_t(Wp) * 128u, 1024, UMMA_LAYOUT_SW128);
        const uint64_t ones_hi = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
        for (int k = k_begin; k < k_end; ++k) {
          const int j = k - k_begin;
          while (loaded < k + nb) {
            ++loaded;
            const int i = loaded - c0;
            mbar_wait(bar_xfull(i % R), (i / R) & 1);
          }
          mbar_wait(bar_yfull(j % kYRing), (j / kYRing) & 1);
          tc_fence_after();
          const uint32_t y_addr = s_y + (j % kYRing) * kYSlot;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            // B: dY rows (q0-1 .. q0+16], three windows one row apart = the taps v = 2, 1, 0
            const uint64_t bdesc = umma_desc(b_hi, y_addr + kk * 2048);
            const uint32_t acc_flag = (j | kk) != 0;
            const uint32_t a0 = s_x + uint32_t(a_row[0]) * 128u, a1 = s_x + uint32_t(a_row[1]) * 128u;
            umma_bf16(tmem, umma_desc(a_hi, a0), bdesc, idesc, acc_flag);
            const uint64_t adesc1 = ones_hi | (uint64_t(((s_ones - a1) >> 4) & 0x3FFF) << 16) | uint64_t((a1 >> 4) & 0x3FFF);
            umma_bf16(tmem + 192, adesc1, bdesc, idesc, acc_flag);
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              a_row[w] += 16;
              a_row[w] -= (a_row[w] >= ring_rows) ? ring_rows : 0;
            }
          }
          umma_commit(bar_xempty(j % R));  // X chunk k-nb (= c0+j) is done
          umma_commit(bar_yempty(j % kYRing));
        }
        umma_commit(bar_done);
      }
    } else {
      // ---------------------------------------------------------------- epilogue: TMEM -> this CTA's partial block
      const int quad = warp & 3;
      const int m = quad * 32 + lane;  // accumulator row
      mbar_wait(bar_done, 0);
      tc_fence_after();
      // accumulator 0: rows (u = m>>6, ci), columns (N atom j -> v = 2-j, co); accumulator 1: rows 0..63 = (u = 2, ci),
      // rows 64..127 = the ones block (row 64, atom 1 = unshifted dY: the bias gradient)
#pragma unroll 1
      for (int acc = 0; acc < 2; ++acc) {
        const int u = acc == 0 ? (m >> 6) : 2;
        const int ci = m & 63;
#pragma unroll 1
        for (int c = 0; c < 192; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem + acc * 192 + c + (uint32_t(quad * 32) << 16), r);
          tmem_ld_wait();
          const int v = 2 - (c >> 6), co0 = c & 63;
          float4* dst = nullptr;
          if (acc == 0 || m < 64) dst = reinterpret_cast<float4*>(part + (size_t(u * 3 + v) * 64 + ci) * 64 + co0);
          else if (m == 64 && v == 1) dst = reinterpret_cast<float4*>(part + 9 * 64 * 64 + co0);
          if (dst) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              dst[q] = make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                   __uint_as_float(r[4 * q + 3]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem);
}

__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  wgrad_body(p, &p.map_x, &p.map_dy, p.partial + size_t(blockIdx.x) * kPartialFloats);
}

// Batched form: blockIdx.y selects the layer (its own pair of tensor maps and workspace slice).  At the 64 x 41 x 41
// training shape a single layer only has ~880 chunks of work: twenty separate launches each pay their fill and drain;
// one launch over all layers keeps every SM busy for several waves of longer CTAs.
constexpr int kMaxWgBatch = 32;
struct alignas(64) WgradBatchParams {
  WgradParams c;  // geometry; c.partial = workspace base
  size_t layer_stride_floats;
  CUtensorMap map_x[kMaxWgBatch];
  CUtensorMap map_dy[kMaxWgBatch];
};
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_batched_kernel(const __grid_constant__ WgradBatchParams bp) {
  const int l = blockIdx.y;
  wgrad_body(bp.c, &bp.map_x[l], &bp.map_dy[l], bp.c.partial + size_t(l) * bp.layer_stride_floats + size_t(blockIdx.x) * kPartialFloats);
}

// Deterministic second pass: dw/dbias = (accumulate ? dw : 0) + sum over CTA partials (fixed order).
// blockIdx.y selects the layer: one launch folds the partial blocks of every layer of a model.  A layer may keep
// only the leading ci_n input / co_n output channels of the 64x64 block (first layer: ci_n = C, last layer:
// co_n = C, computed over zero-padded 64-channel operands); its destination is then the dense [9][ci_n][co_n].
struct WgradDst {
  float* dw;
  float* db;
  int ci_n, co_n;
};
__global__ void __launch_bounds__(128) wgrad_reduce_kernel(const float4* __restrict__ partial_base, size_t layer_stride_f4, int n_part,
                                                           const WgradDst* __restrict__ dsts, WgradDst single, int accumulate) {
  pdl_wait();
  pdl_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // float4 index
  constexpr int kN4 = kPartialFloats / 4;
  if (i >= kN4) return;
  const WgradDst d = dsts ? dsts[blockIdx.y] : single;
  // which output elements does this float4 cover?
  const bool is_bias = i >= 9 * 64 * 64 / 4;
  const int e0 = (is_bias ? i - 9 * 64 * 64 / 4 : i) * 4;
  const int co0 = e0 & 63, ci = (e0 >> 6) & 63, tap = e0 >> 12;
  if (co0 >= d.co_n || (!is_bias && ci >= d.ci_n)) return;
  const float4* partial = partial_base + size_t(blockIdx.y) * layer_stride_f4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int pidx = 0;
  for (; pidx + 8 <= n_part; pidx += 8) {
    float4 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldg(partial + size_t(pidx + j) * kN4 + i);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc.x += v[j].x;
      acc.y += v[j].y;
      acc.z += v[j].z;
      acc.w += v[j].w;
    }
  }
  for (; pidx < n_part; ++pidx) {
    const float4 a = __ldg(partial + size_t(pidx) * kN4 + i);
    acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
  }
  float* dst = is_bias ? d.db + co0 : d.dw + (size_t(tap) * d.ci_n + ci) * d.co_n + co0;
  const float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (co0 + q < d.co_n) dst[q] = accumulate ? dst[q] + r[q] : r[q];
}

static int wgrad_grid(srk_ctx* h, int num_chunks) {
  int g = (num_chunks + 7) / 8;  // >= 8 chunks of MMA work per partial block written
  if (g > h->num_sms) g = h->num_sms;
  if (g < 1) g = 1;
  return g;
}

}  // namespace srk

using namespace srk;

extern "C" size_t srk_conv_wgrad_tc_workspace_bytes(srk_handle_t h, int n_img, int H, int W) {
  if (!h) return 0;
  const FpaGeom g = fpa_geom(n_img, H, W);
  return size_t(wgrad_grid(h, int((g.rows_valid + 127) / 128))) * kPartialFloats * sizeof(float);
}

extern "C" int srk_conv_wgrad_tc(srk_handle_t h, const void* x_fpa, const void* dy_fpa, int n_img, int H, int W, float* dw_hwio,
                                 float* dbias, int accumulate, void* workspace, size_t workspace_bytes, srk_stream_t stream) {
  SRK_REQUIRE(h && x_fpa && dy_fpa && workspace && (dw_hwio == nullptr || dbias != nullptr), "srk_conv_wgrad_tc: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  SRK_REQUIRE((reinterpret_cast<uintptr_t>(dw_hwio) | reinterpret_cast<uintptr_t>(dbias) | reinterpret_cast<uintptr_t>(workspace)) % 16 == 0,
              "srk_conv_wgrad_tc: dw, dbias and workspace must be 16-byte aligned");
  const FpaGeom g = fpa_geom(n_img, H, W);
  SRK_REQUIRE(g.rows_valid < (int64_t(1) << 31), "srk_conv_wgrad_tc: too many rows");
  WgradParams p{};
  p.partial = static_cast<float*>(workspace);
  p.Wp = g.Wp;
  p.num_chunks = int((g.rows_valid + 127) / 128);
  p.nb = (g.Wp + 1 + 127) / 128;
  p.mirror = (g.Wp + 16 + 127) / 128;
  p.ring = kXSlots - p.mirror;
  SRK_REQUIRE(p.ring >= 2 * p.nb + 2, "srk_conv_wgrad_tc: image width %d too large for the flat-stream kernel", W);
  const int grid = wgrad_grid(h, p.num_chunks);
  SRK_REQUIRE(workspace_bytes >= size_t(grid) * kPartialFloats * sizeof(float), "srk_conv_wgrad_tc: workspace too small (%zu B)", workspace_bytes);
  if (int rc = make_tensor_map_2d(h, &p.map_x, x_fpa, uint64_t(g.rows_valid), 64, 128)) return rc;
  if (int rc = make_tensor_map_2d(h, &p.map_dy, dy_fpa, uint64_t(g.rows_valid), 64, kYRows)) return rc;
  if (first_use(h, reinterpret_cast<const void*>(&wgrad_tc_kernel)))
    SRK_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem::kTotal));
  SRK_CHECK_CUDA(launch_pdl(wgrad_tc_kernel, dim3(grid), dim3(kWgThreads), size_t(WgSmem::kTotal), as_stream(stream), p));
  if (dw_hwio) {
    WgradDst single{dw_hwio, dbias, 64, 64};
    SRK_CHECK_CUDA(launch_pdl(wgrad_reduce_kernel, dim3((kPartialFloats / 4 + 127) / 128, 1), dim3(128), 0, as_stream(stream),
                              static_cast<const float4*>(workspace), size_t(0), grid, static_cast<const WgradDst*>(nullptr), single, accumulate));
  }
  return 0;
}

extern "C" int srk_conv_wgrad_tc_batched(srk_handle_t h, const void* const* x_fpas, const void* const* dy_fpas, int n_layers, int n_img, int H,
                                         int W, void* workspace, size_t layer_stride_bytes, const srk_wgrad_dst* dsts_device, int accumulate,
                                         srk_stream_t stream) {
  SRK_REQUIRE(h && x_fpas && dy_fpas && workspace && dsts_device && n_layers > 0, "srk_conv_wgrad_tc_batched: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  SRK_REQUIRE(layer_stride_bytes % 16 == 0 && reinterpret_cast<uintptr_t>(workspace) % 16 == 0, "srk_conv_wgrad_tc_batched: workspace alignment");
  const FpaGeom g = fpa_geom(n_img, H, W);
  SRK_REQUIRE(g.rows_valid < (int64_t(1) << 31), "srk_conv_wgrad_tc_batched: too many rows");
  if (first_use(h, reinterpret_cast<const void*>(&wgrad_tc_batched_kernel)))
    SRK_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem::kTotal));
  const int num_chunks = int((g.rows_valid + 127) / 128);
  for (int l0 = 0; l0 < n_layers; l0 += kMaxWgBatch) {
    const int nl = std::min(kMaxWgBatch, n_layers - l0);
    WgradBatchParams bp{};
    bp.c.partial = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + size_t(l0) * layer_stride_bytes);
    bp.c.Wp = g.Wp;
    bp.c.num_chunks = num_chunks;
    bp.c.nb = (g.Wp + 1 + 127) / 128;
    bp.c.mirror = (g.Wp + 16 + 127) / 128;
    bp.c.ring = kXSlots - bp.c.mirror;
    SRK_REQUIRE(bp.c.ring >= 2 * bp.c.nb + 2, "srk_conv_wgrad_tc_batched: image width %d too large for the flat-stream kernel", W);
    bp.layer_stride_floats = layer_stride_bytes / sizeof(float);
    // CTAs per layer: ONE wave of the device over all layers (long-lived CTAs, no tail wave, few partial blocks to fold):
    // measured 75.6k VDSR patches/s against 71.5k with five waves and 68.0k with eight; at least 8 chunks per CTA
    int cpl = std::min(wgrad_grid(h, num_chunks), std::max(1, h->num_sms / nl));
    SRK_REQUIRE(layer_stride_bytes >= size_t(cpl) * kPartialFloats * sizeof(float), "srk_conv_wgrad_tc_batched: layer stride smaller than one layer's partials");
    for (int l = 0; l < nl; ++l) {
      SRK_REQUIRE(x_fpas[l0 + l] && dy_fpas[l0 + l], "srk_conv_wgrad_tc_batched: null operand for layer %d", l0 + l);
      if (int rc = make_tensor_map_2d(h, &bp.map_x[l], x_fpas[l0 + l], uint64_t(g.rows_valid), 64, 128)) return rc;
      if (int rc = make_tensor_map_2d(h, &bp.map_dy[l], dy_fpas[l0 + l], uint64_t(g.rows_valid), 64, kYRows)) return rc;
    }
    SRK_CHECK_CUDA(launch_pdl(wgrad_tc_batched_kernel, dim3(cpl, nl), dim3(kWgThreads), size_t(WgSmem::kTotal), as_stream(stream), bp));
    SRK_CHECK_CUDA(launch_pdl(wgrad_reduce_kernel, dim3((kPartialFloats / 4 + 127) / 128, nl), dim3(128), 0, as_stream(stream),
                              reinterpret_cast<const float4*>(bp.c.partial), size_t(layer_stride_bytes / 16), cpl,
                              reinterpret_cast<const WgradDst*>(dsts_device) + l0, WgradDst{}, accumulate));
  }
  return 0;
}

extern "C" int srk_wgrad_reduce_many(srk_handle_t h, const void* workspace_base, size_t layer_stride_bytes, int n_layers, int n_img, int H,
                                     int W, const srk_wgrad_dst* dsts_device, int accumulate, srk_stream_t stream) {
  SRK_REQUIRE(h && workspace_base && dsts_device && n_layers > 0, "srk_wgrad_reduce_many: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  SRK_REQUIRE(layer_stride_bytes % 16 == 0, "srk_wgrad_reduce_many: layer stride must be a multiple of 16 bytes");
  static_assert(sizeof(srk_wgrad_dst) == sizeof(WgradDst), "ABI struct mismatch");
  const FpaGeom g = fpa_geom(n_img, H, W);
  const int n_part = wgrad_grid(h, int((g.rows_valid + 127) / 128));
  SRK_REQUIRE(layer_stride_bytes >= size_t(n_part) * kPartialFloats * sizeof(float), "srk_wgrad_reduce_many: layer stride smaller than one layer's partials");
  SRK_CHECK_CUDA(launch_pdl(wgrad_reduce_kernel, dim3((kPartialFloats / 4 + 127) / 128, n_layers), dim3(128), 0, as_stream(stream),
                            static_cast<const float4*>(workspace_base), size_t(layer_stride_bytes / 16), n_part,
                            reinterpret_cast<const WgradDst*>(dsts_device), WgradDst{}, accumulate));
  return 0;
}

// Tensor-core convolution between flat-padded activations (FPA, include/srk.h) for sm_100a.
//
// A KxK stride-1 SAME convolution over an FPA is a sum of shifted GEMMs over ONE flat [rows][CIN] bf16
// matrix: tap (dy,dx) of output row p reads input row p + dy*Wp + dx, and the zero row/column baked into
// the layout supplies the padding.
//
// tcgen05.mma (M=128, SS mode, bf16) costs >= 60 cycles per instruction for any N <= 64 on B200
// (profiles/r1_mma_rate_vs_N.log), so one instruction per tap would cap the 64->64 layer at 53 % of the
// tensor pipe.  Instead ONE instruction covers a whole kernel ROW: for each dy the A window starts at
// p0 - h + dy*Wp (h = K/2) and the B operand is the K taps (dy, -h..h) stacked along N (N = K*NP, e.g. 192).
// Column block dx of the accumulator then holds the contribution of input row i to output row i - dx, so
// the epilogue adds the blocks with a LANE SHIFT:  y[j] = sum_dx D[j + dx][block dx]  (warp shuffles, edge
// lanes exchanged through shared memory).  Tiles therefore advance by 128-(K-1) rows and the K-1 edge
// lanes of each window are recomputed by the neighbouring tile (1.6 % redundancy for K=3).
//
// Each persistent CTA owns a contiguous range of tiles and streams the input through a shared-memory ring
// of 64-row chunks (one TMA load per chunk: every activation byte crosses L2->SMEM once); the A descriptors
// are row-shifted windows of that ring (probe-verified: SW128/SW64 descriptors accept any row shift with
// base_offset 0; the ring's first 128 rows are mirrored behind its last slot so a window never wraps).  The
// weights (K*K blocks of [NP][CIN] bf16, K-major) stay resident in shared memory.  Accumulators live in
// TMEM (2-4 stages) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
//   warp 0      : TMA producer (weights once, then the chunk ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer
//   warp 2      : TMA store issuer (waits for a full staging tile, stores it, frees the staging buffer)
//   warps 4..   : epilogue, 4 warps (one per TMEM lane quadrant) for every 16 output channels, so each
//                 SM sub-partition interleaves NP/16 epilogue warps and hides the tcgen05.ld / shuffle /
//                 shared-memory latencies (a single warp per sub-partition left the tensor pipe 80 % idle:
//                 profiles/r1_ncu_conv_tc_v2.txt): tcgen05.ld -> lane-shift add -> bias/activation/mask ->
//                 bf16 -> swizzled smem -> TMA store, or fp32 NHWC scatter with residual add / panel crop /
//                 pixel shuffle
#include "sm100_ptx.cuh"
#include "srk_common.cuh"

namespace srk {

enum { EPI_FPA = 0, EPI_NHWC = 1 };

constexpr int kChunkRows = 64;
constexpr int kRingSlots = 12;   // ring of 64-row chunks ...
constexpr int kMirrorSlots = 2;  // ... whose first 128 rows are duplicated behind the last slot

struct alignas(64) ConvTcParams {
  CUtensorMap map_in;   // [rows_valid][CIN]  box {CIN, 64}
  CUtensorMap map_w;    // [KS*KS*NP][CIN]    box {CIN, NP}
  CUtensorMap map_out;  // [rows_valid][NP]   box {NP, 128-(KS-1)}   (EPI_FPA)
  const float* bias;    // [NP] or null
  int n_img, H, W, Wp, S;
  int64_t rows_valid;
  int num_tiles;
  int act;
  // EPI_FPA extras
  const __nv_bfloat16* mask_src;
  int mask_kind;
  const __nv_bfloat16* addend_fpa;
  int relu_after_add;
  // EPI_NHWC extras
  float* out;
  const float* addend;
  const srk_panel* panels;
  int cout, shuffle_r, FH, FW;
};

template <int CIN, int NP, int KS>
struct ConvTcCfg {
  static constexpr int kRowBytes = CIN * 2;
  static constexpr int kChunkBytes = kChunkRows * kRowBytes;
  static constexpr int kTaps = KS * KS;
  static constexpr int kHalo = KS / 2;
  static constexpr int kTileStride = 128 - (KS - 1);  // output rows per tile
  static constexpr int kN = KS * NP;                  // MMA N: one kernel row of taps
  static constexpr int kAccStages = (kN > 128) ? 2 : 4;
  static constexpr int kTmemColsRaw = kAccStages * kN;
  static constexpr int kTmemCols = kTmemColsRaw <= 32 ? 32 : kTmemColsRaw <= 64 ? 64 : kTmemColsRaw <= 128 ? 128 : kTmemColsRaw <= 256 ? 256 : 512;
  static constexpr int kWTapBytes = NP * kRowBytes;
  static constexpr int kWBytes = ((kTaps * kWTapBytes + 1023) / 1024) * 1024;
  static constexpr int kRingBytes = (kRingSlots + kMirrorSlots) * kChunkBytes;
  static constexpr int kStageBytes = 2 * 128 * NP * 2;  // output staging (EPI_FPA), double buffered
  static constexpr int kGroups = 4;                 // epilogue column groups (4 warps each): 16 epilogue warps
  static constexpr int kColPass = NP / kGroups;     // accumulator columns per epilogue thread (16 / 8 / 4)
  static constexpr int kEpiThreads = 128 * kGroups;
  static constexpr int kThreads = 128 + kEpiThreads;
  static constexpr int kXchGroupFloats = 2 /*parity*/ * 4 /*quadrants*/ * (KS > 1 ? (KS - 1) * kHalo : 1) * kColPass;
  static constexpr int kXchFloats = kXchGroupFloats * kGroups;
  static constexpr int kOffW = 0;
  static constexpr int kOffRing = kWBytes;
  static constexpr int kOffStage = kOffRing + kRingBytes;
  static constexpr int kOffBias = kOffStage + kStageBytes;
  static constexpr int kOffTab = kOffBias + 256;    // NHWC epilogue: per-channel output offsets
  static constexpr int kOffXch = kOffTab + 256;
  static constexpr int kOffBars = kOffXch + ((kXchFloats * 4 + 15) / 16) * 16;
  static constexpr int kNumBars = 2 * kRingSlots + 1 + 2 * kAccStages + 4;
  static constexpr int kOffTmemSlot = kOffBars + kNumBars * 8;
  static constexpr int kTotal = kOffTmemSlot + 16 + 1024;  // + alignment slack
  static_assert(kN % 16 == 0 && kN <= 256, "invalid UMMA N");
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == SRK_ACT_RELU) return fmaxf(v, 0.f);
  if (act == SRK_ACT_TANH) {
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
  }
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// TMEM -> registers: kColPass fp32 columns of this thread's lane
template <int NCOL>
__device__ __forceinline__ void tmem_load_cols(uint32_t taddr, float (&v)[NCOL]) {
  static_assert(NCOL == 16 || NCOL == 8 || NCOL == 4, "column pass must be 4, 8 or 16");
  uint32_t u[NCOL];
  if constexpr (NCOL == 16) tmem_ld_32x32b_x16(taddr, u);
  else if constexpr (NCOL == 8) tmem_ld_32x32b_x8(taddr, u);
  else tmem_ld_32x32b_x4(taddr, u);
#pragma unroll
  for (int j = 0; j < NCOL; ++j) v[j] = __uint_as_float(u[j]);
}

// Shared-memory addresses the store warp and the epilogue warps work with.
struct EpiCtx {
  uint32_t tmem;
  uint32_t bar_tfull0, bar_tempty0;  // + 8 * accumulator stage
  uint32_t bar_sfull, bar_sfree;  // + 8 * staging buffer
  uint8_t* stage_ptr;
  int stage_stride;               // bytes between the two staging buffers
  float* s_bias;
  int* s_tab;
  float* s_xch;
};

// One thread: wait for a complete staging tile, TMA-store it, hand the staging buffer back.
template <int TS>
__device__ __forceinline__ void conv_store_loop(const ConvTcParams& p, const EpiCtx& e, int t_begin, int t_end) {
  for (int t = t_begin; t < t_end; ++t) {
    const int it = t - t_begin, sb = it & 1, sgen = it >> 1;
    mbar_wait(e.bar_sfull + 8u * sb, sgen & 1);
    tma_store_2d(&p.map_out, 0, TS * t, smem_u32(e.stage_ptr + sb * e.stage_stride));
    tma_store_commit();
    if (it >= 1) {  // the store of tile it-1 has finished READING its staging buffer -> writers of tile it+1 may reuse it
      tma_store_wait_read<1>();
      mbar_arrive(e.bar_sfree + 8u * (sb ^ 1));
    }
  }
  tma_store_wait_all<0>();
}

// Epilogue warp `ewarp` (0..15): TMEM lane quadrant ewarp & 3, column group ewarp >> 2.
template <int NP, int KS, int EPI, int ACC, int KN>
__device__ __forceinline__ void conv_epilogue(const ConvTcParams& p, const EpiCtx& e, int ewarp, int lane, int t_begin, int t_end) {
  constexpr int H_ = KS / 2, TS = 128 - (KS - 1), CP = NP / 4;
  constexpr int kXchGroupFloats = 2 * 4 * (KS > 1 ? (KS - 1) * H_ : 1) * CP;
  const uint32_t tmem = e.tmem;
  float* s_bias = e.s_bias;
  int* s_tab = e.s_tab;
  float* s_xch = e.s_xch;
  auto bar_tfull = [&](int a) { return e.bar_tfull0 + 8u * a; };
  auto bar_tempty = [&](int a) { return e.bar_tempty0 + 8u * a; };
  {
      const int quad = ewarp & 3;   // TMEM lane quadrant this warp may access (hardware: warp index % 4)
      const int grp = ewarp >> 2;   // column group: channels [CP*grp, CP*grp + CP)
      const int row = quad * 32 + lane;  // lane j of the accumulator <-> flat row TS*t - h + j
      const int col0 = grp * CP;
      const bool lane_valid = (row >= H_) && (row < H_ + TS);
      float* xg = s_xch + grp * kXchGroupFloats;
      constexpr int kXq = (KS > 1 ? (KS - 1) * H_ : 1) * CP;  // floats per (parity, quadrant)
      float bias_r[CP];
      int tab_r[CP];
#pragma unroll
      for (int c = 0; c < CP; ++c) {
        bias_r[c] = s_bias[col0 + c];
        tab_r[c] = (EPI == EPI_NHWC) ? s_tab[col0 + c] : 0;
      }
      int xpar = 0;
      // pixel coordinates of this lane's row in the first tile; lanes below the halo never hold a valid output
      const int H1 = p.H + 1;
      int px, pyy, pn;
      {
        const int64_t prow0 = int64_t(TS) * t_begin - H_ + row;
        const uint32_t pr = uint32_t(prow0 < 0 ? 0 : prow0);
        const uint32_t q = pr / uint32_t(p.Wp);
        px = int(pr - q * uint32_t(p.Wp));
        pn = int(q / uint32_t(H1));
        pyy = int(q - uint32_t(pn) * uint32_t(H1));
      }
      const int adv_x = TS % p.Wp, adv_q = TS / p.Wp;
      const int adv_y = adv_q % H1, adv_n = adv_q / H1;
      for (int t = t_begin; t < t_end; ++t) {
        const int it = t - t_begin, acc = it % ACC, accgen = it / ACC;
        mbar_wait(bar_tfull(acc), accgen & 1);
        tc_fence_after();
        const uint32_t taddr = tmem + acc * KN + col0 + (uint32_t(quad * 32) << 16);
        float blk[KS][CP];
#pragma unroll
        for (int b = 0; b < KS; ++b) tmem_load_cols<CP>(taddr + b * NP, blk[b]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(bar_tempty(acc));
        float v[CP];
        if constexpr (KS > 1) {
          // ---- lane-shift add of the KS column blocks: y[j] = sum_dx D[j + dx][block dx]
          // edge lanes publish what the neighbouring quadrants need: block dx<0 from the top lanes, dx>0 from the bottom lanes
          float* xq = xg + (xpar * 4 + quad) * kXq;
#pragma unroll
          for (int b = 0; b < KS; ++b) {
            const int dx = b - H_;
            const int bi = (dx < 0) ? b : b - 1;  // index among the KS-1 shifted blocks
            if (dx < 0 && lane >= 32 + dx) {
              float4* dst = reinterpret_cast<float4*>(xq + (bi * H_ + (lane - (32 + dx))) * CP);
#pragma unroll
              for (int c = 0; c < CP / 4; ++c) dst[c] = make_float4(blk[b][4 * c], blk[b][4 * c + 1], blk[b][4 * c + 2], blk[b][4 * c + 3]);
            }
            if (dx > 0 && lane < dx) {
              float4* dst = reinterpret_cast<float4*>(xq + (bi * H_ + lane) * CP);
#pragma unroll
              for (int c = 0; c < CP / 4; ++c) dst[c] = make_float4(blk[b][4 * c], blk[b][4 * c + 1], blk[b][4 * c + 2], blk[b][4 * c + 3]);
            }
          }
          named_bar_sync(1 + grp, 128);
          // lanes whose shuffle would wrap around the warp first take over the neighbouring quadrant's row
          // (their own value of that block is not needed any more), then ONE rotating shuffle per column
          // serves every lane -- no per-element select, and every output sees the same fp32 addition order
#pragma unroll
          for (int b = 0; b < KS; ++b) {
            const int dx = b - H_;
            if (dx == 0) continue;
            const int bi = (dx < 0) ? b : b - 1;
            const bool edge = (dx < 0) ? (lane >= 32 + dx) : (lane < dx);
            const int nq = (dx < 0) ? quad - 1 : quad + 1;
            if (edge && nq >= 0 && nq < 4) {
              const int li = (dx < 0) ? lane - (32 + dx) : lane;
              const float4* src = reinterpret_cast<const float4*>(xg + (xpar * 4 + nq) * kXq + (bi * H_ + li) * CP);
#pragma unroll
              for (int c = 0; c < CP / 4; ++c) {
                const float4 o = src[c];
                blk[b][4 * c] = o.x;
                blk[b][4 * c + 1] = o.y;
                blk[b][4 * c + 2] = o.z;
                blk[b][4 * c + 3] = o.w;
              }
            }
          }
#pragma unroll
          for (int c = 0; c < CP; ++c) v[c] = blk[H_][c];
#pragma unroll
          for (int b = 0; b < KS; ++b) {
            const int dx = b - H_;
            if (dx == 0) continue;
            const int src_lane = (lane + dx) & 31;  // y[j] += D[j + dx][block dx]
#pragma unroll
            for (int c = 0; c < CP; ++c) v[c] += __shfl_sync(0xffffffffu, blk[b][c], src_lane);
          }
          xpar ^= 1;
        } else {
#pragma unroll
          for (int c = 0; c < CP; ++c) v[c] = blk[0][c];
        }

        // the pixel this lane holds (decoded once per CTA, then advanced by TS rows per tile)
        const int64_t prow = int64_t(TS) * t - H_ + row;
        const bool valid_px = lane_valid && prow < p.rows_valid && (px < p.W) && (pyy > 0);
        bool valid = valid_px;
        const int n = pn, y = pyy - 1, x = px;
        {
          px += adv_x;
          const int cx = px >= p.Wp;
          px -= cx ? p.Wp : 0;
          pyy += adv_y + cx;
          const int cy = pyy >= H1;
          pyy -= cy ? H1 : 0;
          pn += adv_n + cy;
        }

        if constexpr (EPI == EPI_FPA) {
          static_assert(CP >= 8, "FPA epilogue stores 16-byte chunks");
          uint32_t packed[CP / 2];
          if (valid) {
#pragma unroll
            for (int c = 0; c < CP; ++c) v[c] = act_apply(v[c] + bias_r[c], p.act);
            if (p.mask_src) {
              const uint4* m = reinterpret_cast<const uint4*>(p.mask_src + size_t(prow) * NP + col0);
#pragma unroll
              for (int j = 0; j < CP / 8; ++j) {
                const uint4 mv = __ldg(m + j);
                const uint32_t w4[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
                  if (p.mask_kind == SRK_ACT_RELU) {
                    v[j * 8 + e * 2] = f.x > 0.f ? v[j * 8 + e * 2] : 0.f;
                    v[j * 8 + e * 2 + 1] = f.y > 0.f ? v[j * 8 + e * 2 + 1] : 0.f;
                  } else {
                    v[j * 8 + e * 2] *= (1.f - f.x * f.x);
                    v[j * 8 + e * 2 + 1] *= (1.f - f.y * f.y);
                  }
                }
              }
            }
            if (p.addend_fpa) {
              const uint4* a = reinterpret_cast<const uint4*>(p.addend_fpa + size_t(prow) * NP + col0);
#pragma unroll
              for (int j = 0; j < CP / 8; ++j) {
                const uint4 av = __ldg(a + j);
                const uint32_t w4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
                  v[j * 8 + e * 2] += f.x;
                  v[j * 8 + e * 2 + 1] += f.y;
                }
              }
              if (p.relu_after_add) {
#pragma unroll
                for (int c = 0; c < CP; ++c) v[c] = fmaxf(v[c], 0.f);
              }
            }
#pragma unroll
            for (int c = 0; c < CP / 2; ++c) packed[c] = pack_bf16x2(v[2 * c], v[2 * c + 1]);
          } else {
#pragma unroll
            for (int c = 0; c < CP / 2; ++c) packed[c] = 0u;  // pad rows/columns stay exactly zero
          }
          // staging buffer free? (the store of tile it-2, which used the same buffer, has finished reading it)
          const int sb = it & 1, sgen = it >> 1;
          uint8_t* stage_ptr = e.stage_ptr + sb * e.stage_stride;
          mbar_wait(e.bar_sfree + 8u * sb, (sgen & 1) ^ 1);
          if (lane_valid) {
            constexpr int kOutRowBytes = NP * 2;
            const int srow = row - H_;  // staging row = output row within the tile
            const int sw = (NP == 64) ? (srow & 7) : ((srow >> 1) & 3);
#pragma unroll
            for (int j = 0; j < CP / 8; ++j) {
              const int chunk = grp * (CP / 8) + j;  // 16-byte chunk index within the output row
              uint4 q4 = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
              *reinterpret_cast<uint4*>(stage_ptr + srow * kOutRowBytes + ((chunk ^ sw) << 4)) = q4;
            }
          }
          fence_proxy_async_smem();
          mbar_arrive(e.bar_sfull + 8u * sb);
        } else {
          // fp32 NHWC scatter: residual add, panel crop, depth_to_space
          if (valid) {
            int fn = n, fy = y, fx = x;
            if (p.panels) {
              const srk_panel e = p.panels[n];
              valid = (y >= e.own_y0) && (y < e.own_y1) && (x >= e.own_x0) && (x < e.own_x1);
              fn = e.frame;
              fy = e.y0 + y;
              fx = e.x0 + x;
            }
            if (valid) {
              const int r = p.shuffle_r, C = p.cout / (r * r);
              const int64_t OW = int64_t(p.FW) * r;
              const int64_t base = ((int64_t(fn) * p.FH * r + int64_t(fy) * r) * OW + int64_t(fx) * r) * C;
#pragma unroll
              for (int c = 0; c < CP; ++c) {
                const int off = tab_r[c];
                if (off >= 0) {
                  const int64_t idx = base + off;
                  float o = act_apply(v[c] + bias_r[c], p.act);
                  if (p.addend) o += __ldg(p.addend + idx);
                  p.out[idx] = o;
                }
              }
            }
          }
        }
      }
  }
}

template <int CIN, int NP, int KS, int EPI>
__global__ void __launch_bounds__(ConvTcCfg<CIN, NP, KS>::kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  using L = ConvTcCfg<CIN, NP, KS>;
  constexpr uint32_t kLayout = (CIN == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
  constexpr uint32_t kSbo = 8 * L::kRowBytes;
  constexpr int H_ = L::kHalo, TS = L::kTileStride, CP = L::kColPass, ACC = L::kAccStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_w = s_base + L::kOffW;
  const uint32_t s_ring = s_base + L::kOffRing;
  uint8_t* stage_ptr = smem + L::kOffStage;
  float* s_bias = reinterpret_cast<float*>(smem + L::kOffBias);
  int* s_tab = reinterpret_cast<int*>(smem + L::kOffTab);
  float* s_xch = reinterpret_cast<float*>(smem + L::kOffXch);
  const uint32_t s_bars = s_base + L::kOffBars;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmemSlot);
  auto bar_full = [&](int s) { return s_bars + 8u * s; };
  auto bar_empty = [&](int s) { return s_bars + 8u * (kRingSlots + s); };
  const uint32_t bar_wfull = s_bars + 8u * (2 * kRingSlots);
  auto bar_tfull = [&](int a) { return s_bars + 8u * (2 * kRingSlots + 1 + a); };
  auto bar_tempty = [&](int a) { return s_bars + 8u * (2 * kRingSlots + 1 + ACC + a); };
  const uint32_t bar_sfull = s_bars + 8u * (2 * kRingSlots + 1 + 2 * ACC);      // [2] staging tile complete
  const uint32_t bar_sfree = s_bars + 8u * (2 * kRingSlots + 1 + 2 * ACC + 2);  // [2] staging buffer reusable

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // contiguous tile range of this CTA; tile t produces output rows [TS*t, TS*t + TS)
  const int t_begin = int((int64_t(blockIdx.x) * p.num_tiles) / gridDim.x);
  const int t_end = int((int64_t(blockIdx.x + 1) * p.num_tiles) / gridDim.x);
  const int reach = H_ * p.Wp;  // rows of vertical reach
  auto lo_chunk = [&](int t) { return floor_div(TS * t - H_ - reach, kChunkRows); };
  auto hi_chunk = [&](int t) { return floor_div(TS * t - H_ + reach + 127, kChunkRows); };
  const int c0 = lo_chunk(t_begin);  // first chunk this CTA loads (may be negative: TMA zero-fills)
  const int c_last = hi_chunk(t_end - 1);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRingSlots; ++i) {
      mbar_init(bar_full(i), 1);
      mbar_init(bar_empty(i), 1);
    }
    mbar_init(bar_wfull, 1);
    for (int i = 0; i < ACC; ++i) {
      mbar_init(bar_tfull(i), 1);
      mbar_init(bar_tempty(i), L::kEpiThreads);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_sfull + 8u * i, L::kEpiThreads);
      mbar_init(bar_sfree + 8u * i, 1);
    }
    fence_mbar_init();
  }
  if (threadIdx.x < NP) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  if (EPI == EPI_NHWC && threadIdx.x < NP) {
    // output offset of packed channel c relative to pixel (Y*r, X*r, 0): depth_to_space index, -1 = padding channel
    const int c = threadIdx.x, r = p.shuffle_r, C = p.cout / (r * r);
    int off = -1;
    if (c < p.cout) {
      const int ch = c % C, sub = c / C, ddy = sub / r, ddx = sub - ddy * r;
      off = (ddy * p.FW * r + ddx) * C + ch;
    }
    s_tab[c] = off;
  }
  if (warp == 1) tmem_alloc<L::kTmemCols>(smem_u32(tmem_slot));
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_in);
    tma_prefetch_desc(&p.map_w);
    if (EPI == EPI_FPA) tma_prefetch_desc(&p.map_out);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  EpiCtx ec;
  ec.tmem = tmem;
  ec.bar_tfull0 = bar_tfull(0);
  ec.bar_tempty0 = bar_tempty(0);
  ec.bar_sfull = bar_sfull;
  ec.bar_sfree = bar_sfree;
  ec.stage_ptr = stage_ptr;
  ec.stage_stride = L::kStageBytes / 2;
  ec.s_bias = s_bias;
  ec.s_tab = s_tab;
  ec.s_xch = s_xch;

  if (t_begin < t_end) {
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer
      if (lane == 0) {
        mbar_arrive_expect_tx(bar_wfull, L::kTaps * L::kWTapBytes);
        for (int tap = 0; tap < L::kTaps; ++tap) tma_load_2d(s_w + tap * L::kWTapBytes, &p.map_w, 0, tap * NP, bar_wfull);
        for (int c = c0; c <= c_last; ++c) {
          const int i = c - c0, slot = i % kRingSlots, gen = i / kRingSlots;
          mbar_wait(bar_empty(slot), (gen & 1) ^ 1);
          const bool mir = slot < kMirrorSlots;
          mbar_arrive_expect_tx(bar_full(slot), L::kChunkBytes * (mir ? 2 : 1));
          tma_load_2d(s_ring + slot * L::kChunkBytes, &p.map_in, 0, c * kChunkRows, bar_full(slot));
          if (mir) tma_load_2d(s_ring + (kRingSlots + slot) * L::kChunkBytes, &p.map_in, 0, c * kChunkRows, bar_full(slot));
        }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer (one thread)
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, L::kN, 0, 0);
        constexpr uint64_t hi = umma_desc_hi(0, kSbo, kLayout);
        mbar_wait(bar_wfull, 0);
        int loaded = c0 - 1, released = c0;  // chunks <= loaded have landed; chunks < released were handed back
        for (int t = t_begin; t < t_end; ++t) {
          const int it = t - t_begin, acc = it % ACC, accgen = it / ACC;
          const int need = hi_chunk(t);
          while (loaded < need) {
            ++loaded;
            const int i = loaded - c0;
            mbar_wait(bar_full(i % kRingSlots), (i / kRingSlots) & 1);
          }
          mbar_wait(bar_tempty(acc), (accgen & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem + acc * L::kN;
          const int row0 = TS * t - H_ - c0 * kChunkRows;  // window start of dy = 0, relative to the ring origin
#pragma unroll
          for (int r = 0; r < KS; ++r) {
            const int rr = (row0 + (r - H_) * p.Wp) % (kRingSlots * kChunkRows);
            const uint32_t a_addr = s_ring + rr * L::kRowBytes;
            const uint32_t b_addr = s_w + r * KS * L::kWTapBytes;
#pragma unroll
            for (int k = 0; k < CIN / 16; ++k)
              umma_bf16(d_tmem, umma_desc(hi, a_addr + k * 32), umma_desc(hi, b_addr + k * 32), idesc, (r | k) != 0);
          }
          umma_commit(bar_tfull(acc));
          // hand back the chunks no later tile needs
          const int keep_from = (t + 1 < t_end) ? lo_chunk(t + 1) : released;
          while (released < keep_from) {
            umma_commit(bar_empty((released - c0) % kRingSlots));
            ++released;
          }
        }
      }
    } else if (warp == 2) {
      // ------------------------------------------------------------------ TMA store issuer (one thread)
      if (EPI == EPI_FPA && lane == 0) conv_store_loop<TS>(p, ec, t_begin, t_end);
    } else if (warp >= 4) {
      // ------------------------------------------------------------------ epilogue (128 threads per column group)
      conv_epilogue<NP, KS, EPI, ACC, L::kN>(p, ec, warp - 4, lane, t_begin, t_end);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<L::kTmemCols>(tmem);
}

// ======================================================================================= first layer on tensor cores
// Small-Cin first layer (KSxKS, CIN in {1,3} -> 64) as ONE GEMM per 128-pixel tile: K = KS*KS*CIN (padded to a
// multiple of 16), A = im2col rows gathered from the fp32 NHWC frame by four producer warps (one pixel per
// thread, bf16, written straight into the SW128 K-major layout the MMA reads), B = the packed kernel resident
// in smem.  It is an HBM-bound layer (reads 4*CIN B, writes 128 B per pixel); the tensor core only removes the
// 2*K*64 FMAs per pixel that made the CUDA-core version (conv_first_kernel) compute-bound.  Epilogue and store
// path are shared with conv_tc_kernel (KS=1 form: no lane shift).
struct alignas(64) ConvGatherParams {
  ConvTcParams tc;  // map_w, map_out, bias, geometry, act, mask_src/mask_kind, panels
  const float* x;   // fp32 NHWC frames
  int FH, FW;       // frame dims
  int Hin, Win;     // panel-local input window
  int po;           // pad offset: KS/2 (SAME) or 0 (VALID)
};

template <int KS, int CIN>
struct ConvGatherCfg {
  static constexpr int kKT = KS * KS * CIN;
  static constexpr int kKP = (kKT + 15) / 16 * 16;
  static constexpr int kBlocks = (kKP + 63) / 64;       // 64-element K blocks (128-byte rows)
  static constexpr int kABytes = kBlocks * 128 * 128;    // one A stage
  static constexpr int kAStages = kBlocks <= 2 ? 4 : 2;
  static constexpr int kWBytes = kBlocks * 64 * 128;
  static constexpr int kAcc = 4;
  static constexpr int kEpiThreads = 512, kGatherThreads = 128, kGatherGroups = 2;  // each group gathers every other tile
  static constexpr int kThreads = 128 + kEpiThreads + kGatherGroups * kGatherThreads;
  static constexpr int kOffW = 0;
  static constexpr int kOffA = kWBytes;
  static constexpr int kOffStage = kOffA + kAStages * kABytes;
  static constexpr int kOffBias = kOffStage + 2 * 128 * 64 * 2;
  static constexpr int kOffTab = kOffBias + 256;
  static constexpr int kOffXch = kOffTab + 256;
  static constexpr int kOffBars = kOffXch + 64;
  static constexpr int kNumBars = 2 * kAStages + 1 + 2 * kAcc + 4;
  static constexpr int kOffTmemSlot = kOffBars + kNumBars * 8;
  static constexpr int kTotal = kOffTmemSlot + 16 + 1024;
};

template <int KS, int CIN>
__global__ void __launch_bounds__(ConvGatherCfg<KS, CIN>::kThreads, 1) conv_gather_tc_kernel(const __grid_constant__ ConvGatherParams gp) {
  using L = ConvGatherCfg<KS, CIN>;
  constexpr int ACC = L::kAcc, AST = L::kAStages;
  const ConvTcParams& p = gp.tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_w = s_base + L::kOffW;
  uint8_t* a_ptr = smem + L::kOffA;
  const uint32_t s_a = s_base + L::kOffA;
  float* s_bias = reinterpret_cast<float*>(smem + L::kOffBias);
  const uint32_t s_bars = s_base + L::kOffBars;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmemSlot);
  auto bar_afull = [&](int s) { return s_bars + 8u * s; };
  auto bar_aempty = [&](int s) { return s_bars + 8u * (AST + s); };
  const uint32_t bar_wfull = s_bars + 8u * (2 * AST);
  auto bar_tfull = [&](int a) { return s_bars + 8u * (2 * AST + 1 + a); };
  auto bar_tempty = [&](int a) { return s_bars + 8u * (2 * AST + 1 + ACC + a); };
  const uint32_t bar_sfull = s_bars + 8u * (2 * AST + 1 + 2 * ACC);
  const uint32_t bar_sfree = s_bars + 8u * (2 * AST + 1 + 2 * ACC + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t_begin = int((int64_t(blockIdx.x) * p.num_tiles) / gridDim.x);
  const int t_end = int((int64_t(blockIdx.x + 1) * p.num_tiles) / gridDim.x);

  if (threadIdx.x == 0) {
    for (int i = 0; i < AST; ++i) {
      mbar_init(bar_afull(i), L::kGatherThreads);
      mbar_init(bar_aempty(i), 1);
    }
    mbar_init(bar_wfull, 1);
    for (int i = 0; i < ACC; ++i) {
      mbar_init(bar_tfull(i), 1);
      mbar_init(bar_tempty(i), L::kEpiThreads);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_sfull + 8u * i, L::kEpiThreads);
      mbar_init(bar_sfree + 8u * i, 1);
    }
    fence_mbar_init();
  }
  if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  if (warp == 1) tmem_alloc<256>(smem_u32(tmem_slot));
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_w);
    tma_prefetch_desc(&p.map_out);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  EpiCtx ec;
  ec.tmem = tmem;
  ec.bar_tfull0 = bar_tfull(0);
  ec.bar_tempty0 = bar_tempty(0);
  ec.bar_sfull = bar_sfull;
  ec.bar_sfree = bar_sfree;
  ec.stage_ptr = smem + L::kOffStage;
  ec.stage_stride = 128 * 64 * 2;
  ec.s_bias = s_bias;
  ec.s_tab = reinterpret_cast<int*>(smem + L::kOffTab);
  ec.s_xch = reinterpret_cast<float*>(smem + L::kOffXch);

  if (t_begin < t_end) {
    if (warp == 0) {
      if (lane == 0) {  // the packed kernel: kBlocks blocks of [64 co][64 k] bf16
        mbar_arrive_expect_tx(bar_wfull, L::kWBytes);
        for (int b = 0; b < L::kBlocks; ++b) tma_load_2d(s_w + b * 64 * 128, &p.map_w, 0, b * 64, bar_wfull);
      }
    } else if (warp == 1) {
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
        constexpr uint64_t hi = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
        mbar_wait(bar_wfull, 0);
        for (int t = t_begin; t < t_end; ++t) {
          const int it = t - t_begin, acc = it % ACC, st = it % AST;
          mbar_wait(bar_afull(st), (it / AST) & 1);
          mbar_wait(bar_tempty(acc), ((it / ACC) & 1) ^ 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < L::kKP / 16; ++k) {
            const uint32_t a_addr = s_a + st * L::kABytes + (k / 4) * (128 * 128) + (k % 4) * 32;
            const uint32_t b_addr = s_w + (k / 4) * (64 * 128) + (k % 4) * 32;
            umma_bf16(tmem + acc * 64, umma_desc(hi, a_addr), umma_desc(hi, b_addr), idesc, k != 0);
          }
          umma_commit(bar_tfull(acc));
          umma_commit(bar_aempty(st));
        }
      }
    } else if (warp == 2) {
      if (lane == 0) conv_store_loop<128>(p, ec, t_begin, t_end);
    } else if (warp >= 4 && warp < 20) {
      conv_epilogue<64, 1, EPI_FPA, ACC, 64>(p, ec, warp - 4, lane, t_begin, t_end);
    } else if (warp >= 20) {
      // ------------------------------------------------------------------ im2col gather: one pixel row per thread
      // Two groups of 128 threads take alternate tiles so two tiles' worth of loads are in flight.  Loads are
      // unconditional from clamped coordinates and zeroed by a validity select afterwards, so the compiler can
      // issue a whole chunk's loads back to back.
      const int gg = (threadIdx.x - 640) >> 7;   // gather group
      const int r = (threadIdx.x - 640) & 127;   // row of the A tile
      const int H1 = p.H + 1;
      for (int t = t_begin + gg; t < t_end; t += L::kGatherGroups) {
        const int it = t - t_begin, st = it % AST;
        const int64_t prow = int64_t(128) * t + r;
        const uint32_t pr = uint32_t(prow);
        const uint32_t q = pr / uint32_t(p.Wp);
        const int px = int(pr - q * uint32_t(p.Wp));
        const int pn = int(q / uint32_t(H1));
        const int pyy = int(q - uint32_t(pn) * uint32_t(H1));
        const bool valid = prow < p.rows_valid && px < p.W && pyy > 0;
        mbar_wait(bar_aempty(st), ((it / AST) & 1) ^ 1);
        if (valid) {
          int fn = pn, y0 = 0, x0 = 0;
          if (p.panels) {
            const srk_panel e = p.panels[pn];
            fn = e.frame;
            y0 = e.y0;
            x0 = e.x0;
          }
          const int y = pyy - 1 - gp.po, x = px - gp.po;  // top-left source pixel (panel-local)
          const float* frame = gp.x + (int64_t(fn) * gp.FH + y0) * gp.FW * CIN + int64_t(x0) * CIN;
          uint8_t* arow = a_ptr + st * L::kABytes + r * 128;
          // per-tap row/column validity and clamped offsets (KS each)
          int rofs[KS], cofs[KS];
          bool rok[KS], cok[KS];
#pragma unroll
          for (int u = 0; u < KS; ++u) {
            const int sy = y + u, sx = x + u;
            rok[u] = sy >= 0 && sy < gp.Hin;
            cok[u] = sx >= 0 && sx < gp.Win;
            rofs[u] = min(max(sy, 0), gp.Hin - 1) * gp.FW * CIN;
            cofs[u] = min(max(sx, 0), gp.Win - 1) * CIN;
          }
#pragma unroll
          for (int c8 = 0; c8 < L::kKP / 8; ++c8) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = c8 * 8 + j;
              f[j] = 0.f;
              if (k < L::kKT) {
                const int tap = k / CIN, ci = k % CIN, u = tap / KS, v = tap % KS;
                const float val = __ldg(frame + rofs[u] + cofs[v] + ci);
                f[j] = (rok[u] && cok[v]) ? val : 0.f;
              }
            }
            const uint4 q4 = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
            *reinterpret_cast<uint4*>(arow + (c8 / 8) * (128 * 128) + (((c8 % 8) ^ (r & 7)) << 4)) = q4;
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_afull(st));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem);
}

template <int KS, int CIN>
static int launch_conv_gather(srk_ctx* h, ConvGatherParams& gp, const void* w_packed, void* y_fpa, cudaStream_t stream) {
  using L = ConvGatherCfg<KS, CIN>;
  static bool attr_set = false;
  SRK_REQUIRE(L::kTotal <= h->smem_optin, "conv_first_tc: needs %d B smem, device allows %d", L::kTotal, h->smem_optin);
  if (!attr_set) {
    SRK_CHECK_CUDA(cudaFuncSetAttribute(conv_gather_tc_kernel<KS, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  ConvTcParams& p = gp.tc;
  p.num_tiles = int((p.rows_valid + 127) / 128);
  if (int rc = make_tensor_map_2d(&p.map_w, w_packed, uint64_t(L::kBlocks * 64), 64, 64)) return rc;
  if (int rc = make_tensor_map_2d(&p.map_out, y_fpa, uint64_t(p.rows_valid), 64, 128)) return rc;
  const int grid = p.num_tiles < h->num_sms ? p.num_tiles : h->num_sms;
  conv_gather_tc_kernel<KS, CIN><<<grid, L::kThreads, L::kTotal, stream>>>(gp);
  SRK_LAUNCH_CHECK();
  return 0;
}

// --------------------------------------------------------------------------------------- host side
template <int CIN, int NP, int KS, int EPI>
static int launch_conv_tc(srk_ctx* h, ConvTcParams& p, const void* x, const void* w_packed, void* y_fpa, cudaStream_t stream) {
  using L = ConvTcCfg<CIN, NP, KS>;
  static bool attr_set = false;
  SRK_REQUIRE(L::kTotal <= h->smem_optin, "conv_tc: needs %d B smem, device allows %d", L::kTotal, h->smem_optin);
  if (!attr_set) {
    SRK_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<CIN, NP, KS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  p.num_tiles = int((p.rows_valid + L::kTileStride - 1) / L::kTileStride);
  const int span_chunks = (2 * L::kHalo * p.Wp + 128 + kChunkRows - 1) / kChunkRows + 1;
  SRK_REQUIRE(span_chunks <= kRingSlots - 2,
              "conv_tc: image width %d too large for the %dx%d flat-stream kernel (a tile needs %d chunks, ring holds %d); "
              "split the frame into column panels", p.W, KS, KS, span_chunks, kRingSlots);
  if (int rc = make_tensor_map_2d(&p.map_in, x, uint64_t(p.rows_valid), CIN, kChunkRows)) return rc;
  if (int rc = make_tensor_map_2d(&p.map_w, w_packed, uint64_t(KS * KS * NP), CIN, NP)) return rc;
  if (EPI == EPI_FPA) {
    if (int rc = make_tensor_map_2d(&p.map_out, y_fpa, uint64_t(p.rows_valid), NP, L::kTileStride)) return rc;
  }
  const int grid = p.num_tiles < h->num_sms ? p.num_tiles : h->num_sms;
  conv_tc_kernel<CIN, NP, KS, EPI><<<grid, L::kThreads, L::kTotal, stream>>>(p);
  SRK_LAUNCH_CHECK();
  return 0;
}

static int fill_geom(ConvTcParams& p, int n_img, int H, int W) {
  SRK_REQUIRE(n_img > 0 && H > 0 && W > 0, "conv_tc: bad geometry n_img=%d H=%d W=%d", n_img, H, W);
  const FpaGeom g = fpa_geom(n_img, H, W);
  SRK_REQUIRE(g.rows_valid < (int64_t(1) << 30), "conv_tc: %lld rows exceed the 2^30 row limit", (long long)g.rows_valid);
  p.n_img = n_img;
  p.H = H;
  p.W = W;
  p.Wp = g.Wp;
  p.S = g.S;
  p.rows_valid = g.rows_valid;
  return 0;
}

}  // namespace srk

using namespace srk;

extern "C" int srk_conv_tc(srk_handle_t h, const void* x_fpa, int cin_p, const void* w_packed, const float* bias, int k,
                           int cout_p, int act, int n_img, int H, int W, void* y_fpa, const void* mask_src,
                           int mask_kind, const void* addend_fpa, int relu_after_add, srk_stream_t stream) {
  SRK_REQUIRE(h && x_fpa && w_packed && y_fpa, "srk_conv_tc: null argument");
  ConvTcParams p{};
  if (int rc = fill_geom(p, n_img, H, W)) return rc;
  p.bias = bias;
  p.act = act;
  p.mask_src = static_cast<const __nv_bfloat16*>(mask_src);
  p.mask_kind = mask_kind;
  p.addend_fpa = static_cast<const __nv_bfloat16*>(addend_fpa);
  p.relu_after_add = relu_after_add;
  cudaStream_t s = as_stream(stream);
#define SRK_CASE(CIN, NP, KS) \
  if (cin_p == CIN && cout_p == NP && k == KS) return launch_conv_tc<CIN, NP, KS, EPI_FPA>(h, p, x_fpa, w_packed, y_fpa, s);
  SRK_CASE(64, 64, 3)
  SRK_CASE(64, 32, 3)
  SRK_CASE(64, 64, 1)
  SRK_CASE(64, 32, 1)
  SRK_CASE(32, 64, 3)
  SRK_CASE(32, 64, 1)
  SRK_CASE(32, 32, 3)
#undef SRK_CASE
  set_error("srk_conv_tc: unsupported (cin_p=%d, cout_p=%d, k=%d)", cin_p, cout_p, k);
  return -1;
}

extern "C" int srk_conv_tc_last(srk_handle_t h, const void* x_fpa, int cin_p, const void* w_packed, const float* bias,
                                int k, int cout, int cout_p, int act, int n_img, int H, int W, const srk_panel* panels,
                                int n_frames, int FH, int FW, int shuffle_r, const float* addend, float* out,
                                srk_stream_t stream) {
  SRK_REQUIRE(h && x_fpa && w_packed && out, "srk_conv_tc_last: null argument");
  SRK_REQUIRE(shuffle_r >= 1 && cout % (shuffle_r * shuffle_r) == 0, "srk_conv_tc_last: cout %d not divisible by r^2 (r=%d)", cout, shuffle_r);
  SRK_REQUIRE(cout <= cout_p, "srk_conv_tc_last: cout %d > cout_p %d", cout, cout_p);
  SRK_REQUIRE(panels || (n_frames == n_img && FH == H && FW == W), "srk_conv_tc_last: without panels the frame must equal the FPA geometry");
  ConvTcParams p{};
  if (int rc = fill_geom(p, n_img, H, W)) return rc;
  p.bias = bias;
  p.act = act;
  p.out = out;
  p.addend = addend;
  p.panels = panels;
  p.cout = cout;
  p.shuffle_r = shuffle_r;
  p.FH = FH;
  p.FW = FW;
  cudaStream_t s = as_stream(stream);
#define SRK_CASE(CIN, NP, KS) \
  if (cin_p == CIN && cout_p == NP && k == KS) return launch_conv_tc<CIN, NP, KS, EPI_NHWC>(h, p, x_fpa, w_packed, nullptr, s);
  SRK_CASE(64, 16, 3)
  SRK_CASE(32, 16, 3)
  SRK_CASE(32, 32, 3)
  SRK_CASE(32, 16, 5)
  SRK_CASE(32, 64, 3)
  SRK_CASE(64, 32, 3)
#undef SRK_CASE
  set_error("srk_conv_tc_last: unsupported (cin_p=%d, cout_p=%d, k=%d)", cin_p, cout_p, k);
  return -1;
}

extern "C" int srk_conv_first_tc(srk_handle_t h, const float* x, int n_frames, int FH, int FW, int cin, const void* w_packed,
                                 const float* bias, int k, int pad_mode, int act, const srk_panel* panels, int n_img, int H, int W,
                                 void* y_fpa, const void* mask_src, int mask_kind, srk_stream_t stream) {
  SRK_REQUIRE(h && x && w_packed && y_fpa, "srk_conv_first_tc: null argument");
  const int halo = (pad_mode == SRK_PAD_VALID) ? k - 1 : 0;
  SRK_REQUIRE(panels || (n_frames == n_img && FH == H + halo && FW == W + halo),
              "srk_conv_first_tc: without panels the frame (%dx%d) must match the output geometry (%dx%d, k=%d)", FH, FW, H, W, k);
  ConvGatherParams gp{};
  if (int rc = fill_geom(gp.tc, n_img, H, W)) return rc;
  gp.tc.bias = bias;
  gp.tc.act = act;
  gp.tc.mask_src = static_cast<const __nv_bfloat16*>(mask_src);
  gp.tc.mask_kind = mask_kind;
  gp.tc.panels = panels;
  gp.x = x;
  gp.FH = FH;
  gp.FW = FW;
  gp.Hin = H + halo;
  gp.Win = W + halo;
  gp.po = (pad_mode == SRK_PAD_VALID) ? 0 : k / 2;
  cudaStream_t s = as_stream(stream);
#define SRK_CASE(KS, CIN) \
  if (k == KS && cin == CIN) return launch_conv_gather<KS, CIN>(h, gp, w_packed, y_fpa, s);
  SRK_CASE(3, 1) SRK_CASE(3, 3) SRK_CASE(5, 1) SRK_CASE(5, 3) SRK_CASE(9, 1) SRK_CASE(9, 3)
#undef SRK_CASE
  set_error("srk_conv_first_tc: unsupported (k=%d, cin=%d)", k, cin);
  return -1;
}

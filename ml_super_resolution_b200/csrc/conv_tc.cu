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
// of 64- or 128-row chunks (one TMA load per chunk: every activation byte crosses L2->SMEM once); the A descriptors
// are row-shifted windows of that ring (probe-verified: SW128/SW64 descriptors accept any row shift with
// base_offset 0; the ring's first 128 rows are mirrored behind its last slot so a window never wraps).  The
// weights (K*K blocks of [NP][CIN] bf16, K-major) stay resident in shared memory.  Accumulators live in
// TMEM (2-8 stages) so the epilogue of tile i overlaps the MMAs of the tiles behind it.
//
// Warp roles (20 warps; the single-thread roles sit above the epilogue warps):
//   warps 0..15 : epilogue, in SETS sets that take tiles round-robin (2 sets x 4 TMEM lane quadrants x 2 column halves for
//                 64-wide outputs, 4 sets x 4 quadrant warps for 32/16-wide ones): tcgen05.ld -> lane-shift add ->
//                 bias/activation/mask -> bf16 -> swizzled smem -> TMA store; or fp32 NHWC with residual add / panel crop /
//                 pixel shuffle, transposed through shared memory so that the global stores are coalesced
//   warp 16     : TMA producer (chunk ring; a chunk's bytes are counted on the "ready" barrier of the first tile reading it)
//   warps 17,19 : tcgen05.mma issuers taking alternate tiles (warp-convergent loops, one elected lane issues)
//   warp 18     : TMA store issuer (waits for a full staging tile, stores it, frees the staging buffer)
// The hand-over protocol (per-tile ready / done barriers) is described next to the kernel; profiles/r1_trace_conv_tc.txt
// holds the in-kernel timelines it was derived from.
#include <cstdlib>

#include "sm100_ptx.cuh"
#include "srk_common.cuh"

namespace srk {

enum { EPI_FPA = 0, EPI_NHWC = 1 };

// Development-only timeline (build with -DSRK_TRACE -> libsrk_trace.so, tools/trace_conv.py): CTA 0 records clock64() at
// the hand-over points of its warp roles for its first 256 tiles / chunks.
#ifdef SRK_TRACE
__device__ unsigned long long g_trace[24 * 256];
#define SRK_TRACE_EV(ev, idx)                                                              \
  do {                                                                                     \
    if (blockIdx.x == 0 && (idx) >= 0 && (idx) < 256) g_trace[(ev) * 256 + (idx)] = clock64(); \
  } while (0)
#define SRK_ABLATE(p, bit) ((p).dbg & (bit))  // env SRK_DBG: 1 no lane shift, 2 no staging stores, 4 no tmem loads, 32 no tanh, 64 no bias
#else
#define SRK_TRACE_EV(ev, idx) \
  do {                        \
  } while (0)
#define SRK_ABLATE(p, bit) false
#endif

constexpr int kMaxRingRows = 2560;  // ring capacity cap: the per-tile barriers (32, round-robin) assume < 21 tiles of rows
constexpr int kSmemBudget = 227 * 1024;

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
  int dbg;  // development switches (env SRK_DBG): bit0 skip lane exchange+shuffles, bit1 skip staging store, bit2 skip tmem loads
};

template <int CIN, int NP, int KS, int EPI = EPI_FPA>
struct ConvTcCfg {
  static constexpr int kRowBytes = CIN * 2;
  // ring granularity = rows per TMA load.  Every load costs the producer thread ~4 slow single-thread operations
  // (slot-free wait, expect_tx, TMA issue, ready arrive), so chunks are 128 rows (about one per tile) wherever the ring
  // still holds a 254-wide panel's 2*Wp+128 rows plus look-ahead; the widest config keeps 64 for the finer occupancy.
  static constexpr int kChunkRows = (NP == 64 && CIN == 64 && KS == 3) ? 64 : 128;
  static constexpr int kMirrorSlots = 128 / kChunkRows;  // the ring's first 128 rows are duplicated behind its last slot
  static constexpr int kMaxRingSlots = kMaxRingRows / kChunkRows;
  static constexpr int kChunkBytes = kChunkRows * kRowBytes;
  static constexpr int kTaps = KS * KS;
  static constexpr int kHalo = KS / 2;
  static constexpr int kTileStride = 128 - (KS - 1);  // output rows per tile
  static constexpr int kN = KS * NP;                  // MMA N: one kernel row of taps
  static constexpr int kAccStages = (kN > 128) ? 2 : (kN > 64 ? 4 : 8);
  static constexpr int kTmemColsRaw = kAccStages * kN;
  static constexpr int kTmemCols = kTmemColsRaw <= 32 ? 32 : kTmemColsRaw <= 64 ? 64 : kTmemColsRaw <= 128 ? 128 : kTmemColsRaw <= 256 ? 256 : 512;
  static constexpr int kWTapBytes = NP * kRowBytes;
  static constexpr int kWBytes = ((kTaps * kWTapBytes + 1023) / 1024) * 1024;
  // epilogue shape: kSets tiles are in flight at once (each set = 4 lane-quadrant warps x kHalves column halves); the
  // narrow layers are bound by the per-tile latency chain, not by issue slots, so they run 4 sets of 4 warps
  static constexpr int kSets = (NP <= 32) ? 4 : 2;
  static constexpr int kHalves = (NP <= 32) ? 1 : 2;
  static constexpr int kStageBufs = (NP == 64 && CIN == 64 && KS == 3) ? 1 : kSets;  // the widest config spends its smem on the ring
  // EPI_NHWC with NP <= 32: per (tile set, lane quadrant) a double-buffered 32-pixel x NP fp32 transpose tile (row pitch
  // NP+1 words: conflict-free) plus 32 output base indices, so that the fp32 NHWC stores go out coalesced
  static constexpr bool kCoopStore = (EPI == EPI_NHWC) && (NP <= 32);  // (needs kHalves == 1: one warp owns 32 pixels)
  static constexpr int kCoopBufBytes = 32 * (NP + 1) * 4 + 32 * 8;
  static constexpr int kStageBytes = (EPI == EPI_NHWC) ? (kCoopStore ? kSets * 4 * kCoopBufBytes : 0)
                                                       : kStageBufs * 128 * NP * 2;  // output staging (EPI_FPA)
  static constexpr int kGroups = kSets * kHalves;   // lane-exchange groups (4 quadrant warps each): 16 epilogue warps
  static constexpr int kColPass = 16;               // accumulator columns a thread handles per pass
  static_assert(kGroups == 4, "named barriers 1..4 serve the exchange groups");
  static_assert(kAccStages % kSets == 0, "an accumulator stage is always drained by the same epilogue set (barrier parities are waited in that set's tile order)");
  static constexpr int kEpiThreads = 128 * kGroups;
  static constexpr int kIssuers = 2;                // MMA issuer warps (tiles round-robin); must divide kAccStages so that
                                                    // an accumulator stage is always driven by the same issuer (its
                                                    // "empty" parity wait would alias otherwise)
  static constexpr int kCtlWarps = 2 + kIssuers;    // + TMA producer, TMA store issuer
  static constexpr int kTileBars = 32;              // per-tile "ready" / "done" barriers, used round-robin
  static constexpr int kThreads = kEpiThreads + 32 * kCtlWarps;
  static constexpr int kXchGroupFloats = 2 /*parity*/ * 4 /*quadrants*/ * (KS > 1 ? (KS - 1) * kHalo : 1) * kColPass;  // == kXgrp / 4 in conv_epilogue
  static constexpr int kXchFloats = kXchGroupFloats * kGroups;
  static constexpr int kXchBytes = ((kXchFloats * 4 + 15) / 16) * 16;
  // every byte left after weights / staging / bookkeeping goes to the input ring: look-ahead is what hides HBM latency
  static constexpr int kFixedBytes = kWBytes + kStageBytes + 256 + 256 + kXchBytes + (32 + 2 * kTileBars) * 8 + 16 + 1024;
  static constexpr int kRingSlotsRaw = (kSmemBudget - kFixedBytes) / kChunkBytes - kMirrorSlots;
  static constexpr int kRingSlots = kRingSlotsRaw > kMaxRingSlots ? kMaxRingSlots : kRingSlotsRaw;
  static constexpr int kRingBytes = (kRingSlots + kMirrorSlots) * kChunkBytes;
  static constexpr int kOffW = 0;
  static constexpr int kOffRing = kWBytes;
  static constexpr int kOffStage = kOffRing + kRingBytes;
  static constexpr int kOffBias = kOffStage + kStageBytes;
  static constexpr int kOffTab = kOffBias + 256;    // NHWC epilogue: per-channel output offsets
  static constexpr int kOffXch = kOffTab + 256;
  static constexpr int kOffBars = kOffXch + kXchBytes;
  static constexpr int kNumBars = 1 + 2 * kAccStages + 8 + 2 * kTileBars;
  static constexpr int kOffTmemSlot = kOffBars + kNumBars * 8;
  static constexpr int kTotal = kOffTmemSlot + 16 + 1024;  // + alignment slack
  static_assert(kTotal <= kSmemBudget && kRingSlots * kChunkRows >= 640, "shared-memory plan does not fit");
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
  uint32_t bar_sfull, bar_sfree;     // + 8 * staging buffer / + 8 * epilogue set
  uint32_t stage_addr;               // shared-window address of staging buffer 0
  int stage_stride;                  // bytes between the staging buffers
  uint32_t xch_addr;                 // shared-window address of the lane-exchange area
  float* s_bias;
  int* s_tab;
};

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// One thread: wait for a complete staging tile, TMA-store it, hand the staging buffers back.
// Tile `it` (CTA-local index) uses staging buffer it % BUFS and is produced by epilogue set it % SETS.  "Buffer free"
// is signalled on a barrier OWNED BY THE SET that writes next into it (tile it + BUFS), so a set only ever waits for
// the next completion of its own barrier: the parity can never alias, however far the sets drift apart.
template <int TS, int SETS, int BUFS>
__device__ __forceinline__ void conv_store_loop(const ConvTcParams& p, const EpiCtx& e, int t_begin, int t_end) {
  for (int t = t_begin; t < t_end; ++t) {
    const int it = t - t_begin, sb = it % BUFS, sgen = it / BUFS;
    mbar_wait(e.bar_sfull + 8u * sb, sgen & 1);
    SRK_TRACE_EV(8, it);
    tma_store_2d(&p.map_out, 0, TS * t, e.stage_addr + sb * e.stage_stride);
    tma_store_commit();
    if (it >= BUFS - 1) {  // the store of tile it-(BUFS-1) has finished READING its buffer -> tile it+1 may overwrite it
      tma_store_wait_read<BUFS - 1>();
      mbar_arrive(e.bar_sfree + 8u * ((it + 1) % SETS));
    }
    SRK_TRACE_EV(9, it);
  }
  tma_store_wait_all<0>();
}

// Epilogue.  The 16 epilogue warps form SETS sets that take tiles round-robin, so SETS tiles' epilogue chains
// (tcgen05.ld -> lane exchange -> shuffles -> pack -> staging) are in flight at once: with a single set the chain
// latency of ~1.4 us per tile, not the tensor pipe, bounded the kernel.  Within a set: TMEM lane quadrant
// ewarp & 3, column half (HALVES == 2 only); a thread covers its NP / HALVES output channels in passes of CP
// columns (keeps the live accumulator registers at 3*CP).
//   per pass: KS tcgen05.ld of CP columns -> lane-shift add (rotating shuffles; the edge lanes of a warp first swap in
//   the neighbouring quadrant's row through shared memory) -> packed fp32x2 bias add -> bf16x2 pack -> ReLU / ReLU'
//   mask on the packed pairs;  per tile: swizzled 16-byte stores into the staging tile (TMA-stored by warp 2).
template <int NP, int KS, int EPI, int ACC, int KN, int SETS, int HALVES, int CP, int BUFS>
__device__ __forceinline__ void conv_epilogue(const ConvTcParams& p, const EpiCtx& e, int ewarp, int lane, int t_begin, int t_end) {
  constexpr int H_ = KS / 2, TS = 128 - (KS - 1);
  constexpr int PASSES = NP / (HALVES * CP);
  static_assert(SETS * HALVES == 4 && PASSES * HALVES * CP == NP, "epilogue shape");
  constexpr int kNB = (KS > 1) ? KS - 1 : 1;           // shifted blocks
  constexpr int kXq = kNB * (H_ > 0 ? H_ : 1) * CP * 4;  // bytes per (parity, quadrant)
  constexpr int kXpar = 4 * kXq;                       // bytes per parity
  constexpr int kXgrp = 2 * kXpar;                     // bytes per exchange group
  const uint32_t tmem = e.tmem;
  const int quad = ewarp & 3;          // TMEM lane quadrant this warp may access (hardware: warp index % 4)
  const int half = (HALVES == 2) ? (ewarp >> 2) & 1 : 0;  // which half of the output channels
  const int set = ewarp / (4 * HALVES);  // which tiles: local tile index it with it % SETS == set
  const int grp = set * HALVES + half;   // exchange group / named barrier: the 4 quadrant warps that share columns and tile
  const int row = quad * 32 + lane;    // lane j of the accumulator <-> flat row TS*t - h + j
  const bool lane_valid = (row >= H_) && (row < H_ + TS);
  const uint32_t xg = e.xch_addr + grp * kXgrp;
  // per-thread constant addresses of the lane exchange (parity 0) ...
  uint32_t pub_addr[KS], sub_addr[KS];
  bool do_pub[KS], do_sub[KS];
#pragma unroll
  for (int b = 0; b < KS; ++b) {
    const int dx = b - H_;
    const int bi = (dx < 0) ? b : b - 1;
    const bool edge = (dx < 0) ? (lane >= 32 + dx) : (dx > 0 ? lane < dx : false);
    const int li = (dx < 0) ? lane - (32 + dx) : lane;
    const int nq = (dx < 0) ? quad - 1 : quad + 1;
    do_pub[b] = edge;
    do_sub[b] = edge && nq >= 0 && nq < 4;
    pub_addr[b] = xg + quad * kXq + (bi * H_ + li) * CP * 4;
    sub_addr[b] = xg + nq * kXq + (bi * H_ + li) * CP * 4;
  }
  // ... and of this thread's first 16-byte chunk in staging buffer 0 (chunk index XOR swizzle applied per chunk)
  const int srow = row - H_;
  const int sw = (NP == 64) ? (srow & 7) : ((srow >> 1) & 3);
  const uint32_t st_row = e.stage_addr + srow * (NP * 2);
  // pixel coordinates of this lane's row in this set's first tile, then advanced by SETS*TS rows per processed tile
  const int H1 = p.H + 1;
  int px, pyy, pn;
  {
    const int64_t prow0 = int64_t(TS) * (t_begin + set) - H_ + row;
    const uint32_t pr = uint32_t(prow0 < 0 ? 0 : prow0);
    const uint32_t q = pr / uint32_t(p.Wp);
    px = int(pr - q * uint32_t(p.Wp));
    pn = int(q / uint32_t(H1));
    pyy = int(q - uint32_t(pn) * uint32_t(H1));
  }
  const int adv_x = (SETS * TS) % p.Wp, adv_q = (SETS * TS) / p.Wp;
  const int adv_y = adv_q % H1, adv_n = adv_q / H1;
  const int act = p.act;
  const __nv_bfloat162 zero2 = __floats2bfloat162_rn(0.f, 0.f);
  uint32_t xpar = 0;  // byte offset of the exchange parity in use
  // cooperative NHWC store (EPI_NHWC, NP <= 32): the two column-half warps of a (set, quadrant) pair transpose
  // their 32 pixels x NP channels through shared memory and then write each output row segment with
  // consecutive threads on consecutive floats (a lane-per-pixel scatter touches 3x the sectors)
  constexpr bool kCoop = (EPI == EPI_NHWC) && (NP <= 32);
  static_assert(!kCoop || HALVES == 1, "the coalesced NHWC store is warp-synchronous: one warp owns its 32 pixels");
  constexpr int kTP = NP + 1;
  constexpr int kCoopBuf = 32 * kTP * 4 + 32 * 8;
  const uint32_t tb = e.stage_addr + (set * 4 + quad) * kCoopBuf;
  const int sr = (EPI == EPI_NHWC) ? p.shuffle_r : 1;
  const int rC = (EPI == EPI_NHWC) ? p.cout / sr : 1;             // floats per pixel per output row
  const uint32_t div_m = 65536u / uint32_t(rC) + 1u;              // el / rC == (el * div_m) >> 16 for el < 1024
  const int64_t orow = (EPI == EPI_NHWC) ? int64_t(p.FW) * rC : 0;  // floats per output row

  for (int t = t_begin + set; t < t_end; t += SETS) {
    const int it = t - t_begin, acc = it % ACC, accgen = it / ACC;
    // the pixel this lane holds
    const int64_t prow = int64_t(TS) * t - H_ + row;
    bool valid = lane_valid && prow < p.rows_valid && (px < p.W) && (pyy > 0);
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
    mbar_wait(e.bar_tfull0 + 8u * acc, accgen & 1);
    tc_fence_after();
    if (quad == 0 && half == 0 && lane == 0) SRK_TRACE_EV(4, it);
    uint32_t packed[PASSES][CP >= 2 ? CP / 2 : 1];

#pragma unroll
    for (int pass = 0; pass < PASSES; ++pass) {
      const int col0 = (half * PASSES + pass) * CP;
      const uint32_t taddr = tmem + acc * KN + col0 + (uint32_t(quad * 32) << 16);
      // data-gradient mask (saved activation of the producing layer): requested before the accumulator is read so that its
      // L2 latency hides under the TMEM loads, the lane exchange and the shuffles
      uint4 mpre[CP >= 8 ? CP / 8 : 1];
      const bool mask_fast = EPI == EPI_FPA && p.mask_src && p.mask_kind == SRK_ACT_RELU && !p.addend_fpa && valid;
      if (mask_fast) {
        const uint4* m = reinterpret_cast<const uint4*>(p.mask_src + size_t(prow) * NP + col0);
#pragma unroll
        for (int j = 0; j < CP / 8; ++j) mpre[j] = __ldg(m + j);
      }
      float blk[KS][CP];
#pragma unroll
      for (int b = 0; b < KS; ++b) {
        if (SRK_ABLATE(p, 4)) {
#pragma unroll
          for (int c = 0; c < CP; ++c) blk[b][c] = float(lane + c);
        } else {
          tmem_load_cols<CP>(taddr + b * NP, blk[b]);
        }
      }
      tmem_ld_wait();
      if (pass == 0 && quad == 0 && half == 0 && lane == 0) SRK_TRACE_EV(10, it);
      if (pass == PASSES - 1) {  // this thread has read everything it needs from the accumulator stage
        tc_fence_before();
        mbar_arrive(e.bar_tempty0 + 8u * acc);
        if (quad == 0 && half == 0 && lane == 0) SRK_TRACE_EV(5, it);
      }
      float v[CP];
#pragma unroll
      for (int c = 0; c < CP; ++c) v[c] = blk[H_][c];
      if (KS > 1 && !SRK_ABLATE(p, 1)) {
        // ---- lane-shift add of the KS column blocks: y[j] = sum_dx D[j + dx][block dx]
#pragma unroll
        for (int b = 0; b < KS; ++b) {
          if (b == H_) continue;
          if (do_pub[b]) {
#pragma unroll
            for (int c = 0; c < CP / 4; ++c) sts128(pub_addr[b] + xpar + c * 16, blk[b][4 * c], blk[b][4 * c + 1], blk[b][4 * c + 2], blk[b][4 * c + 3]);
          }
        }
        if (pass == 0 && quad == 0 && half == 0 && lane == 0) SRK_TRACE_EV(11, it);
        named_bar_sync(1 + grp, 128);
        if (pass == 0 && quad == 0 && half == 0 && lane == 0) SRK_TRACE_EV(12, it);
        // the edge lanes take over the neighbouring quadrant's row (their own value of that block is not needed any
        // more); then ONE rotating shuffle per column serves every lane: no per-element select, and every output
        // sees the same fp32 addition order (tiled == un-tiled bit for bit)
#pragma unroll
        for (int b = 0; b < KS; ++b) {
          if (b == H_) continue;
          if (do_sub[b]) {
#pragma unroll
            for (int c = 0; c < CP / 4; ++c) {
              const float4 o = lds128(sub_addr[b] + xpar + c * 16);
              blk[b][4 * c] = o.x;
              blk[b][4 * c + 1] = o.y;
              blk[b][4 * c + 2] = o.z;
              blk[b][4 * c + 3] = o.w;
            }
          }
        }
#pragma unroll
        for (int b = 0; b < KS; ++b) {
          if (b == H_) continue;
          const int src_lane = (lane + (b - H_)) & 31;
          float sh[CP];
#pragma unroll
          for (int c = 0; c < CP; ++c) sh[c] = __shfl_sync(0xffffffffu, blk[b][c], src_lane);
#pragma unroll
          for (int c = 0; c < CP; c += 2) {
            const float2 r = __fadd2_rn(make_float2(v[c], v[c + 1]), make_float2(sh[c], sh[c + 1]));
            v[c] = r.x;
            v[c + 1] = r.y;
          }
        }
        xpar ^= kXpar;
        if (pass == 0 && quad == 0 && half == 0 && lane == 0) SRK_TRACE_EV(13, it);
      }
      // bias (packed fp32x2 adds; read from shared memory to keep registers for the accumulator blocks) and tanh
      if (!SRK_ABLATE(p, 64))
#pragma unroll
      for (int c = 0; c < CP; c += 2) {
        const float2 bb = *reinterpret_cast<const float2*>(e.s_bias + col0 + c);
        const float2 r = __fadd2_rn(make_float2(v[c], v[c + 1]), bb);
        v[c] = r.x;
        v[c + 1] = r.y;
      }
      if (act == SRK_ACT_TANH && !SRK_ABLATE(p, 32)) {
#pragma unroll
        for (int c = 0; c < CP; ++c) v[c] = act_apply(v[c], SRK_ACT_TANH);
      }

      if constexpr (EPI == EPI_FPA) {
        static_assert(CP >= 8, "FPA epilogue stores 16-byte chunks");
        uint32_t* pk = packed[pass];
        if (valid) {
          if (p.addend_fpa || (p.mask_src && p.mask_kind != SRK_ACT_RELU)) {
            // rare forms (EnhanceNet block residual, tanh' mask): fp32 path
            if (act == SRK_ACT_RELU) {
#pragma unroll
              for (int c = 0; c < CP; ++c) v[c] = fmaxf(v[c], 0.f);
            }
            // relu_after_add == 2: data gradient through a residual junction, (conv + addend) * act'(mask_src);
            // otherwise conv * act'(mask_src) [+ addend [-> ReLU]]
            const bool add_first = p.addend_fpa && p.relu_after_add == 2;
            auto add_addend = [&]() {
              const uint4* a = reinterpret_cast<const uint4*>(p.addend_fpa + size_t(prow) * NP + col0);
#pragma unroll
              for (int j = 0; j < CP / 8; ++j) {
                const uint4 av = __ldg(a + j);
                const uint32_t w4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[q]));
                  v[j * 8 + q * 2] += f.x;
                  v[j * 8 + q * 2 + 1] += f.y;
                }
              }
            };
            if (add_first) add_addend();
            if (p.mask_src) {
              const uint4* m = reinterpret_cast<const uint4*>(p.mask_src + size_t(prow) * NP + col0);
              const bool tanh_kind = p.mask_kind == SRK_ACT_TANH;
#pragma unroll
              for (int j = 0; j < CP / 8; ++j) {
                const uint4 mv = __ldg(m + j);
                const uint32_t w4[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[q]));
                  v[j * 8 + q * 2] *= tanh_kind ? (1.f - f.x * f.x) : (f.x > 0.f ? 1.f : 0.f);
                  v[j * 8 + q * 2 + 1] *= tanh_kind ? (1.f - f.y * f.y) : (f.y > 0.f ? 1.f : 0.f);
                }
              }
            }
            if (p.addend_fpa && !add_first) {
              add_addend();
              if (p.relu_after_add) {
#pragma unroll
                for (int c = 0; c < CP; ++c) v[c] = fmaxf(v[c], 0.f);
              }
            }
#pragma unroll
            for (int c = 0; c < CP / 2; ++c) pk[c] = pack_bf16x2(v[2 * c], v[2 * c + 1]);
          } else {
            // common forms: ReLU and the ReLU' mask act on the packed bf16 pairs (both commute with the rounding)
#pragma unroll
            for (int c = 0; c < CP / 2; ++c) pk[c] = pack_bf16x2(v[2 * c], v[2 * c + 1]);
            if (act == SRK_ACT_RELU) {
#pragma unroll
              for (int c = 0; c < CP / 2; ++c) {
                const __nv_bfloat162 h = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&pk[c]), zero2);
                pk[c] = *reinterpret_cast<const uint32_t*>(&h);
              }
            }
            if (mask_fast) {
#pragma unroll
              for (int j = 0; j < CP / 8; ++j) {
                const uint4 mv = mpre[j];
                const uint32_t w4[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const __nv_bfloat162 gt = __hgt2(*reinterpret_cast<const __nv_bfloat162*>(&w4[q]), zero2);  // 1.0 / 0.0 per half
                  const __nv_bfloat162 h = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&pk[j * 4 + q]), gt);
                  pk[j * 4 + q] = *reinterpret_cast<const uint32_t*>(&h);
                }
              }
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < CP / 2; ++c) pk[c] = 0u;  // pad rows/columns stay exactly zero
        }
      } else {
        // fp32 NHWC: residual add, panel crop, depth_to_space
        if (act == SRK_ACT_RELU) {
#pragma unroll
          for (int c = 0; c < CP; ++c) v[c] = fmaxf(v[c], 0.f);
        }
        bool st = valid;
        int fn = n, fy = y, fx = x;
        if (st && p.panels) {
          const srk_panel pe = p.panels[n];
          st = (y >= pe.own_y0) && (y < pe.own_y1) && (x >= pe.own_x0) && (x < pe.own_x1);
          fn = pe.frame;
          fy = pe.y0 + y;
          fx = pe.x0 + x;
        }
        const int64_t base = (int64_t(fn) * p.FH + fy) * sr * orow + int64_t(fx) * rC;
        if constexpr (kCoop) {
#pragma unroll
          for (int c = 0; c < CP; ++c)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(tb + (lane * kTP + col0 + c) * 4), "f"(v[c]) : "memory");
          if (pass == 0)
            asm volatile("st.shared.b64 [%0], %1;" ::"r"(tb + 32 * kTP * 4 + lane * 8), "l"(st ? base : int64_t(-1)) : "memory");
        } else if (st) {
#pragma unroll
          for (int c = 0; c < CP; ++c) {
            const int off = e.s_tab[col0 + c];
            if (off >= 0) {
              const int64_t idx = base + off;
              float o = v[c];
              if (p.addend) o += __ldg(p.addend + idx);
              p.out[idx] = o;
            }
          }
        }
      }
      if (pass == 0 && quad == 0 && half == 0 && lane == 0) SRK_TRACE_EV(14, it);
    }  // pass
    if (quad == 0 && half == 0 && lane == 0) SRK_TRACE_EV(6, it);

    if constexpr (kCoop) {
      __syncwarp();  // all NP channels of this warp's 32 pixels are in the transpose tile
      // Output "row" j = dy * rC + k covers, for the 32 pixels, elements {32k + lane} of the rC-float runs that output
      // row dy receives: consecutive lanes write consecutive floats.  Four rows are batched so that the shared-memory
      // loads of a batch are all in flight before the first global store depends on one.
      constexpr int U = 4;
      for (int j0 = 0; j0 < p.cout; j0 += U) {
        int64_t pb[U];
        float o[U];
        int rem_[U], dy_[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = min(j0 + u, p.cout - 1);
          const uint32_t dy = (uint32_t(j) * div_m) >> 16;
          const int el = 32 * (j - int(dy) * rC) + lane;
          const uint32_t pi = (uint32_t(el) * div_m) >> 16;
          rem_[u] = el - int(pi) * rC;
          dy_[u] = int(dy);
          asm volatile("ld.shared.b64 %0, [%1];" : "=l"(pb[u]) : "r"(tb + 32 * kTP * 4 + pi * 8));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o[u]) : "r"(tb + (pi * kTP + dy * rC + rem_[u]) * 4));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (j0 + u < p.cout && pb[u] >= 0) {
            const int64_t idx = pb[u] + dy_[u] * orow + rem_[u];
            float ov = o[u];
            if (p.addend) ov += __ldg(p.addend + idx);
            p.out[idx] = ov;
          }
        }
      }
      __syncwarp();  // reads done before the next tile's writes
    }
    if constexpr (EPI == EPI_FPA) {
      // staging buffer free? (the store that last used this buffer has finished reading it: signalled on this set's barrier)
      const int sb = it % BUFS;
      if (it >= BUFS) mbar_wait(e.bar_sfree + 8u * set, ((it - BUFS) / SETS) & 1);
      if (lane_valid && !SRK_ABLATE(p, 2)) {
#pragma unroll
        for (int pass = 0; pass < PASSES; ++pass)
#pragma unroll
          for (int j = 0; j < CP / 8; ++j) {
            const int chunk = (half * PASSES + pass) * (CP / 8) + j;
            sts128u(st_row + sb * e.stage_stride + ((chunk ^ sw) << 4), packed[pass][4 * j], packed[pass][4 * j + 1], packed[pass][4 * j + 2],
                    packed[pass][4 * j + 3]);
          }
      }
      fence_proxy_async_smem();
      mbar_arrive(e.bar_sfull + 8u * sb);
    }
    if (quad == 0 && half == 0 && lane == 0) SRK_TRACE_EV(7, it);
  }
}

template <int CIN, int NP, int KS, int EPI>
__global__ void __launch_bounds__(ConvTcCfg<CIN, NP, KS, EPI>::kThreads, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  using L = ConvTcCfg<CIN, NP, KS, EPI>;
  constexpr int kRingSlots = L::kRingSlots, kChunkRows = L::kChunkRows, kMirrorSlots = L::kMirrorSlots;
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
  constexpr int TB = L::kTileBars;
  const uint32_t bar_wfull = s_bars;                                        // weights loaded
  auto bar_tfull = [&](int a) { return s_bars + 8u * (1 + a); };            // accumulator stage complete
  auto bar_tempty = [&](int a) { return s_bars + 8u * (1 + ACC + a); };     // accumulator stage drained
  const uint32_t bar_sfull = s_bars + 8u * (1 + 2 * ACC);      // [4] staging tile complete (per buffer)
  const uint32_t bar_sfree = s_bars + 8u * (1 + 2 * ACC + 4);  // [4] staging buffer reusable (per set)
  auto bar_tready = [&](uint32_t it) { return s_bars + 8u * (1 + 2 * ACC + 8 + (it % TB)); };       // tile's last rows have landed
  auto bar_tdone = [&](uint32_t it) { return s_bars + 8u * (1 + 2 * ACC + 8 + TB + (it % TB)); };   // tile's MMAs have completed

  // Warp roles.  The SM sub-partition arbiter favours the HIGHEST warp id (B300_MICROARCH: hi-wid-first), so the
  // single-thread roles sit above the 16 epilogue warps: 16 = TMA producer, 17 / 19 = MMA issuers, 18 = TMA store
  // issuer.
  //
  // Hand-over protocol.  Every mbarrier operation of an issuer costs ~200 cycles (its instruction stream is one long
  // dependent chain competing with 16 busy epilogue warps), and per-chunk waits / hand-backs made that the kernel's
  // critical path (in-kernel timeline, profiles/r1_trace_conv_tc.txt).  So the issuers only touch PER-TILE barriers:
  //   producer : loads 64-row chunks into the ring.  A chunk's bytes are counted on the "tile ready" barrier of the FIRST
  //              tile that reads it (expect_tx per chunk, one arrive when the tile's last chunk has been issued), so a
  //              tile's rows are all in once the ready barriers of every tile up to it have completed.  A slot is reused
  //              once every tile that read its previous chunk is done ("tile done" barriers, waited in tile order);
  //   issuer   : wait "tile ready" of its tile and of the other issuer's tile before it + "accumulator drained"
  //              -> 12 MMAs -> commit "accumulator complete" + "tile done".
  // The per-tile barriers are used round-robin (32 of each); a tile 32 ahead cannot complete before the waiter has seen
  // the current phase because the ring holds fewer than 21 tiles of rows.
  constexpr int kEpiWarps = L::kEpiThreads / 32;
  const int warp = int(threadIdx.x >> 5) - kEpiWarps;  // control warp index 0..3; negative: epilogue warp
  const int ewarp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // contiguous tile range of this CTA; tile t produces output rows [TS*t, TS*t + TS)
  const int t_begin = int((int64_t(blockIdx.x) * p.num_tiles) / gridDim.x);
  const int t_end = int((int64_t(blockIdx.x + 1) * p.num_tiles) / gridDim.x);
  const int reach = H_ * p.Wp;  // rows of vertical reach
  auto lo_chunk = [&](int t) { return floor_div(TS * t - H_ - reach, kChunkRows); };
  auto hi_chunk = [&](int t) { return floor_div(TS * t - H_ + reach + 127, kChunkRows); };
  const int c0 = lo_chunk(t_begin);  // first chunk this CTA loads (may be negative: TMA zero-fills)
  const int c_last = hi_chunk(t_end - 1);

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < TB; ++i) {
      mbar_init(bar_tready(i), 1);
      mbar_init(bar_tdone(i), 1);
    }
    mbar_init(bar_wfull, 1);
    for (int i = 0; i < ACC; ++i) {
      mbar_init(bar_tfull(i), 1);
      mbar_init(bar_tempty(i), L::kEpiThreads / L::kSets);  // one epilogue set per tile
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_sfull + 8u * i, L::kEpiThreads / L::kSets);
      mbar_init(bar_sfree + 8u * i, 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<L::kTmemCols>(smem_u32(tmem_slot));
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_in);
    tma_prefetch_desc(&p.map_w);
    if (EPI == EPI_FPA) tma_prefetch_desc(&p.map_out);
  }
  if (threadIdx.x == 0) SRK_TRACE_EV(20, 0);
  pdl_wait();               // predecessor kernels are complete and their writes visible from here on
  pdl_launch_dependents();  // the next kernel may start filling SMs as our CTAs retire
  if (threadIdx.x == 0) SRK_TRACE_EV(20, 1);
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
  ec.stage_addr = smem_u32(stage_ptr);
  ec.stage_stride = 128 * NP * 2;
  ec.s_bias = s_bias;
  ec.s_tab = s_tab;
  ec.xch_addr = smem_u32(s_xch);

  if (t_begin < t_end) {
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer
      if (lane == 0) {
        mbar_arrive_expect_tx(bar_wfull, L::kTaps * L::kWTapBytes);
        for (int tap = 0; tap < L::kTaps; ++tap) tma_load_2d(s_w + tap * L::kWTapBytes, &p.map_w, 0, tap * NP, bar_wfull);
        // tiles [0, done_n) of this CTA are known to be complete; lo_row_done = first row tile done_n reads.
        // own_it = the tile whose "ready" barrier counts the chunk being issued; hi_row_own = last row that tile reads.
        const int n_tiles = t_end - t_begin;
        int done_n = 0, own_it = 0;
        int lo_row_done = TS * t_begin - H_ - reach;
        int hi_row_own = TS * t_begin - H_ + reach + 127;
        for (int c = c0; c <= c_last; ++c) {
          const int i = c - c0, slot = i % kRingSlots;
          SRK_TRACE_EV(16, i);
          if (i >= kRingSlots) {
            // the chunk this one replaces ends at row (c - kRingSlots) * 64 + 63: wait for every tile that reads it
            const int old_end = (c - kRingSlots) * kChunkRows + kChunkRows - 1;
            while (done_n < n_tiles && lo_row_done <= old_end) {
              mbar_wait(bar_tdone(uint32_t(done_n)), (uint32_t(done_n) / TB) & 1);
              ++done_n;
              lo_row_done += TS;
            }
          }
          SRK_TRACE_EV(17, i);
          while (own_it < n_tiles - 1 && hi_row_own < c * kChunkRows) {  // tile own_it reads nothing from this chunk on
            mbar_arrive(bar_tready(uint32_t(own_it)));
            ++own_it;
            hi_row_own += TS;
          }
          SRK_TRACE_EV(0, i);
          const bool mir = slot < kMirrorSlots;
          const uint32_t bar = bar_tready(uint32_t(own_it));
          mbar_expect_tx(bar, L::kChunkBytes * (mir ? 2 : 1));
          SRK_TRACE_EV(18, i);
          tma_load_2d(s_ring + slot * L::kChunkBytes, &p.map_in, 0, c * kChunkRows, bar);
          SRK_TRACE_EV(19, i);
          if (mir) tma_load_2d(s_ring + (kRingSlots + slot) * L::kChunkBytes, &p.map_in, 0, c * kChunkRows, bar);
        }
        for (; own_it < n_tiles; ++own_it) mbar_arrive(bar_tready(uint32_t(own_it)));
      }
    } else if (warp == 1 || warp == 3) {
      // ------------------------------------------------------------------ MMA issuers take tiles round-robin
      {
        // Each warp executes its loop convergently so that addresses and descriptors live in UNIFORM registers (no R2UR
        // hops before every UTCHMMA); only the tcgen05 instructions themselves are predicated on one elected lane.
        // UTCHMMA issue blocks while the tensor queue is full, so an issuer's bookkeeping does not overlap its own MMAs:
        // two issuers alternate tiles and one warp's bookkeeping runs under the other warp's MMAs.  The weight
        // descriptors are built once and a window's descriptor is one 32-bit add away from the previous one.
        constexpr uint32_t idesc = umma_idesc_bf16(128, L::kN, 0, 0);
        constexpr uint64_t hi = umma_desc_hi(0, kSbo, kLayout);
        constexpr uint32_t hi32 = uint32_t(hi >> 32);
        constexpr int kRingRows = kRingSlots * kChunkRows;
        constexpr int KK = CIN / 16;
        constexpr int NI = L::kIssuers;
        constexpr int kStep = NI * TS;  // rows between two tiles of the same issuer
        static_assert(kStep < kRingRows && (ACC & (ACC - 1)) == 0 && ACC % NI == 0, "issuer stride / accumulator stage arithmetic");
        const int mw = warp >> 1;  // issuer index
        const int t_first = t_begin + mw;
        mbar_wait(bar_wfull, 0);
        uint32_t b_lo[KS * KK];  // low words of the weight descriptors: constant for the whole kernel
#pragma unroll
        for (int r = 0; r < KS; ++r)
#pragma unroll
          for (int k = 0; k < KK; ++k) b_lo[r * KK + k] = ((s_w + r * KS * L::kWTapBytes + k * 32) >> 4) & 0x3FFF;
        int win[KS];  // window start rows relative to the ring origin, already wrapped
#pragma unroll
        for (int r = 0; r < KS; ++r) win[r] = (TS * t_first - H_ - c0 * kChunkRows + (r - H_) * p.Wp) % kRingRows;
        for (int t = t_first; t < t_end; t += NI) {
          const uint32_t it = uint32_t(t - t_begin);
          const uint32_t acc = it % ACC, acc_par = ((it / ACC) & 1) ^ 1;  // accumulator stage, parity of its "empty" barrier
          if (it > 0) mbar_wait(bar_tready(it - 1), ((it - 1) / TB) & 1);  // chunks counted on the other issuer's tile
          mbar_wait(bar_tready(it), (it / TB) & 1);
          if (lane == 0) SRK_TRACE_EV(1, t - t_begin);
          mbar_wait(bar_tempty(acc), acc_par);
          tc_fence_after();
          if (lane == 0) SRK_TRACE_EV(2, t - t_begin);
          const uint32_t d_tmem = tmem + acc * L::kN;
#pragma unroll
          for (int r = 0; r < KS; ++r) {
            const uint32_t a_lo = ((s_ring + win[r] * L::kRowBytes) >> 4) & 0x3FFF;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < KK; ++k) {
                const uint64_t adesc = (uint64_t(hi32) << 32) | uint64_t(a_lo + 2 * k);
                const uint64_t bdesc = (uint64_t(hi32) << 32) | uint64_t(b_lo[r * KK + k]);
                if (r == 0 && k == 0) umma_bf16(d_tmem, adesc, bdesc, idesc, 0u);
                else umma_bf16(d_tmem, adesc, bdesc, idesc, 1u);
              }
            }
            __syncwarp();
            win[r] += kStep;
            win[r] -= (win[r] >= kRingRows) ? kRingRows : 0;
          }
          if (elect_one()) {
            umma_commit(bar_tfull(acc));
            umma_commit(bar_tdone(it));
          }
          __syncwarp();
          if (lane == 0) SRK_TRACE_EV(3, t - t_begin);
        }
      }
    } else if (warp == 2) {
      // ------------------------------------------------------------------ TMA store issuer (one thread)
      if (EPI == EPI_FPA && lane == 0) conv_store_loop<TS, L::kSets, L::kStageBufs>(p, ec, t_begin, t_end);
    } else if (warp < 0) {
      // ------------------------------------------------------------------ epilogue (128 threads per column group)
      conv_epilogue<NP, KS, EPI, ACC, L::kN, L::kSets, L::kHalves, CP, L::kStageBufs>(p, ec, ewarp, lane, t_begin, t_end);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) SRK_TRACE_EV(20, 2);
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
  // Gather groups take tiles round-robin; three groups (12 warps, 32 warps per CTA) where one 16 KB im2col stage per tile
  // leaves room for six stages.  A stage must always be filled by the same group and consumed by the same MMA issuer
  // (their barrier parities are waited in that role's own tile order): kAStages is a multiple of both counts.
  static constexpr int kGatherGroups = kBlocks == 1 ? 3 : 2;
  static constexpr int kAStages = kBlocks == 1 ? 6 : (kBlocks == 2 ? 4 : 2);
  static_assert(kAStages % kGatherGroups == 0 && kAStages % 2 == 0, "stage ownership");
  static constexpr int kWBytes = kBlocks * 64 * 128;
  static constexpr int kAcc = 4;
  static constexpr int kEpiThreads = 512, kGatherThreads = 128;
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
      mbar_init(bar_tempty(i), L::kEpiThreads / 2);  // one epilogue set (8 warps) per tile
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_sfull + 8u * i, L::kEpiThreads / 2);
      mbar_init(bar_sfree + 8u * i, 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(smem_u32(tmem_slot));
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_w);
    tma_prefetch_desc(&p.map_out);
  }
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
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
  ec.stage_addr = s_base + L::kOffStage;
  ec.stage_stride = 128 * 64 * 2;
  ec.s_bias = s_bias;
  ec.s_tab = reinterpret_cast<int*>(smem + L::kOffTab);
  ec.xch_addr = s_base + L::kOffXch;

  if (t_begin < t_end) {
    if (warp == 0) {
      if (lane == 0) {  // the packed kernel: kBlocks blocks of [64 co][64 k] bf16
        mbar_arrive_expect_tx(bar_wfull, L::kWBytes);
        for (int b = 0; b < L::kBlocks; ++b) tma_load_2d(s_w + b * 64 * 128, &p.map_w, 0, b * 64, bar_wfull);
      }
    } else if (warp == 1 || warp == 3) {
      // two MMA issuers take alternate tiles (AST and ACC are even: an A stage / accumulator stage always belongs to the same
      // issuer): one warp's barrier waits and commits run under the other's MMAs, as in conv_tc_kernel
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
        constexpr uint64_t hi = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
        static_assert(ACC % 2 == 0 && AST % 2 == 0, "stage ownership per issuer");
        mbar_wait(bar_wfull, 0);
        for (int t = t_begin + (warp >> 1); t < t_end; t += 2) {
          const int it = t - t_begin, acc = it % ACC, st = it % AST;
          mbar_wait(bar_afull(st), (it / AST) & 1);
          SRK_TRACE_EV(1, it);
          mbar_wait(bar_tempty(acc), ((it / ACC) & 1) ^ 1);
          tc_fence_after();
          SRK_TRACE_EV(2, it);
#pragma unroll
          for (int k = 0; k < L::kKP / 16; ++k) {
            const uint32_t a_addr = s_a + st * L::kABytes + (k / 4) * (128 * 128) + (k % 4) * 32;
            const uint32_t b_addr = s_w + (k / 4) * (64 * 128) + (k % 4) * 32;
            umma_bf16(tmem + acc * 64, umma_desc(hi, a_addr), umma_desc(hi, b_addr), idesc, k != 0);
          }
          umma_commit(bar_tfull(acc));
          umma_commit(bar_aempty(st));
          SRK_TRACE_EV(3, it);
        }
      }
    } else if (warp == 2) {
      if (lane == 0) conv_store_loop<128, 2, 2>(p, ec, t_begin, t_end);
    } else if (warp >= 4 && warp < 20) {
      conv_epilogue<64, 1, EPI_FPA, ACC, 64, 2, 2, 16, 2>(p, ec, warp - 4, lane, t_begin, t_end);
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
        if (r == 0) SRK_TRACE_EV(16, it);
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
          const bool interior = y >= 0 && y + KS <= gp.Hin && x >= 0 && x + KS <= gp.Win;
          if (interior) {
            // every tap lies inside the panel window (all but the border pixels): plain loads at constant offsets from
            // KS row pointers, no clamps and no selects
            const float* row0 = frame + (int64_t(y) * gp.FW + x) * CIN;
            const int rs = gp.FW * CIN;
#pragma unroll
            for (int c8 = 0; c8 < L::kKP / 8; ++c8) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int k = c8 * 8 + j;
                f[j] = 0.f;
                if (k < L::kKT) {
                  const int tap = k / CIN, ci = k % CIN, u = tap / KS, v = tap % KS;
                  f[j] = __ldg(row0 + u * rs + v * CIN + ci);
                }
              }
              const uint4 q4 = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
              *reinterpret_cast<uint4*>(arow + (c8 / 8) * (128 * 128) + (((c8 % 8) ^ (r & 7)) << 4)) = q4;
            }
          } else {
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
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_afull(st));
        if (r == 0) SRK_TRACE_EV(17, it);
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
  SRK_REQUIRE(L::kTotal <= h->smem_optin, "conv_first_tc: needs %d B smem, device allows %d", L::kTotal, h->smem_optin);
  if (first_use(h, reinterpret_cast<const void*>(&conv_gather_tc_kernel<KS, CIN>)))
    SRK_CHECK_CUDA(cudaFuncSetAttribute(conv_gather_tc_kernel<KS, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
  ConvTcParams& p = gp.tc;
  p.num_tiles = int((p.rows_valid + 127) / 128);
  if (int rc = make_tensor_map_2d(h, &p.map_w, w_packed, uint64_t(L::kBlocks * 64), 64, 64)) return rc;
  if (int rc = make_tensor_map_2d(h, &p.map_out, y_fpa, uint64_t(p.rows_valid), 64, 128)) return rc;
  const int grid = p.num_tiles < h->num_sms ? p.num_tiles : h->num_sms;
  SRK_CHECK_CUDA(launch_pdl(conv_gather_tc_kernel<KS, CIN>, dim3(grid), dim3(L::kThreads), L::kTotal, stream, gp));
  return 0;
}

// --------------------------------------------------------------------------------------- host side
template <int CIN, int NP, int KS, int EPI>
static int launch_conv_tc(srk_ctx* h, ConvTcParams& p, const void* x, const void* w_packed, void* y_fpa, cudaStream_t stream) {
  using L = ConvTcCfg<CIN, NP, KS, EPI>;
  SRK_REQUIRE(L::kTotal <= h->smem_optin, "conv_tc: needs %d B smem, device allows %d", L::kTotal, h->smem_optin);
  if (first_use(h, reinterpret_cast<const void*>(&conv_tc_kernel<CIN, NP, KS, EPI>)))
    SRK_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<CIN, NP, KS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
  p.num_tiles = int((p.rows_valid + L::kTileStride - 1) / L::kTileStride);
  constexpr int kChunkRows = L::kChunkRows;
  const int span_chunks = (2 * L::kHalo * p.Wp + 128 + kChunkRows - 1) / kChunkRows + 1;
  SRK_REQUIRE(span_chunks + 2 <= L::kRingSlots,
              "conv_tc: image width %d too large for the %dx%d flat-stream kernel (a tile needs %d chunks resident, ring holds %d); "
              "split the frame into column panels", p.W, KS, KS, span_chunks, L::kRingSlots);
  if (int rc = make_tensor_map_2d(h, &p.map_in, x, uint64_t(p.rows_valid), CIN, kChunkRows)) return rc;
  if (int rc = make_tensor_map_2d(h, &p.map_w, w_packed, uint64_t(KS * KS * NP), CIN, NP)) return rc;
  if (EPI == EPI_FPA) {
    if (int rc = make_tensor_map_2d(h, &p.map_out, y_fpa, uint64_t(p.rows_valid), NP, L::kTileStride)) return rc;
  }
  const int grid = p.num_tiles < h->num_sms ? p.num_tiles : h->num_sms;
  SRK_CHECK_CUDA(launch_pdl(conv_tc_kernel<CIN, NP, KS, EPI>, dim3(grid), dim3(L::kThreads), L::kTotal, stream, p));
  return 0;
}

static int fill_geom(ConvTcParams& p, int n_img, int H, int W) {
#ifdef SRK_TRACE
  {
    const char* e = getenv("SRK_DBG");  // development builds only: epilogue ablation switches
    p.dbg = e ? atoi(e) : 0;
  }
#else
  p.dbg = 0;
#endif
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

#ifdef SRK_TRACE
extern "C" int srk_debug_trace_read(unsigned long long* host_dst) {  // development builds only, not part of include/srk.h
  return int(cudaMemcpyFromSymbol(host_dst, g_trace, sizeof(g_trace)));
}
#endif

extern "C" int srk_conv_tc(srk_handle_t h, const void* x_fpa, int cin_p, const void* w_packed, const float* bias, int k,
                           int cout_p, int act, int n_img, int H, int W, void* y_fpa, const void* mask_src,
                           int mask_kind, const void* addend_fpa, int relu_after_add, srk_stream_t stream) {
  SRK_REQUIRE(h && x_fpa && w_packed && y_fpa, "srk_conv_tc: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  ConvTcParams p{};
  if (int rc = fill_geom(p, n_img, H, W)) return rc;
  p.bias = bias;
  p.act = act;
  p.mask_src = static_cast<const __nv_bfloat16*>(mask_src);
  p.mask_kind = mask_kind;
  p.addend_fpa = static_cast<const __nv_bfloat16*>(addend_fpa);
  p.relu_after_add = relu_after_add;
  cudaStream_t s = as_stream(stream);
  // wide frames: the column-strip form of the plain 3x3 64->64 layer (no lane shift in the epilogue, conv_strip.cu)
  if (cin_p == 64 && cout_p == 64 && k == 3 && (!mask_src || mask_kind == SRK_ACT_RELU) && !addend_fpa && (act == SRK_ACT_NONE || act == SRK_ACT_RELU) &&
      (h->conv_form == SRK_CONV_FORM_STRIP || (h->conv_form == SRK_CONV_FORM_AUTO && conv_strip_applicable(h, n_img, H, W))))
    return launch_conv_strip(h, x_fpa, w_packed, bias, act, n_img, H, W, y_fpa, mask_src, s);
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
  if (int rc_dev = check_device(h)) return rc_dev;
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
  SRK_CASE(64, 16, 5)
  SRK_CASE(32, 64, 3)
  SRK_CASE(64, 32, 3)
  SRK_CASE(64, 64, 3)  // ESPCN 4x RGB training: 48 outputs from the 64-wide (zero-padded) f2 activations
#undef SRK_CASE
  set_error("srk_conv_tc_last: unsupported (cin_p=%d, cout_p=%d, k=%d)", cin_p, cout_p, k);
  return -1;
}

extern "C" int srk_conv_first_tc(srk_handle_t h, const float* x, int n_frames, int FH, int FW, int cin, const void* w_packed,
                                 const float* bias, int k, int pad_mode, int act, const srk_panel* panels, int n_img, int H, int W,
                                 void* y_fpa, const void* mask_src, int mask_kind, srk_stream_t stream) {
  SRK_REQUIRE(h && x && w_packed && y_fpa, "srk_conv_first_tc: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
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

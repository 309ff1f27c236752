// Whole-ESPCN inference in ONE persistent kernel for sm_100a: f1 (5x5 C->64, tanh) -> f2 (3x3 64->32, tanh) -> f3 (3x3 32->C*r^2)
// -> depth_to_space, with the two intermediate activations living only in TENSOR MEMORY.
//   replaces the three tf.nn.conv2d + bias_add + tanh of espcn/espcn/model_espcn.py:117-134 and the host un-pack of
//   espcn/espcn/experiment_test.py:173-177 (optionally also its saturate_cast to uint8, :179-184).
//
// Why: layer by layer the model moves 3.5 GB per 4 x 1080p-LR frames (the 64- and 32-channel activations are written and read
// back); fused it moves the input (4 B / LR pixel) and the output only, and the bound becomes the SM itself (tensor pipe,
// MUFU.TANH and issue slots, roughly balanced).
//
// Geometry.  A CTA walks COLUMN STRIPS of 120 output pixels from top to bottom, one LR row per step.  TMEM lane i of a step is
// the pixel x = 120*strip - 4 + i of the row being processed (128 lanes: 120 stored + a 4-pixel apron each side, which is the
// receptive field the three layers eat: 1 lane per 3x3 layer through the lane-shift add below, and the 5x5 layer is gathered
// for all 128 lanes).  Per virtual row v (rows of one strip segment, +4 rows of vertical apron) three GEMM stages run, each one
// row behind the stage that feeds it:
//   S1  a1[v]  = tanh(im2col(lr rows y-2..y+2) x W1 + b1)       A (im2col, bf16) built in smem by the gather warps, SS-mode MMA
//                                                               128 x 64 x (25C+1): the bias rides in a spare K slot (A column of
//                                                               ones x a W1 row holding b1); D1 -> EPI1 -> bf16 into TMEM ring A1
//   S2  a2[v]  = tanh(sum_dy a1[v-2+dy] x W2[dy] + b2)          TS-mode MMA: A = a1 rows straight from TMEM, B = the three taps
//                                                               (dy, -1..1) stacked along N (N = 96); D2 block dx holds the
//                                                               contribution of input lane i to output lane i - dx, so EPI2 adds
//                                                               the blocks with a LANE SHIFT (as conv_tc.cu) -> bf16 into ring A2
//   S3  out[v] = sum_dy a2[v-2+dy] x W3[dy] + b3                TS-mode MMA, N = 3*NP3; EPI3 lane-shift adds, transposes through
//                                                               shared memory and stores the r HR rows with 16-byte stores
// Rows/columns outside the image are written as ZERO activations (SAME padding pads every layer's input with zeros, not with
// tanh(bias)).  Segment boundaries need no special casing: every stage runs for every virtual row and results whose inputs
// straddle two segments are simply never consumed (masked in EPI2, not stored in EPI3).
//
// Warp roles (4S+8 warps, S = 4): warps 0..4S-1 epilogue in S SETS of four quadrant warps (warp % 4 = TMEM lane quadrant) that take the
// virtual rows round-robin; a set runs EPI3(v-2S), EPI2(v-S), EPI1(v) back to back (oldest stage first: what it waits for never
// depends on its own later work), so an epilogue warp is busy most of the time and the only
// mbarrier waits it performs are on tensor-pipe commits.  (The first version gave every stage its own warps: 24 of 28 warps sat
// in try_wait loops, woken by every barrier event of the CTA; their polling was 70 % of all issued instructions and seven
// distinct instruction streams per scheduler thrashed the instruction cache -- profiles/r2_ncu_espcn_fused_v1.txt.)
// then 4 gather warps (cp.async input ring -> im2col), one MMA issuer warp per stage, one set-up warp (TMEM allocation, weights by TMA).
// Hand-over (all indexed by the virtual row v; the issuers do ONE wait and ONE commit per row -- every mbarrier operation of a
// control warp costs ~200 cycles while the epilogue warps are busy, profiles/r1_trace_conv_tc.txt):
//   G1[v%2] <- gather(v) built A  +  EPI1(v-1) drained D1     => MMA1(v)   commit C1[v%2]  (EPI1(v) reads D1; gather(v+2) reuses A)
//   G2[v%2] <- EPI1(v) wrote a1[v] + EPI2(v-ND2) drained D2   => MMA2(v)   commit C2[v%4]  (EPI2(v) reads D2; EPI1 reuses a1 slot)
//   G3[v%2] <- EPI2(v) wrote a2[v] + EPI3(v-1) drained D3     => MMA3(v)   commit C3[v%4]  (EPI3(v) reads D3; EPI2 reuses a2 slot)
// Measured building blocks: profiles/r2_ts_probe.log (TS-mode layout and rate: 49 cycles per M=128 K=16 instruction for N <= 96);
// the in-kernel timelines that shaped the epilogues: profiles/r2_trace_espcn_fused.txt.
#pragma once
#include <algorithm>

#include "sm100_ptx.cuh"
#include "srk_common.cuh"
#include "strip_walk.cuh"

namespace srk {

// Development-only timeline (-DSRK_TRACE build, tools/trace_espcn.py): CTA 0 records clock64() at the hand-over points of
// every role for its first 256 virtual rows.
#ifdef SRK_TRACE
static __device__ unsigned long long g_ef_trace[24 * 256];  // one copy per translation unit
#define EF_EV(ev, v)                                                            \
  do {                                                                          \
    if (blockIdx.x == 0 && (v) < 256) g_ef_trace[(ev) * 256 + (v)] = clock64(); \
  } while (0)
#else
#define EF_EV(ev, v) \
  do {               \
  } while (0)
#endif

#ifndef SRK_EF_SETS
#define SRK_EF_SETS 4
#endif

constexpr int kEfStripW = 120;  // stored pixels per strip
constexpr int kEfLane0 = 4;     // lane of the strip's first stored pixel (a multiple of 4 keeps warp segments 16-byte aligned)

struct alignas(64) EspcnFusedParams {
  CUtensorMap map_w1;  // [kB1*64][64]  box {64, 64}    SRK_PACK_FIRST
  CUtensorMap map_w2;  // [9*32][64]    box {64, 32}    SRK_PACK_FWD np 32, cinp 64
  CUtensorMap map_w3;  // [9*NP3][32]   box {32, NP3}   SRK_PACK_FWD np NP3, cinp 32
  const float* lr;     // fp32 NHWC [n][H][W][C]
  const float* b1;     // [64]
  const float* b2;     // [32]
  const float* b3;     // [cout]
  void* out;           // fp32 or uint8 [n][H*r][W*r][C] (shuffled) or [n][H][W][C*r^2] (packed)
  int n, H, W;
  int y0, hb;          // rows [y0, y0+hb) of every frame are produced (row-band sharding across GPUs)
  int strips;          // per frame
  long long units;     // n * strips * hb   (one unit = one LR row of one strip)
  int out_kind;        // SRK_OUT_F32 | SRK_OUT_U8
};

template <int C, int R, bool SHUF>
struct EspcnCfg {
  static constexpr int kCout = C * R * R;
  static constexpr int kSets = (C == 1) ? 4 : ((kCout > 32) ? 2 : 3);  // epilogue sets (4 quadrant warps each) taking virtual rows round-robin
  static constexpr int NP3 = (kCout + 15) / 16 * 16;
  static constexpr int kRows = SHUF ? R : 1;          // output rows per LR row
  static constexpr int kRC = kCout / kRows;           // output elements per LR pixel per output row
  static constexpr int kK1 = 25 * C;                  // im2col depth; K slot kK1 carries the bias
  static constexpr int kK1Steps = (kK1 + 1 + 15) / 16;
  static constexpr int kB1 = (kK1 + 1 + 63) / 64;
  // Y-channel form (C == 1, the BASELINE configuration): f2 and f3 read their A operands (a1, a2) from SHARED memory (SS mode)
  // so that the horizontal taps are row-shifted operand descriptors instead of a lane-shift in the epilogue, and stack the
  // three VERTICAL taps along N into a rotating window of three accumulators (see MMA2 below); the biases are one more MMA
  // each.  The epilogues shrink to load / tanh / pack / store and get DEDICATED warps per stage (two EPI1 sets alternating
  // rows, four EPI2 warps, four EPI3 warps), because a rotating window is single-buffered: the next row's MMAs start when the
  // finished row has been read out, so the read-out must not queue behind other work.  RGB keeps the TS-mode form (its im2col
  // tile and transpose buffers leave no room for the a1 / a2 rings).
  static constexpr bool kSS = (C == 1);
  static constexpr int kND1 = kSS ? 2 : 1;           // D1 accumulators
  static constexpr int kNA1S = 3, kNA2S = 3;         // a1 / a2 ring slots in shared memory (kSS)
  static constexpr int kA1Bytes = 128 * 128, kA2Bytes = 128 * 64;
  static constexpr int kW1Bytes = kB1 * 64 * 128;
  static constexpr int kW2Blocks = kSS ? 15 : 9;     // kSS: per dx the blocks [dy2 dy1 dy0 dy2 dy1] (every rotation is a window of 3)
  static constexpr int kW2Bytes = kW2Blocks * 32 * 128;
  static constexpr int kW3Bytes = kW2Blocks * NP3 * 64;
  static constexpr int kAS = 2;                       // im2col stages
  static constexpr int kABytes = kB1 * 128 * 128;
  static constexpr int kInElems = 132 * C;            // one input row of the strip: x = 120*s - 6 .. 120*s + 125
  static constexpr int kInRowBytes = (kInElems * 4 + 15) / 16 * 16;
  static constexpr int kInSlots = (C == 1) ? 32 : 20, kPrefetch = (C == 1) ? 12 : 10;  // input ring: 5 rows in use + 12 in flight + 5 (segment jump) < 32
  // tensor memory columns
  static constexpr int kN2 = 96, kN3 = 3 * NP3;
  static constexpr int kND2 = kSS ? 1 : ((NP3 <= 32) ? 2 : 1);
  static constexpr int kNA1 = kSS ? 0 : ((NP3 == 32) ? 3 : 4);
  static constexpr int kColD1 = 0;
  static constexpr int kColA1 = 64 * kND1;
  static constexpr int kColD2 = kColA1 + kNA1 * 32;
  static constexpr int kColA2 = kColD2 + kND2 * kN2;
  static_assert(!kSS || (kK1 + 1 <= 32 && NP3 == 16), "kSS: the bias K-steps of the first-layer tiles must be free; 1 KB weight blocks");
  static constexpr int kColD3 = kColA2 + (kSS ? 0 : 4 * 16);
  static_assert(kColD3 + kN3 <= 512, "tensor memory plan does not fit");
  // shared memory
  static constexpr int kTBufBytes = 32 * kCout * 4;   // per EPI3 warp: its 32 pixels' outputs in OUTPUT order [row][pixel*kRC + e]
  static constexpr int kXchBytes = 2 * 4 * 2 * 16 * 4;  // [parity][quadrant][block][16 columns]
  static constexpr int kOffW1 = 0;
  static constexpr int kOffW2 = kOffW1 + kW1Bytes;
  static constexpr int kOffW3 = kOffW2 + kW2Bytes;
  static constexpr int kOffA = (kOffW3 + kW3Bytes + 1023) / 1024 * 1024;
  static constexpr int kOffA1 = kOffA + kAS * kABytes;  // 1024-aligned (kABytes is a multiple of 16 KB); rows -1 / 128 of a slot are its neighbours' memory
  static constexpr int kOffA2 = kOffA1 + (kSS ? kNA1S * kA1Bytes : 0);
  static constexpr int kOffIn = kOffA2 + (kSS ? kNA2S * kA2Bytes : 0);
  static constexpr int kOffTBuf = kOffIn + kInSlots * kInRowBytes;
  static constexpr int kOffXch = kOffTBuf + 4 * (kSS ? 1 : kSets) * kTBufBytes;    // one lane-exchange area per epilogue set
  static constexpr int kOffBias = kOffXch + kSets * kXchBytes;     // b2[32] b3[NP3]
  static constexpr int kOffTab = kOffBias + (32 + NP3) * 4;   // segment table
  static constexpr int kOffBars = (kOffTab + kEfTabInts * 4 + 7) / 8 * 8;
  static constexpr int kNumBars = 1 + kAS + 4 + 4 + 4 + 4 + 4;
  static constexpr int kOffTmemSlot = kOffBars + kNumBars * 8;
  static constexpr int kUsed = kOffTmemSlot + 16 + 1024;
  static constexpr int kTotal = kUsed < 120 * 1024 ? 120 * 1024 : kUsed;  // > half an SM's shared memory: one CTA (one 512-column TMEM owner) per SM
  static constexpr int kThreads = (4 * kSets + 8) * 32;
  static_assert(kTotal <= 227 * 1024, "shared-memory plan does not fit");
};

// ---- small helpers
__device__ __forceinline__ float ef_tanh(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ uint32_t ef_pack(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void ef_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ float ef_lds(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void ef_sts(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void ef_sts128(uint32_t a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void ef_sts128u(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ float4 ef_lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t ef_u8(float v) {  // tf.saturate_cast(v * 127.5 + 127.5, uint8): clamp, then truncate
  const float q = __fadd_rn(__fmul_rn(v, 127.5f), 127.5f);
  return uint32_t(fminf(fmaxf(q, 0.f), 255.f));
}

// t[j] = D[j-1][block 0] + D[j+1][block 2] over the 128 lanes of a tile for NC live columns: the edge lanes of each quadrant
// warp swap in the neighbouring quadrant's row through shared memory (named barrier `bar`, 128 threads), then one rotating
// shuffle per column serves every lane.  `xq` = the group's exchange area [quadrant][block][16], already offset by the parity
// in use.  Plain shared-memory accesses and shuffles issued back to back: the compiler is free to overlap their latencies.
template <int NC>
__device__ __forceinline__ void ef_lane_shift(float (&b0)[16], float (&b2)[16], float (&t)[16], float* xq, int quad, int lane, int bar) {
  constexpr int NQ = (NC + 3) / 4;
  if (lane == 31) {
#pragma unroll
    for (int c = 0; c < NQ; ++c) *reinterpret_cast<float4*>(xq + (quad * 2 + 0) * 16 + 4 * c) = make_float4(b0[4 * c], b0[4 * c + 1], b0[4 * c + 2], b0[4 * c + 3]);
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < NQ; ++c) *reinterpret_cast<float4*>(xq + (quad * 2 + 1) * 16 + 4 * c) = make_float4(b2[4 * c], b2[4 * c + 1], b2[4 * c + 2], b2[4 * c + 3]);
  }
  ef_bar_sync(bar, 128);
  if (lane == 31 && quad > 0) {
#pragma unroll
    for (int c = 0; c < NQ; ++c) {
      const float4 o = *reinterpret_cast<const float4*>(xq + ((quad - 1) * 2 + 0) * 16 + 4 * c);
      b0[4 * c] = o.x, b0[4 * c + 1] = o.y, b0[4 * c + 2] = o.z, b0[4 * c + 3] = o.w;
    }
  }
  if (lane == 0 && quad < 3) {
#pragma unroll
    for (int c = 0; c < NQ; ++c) {
      const float4 o = *reinterpret_cast<const float4*>(xq + ((quad + 1) * 2 + 1) * 16 + 4 * c);
      b2[4 * c] = o.x, b2[4 * c + 1] = o.y, b2[4 * c + 2] = o.z, b2[4 * c + 3] = o.w;
    }
  }
  const int up = (lane + 31) & 31, dn = (lane + 1) & 31;
  float l[NC], r[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) l[c] = __shfl_sync(0xffffffffu, b0[c], up);
#pragma unroll
  for (int c = 0; c < NC; ++c) r[c] = __shfl_sync(0xffffffffu, b2[c], dn);
#pragma unroll
  for (int c = 0; c < NC; ++c) t[c] = l[c] + r[c];
}

// Element-by-element store of the valid part [lo, hi) of a 4-element group (out of line: rare, and the hot loops must stay small).
static __device__ __noinline__ void ef_store_partial(void* out, bool u8, int64_t idx, float4 val, int lo, int hi) {
  const float vv[4] = {val.x, val.y, val.z, val.w};
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    if (i >= lo && i < hi) {
      if (u8) static_cast<uint8_t*>(out)[idx + i] = uint8_t(ef_u8(vv[i]));
      else static_cast<float*>(out)[idx + i] = vv[i];
    }
  }
}

template <int C, int R, bool SHUF>
__global__ void __launch_bounds__(EspcnCfg<C, R, SHUF>::kThreads, 1) espcn_fused_kernel(const __grid_constant__ EspcnFusedParams p) {
  using L = EspcnCfg<C, R, SHUF>;
  constexpr int S = L::kSets, W0 = 4 * S;  // W0 = first non-epilogue warp
  constexpr int AS = L::kAS, NA1 = L::kNA1, ND2 = L::kND2, NR = L::kInSlots, NP3 = L::NP3, COUT = L::kCout;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET from the shared array (not by rounding a generic pointer): the compiler keeps the address space
  // and emits LDS/STS instead of generic loads with their descriptor set-up
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t s_base = smem_u32(smem);
  int* const s_tab = reinterpret_cast<int*>(smem + L::kOffTab);
  const uint32_t s_w1 = s_base + L::kOffW1, s_w2 = s_base + L::kOffW2, s_w3 = s_base + L::kOffW3;
  const uint32_t s_a = s_base + L::kOffA, s_a1 = s_base + L::kOffA1, s_a2 = s_base + L::kOffA2, s_in = s_base + L::kOffIn;
  float* s_bias = reinterpret_cast<float*>(smem + L::kOffBias);
  const uint32_t s_bars = s_base + L::kOffBars;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmemSlot);
  const uint32_t bar_w = s_bars;
  auto G1 = [&](int i) { return s_bars + 8u * (1 + i); };
  // Commit barriers are rings of FOUR indexed by v & 3: a waiter for row v must know that the ring slot's previous phase (row
  // v - 4) is complete, or its parity test passes early; every waiter's preceding wait was for a row >= v - S, S <= 4.
  auto C1 = [&](int i) { return s_bars + 8u * (1 + AS + i); };
  auto G2 = [&](int i) { return s_bars + 8u * (1 + AS + 4 + i); };  // two in use (TS form) or four (kSS: EPI1 runs up to three rows ahead of MMA2)
  auto C2 = [&](int i) { return s_bars + 8u * (1 + AS + 8 + i); };
  auto G3 = [&](int i) { return s_bars + 8u * (1 + AS + 12 + i); };  // likewise two or four
  auto C3 = [&](int i) { return s_bars + 8u * (1 + AS + 16 + i); };
  static_assert(L::kSets <= 4, "commit-barrier rings hold four phases");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t u0 = uint32_t((p.units * blockIdx.x) / gridDim.x), u1 = uint32_t((p.units * (blockIdx.x + 1)) / gridDim.x);

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int i = 0; i < AS; ++i) {
      mbar_init(G1(i), 4 + 4);  // one arrival per warp (every arrival wakes every sleeping waiter of the CTA): gather + one epilogue set
    }
    for (int i = 0; i < 4; ++i) mbar_init(G2(i), 4 + 4);  // EPI1 of one set + EPI2 (drain) of a set
    for (int i = 0; i < 4; ++i) mbar_init(G3(i), 4 + 4);  // EPI2 + EPI3 (drain)
    for (int i = 0; i < 4; ++i) {
      mbar_init(C1(i), 1);
      mbar_init(C2(i), 1);
      mbar_init(C3(i), 1);
    }
    fence_mbar_init();
  }
  if (warp == W0 + 7) {
    tmem_alloc<512>(smem_u32(tmem_slot));
    if (lane == 0) {
      tma_prefetch_desc(&p.map_w1);
      tma_prefetch_desc(&p.map_w2);
      tma_prefetch_desc(&p.map_w3);
    }
  }
  pdl_wait();
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 32 + NP3; i += blockDim.x) s_bias[i] = (i < 32) ? p.b2[i] : (i - 32 < COUT ? p.b3[i - 32] : 0.f);
  if (threadIdx.x == 32) ef_build_segments(s_tab, u0, u1, p.hb, p.y0, p.strips, 4);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int V = s_tab[s_tab[71]];  // virtual rows of this CTA

  if (V > 0) {
    if (warp == W0 + 7) {
      // ---------------------------------------------------------------- weights: one TMA burst, resident for the whole kernel
      if (lane == 0) {
        mbar_arrive_expect_tx(bar_w, L::kW1Bytes + L::kW2Bytes + L::kW3Bytes);
        for (int b = 0; b < L::kB1; ++b) tma_load_2d(s_w1 + b * 8192, &p.map_w1, 0, b * 64, bar_w);
        if constexpr (L::kSS) {
          for (int dx = 0; dx < 3; ++dx)
            for (int j = 0; j < 5; ++j) tma_load_2d(s_w2 + (dx * 5 + j) * 4096, &p.map_w2, 0, (((j < 3) ? 2 - j : 5 - j) * 3 + dx) * 32, bar_w);
        } else {
          for (int t = 0; t < 9; ++t) tma_load_2d(s_w2 + t * 4096, &p.map_w2, 0, t * 32, bar_w);
        }
        if constexpr (L::kSS) {
          for (int dx = 0; dx < 3; ++dx)
            for (int j = 0; j < 5; ++j) tma_load_2d(s_w3 + (dx * 5 + j) * NP3 * 64, &p.map_w3, 0, (((j < 3) ? 2 - j : 5 - j) * 3 + dx) * NP3, bar_w);
        } else {
          for (int t = 0; t < 9; ++t) tma_load_2d(s_w3 + t * NP3 * 64, &p.map_w3, 0, t * NP3, bar_w);
        }
      }
#ifdef SRK_TRACE
      // timeline only: lanes 0..2 watch the commit barriers of the three MMA stages and record when each row's MMAs completed
      if (lane < 3 && blockIdx.x == 0) {
        const uint32_t c0 = lane == 0 ? C1(0) : (lane == 1 ? C2(0) : C3(0));
        for (int v = 0; v < 256; ++v) {
          mbar_wait(c0 + 8u * (v & 3), (v >> 2) & 1);
          EF_EV(21 + lane, v);
        }
      }
#endif
    } else if (warp == W0 + 4) {
      // ---------------------------------------------------------------- MMA1: a1 accumulator = im2col x W1 (SS mode)
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
      constexpr uint64_t hi = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
      mbar_wait(bar_w, 0);
      {
        // the bias rides in K slot kK1 of the packed kernel (zero padding until now): W1[co][kK1] = b1[co] (SW128 K-major)
        constexpr int kb = L::kK1 / 64, kk = L::kK1 % 64;
        for (int co = lane; co < 64; co += 32) {
          const __nv_bfloat16 bv = __float2bfloat16_rn(p.b1[co]);
          *reinterpret_cast<__nv_bfloat16*>(smem + L::kOffW1 + kb * 8192 + co * 128 + (((kk / 8) ^ (co & 7)) << 4) + (kk % 8) * 2) = bv;
        }
        if constexpr (L::kSS) {
          // f2's bias is one more MMA (K-step 2 of these tiles: A = the im2col tile's constant columns 32, 33 = 1, B = rows
          // n < 32 of this tile with b2[n] split into bf16 high and low parts at columns 32, 33: exact to 2^-17)
          const float b = p.b2[lane];
          const __nv_bfloat16 bh = __float2bfloat16_rn(b), bl = __float2bfloat16_rn(b - __bfloat162float(bh));
          __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(smem + L::kOffW1 + lane * 128 + ((4 ^ (lane & 7)) << 4));
          d[0] = bh;
          d[1] = bl;
          if (lane < NP3) {  // f3's bias likewise: K-step 3, columns 48, 49, rows n < NP3
            const float b3v = lane < COUT ? p.b3[lane] : 0.f;
            const __nv_bfloat16 ch = __float2bfloat16_rn(b3v), cl = __float2bfloat16_rn(b3v - __bfloat162float(ch));
            __nv_bfloat16* d3 = reinterpret_cast<__nv_bfloat16*>(smem + L::kOffW1 + lane * 128 + ((6 ^ (lane & 7)) << 4));
            d3[0] = ch;
            d3[1] = cl;
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
      }
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        const int st = v % AS;
        mbar_wait(G1(st), (v / AS) & 1);
        tc_fence_after();
        if (lane == 0) EF_EV(3, v);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < L::kK1Steps; ++k) {
            const uint32_t a_addr = s_a + st * L::kABytes + (k / 4) * (128 * 128) + (k % 4) * 32;
            const uint32_t b_addr = s_w1 + (k / 4) * (64 * 128) + (k % 4) * 32;
            umma_bf16(tmem + L::kColD1 + (v % L::kND1) * 64, umma_desc(hi, a_addr), umma_desc(hi, b_addr), idesc, k != 0);
          }
          umma_commit(C1(v & 3));
        }
        __syncwarp();
        if (lane == 0) EF_EV(4, v);
      }
    } else if (warp == W0 + 5) {
      // ---------------------------------------------------------------- MMA2: a2 accumulator = sum_dy a1[v-2+dy] (TMEM) x W2 row dy
      constexpr uint32_t idesc = umma_idesc_bf16(128, L::kN2, 0, 0);
      constexpr uint64_t hi = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
      mbar_wait(bar_w, 0);
      if constexpr (L::kSS) {
        // Row u of a1 (shared memory, SW128 K-major, one 128-byte row per lane) feeds the three output rows v = u, u+1, u+2
        // (vertical tap dy = u + 2 - v).  Their accumulators are the three 32-column slots of ONE 96-column window, row v in
        // slot v % 3, so a single N = 96 instruction per (dx, K-step) serves all three; which tap each slot needs rotates
        // with u % 3, and the weight blocks are stored [dy2 dy1 dy0 dy2 dy1] so that every rotation is a contiguous window.
        // The horizontal tap dx is a ROW-SHIFTED A descriptor (start address -/+ 128 bytes: lane j reads pixel j + dx - 1;
        // the two rows beyond the slot are neighbouring memory and only reach the discarded apron lanes 0 and 127).
        // Row u+2 is opened by the bias instruction (accumulate off), closed by row u+2's own instructions.
        constexpr uint32_t idesc32 = umma_idesc_bf16(128, 32, 0, 0);
        int m3 = 0, sl = 0;  // v % 3, v % kNA1S
#pragma unroll 1
        for (int v = 0; v < V; ++v) {
          mbar_wait(G2(v & 3), (v >> 2) & 1);
          tc_fence_after();
          if (lane == 0) EF_EV(9, v);
          const uint32_t a0 = s_a1 + sl * L::kA1Bytes - 128;
          const uint32_t b0 = s_w2 + ((2 * m3) % 3) * 4096;
          const uint32_t dinit = tmem + L::kColD2 + ((m3 + 2) % 3) * 32;
          if (elect_one()) {
            umma_bf16(dinit, umma_desc(hi, s_a + 64), umma_desc(hi, s_w1 + 64), idesc32, 0);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16(tmem + L::kColD2, umma_desc(hi, a0 + dx * 128 + k * 32), umma_desc(hi, b0 + dx * 5 * 4096 + k * 32), idesc, 1);
            }
            umma_commit(C2(v & 3));
          }
          __syncwarp();
          if (lane == 0) EF_EV(10, v);
          m3 = (m3 == 2) ? 0 : m3 + 1;
          sl = (sl == L::kNA1S - 1) ? 0 : sl + 1;
        }
      } else {
#pragma unroll 1
        for (int v = 0; v < V; ++v) {
          mbar_wait(G2(v & 1), (v >> 1) & 1);
          tc_fence_after();
          if (lane == 0) EF_EV(9, v);
          const uint32_t d = tmem + L::kColD2 + (v % ND2) * L::kN2;
          if (elect_one()) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint32_t a = tmem + L::kColA1 + ((v + dy + 2 * NA1 - 2) % NA1) * 32;
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_bf16_ts(d, a + k * 8, umma_desc(hi, s_w2 + dy * 3 * 4096 + k * 32), idesc, (dy | k) != 0);
            }
            umma_commit(C2(v & 3));
          }
          __syncwarp();
          if (lane == 0) EF_EV(10, v);
        }
      }
    } else if (warp == W0 + 6) {
      // ---------------------------------------------------------------- MMA3: out accumulator = sum_dy a2[v-2+dy] (TMEM) x W3 row dy
      constexpr uint32_t idesc = umma_idesc_bf16(128, L::kN3, 0, 0);
      constexpr uint64_t hi = umma_desc_hi(0, 512, UMMA_LAYOUT_SW64);
      mbar_wait(bar_w, 0);
      if constexpr (L::kSS) {
        // as MMA2: a2 rows in shared memory (SW64 K-major, 64-byte rows), window of three NP3-column accumulators, bias instruction
        constexpr uint32_t idesc16 = umma_idesc_bf16(128, NP3, 0, 0);
        constexpr uint64_t hi128 = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
        int m3 = 0, sl = 0;
#pragma unroll 1
        for (int v = 0; v < V; ++v) {
          mbar_wait(G3(v & 3), (v >> 2) & 1);
          tc_fence_after();
          if (lane == 0) EF_EV(15, v);
          const uint32_t a0 = s_a2 + sl * L::kA2Bytes - 64;
          const uint32_t b0 = s_w3 + ((2 * m3) % 3) * (NP3 * 64);
          const uint32_t dinit = tmem + L::kColD3 + ((m3 + 2) % 3) * NP3;
          if (elect_one()) {
            umma_bf16(dinit, umma_desc(hi128, s_a + 96), umma_desc(hi128, s_w1 + 96), idesc16, 0);
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
              for (int k = 0; k < 2; ++k)
                umma_bf16(tmem + L::kColD3, umma_desc(hi, a0 + dx * 64 + k * 32), umma_desc(hi, b0 + dx * 5 * NP3 * 64 + k * 32), idesc, 1);
            }
            umma_commit(C3(v & 3));
          }
          __syncwarp();
          if (lane == 0) EF_EV(16, v);
          m3 = (m3 == 2) ? 0 : m3 + 1;
          sl = (sl == L::kNA2S - 1) ? 0 : sl + 1;
        }
      } else {
#pragma unroll 1
        for (int v = 0; v < V; ++v) {
          mbar_wait(G3(v & 1), (v >> 1) & 1);
          tc_fence_after();
          if (lane == 0) EF_EV(15, v);
          if (elect_one()) {
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint32_t a = tmem + L::kColA2 + ((v + dy + 2) & 3) * 16;
#pragma unroll
              for (int k = 0; k < 2; ++k) umma_bf16_ts(tmem + L::kColD3, a + k * 8, umma_desc(hi, s_w3 + dy * 3 * NP3 * 64 + k * 32), idesc, (dy | k) != 0);
            }
            umma_commit(C3(v & 3));
          }
          __syncwarp();
          if (lane == 0) EF_EV(16, v);
        }
      }
    } else if (warp >= W0 && warp < W0 + 4) {
      // ---------------------------------------------------------------- gather: cp.async input ring -> im2col rows of A (bf16, SW128)
      const int gt = threadIdx.x - W0 * 32;  // row of the A tile = TMEM lane = pixel x = 120*s - 4 + gt
      // input rows form their own sequence: segment i owns the input rows [v0_i + 4i, v1_i + 4(i+1)), image row ya - 4 + j
      EfSeg pf, w;
      ef_seg_load(pf, s_tab, 0);
      ef_seg_load(w, s_tab, 0);
      const int total_in = V + 4 * s_tab[71];
      int issued = 0;
      auto issue_row = [&]() {
        const int slot = issued % NR;
        while (issued >= pf.v1 + 4 * (pf.i + 1)) ef_seg_load(pf, s_tab, pf.i + 1);
        const int y = pf.ya - 4 + (issued - (pf.v0 + 4 * pf.i));
        const int xe0 = (pf.s * kEfStripW - kEfLane0 - 2) * C;  // element index (x*C + ci) of ring element 0
        const bool yok = (y >= 0) && (y < p.H);
        const float* row = p.lr + (int64_t(pf.n) * p.H + (yok ? y : 0)) * int64_t(p.W) * C;
        const uint32_t dst = s_in + slot * L::kInRowBytes;
#pragma unroll
        for (int i = 0; i < (L::kInElems + 127) / 128; ++i) {
          const int e = gt + 128 * i;
          if (e < L::kInElems) {
            const int ge = xe0 + e;
            const bool ok = yok && ge >= 0 && ge < p.W * C;
            cp_async_4(dst + e * 4, row + (ok ? ge : 0), ok ? 4u : 0u);
          }
        }
        ++issued;
      };
      // One cp.async group per step.  After step v's issue the rows up to need(v+1) + kPrefetch are under way, and need() grows
      // by at most 9 over five steps (one row per step, five at a segment start, segments last >= 5 steps), so the rows step v
      // reads were committed by step v-6 at the latest: wait_group<5> + the group barrier makes them visible.
      static_assert(L::kPrefetch >= 9 && L::kInSlots >= L::kPrefetch + 10, "input ring look-ahead");
      if constexpr (L::kSS) {
        // constant columns of the im2col tiles (K-steps 2, 3; the gather only ever writes K-steps 0, 1): A[.][32] = A[.][33] = 1
        // is the A operand of the bias instructions (bf16 high + low part of the bias in the B rows)
        for (int st = 0; st < AS; ++st) {
          const uint32_t arow = s_a + st * L::kABytes + gt * 128;
          ef_sts128u(arow + ((4 ^ (gt & 7)) << 4), 0x3F803F80u, 0u, 0u, 0u);
          ef_sts128u(arow + ((5 ^ (gt & 7)) << 4), 0u, 0u, 0u, 0u);
          ef_sts128u(arow + ((6 ^ (gt & 7)) << 4), 0x3F803F80u, 0u, 0u, 0u);
          ef_sts128u(arow + ((7 ^ (gt & 7)) << 4), 0u, 0u, 0u, 0u);
        }
      }
      for (int i = 0; i < 5 + L::kPrefetch && issued < total_in; ++i) issue_row();
      cp_async_commit();
      for (int i = 0; i < 5; ++i) cp_async_commit();  // (empty groups: uniform accounting from step 0 on)
#pragma unroll 1
      for (int v = 0; v < V; ++v) {
        ef_seg_seek(w, s_tab, v);
        const int r0 = v + 4 * w.i;  // input rows [r0, r0 + 5) of the input sequence
        cp_async_wait<5>();
        const int st = v % AS;
        if (gt == 0) EF_EV(0, v);
        if (gt < 32 && v >= AS) mbar_wait(C1((v - AS) & 3), ((v - AS) >> 2) & 1);  // MMA1(v - AS) has consumed this A stage (one warp polls)
        ef_bar_sync(8, 128);  // every thread's part of those rows has landed; everyone is done reading the previous step's rows
        if (gt == 0) EF_EV(1, v);
        const uint32_t arow = s_a + st * L::kABytes + gt * 128;
#pragma unroll
        for (int c8 = 0; c8 < 2 * L::kK1Steps; ++c8) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = c8 * 8 + i;  // K index (u*5 + v)*C + ci = u*5C + (v*C + ci); slot kK1 = 1 (bias)
            f[i] = (k == L::kK1) ? 1.f : 0.f;
            if (k < L::kK1) {
              const int u = k / (5 * C), t = k % (5 * C);
              f[i] = ef_lds(s_in + ((r0 + u) % NR) * L::kInRowBytes + (gt * C + t) * 4);
            }
          }
          ef_sts128u(arow + (c8 / 8) * (128 * 128) + (((c8 % 8) ^ (gt & 7)) << 4), ef_pack(f[0], f[1]), ef_pack(f[2], f[3]), ef_pack(f[4], f[5]),
                     ef_pack(f[6], f[7]));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(G1(st));
        if (gt == 0) EF_EV(2, v);
        // refill: the slots overwritten hold rows at least kInSlots - kPrefetch - 5 older than this step's newest row, which
        // every thread finished reading before this step's group barrier
        const int next_need = (v + 1 == w.v1) ? r0 + 5 + 5 : r0 + 5 + 1;  // the next step starts a new segment: five new rows at once
        while (issued < next_need + L::kPrefetch && issued < total_in) issue_row();
        cp_async_commit();
      }
      cp_async_wait<0>();
    } else if (warp < W0) {
      if constexpr (L::kSS) {
        // ---------------------------------------------------------------- epilogues of the Y-channel form: dedicated warps per stage
        constexpr int RC = L::kRC, ROWS = L::kRows;
        constexpr int kVec = 8 * RC;                    // 16-byte groups per output row of this warp's 32 pixels
        constexpr int kIter = (kVec + 31) / 32;
        const int quad = warp & 3, grp = warp >> 2, gl = quad * 32 + lane;
        const uint32_t lane_addr = uint32_t(quad * 32) << 16;
        const bool tr = (quad == 0 && lane == 0);  // the thread that records the timeline
        EfSeg w;
        ef_seg_load(w, s_tab, 0);
        if (grp < 2) {
          // ============================================================== EPI1 (two sets, rows v = grp, grp + 2, ...):
          // a1 = tanh(D1) -> bf16 -> shared-memory ring (SW128 K-major: lane = tile row, eight 16-byte chunks of eight channels)
          if (lane == 0) mbar_arrive(G1(grp));  // stands in for EPI1(grp - 2): both D1 accumulators start out drained
          int sl = grp % L::kNA1S;
#pragma unroll 1
          for (int v = grp; v < V; v += 2) {
            ef_seg_seek(w, s_tab, v);
            const int x = w.s * kEfStripW - kEfLane0 + gl, y = w.ya - 2 + (v - w.v0);
            const bool valid = (x >= 0) && (x < p.W) && (y >= 0) && (y < p.H);
            mbar_wait(C1(v & 3), (v >> 2) & 1);
            tc_fence_after();
            if (tr) EF_EV(5, v);
            if (v >= L::kNA1S) {  // the slot's previous row was read by MMA2(v - kNA1S)
              const int vv = v - L::kNA1S;
              mbar_wait(C2(vv & 3), (vv >> 2) & 1);
            }
            if (tr) EF_EV(7, v);
            const bool all_valid = __all_sync(0xffffffffu, valid);
            const uint32_t dst = s_a1 + sl * L::kA1Bytes + gl * 128;
            const uint32_t src = tmem + L::kColD1 + (v & 1) * 64 + lane_addr;
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {  // (rolled: instruction-cache footprint)
              uint32_t u[32], pk[16];
              tmem_ld_32x32b_x32(src + hf * 32, u);
              tmem_ld_wait();
              if (hf == 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(G1(v & 1));  // D1 accumulator drained: MMA1(v + 2) may overwrite it
                if (tr) EF_EV(6, v);
              }
#pragma unroll
              for (int c = 0; c < 16; ++c) pk[c] = ef_pack(ef_tanh(__uint_as_float(u[2 * c])), ef_tanh(__uint_as_float(u[2 * c + 1])));
              if (!all_valid) {
#pragma unroll
                for (int c = 0; c < 16; ++c) pk[c] = valid ? pk[c] : 0u;
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) ef_sts128u(dst + (((hf * 4 + q) ^ (gl & 7)) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(G2(v & 3));  // a1[v] is in shared memory
            if (tr) EF_EV(8, v);
            sl = (sl + 2) % L::kNA1S;
          }
        } else if (grp == 2) {
          // ============================================================== EPI2 (every row): the accumulator slot holds the nine
          // taps + bias; a2 = tanh(.) -> bf16 -> shared-memory ring (SW64 K-major: 64-byte rows, four 16-byte chunks)
          if (lane == 0) mbar_arrive(G2(0));  // stands in for EPI2(-1)
          int m3 = 0, sl = 0;
#pragma unroll 1
          for (int v2 = 0; v2 < V; ++v2) {
            ef_seg_seek(w, s_tab, v2);
            const int j2 = v2 - w.v0;
            const int x = w.s * kEfStripW - kEfLane0 + gl, y = w.ya - 3 + j2;
            const bool valid = (j2 >= 2) && (x >= 0) && (x < p.W) && (y >= 0) && (y < p.H);
            const bool all_valid = __all_sync(0xffffffffu, valid);
            mbar_wait(C2(v2 & 3), (v2 >> 2) & 1);
            tc_fence_after();
            if (tr) EF_EV(11, v2);
            uint32_t u[32], pk[16];
            tmem_ld_32x32b_x32(tmem + L::kColD2 + m3 * 32 + lane_addr, u);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(G2((v2 + 1) & 3));  // slot read out: MMA2(v2 + 1) may open row v2 + 3 in it
            if (tr) EF_EV(12, v2);
#pragma unroll
            for (int c = 0; c < 16; ++c) pk[c] = ef_pack(ef_tanh(__uint_as_float(u[2 * c])), ef_tanh(__uint_as_float(u[2 * c + 1])));
            if (!all_valid) {
#pragma unroll
              for (int c = 0; c < 16; ++c) pk[c] = valid ? pk[c] : 0u;
            }
            if (v2 >= L::kNA2S) {  // the slot's previous row was read by MMA3(v2 - kNA2S)
              const int vv = v2 - L::kNA2S;
              mbar_wait(C3(vv & 3), (vv >> 2) & 1);
            }
            if (tr) EF_EV(13, v2);
            const uint32_t dst = s_a2 + sl * L::kA2Bytes + gl * 64;
#pragma unroll
            for (int q = 0; q < 4; ++q) ef_sts128u(dst + ((q ^ ((gl >> 1) & 3)) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(G3(v2 & 3));  // a2[v2] is in shared memory
            if (tr) EF_EV(14, v2);
            m3 = (m3 == 2) ? 0 : m3 + 1;
            sl = (sl == L::kNA2S - 1) ? 0 : sl + 1;
          }
        } else {
          // ============================================================== EPI3 (every row): out = accumulator slot (taps + bias)
          // -> depth_to_space through a per-warp transpose tile -> global, 16-byte stores
          if (lane == 0) mbar_arrive(G3(0));  // stands in for EPI3(-1)
          float* const tb = reinterpret_cast<float*>(smem + L::kOffTBuf + quad * L::kTBufBytes);
          const int64_t orow = int64_t(p.W) * RC;  // elements per output row
          const bool vec_ok = (orow & 3) == 0;       // every warp segment starts on a 16-byte boundary
          const bool u8 = p.out_kind == SRK_OUT_U8;
          int m3 = 0;
#pragma unroll 1
          for (int v3 = 0; v3 < V; ++v3) {
            mbar_wait(C3(v3 & 3), (v3 >> 2) & 1);
            tc_fence_after();
            if (tr) EF_EV(17, v3);
            uint32_t u[16];
            tmem_ld_32x32b_x16(tmem + L::kColD3 + m3 * NP3 + lane_addr, u);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(G3((v3 + 1) & 3));  // slot read out: MMA3(v3 + 1) may open row v3 + 3 in it
            if (tr) EF_EV(18, v3);
            m3 = (m3 == 2) ? 0 : m3 + 1;
            ef_seg_seek(w, s_tab, v3);
            const int j3 = v3 - w.v0;
            if (j3 >= 4) {
              // transpose tile in OUTPUT order: channel c = (dy*R + dx)*C + ch goes to output row dy, element pixel*RC + (dx*C + ch)
#pragma unroll
              for (int c = 0; c < COUT; ++c) tb[((c / RC) * 32 + lane) * RC + (c % RC)] = __uint_as_float(u[c]);
              __syncwarp();
              if (tr) EF_EV(19, v3);
              // The warp's 32 pixels own, in each of the ROWS output rows, one contiguous run of 32*RC elements that starts on a
              // 16-byte boundary (x of lane 0 is a multiple of 4): 16-byte stores, consecutive lanes on consecutive groups.
              const int wvalid = min(kEfStripW, p.W - w.s * kEfStripW);
              const int e_lo = max(0, kEfLane0 - quad * 32) * RC;                           // this warp's valid element range
              const int e_hi = max(0, min(32, kEfLane0 + wvalid - quad * 32)) * RC;         // of a 32*RC-element row segment
              const int64_t base = (int64_t(w.n) * p.H + (w.ya + j3 - 4)) * ROWS * orow + int64_t(w.s * kEfStripW - kEfLane0 + quad * 32) * RC;
#pragma unroll
              for (int dy = 0; dy < ROWS; ++dy) {
#pragma unroll
                for (int it = 0; it < kIter; ++it) {
                  const int e = 4 * (lane + 32 * it);
                  if (kVec % 32 == 0 || e < 4 * kVec) {
                    const float4 val = *reinterpret_cast<const float4*>(tb + dy * 32 * RC + e);
                    const int64_t idx = base + dy * orow + e;
                    if (vec_ok && e >= e_lo && e + 4 <= e_hi) {
                      if (u8) *reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(p.out) + idx) = ef_u8(val.x) | (ef_u8(val.y) << 8) | (ef_u8(val.z) << 16) | (ef_u8(val.w) << 24);
                      else *reinterpret_cast<float4*>(static_cast<float*>(p.out) + idx) = val;
                    } else if (e + 4 > e_lo && e < e_hi) {  // strip edge of a frame whose width is not a multiple of 4
                      ef_store_partial(p.out, u8, idx, val, e_lo - e, e_hi - e);
                    }
                  }
                }
              }
              __syncwarp();
            }
            if (tr) EF_EV(20, v3);
          }
        }
      } else {
      // ---------------------------------------------------------------- epilogue: set `set` takes the virtual rows v = set, set+S, ...;
      // per iteration it drains the oldest stage first: EPI3(v-2S), EPI2(v-S), EPI1(v)
      constexpr int RC = L::kRC, ROWS = L::kRows;
      constexpr int kVec = 8 * RC;                    // 16-byte groups per output row of this warp's 32 pixels
      constexpr int kIter = (kVec + 31) / 32;
      const int quad = warp & 3, set = warp >> 2, gl = quad * 32 + lane;
      const uint32_t lane_addr = uint32_t(quad * 32) << 16;
      float* const xch = reinterpret_cast<float*>(smem + L::kOffXch + set * L::kXchBytes);
      float* const tb = reinterpret_cast<float*>(smem + L::kOffTBuf + (set * 4 + quad) * L::kTBufBytes);
      const float4* const bias2 = reinterpret_cast<const float4*>(s_bias);
      const int bar_id = 1 + set;
      const bool tr = (quad == 0 && lane == 0);  // the thread that records the timeline
      int xpar = 0;
      const int64_t orow = int64_t(p.W) * RC;  // elements per output row
      const bool vec_ok = (orow & 3) == 0;       // every warp segment starts on a 16-byte boundary
      const bool u8 = p.out_kind == SRK_OUT_U8;
      // one walker per stage (the stages of an iteration work on rows v, v-S, v-2S), each advanced S rows per iteration
      EfSeg w1, w2, w3;
      ef_seg_load(w1, s_tab, 0);
      w2 = w1;
      w3 = w1;
      // the accumulators start out drained: stand in for the drain arrivals of the rows before row 0
      if (lane == 0) {
        if (set == S - 1) mbar_arrive(G1(0));                                        // EPI1(-1)
        for (int i = 0; i < ND2; ++i)
          if ((i - ND2 + 2 * S) % S == set) mbar_arrive(G2(i));                      // EPI2(i - ND2)
        if (set == S - 1) mbar_arrive(G3(0));                                        // EPI3(-1)
      }
#pragma unroll 1
      for (int v = set; v - 2 * S < V; v += S) {
        // ================================================================ EPI3(v-2S): out = lane-shift(D3) + b3 -> depth_to_space -> global
        if (v >= 2 * S) {
          const int v3 = v - 2 * S;
          mbar_wait(C3(v3 & 3), (v3 >> 2) & 1);
          tc_fence_after();
          if (tr) EF_EV(17, v3);
#pragma unroll
          for (int pass = 0; pass < NP3 / 16; ++pass) {
            constexpr int kLiveMax = 16;
            const int live = (COUT - pass * 16) < kLiveMax ? (COUT - pass * 16) : kLiveMax;  // compile-time after unrolling
            uint32_t u0r[16], u1r[16], u2r[16];
            tmem_ld_32x32b_x16(tmem + L::kColD3 + pass * 16 + lane_addr, u0r);
            tmem_ld_32x32b_x16(tmem + L::kColD3 + 2 * NP3 + pass * 16 + lane_addr, u2r);
            tmem_ld_32x32b_x16(tmem + L::kColD3 + NP3 + pass * 16 + lane_addr, u1r);
            tmem_ld_wait();
            if (pass == NP3 / 16 - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(G3((v3 + 1) & 1));  // D3 drained: MMA3(v3+1) may overwrite it
              if (tr) EF_EV(18, v3);
            }
            float b0[16], b2[16], t[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) b0[c] = __uint_as_float(u0r[c]), b2[c] = __uint_as_float(u2r[c]);
            if (live >= 13) ef_lane_shift<16>(b0, b2, t, xch + xpar, quad, lane, bar_id);
            else if (live >= 10) ef_lane_shift<12>(b0, b2, t, xch + xpar, quad, lane, bar_id);
            else if (live >= 5) ef_lane_shift<9>(b0, b2, t, xch + xpar, quad, lane, bar_id);
            else ef_lane_shift<4>(b0, b2, t, xch + xpar, quad, lane, bar_id);
            xpar ^= 128;
            // transpose tile in OUTPUT order: channel c = (dy*R + dx)*C + ch goes to output row dy, element pixel*RC + (dx*C + ch)
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const int ch = pass * 16 + c;
              if (ch < COUT) {
                const int orow_i = ch / RC, rem = ch % RC;
                tb[(orow_i * 32 + lane) * RC + rem] = (__uint_as_float(u1r[c]) + t[c]) + s_bias[32 + ch];
              }
            }
          }
          __syncwarp();
          if (tr) EF_EV(19, v3);
          ef_seg_seek(w3, s_tab, v3);
          const int j3 = v3 - w3.v0;
          if (j3 >= 4) {
            // The warp's 32 pixels own, in each of the ROWS output rows, one contiguous run of 32*RC elements that starts on a
            // 16-byte boundary (x of lane 0 is a multiple of 4): 16-byte stores, consecutive lanes on consecutive groups.
            const int wvalid = min(kEfStripW, p.W - w3.s * kEfStripW);
            const int e_lo = max(0, kEfLane0 - quad * 32) * RC;                           // this warp's valid element range
            const int e_hi = max(0, min(32, kEfLane0 + wvalid - quad * 32)) * RC;         // of a 32*RC-element row segment
            const int64_t base = (int64_t(w3.n) * p.H + (w3.ya + j3 - 4)) * ROWS * orow + int64_t(w3.s * kEfStripW - kEfLane0 + quad * 32) * RC;
#pragma unroll
            for (int dy = 0; dy < ROWS; ++dy) {
#pragma unroll
              for (int it = 0; it < kIter; ++it) {
                const int e = 4 * (lane + 32 * it);
                if (kVec % 32 == 0 || e < 4 * kVec) {
                  const float4 val = *reinterpret_cast<const float4*>(tb + dy * 32 * RC + e);
                  const int64_t idx = base + dy * orow + e;
                  if (vec_ok && e >= e_lo && e + 4 <= e_hi) {
                    if (u8) *reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(p.out) + idx) = ef_u8(val.x) | (ef_u8(val.y) << 8) | (ef_u8(val.z) << 16) | (ef_u8(val.w) << 24);
                    else *reinterpret_cast<float4*>(static_cast<float*>(p.out) + idx) = val;
                  } else if (e + 4 > e_lo && e < e_hi) {  // strip edge of a frame whose width is not a multiple of 4
                    ef_store_partial(p.out, u8, idx, val, e_lo - e, e_hi - e);
                  }
                }
              }
            }
          }
          __syncwarp();
          if (tr) EF_EV(20, v3);
        }
        // ================================================================ EPI2(v-S): a2 = tanh(lane-shift(D2) + b2) -> bf16 -> TMEM ring A2
        if (v >= S && v - S < V) {
          const int v2 = v - S;
          ef_seg_seek(w2, s_tab, v2);
          const int j2 = v2 - w2.v0;
          const int x = w2.s * kEfStripW - kEfLane0 + gl, y = w2.ya - 3 + j2;
          const bool valid = (j2 >= 2) && (x >= 0) && (x < p.W) && (y >= 0) && (y < p.H);
          mbar_wait(C2(v2 & 3), (v2 >> 2) & 1);
          tc_fence_after();
          if (tr) EF_EV(11, v2);
          const uint32_t d = tmem + L::kColD2 + (v2 % ND2) * L::kN2 + lane_addr;
          if (v2 >= 2) {  // the slot's previous row was last read by MMA3(v2 - 2)
            mbar_wait(C3((v2 - 2) & 3), ((v2 - 2) >> 2) & 1);
            tc_fence_after();
          }
          if (tr) EF_EV(13, v2);
          const bool all_valid = __all_sync(0xffffffffu, valid);
#pragma unroll 1
          for (int pass = 0; pass < 2; ++pass) {  // 16 output channels per pass (rolled: instruction-cache footprint)
            uint32_t u0r[16], u1r[16], u2r[16], pk[8];
            tmem_ld_32x32b_x16(d + pass * 16, u0r);
            tmem_ld_32x32b_x16(d + 64 + pass * 16, u2r);
            tmem_ld_32x32b_x16(d + 32 + pass * 16, u1r);
            tmem_ld_wait();
            float b0[16], b2[16], t[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) b0[c] = __uint_as_float(u0r[c]), b2[c] = __uint_as_float(u2r[c]);
            ef_lane_shift<16>(b0, b2, t, xch + xpar, quad, lane, bar_id);
            xpar ^= 128;
#pragma unroll
            for (int c = 0; c < 16; c += 4) {
              const float4 b = bias2[pass * 4 + c / 4];
              const float2 s01 = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(u1r[c]), __uint_as_float(u1r[c + 1])), make_float2(t[c], t[c + 1])), make_float2(b.x, b.y));
              const float2 s23 = __fadd2_rn(__fadd2_rn(make_float2(__uint_as_float(u1r[c + 2]), __uint_as_float(u1r[c + 3])), make_float2(t[c + 2], t[c + 3])), make_float2(b.z, b.w));
              pk[c / 2] = ef_pack(ef_tanh(s01.x), ef_tanh(s01.y));
              pk[c / 2 + 1] = ef_pack(ef_tanh(s23.x), ef_tanh(s23.y));
            }
            if (!all_valid) {
#pragma unroll
              for (int c = 0; c < 8; ++c) pk[c] = valid ? pk[c] : 0u;
            }
            tmem_st_32x32b_x8(tmem + L::kColA2 + (v2 & 3) * 16 + pass * 8 + lane_addr, pk);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(G2((v2 + ND2) & 1));  // D2 buffer drained: MMA2(v2 + ND2) may overwrite it
          if (tr) EF_EV(12, v2);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(G3(v2 & 1));
          if (tr) EF_EV(14, v2);
        }
        // ================================================================ EPI1(v): a1 = tanh(D1) -> bf16 -> TMEM ring A1
        if (v < V) {
          ef_seg_seek(w1, s_tab, v);
          const int x = w1.s * kEfStripW - kEfLane0 + gl, y = w1.ya - 2 + (v - w1.v0);
          const bool valid = (x >= 0) && (x < p.W) && (y >= 0) && (y < p.H);
          mbar_wait(C1(v & 3), (v >> 2) & 1);
          tc_fence_after();
          if (tr) EF_EV(5, v);
          if (v >= NA1 - 2) {  // the slot's previous row was last read by MMA2(v - (NA1 - 2))
            const int vv = v - (NA1 - 2);
            mbar_wait(C2(vv & 3), (vv >> 2) & 1);
            tc_fence_after();
          }
          if (tr) EF_EV(7, v);
          const bool all_valid = __all_sync(0xffffffffu, valid);
          const uint32_t dst = tmem + L::kColA1 + (v % NA1) * 32 + lane_addr;
#pragma unroll 1
          for (int hf = 0; hf < 2; ++hf) {  // (rolled: instruction-cache footprint)
            uint32_t u[32], pk[16];
            tmem_ld_32x32b_x32(tmem + L::kColD1 + hf * 32 + lane_addr, u);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) pk[c] = ef_pack(ef_tanh(__uint_as_float(u[2 * c])), ef_tanh(__uint_as_float(u[2 * c + 1])));
            if (!all_valid) {
#pragma unroll
              for (int c = 0; c < 16; ++c) pk[c] = valid ? pk[c] : 0u;
            }
            tmem_st_32x32b_x16(dst + hf * 16, pk);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(G1((v + 1) % AS));  // D1 drained: MMA1(v+1) may overwrite it
          if (tr) EF_EV(6, v);
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(G2(v & 1));
          if (tr) EF_EV(8, v);
        }
      }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W0 + 7) tmem_dealloc<512>(tmem);
}

template <int C, int R, bool SHUF>
int launch_espcn_fused(srk_ctx* h, EspcnFusedParams& p, const void* w1p, const void* w2p, const void* w3p, cudaStream_t stream) {
  using L = EspcnCfg<C, R, SHUF>;
  SRK_REQUIRE(L::kTotal <= h->smem_optin, "srk_espcn_forward: needs %d B smem, device allows %d", L::kTotal, h->smem_optin);
  if (first_use(h, reinterpret_cast<const void*>(&espcn_fused_kernel<C, R, SHUF>)))
    SRK_CHECK_CUDA(cudaFuncSetAttribute(espcn_fused_kernel<C, R, SHUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
  if (int rc = make_tensor_map_2d(h, &p.map_w1, w1p, uint64_t(L::kB1 * 64), 64, 64)) return rc;
  if (int rc = make_tensor_map_2d(h, &p.map_w2, w2p, uint64_t(9 * 32), 64, 32)) return rc;
  if (int rc = make_tensor_map_2d(h, &p.map_w3, w3p, uint64_t(9 * L::NP3), 32, L::NP3)) return rc;
  const int grid = int(std::min<long long>(p.units, h->num_sms));
  SRK_CHECK_CUDA(launch_pdl(espcn_fused_kernel<C, R, SHUF>, dim3(grid), dim3(L::kThreads), size_t(L::kTotal), stream, p));
  return 0;
}

// one translation unit per channel count (espcn_fused_c1.cu, espcn_fused_c3.cu): the six (r, shuffle) forms of each compile in parallel
int launch_espcn_fused_c1(srk_ctx* h, EspcnFusedParams& p, int r, bool shuffle, const void* w1p, const void* w2p, const void* w3p, cudaStream_t stream);
int launch_espcn_fused_c3(srk_ctx* h, EspcnFusedParams& p, int r, bool shuffle, const void* w1p, const void* w2p, const void* w3p, cudaStream_t stream);

}  // namespace srk

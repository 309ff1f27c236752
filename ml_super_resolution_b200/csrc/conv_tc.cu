// Tensor-core convolution between flat-padded activations (FPA, include/srk.h) for sm_100a.
//
// A KxK stride-1 SAME convolution over an FPA is K*K shifted GEMMs over ONE flat [rows][CIN] bf16
// matrix: tap (dy,dx) of output row p reads input row p + dy*Wp + dx, and the zero row/column baked
// into the layout supplies the padding.  Each persistent CTA owns a contiguous range of 128-row
// tiles and streams the input through a shared-memory ring of 128-row chunks (one TMA load per
// chunk => every activation byte crosses L2->SMEM once); the K*K taps are tcgen05.mma instructions
// whose A descriptors point at row-shifted windows of that ring (probe-verified: SW128/SW64
// descriptors accept any 128 B / 64 B row shift with base_offset 0).  The weights (K*K blocks of
// [NP][CIN] bf16, K-major) stay resident in shared memory for the whole kernel.  Accumulators live
// in TMEM (4 stages) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
//   warp 0     : TMA producer (weights once, then the chunk ring)
//   warp 1     : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5 : epilogue (tcgen05.ld -> bias/activation/mask -> bf16 -> swizzled smem -> TMA store,
//                or fp32 NHWC scatter with residual add / pixel shuffle for the last layer)
#include "sm100_ptx.cuh"
#include "srk_common.cuh"

namespace srk {

enum { EPI_FPA = 0, EPI_NHWC = 1 };

constexpr int kRing = 7;       // chunk slots (plus one mirror slot)
constexpr int kAccStages = 4;  // TMEM accumulator stages
constexpr int kConvThreads = 192;

struct alignas(64) ConvTcParams {
  CUtensorMap map_in;   // [rows_valid][CIN]  box {CIN,128}
  CUtensorMap map_w;    // [KS*KS*NP][CIN]    box {CIN,NP}
  CUtensorMap map_out;  // [rows_valid][NP]   box {NP,128}   (EPI_FPA)
  const float* bias;    // [NP] or null
  int n_img, H, W, Wp, S;
  int64_t rows_valid;
  int num_tiles;
  int nb;  // chunks of look-behind / look-ahead a tile needs: ceil((KS/2)*(Wp+1)/128)
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
struct ConvTcSmem {
  static constexpr int kRowBytes = CIN * 2;
  static constexpr int kChunkBytes = 128 * kRowBytes;
  static constexpr int kTaps = KS * KS;
  static constexpr int kWTapBytes = NP * kRowBytes;
  static constexpr int kWBytes = ((kTaps * kWTapBytes + 1023) / 1024) * 1024;
  static constexpr int kRingBytes = (kRing + 1) * kChunkBytes;
  static constexpr int kStageBytes = 128 * NP * 2;  // output staging (EPI_FPA)
  static constexpr int kOffW = 0;
  static constexpr int kOffRing = kWBytes;
  static constexpr int kOffStage = kOffRing + kRingBytes;
  static constexpr int kOffBias = kOffStage + kStageBytes;
  static constexpr int kOffBars = kOffBias + 256;
  static constexpr int kNumBars = 2 * kRing + 1 + 2 * kAccStages;
  static constexpr int kOffTmemSlot = kOffBars + kNumBars * 8;
  static constexpr int kTotal = kOffTmemSlot + 16 + 1024;  // + alignment slack
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

template <int CIN, int NP, int KS, int EPI>
__global__ void __launch_bounds__(kConvThreads, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
  using L = ConvTcSmem<CIN, NP, KS>;
  constexpr uint32_t kLayout = (CIN == 64) ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64;
  constexpr uint32_t kSbo = 8 * L::kRowBytes;
  constexpr int kTmemCols = (kAccStages * NP < 32) ? 32 : kAccStages * NP;
  static_assert((kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns must be a power of two");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_w = s_base + L::kOffW;
  const uint32_t s_ring = s_base + L::kOffRing;
  uint8_t* stage_ptr = smem + L::kOffStage;
  float* s_bias = reinterpret_cast<float*>(smem + L::kOffBias);
  const uint32_t s_bars = s_base + L::kOffBars;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmemSlot);
  auto bar_full = [&](int s) { return s_bars + 8u * s; };
  auto bar_empty = [&](int s) { return s_bars + 8u * (kRing + s); };
  const uint32_t bar_wfull = s_bars + 8u * (2 * kRing);
  auto bar_tfull = [&](int a) { return s_bars + 8u * (2 * kRing + 1 + a); };
  auto bar_tempty = [&](int a) { return s_bars + 8u * (2 * kRing + 1 + kAccStages + a); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // contiguous tile range of this CTA
  const int t_begin = int((int64_t(blockIdx.x) * p.num_tiles) / gridDim.x);
  const int t_end = int((int64_t(blockIdx.x + 1) * p.num_tiles) / gridDim.x);
  const int nb = p.nb;
  const int c0 = t_begin - nb;         // first chunk this CTA loads (may be negative: TMA zero-fills)
  const int c_last = t_end - 1 + nb;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; ++i) {
      mbar_init(bar_full(i), 1);
      mbar_init(bar_empty(i), 1);
    }
    mbar_init(bar_wfull, 1);
    for (int i = 0; i < kAccStages; ++i) {
      mbar_init(bar_tfull(i), 1);
      mbar_init(bar_tempty(i), 128);
    }
    fence_mbar_init();
  }
  if (threadIdx.x < NP) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  if (warp == 1) tmem_alloc<kTmemCols>(smem_u32(tmem_slot));
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_in);
    tma_prefetch_desc(&p.map_w);
    if (EPI == EPI_FPA) tma_prefetch_desc(&p.map_out);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (t_begin < t_end) {
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer
      if (lane == 0) {
        mbar_arrive_expect_tx(bar_wfull, L::kTaps * L::kWTapBytes);
        for (int tap = 0; tap < L::kTaps; ++tap) tma_load_2d(s_w + tap * L::kWTapBytes, &p.map_w, 0, tap * NP, bar_wfull);
        for (int c = c0; c <= c_last; ++c) {
          const int i = c - c0, slot = i % kRing, gen = i / kRing;
          mbar_wait(bar_empty(slot), (gen & 1) ^ 1);
          mbar_arrive_expect_tx(bar_full(slot), L::kChunkBytes * (slot == 0 ? 2 : 1));
          tma_load_2d(s_ring + slot * L::kChunkBytes, &p.map_in, 0, c * 128, bar_full(slot));
          // mirror of slot 0 behind the last slot: a 128-row window starting in slot kRing-1 stays contiguous
          if (slot == 0) tma_load_2d(s_ring + kRing * L::kChunkBytes, &p.map_in, 0, c * 128, bar_full(slot));
        }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer (one thread)
      if (lane == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, NP, 0, 0);
        constexpr uint64_t hi = umma_desc_hi(0, kSbo, kLayout);
        mbar_wait(bar_wfull, 0);
        int loaded = c0 - 1;
        for (int t = t_begin; t < t_end; ++t) {
          const int it = t - t_begin, acc = it % kAccStages, accgen = it / kAccStages;
          while (loaded < t + nb) {
            ++loaded;
            const int i = loaded - c0;
            mbar_wait(bar_full(i % kRing), (i / kRing) & 1);
          }
          mbar_wait(bar_tempty(acc), (accgen & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem + acc * NP;
          const int row0 = (t - c0) * 128;
#pragma unroll 1
          for (int tap = 0; tap < L::kTaps; ++tap) {
            const int dy = tap / KS - KS / 2, dx = tap % KS - KS / 2;
            const int rr = (row0 + dy * p.Wp + dx) % (kRing * 128);
            const uint32_t a_addr = s_ring + rr * L::kRowBytes;
            const uint32_t b_addr = s_w + tap * L::kWTapBytes;
#pragma unroll
            for (int k = 0; k < CIN / 16; ++k)
              umma_bf16(d_tmem, umma_desc(hi, a_addr + k * 32), umma_desc(hi, b_addr + k * 32), idesc, (tap | k) != 0);
          }
          umma_commit(bar_tfull(acc));
          umma_commit(bar_empty(it % kRing));  // chunk t-nb (= c0+it) is no longer needed by later tiles
        }
      }
    } else {
      // ------------------------------------------------------------------ epilogue (128 threads)
      const int quad = warp & 3;  // TMEM lane quadrant this warp may access
      const int row = quad * 32 + lane;
      const int etid = threadIdx.x - 64;
      for (int t = t_begin; t < t_end; ++t) {
        const int it = t - t_begin, acc = it % kAccStages, accgen = it / kAccStages;
        mbar_wait(bar_tfull(acc), accgen & 1);
        tc_fence_after();
        float v[NP];
        {
          const uint32_t taddr = tmem + acc * NP + (uint32_t(quad * 32) << 16);
          if constexpr (NP >= 32) {
#pragma unroll
            for (int c = 0; c < NP; c += 32) {
              uint32_t u[32];
              tmem_ld_32x32b_x32(taddr + c, u);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) v[c + j] = __uint_as_float(u[j]);
            }
          } else {
            uint32_t u[16];
            tmem_ld_32x32b_x16(taddr, u);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(u[j]);
          }
        }
        tc_fence_before();
        mbar_arrive(bar_tempty(acc));

        // decode the pixel this row holds
        const int64_t prow = int64_t(t) * 128 + row;
        bool valid = prow < p.rows_valid;
        int n = 0, y = 0, x = 0;
        if (valid) {
          const uint32_t pr = uint32_t(prow);
          const uint32_t q = pr / uint32_t(p.Wp);
          x = int(pr - q * uint32_t(p.Wp));
          n = int(q / uint32_t(p.H + 1));
          const int yy = int(q - uint32_t(n) * uint32_t(p.H + 1));
          y = yy - 1;
          valid = (x < p.W) && (yy > 0);
        }

        if constexpr (EPI == EPI_FPA) {
          uint32_t packed[NP / 2];
          if (valid) {
#pragma unroll
            for (int c = 0; c < NP; ++c) v[c] = act_apply(v[c] + s_bias[c], p.act);
            if (p.mask_src) {
              const uint4* m = reinterpret_cast<const uint4*>(p.mask_src + size_t(prow) * NP);
#pragma unroll
              for (int j = 0; j < NP / 8; ++j) {
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
              const uint4* a = reinterpret_cast<const uint4*>(p.addend_fpa + size_t(prow) * NP);
#pragma unroll
              for (int j = 0; j < NP / 8; ++j) {
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
                for (int c = 0; c < NP; ++c) v[c] = fmaxf(v[c], 0.f);
              }
            }
#pragma unroll
            for (int c = 0; c < NP / 2; ++c) packed[c] = pack_bf16x2(v[2 * c], v[2 * c + 1]);
          } else {
#pragma unroll
            for (int c = 0; c < NP / 2; ++c) packed[c] = 0u;  // pad rows/columns stay exactly zero
          }
          // staging buffer free? (previous tile's TMA store finished reading it)
          if (etid == 0) tma_store_wait_read<0>();
          named_bar_sync(1, 128);
          constexpr int kOutRowBytes = NP * 2;
          const int sw = (NP == 64) ? (row & 7) : ((row >> 1) & 3);
#pragma unroll
          for (int j = 0; j < NP / 8; ++j) {
            uint4 q4 = make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]);
            *reinterpret_cast<uint4*>(stage_ptr + row * kOutRowBytes + ((j ^ sw) << 4)) = q4;
          }
          fence_proxy_async_smem();
          named_bar_sync(2, 128);
          if (etid == 0) {
            tma_store_2d(&p.map_out, 0, t * 128, smem_u32(stage_ptr));
            tma_store_commit();
          }
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
              const int64_t base = (int64_t(fn) * p.FH * r + int64_t(fy) * r) * OW + int64_t(fx) * r;
#pragma unroll
              for (int c = 0; c < NP; ++c) {
                if (c < p.cout) {
                  const int ch = c % C, sub = c / C;
                  const int ddy = sub / r, ddx = sub - ddy * r;
                  const int64_t idx = (base + int64_t(ddy) * OW + ddx) * C + ch;
                  float o = act_apply(v[c] + s_bias[c], p.act);
                  if (p.addend) o += __ldg(p.addend + idx);
                  p.out[idx] = o;
                }
              }
            }
          }
        }
      }
      if (EPI == EPI_FPA && etid == 0) tma_store_wait_all<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<kTmemCols>(tmem);
}

// --------------------------------------------------------------------------------------- host side
template <int CIN, int NP, int KS, int EPI>
static int launch_conv_tc(srk_ctx* h, ConvTcParams& p, const void* x, const void* w_packed, void* y_fpa,
                          cudaStream_t stream) {
  using L = ConvTcSmem<CIN, NP, KS>;
  static bool attr_set = false;
  if (!attr_set) {
    SRK_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<CIN, NP, KS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_set = true;
  }
  SRK_REQUIRE(L::kTotal <= h->smem_optin, "conv_tc: needs %d B smem, device allows %d", L::kTotal, h->smem_optin);
  if (int rc = make_tensor_map_2d(&p.map_in, x, uint64_t(p.rows_valid), CIN, 128)) return rc;
  if (int rc = make_tensor_map_2d(&p.map_w, w_packed, uint64_t(KS * KS * NP), CIN, NP)) return rc;
  if (EPI == EPI_FPA) {
    if (int rc = make_tensor_map_2d(&p.map_out, y_fpa, uint64_t(p.rows_valid), NP, 128)) return rc;
  }
  const int grid = p.num_tiles < h->num_sms ? p.num_tiles : h->num_sms;
  conv_tc_kernel<CIN, NP, KS, EPI><<<grid, kConvThreads, L::kTotal, stream>>>(p);
  SRK_LAUNCH_CHECK();
  return 0;
}

static int fill_geom(ConvTcParams& p, int n_img, int H, int W, int k) {
  SRK_REQUIRE(n_img > 0 && H > 0 && W > 0, "conv_tc: bad geometry n_img=%d H=%d W=%d", n_img, H, W);
  const FpaGeom g = fpa_geom(n_img, H, W);
  SRK_REQUIRE(g.rows_valid < (int64_t(1) << 31), "conv_tc: %lld rows exceed the 2^31 row limit", (long long)g.rows_valid);
  p.n_img = n_img;
  p.H = H;
  p.W = W;
  p.Wp = g.Wp;
  p.S = g.S;
  p.rows_valid = g.rows_valid;
  p.num_tiles = int((g.rows_valid + 127) / 128);
  const int reach = (k / 2) * (g.Wp + 1);
  p.nb = (reach + 127) / 128;
  SRK_REQUIRE(2 * p.nb + 2 <= kRing, "conv_tc: image width %d too large for the %dx%d flat-stream kernel (reach %d rows > %d); "
              "split the frame into column panels", W, k, k, reach, ((kRing - 2) / 2) * 128);
  return 0;
}

}  // namespace srk

using namespace srk;

extern "C" int srk_conv_tc(srk_handle_t h, const void* x_fpa, int cin_p, const void* w_packed, const float* bias, int k,
                           int cout_p, int act, int n_img, int H, int W, void* y_fpa, const void* mask_src,
                           int mask_kind, const void* addend_fpa, int relu_after_add, srk_stream_t stream) {
  SRK_REQUIRE(h && x_fpa && w_packed && y_fpa, "srk_conv_tc: null argument");
  ConvTcParams p{};
  if (int rc = fill_geom(p, n_img, H, W, k)) return rc;
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
  if (int rc = fill_geom(p, n_img, H, W, k)) return rc;
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

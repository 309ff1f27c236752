// Generic tcgen05 GEMM for sm_100a:  D[b][M][N] = act(A[b][M][K] x B[b][N][K]^T + bias[N]),  both operands K-major
// (row-major with K contiguous), bf16 x bf16 or tf32 x tf32 (fp32 storage) -> fp32 accumulation in tensor memory.
//
// This is the general-shape workhorse next to the specialised flat-stream kernels (conv_tc.cu / wgrad_tc.cu / espcn_fused.cuh):
//   * every layer of EnhanceNet's TRAINING losses -- the discriminator's 3x3 stride-1 / stride-2 convolutions
//     of 32..512 channels, its two dense layers, VGG-19's sixteen convolutions, the 16x16-patch Gram matrices -- forward, data
//     gradient and weight gradient, each as im2col / transpose (bandwidth kernels below) + this GEMM
//       enet/enet/model_enet.py:118-261, enet/enet/model_vgg.py:11-99;
//   * the tf32 form of the convolutions north_star names (kind::tf32: fp32 activations and weights, 10-bit mantissa products,
//     fp32 accumulation; exact tanhf in the epilogue) -- vdsr/vdsr/experiment_train.py:17-26 runs fp32 everywhere.
//
// One 128 x BN output tile per CTA (BN = N rounded up to 16, at most 256; larger N: more CTAs along y), K streamed through a
// 4-stage TMA ring of 128-byte swizzled rows (64 bf16 or 32 tf32 per row and stage), four M=128 MMAs per stage.
// Warp roles: 0 TMA producer, 1 MMA issuer (+ TMEM allocation), 2..5 epilogue (TMEM lane quadrant = warp % 4).
// Ragged M, N, K are handled by the tensor maps (out-of-bounds reads are zero) and by guarded stores.
#include <algorithm>

#include "sm100_ptx.cuh"
#include "srk_common.cuh"

namespace srk {

constexpr int kGemmStages = 4;
constexpr int kGemmThreads = 192;

struct alignas(64) GemmParams {
  CUtensorMap map_a;  // [batch][M][K]  box {128 B, 128 rows, 1}
  CUtensorMap map_b;  // [batch][N][K]  box {128 B, BN rows, 1}
  void* d;
  const float* bias;
  long long ldd, stride_d;  // elements
  int M, N, K, BN, k_tiles;
  int act, out_bf16, tf32;
  float leaky;
};

__device__ __forceinline__ float gemm_act(float v, int act, float leaky) {
  switch (act) {
    case SRK_ACT_RELU: return fmaxf(v, 0.f);
    case SRK_ACT_TANH: return tanhf(v);
    case SRK_ACT_LEAKY_RELU: return v > 0.f ? v : leaky * v;
    case SRK_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t s_base = smem_u32(smem);
  const uint32_t stage_bytes = 16384u + uint32_t(p.BN) * 128u;
  const uint32_t s_bars = s_base + kGemmStages * stage_bytes;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kGemmStages * stage_bytes + 128);
  auto full = [&](int s) { return s_bars + 8u * s; };
  auto empty = [&](int s) { return s_bars + 8u * (kGemmStages + s); };
  const uint32_t acc_full = s_bars + 8u * (2 * kGemmStages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * p.BN, b = blockIdx.z;
  uint32_t tcols = 32;
  while (tcols < uint32_t(p.BN)) tcols <<= 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&p.map_a);
    tma_prefetch_desc(&p.map_b);
  }
  if (warp == 1) tmem_alloc_n(smem_u32(tmem_slot), tcols);
  pdl_wait();
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kt = 0; kt < p.k_tiles; ++kt) {
        const int s = kt % kGemmStages;
        if (kt >= kGemmStages) mbar_wait(empty(s), ((kt / kGemmStages) - 1) & 1);
        mbar_arrive_expect_tx(full(s), stage_bytes);
        const int kc = kt * (p.tf32 ? 32 : 64);
        tma_load_3d(s_base + s * stage_bytes, &p.map_a, kc, m0, b, full(s));
        tma_load_3d(s_base + s * stage_bytes + 16384u, &p.map_b, kc, n0, b, full(s));
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = p.tf32 ? umma_idesc_tf32(128, uint32_t(p.BN)) : umma_idesc_bf16(128, uint32_t(p.BN), 0, 0);
    constexpr uint64_t hi = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
    for (int kt = 0; kt < p.k_tiles; ++kt) {
      const int s = kt % kGemmStages;
      mbar_wait(full(s), (kt / kGemmStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a = s_base + s * stage_bytes, bb = a + 16384u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (p.tf32) umma_tf32(tmem, umma_desc(hi, a + k * 32), umma_desc(hi, bb + k * 32), idesc, (kt | k) != 0);
          else umma_bf16(tmem, umma_desc(hi, a + k * 32), umma_desc(hi, bb + k * 32), idesc, (kt | k) != 0);
        }
        umma_commit(empty(s));
        if (kt == p.k_tiles - 1) umma_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- epilogue: one accumulator row per thread
    const int quad = warp & 3;
    const int m = m0 + quad * 32 + lane;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const long long row = (long long)b * p.stride_d + (long long)m * p.ldd;
    for (int c0 = 0; c0 < p.BN; c0 += 16) {
      uint32_t u[16];
      tmem_ld_32x32b_x16(tmem + c0 + (uint32_t(quad * 32) << 16), u);
      tmem_ld_wait();
      if (m < p.M) {
        float v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int n = n0 + c0 + c;
          v[c] = gemm_act(__uint_as_float(u[c]) + ((p.bias && n < p.N) ? __ldg(p.bias + n) : 0.f), p.act, p.leaky);
        }
        const int nb = n0 + c0;
        if (p.out_bf16) {
          __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.d) + row + nb;
          if (nb + 16 <= p.N && ((row + nb) & 7) == 0) {
            uint32_t pk[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * c], v[2 * c + 1]);
              pk[c] = *reinterpret_cast<uint32_t*>(&h2);
            }
            reinterpret_cast<uint4*>(o)[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            reinterpret_cast<uint4*>(o)[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          } else {
            for (int c = 0; c < 16; ++c)
              if (nb + c < p.N) o[c] = __float2bfloat16_rn(v[c]);
          }
        } else {
          float* o = static_cast<float*>(p.d) + row + nb;
          if (nb + 16 <= p.N && ((row + nb) & 3) == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) reinterpret_cast<float4*>(o)[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          } else {
            for (int c = 0; c < 16; ++c)
              if (nb + c < p.N) o[c] = v[c];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc_n(tmem, tcols);
}

// ------------------------------------------------------------------------------------------------ im2col / col2im
// col[m][(u*k + v)*C + c] = x[n, oy*s + u - pt, ox*s + v - pl, c] (0 outside), m = (n*Ho + oy)*Wo + ox; rows are Kp long
// (columns >= k*k*C are zero).  transposed != 0 writes colT[kk][m] (row length Mp) instead: the K-major operand of the weight
// gradient.  TF 'SAME': Ho = ceil(H/s), pad_total = max((Ho-1)*s + k - H, 0), pad_before = pad_total / 2 (the extra pixel goes
// AFTER: asymmetric for stride 2 on even sizes); 'VALID': Ho = (H - k)/s + 1, no padding.
template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const T* __restrict__ x, int n, int H, int W, int C, int k, int s, int Ho, int Wo, int pt, int pl,
                                                     T* __restrict__ col, int Kp, long long Mp, int transposed) {
  const long long M = (long long)n * Ho * Wo;
  const long long total = M * Kp;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long m;
    int kk;
    if (transposed) {  // consecutive threads walk m: coalesced writes of colT
      kk = int(i / M);
      m = i - (long long)kk * M;
    } else {
      m = i / Kp;
      kk = int(i - m * Kp);
    }
    T v = T(0.f);
    if (kk < k * k * C) {
      const int c = kk % C, tap = kk / C, u = tap / k, vv = tap - u * k;
      const int ox = int(m % Wo), oy = int((m / Wo) % Ho), img = int(m / ((long long)Wo * Ho));
      const int y = oy * s + u - pt, xx = ox * s + vv - pl;
      if (y >= 0 && y < H && xx >= 0 && xx < W) v = x[(((long long)img * H + y) * W + xx) * C + c];
    }
    if (transposed) col[(long long)kk * Mp + m] = v;
    else col[i] = v;
  }
}

// dx[n,y,x,c] = sum over the (u,v) with (y + pt - u) % s == 0 and (x + pl - v) % s == 0 of dcol[m(oy,ox)][(u*k+v)*C + c]
// (gather form: no atomics, deterministic).  dcol is fp32 or bf16 [M][Kp]; dx the same type.
template <typename T>
__global__ void __launch_bounds__(256) col2im_kernel(const T* __restrict__ dcol, int n, int H, int W, int C, int k, int s, int Ho, int Wo, int pt,
                                                     int pl, int Kp, T* __restrict__ dx) {
  const long long total = (long long)n * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % C);
    const int xx = int((i / C) % W), y = int((i / ((long long)C * W)) % H), img = int(i / ((long long)C * W * H));
    float acc = 0.f;
    for (int u = 0; u < k; ++u) {
      const int ty = y + pt - u;
      if (ty < 0 || ty % s) continue;
      const int oy = ty / s;
      if (oy >= Ho) continue;
      for (int v = 0; v < k; ++v) {
        const int tx = xx + pl - v;
        if (tx < 0 || tx % s) continue;
        const int ox = tx / s;
        if (ox >= Wo) continue;
        acc += float(dcol[(((long long)img * Ho + oy) * Wo + ox) * Kp + (u * k + v) * C + c]);
      }
    }
    dx[i] = T(acc);
  }
}

// batched 2-D transpose: y[b][c][r] = x[b][r][c]  (x rows ldx apart, y rows ldy apart), through a padded shared tile
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T* __restrict__ x, int R, int Ccols, long long ldx, long long bsx, T* __restrict__ y,
                                                        long long ldy, long long bsy) {
  __shared__ T tile[32][33];
  const int b = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8)
    if (r0 + j < R && c0 + tx < Ccols) tile[j][tx] = x[b * bsx + (long long)(r0 + j) * ldx + c0 + tx];
  __syncthreads();
  for (int j = ty; j < 32; j += 8)
    if (c0 + j < Ccols && r0 + tx < R) y[b * bsy + (long long)(c0 + j) * ldy + r0 + tx] = tile[tx][j];
}

static int grid_for(srk_ctx* h, long long total) { return int(std::min<long long>((total + 255) / 256, (long long)h->num_sms * 16)); }

static void same_geometry(int H, int k, int s, int pad_mode, int& Ho, int& pt) {
  if (pad_mode == SRK_PAD_VALID) {
    Ho = (H - k) / s + 1;
    pt = 0;
  } else {
    Ho = (H + s - 1) / s;
    const int total = std::max((Ho - 1) * s + k - H, 0);
    pt = total / 2;
  }
}

}  // namespace srk

using namespace srk;

extern "C" int srk_gemm_tc(srk_handle_t h, const void* A, const void* B, void* D, int M, int N, int K, int batch, long long lda, long long ldb,
                           long long ldd, long long stride_a, long long stride_b, long long stride_d, const float* bias, int act, float leaky,
                           int in_dtype, int out_dtype, srk_stream_t stream) {
  SRK_REQUIRE(h && A && B && D && M > 0 && N > 0 && K > 0 && batch > 0, "srk_gemm_tc: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  SRK_REQUIRE((in_dtype == SRK_DT_BF16 || in_dtype == SRK_DT_TF32) && (out_dtype == SRK_DT_BF16 || out_dtype == SRK_DT_F32), "srk_gemm_tc: dtypes");
  const uint32_t eb = in_dtype == SRK_DT_TF32 ? 4 : 2;
  GemmParams p{};
  p.BN = std::min(256, (N + 15) / 16 * 16);
  const int kt_elems = 128 / int(eb);
  p.k_tiles = (K + kt_elems - 1) / kt_elems;
  if (int rc = make_tensor_map_3d(h, &p.map_a, A, eb, uint64_t(M), uint64_t(K), uint64_t(batch), uint64_t(lda) * eb, uint64_t(stride_a) * eb, 128)) return rc;
  if (int rc = make_tensor_map_3d(h, &p.map_b, B, eb, uint64_t(N), uint64_t(K), uint64_t(batch), uint64_t(ldb) * eb, uint64_t(stride_b) * eb, uint32_t(p.BN))) return rc;
  p.d = D;
  p.bias = bias;
  p.ldd = ldd;
  p.stride_d = stride_d;
  p.M = M;
  p.N = N;
  p.K = K;
  p.act = act;
  p.leaky = leaky;
  p.out_bf16 = out_dtype == SRK_DT_BF16;
  p.tf32 = in_dtype == SRK_DT_TF32;
  const int smem = kGemmStages * (16384 + p.BN * 128) + 256 + 1024;
  SRK_REQUIRE(smem <= h->smem_optin, "srk_gemm_tc: needs %d B smem", smem);
  if (first_use(h, reinterpret_cast<const void*>(&gemm_tc_kernel)))
    SRK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmStages * (16384 + 256 * 128) + 256 + 1024));
  const dim3 grid((M + 127) / 128, (N + p.BN - 1) / p.BN, batch);
  SRK_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "srk_gemm_tc: grid too large (N tiles %u, batch %u)", grid.y, grid.z);
  SRK_CHECK_CUDA(launch_pdl(gemm_tc_kernel, grid, dim3(kGemmThreads), size_t(smem), as_stream(stream), p));
  return 0;
}

extern "C" int srk_conv_out_size(int in, int k, int stride, int pad_mode) {
  int o, pt;
  same_geometry(in, k, stride, pad_mode, o, pt);
  return o;
}

extern "C" int srk_im2col(srk_handle_t h, const void* x, int dtype, int n, int H, int W, int C, int k, int stride, int pad_mode, void* col, int Kp,
                          long long Mp, int transposed, srk_stream_t stream) {
  SRK_REQUIRE(h && x && col && n > 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0 && Kp >= k * k * C, "srk_im2col: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  int Ho, Wo, pt, pl;
  same_geometry(H, k, stride, pad_mode, Ho, pt);
  same_geometry(W, k, stride, pad_mode, Wo, pl);
  const long long M = (long long)n * Ho * Wo;
  SRK_REQUIRE(!transposed || Mp >= M, "srk_im2col: Mp %lld < M %lld", Mp, M);
  const int grid = grid_for(h, M * Kp);
  if (dtype == SRK_DT_BF16)
    im2col_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), n, H, W, C, k, stride, Ho, Wo, pt, pl,
                                                                     static_cast<__nv_bfloat16*>(col), Kp, Mp, transposed);
  else
    im2col_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), n, H, W, C, k, stride, Ho, Wo, pt, pl, static_cast<float*>(col), Kp, Mp,
                                                             transposed);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_col2im(srk_handle_t h, const void* dcol, int dtype, int n, int H, int W, int C, int k, int stride, int pad_mode, int Kp, void* dx,
                          srk_stream_t stream) {
  SRK_REQUIRE(h && dcol && dx && n > 0 && H > 0 && W > 0 && C > 0 && k > 0 && stride > 0 && Kp >= k * k * C, "srk_col2im: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  int Ho, Wo, pt, pl;
  same_geometry(H, k, stride, pad_mode, Ho, pt);
  same_geometry(W, k, stride, pad_mode, Wo, pl);
  const int grid = grid_for(h, (long long)n * H * W * C);
  if (dtype == SRK_DT_BF16)
    col2im_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(dcol), n, H, W, C, k, stride, Ho, Wo, pt, pl, Kp,
                                                                     static_cast<__nv_bfloat16*>(dx));
  else
    col2im_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(dcol), n, H, W, C, k, stride, Ho, Wo, pt, pl, Kp, static_cast<float*>(dx));
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_transpose(srk_handle_t h, const void* x, int dtype, int batch, int R, int Ccols, long long ldx, long long stride_x, void* y,
                             long long ldy, long long stride_y, srk_stream_t stream) {
  SRK_REQUIRE(h && x && y && batch > 0 && R > 0 && Ccols > 0 && ldx >= Ccols && ldy >= R, "srk_transpose: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  SRK_REQUIRE(batch <= 65535 && (R + 31) / 32 <= 65535, "srk_transpose: too many tiles");
  const dim3 grid((Ccols + 31) / 32, (R + 31) / 32, batch);
  if (dtype == SRK_DT_BF16)
    transpose_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), R, Ccols, ldx, stride_x,
                                                                        static_cast<__nv_bfloat16*>(y), ldy, stride_y);
  else
    transpose_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), R, Ccols, ldx, stride_x, static_cast<float*>(y), ldy, stride_y);
  SRK_LAUNCH_CHECK();
  return 0;
}

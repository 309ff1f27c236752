// CUDA-core kernels around the tensor-core conv: weight packing, the small-Cin first layer
// (fp32 NHWC frame -> FPA), FPA <-> NHWC converters, and the first/last-layer weight gradients.
// These layers carry <1% of the FLOPs and are HBM/L2-bound byte shuffles; they run on fp32 FMAs.
#include <algorithm>

#include "srk_common.cuh"

namespace srk {

// ------------------------------------------------------------------------------------ weight packing
// packed[tap][n][kk] bf16, n < np, kk < cinp.  FWD: n=co, kk=ci, tap=(u,v).  DGRAD: n=ci, kk=co, tap=(k-1-u,k-1-v).
__global__ void pack_weights_kernel(const float* __restrict__ w, int k, int cin, int cout, int mode, int np, int cinp,
                                    __nv_bfloat16* __restrict__ out) {
  const int total = k * k * np * cinp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int kk = i % cinp;
    const int n = (i / cinp) % np;
    const int tap = i / (cinp * np);
    int u = tap / k, v = tap % k;
    float val = 0.f;
    if (mode == SRK_PACK_FWD) {
      if (n < cout && kk < cin) val = w[((u * k + v) * cin + kk) * cout + n];
    } else {
      u = k - 1 - u;
      v = k - 1 - v;
      if (n < cin && kk < cout) val = w[((u * k + v) * cin + n) * cout + kk];
    }
    out[i] = __float2bfloat16_rn(val);
  }
}

// First-layer forms (srk_conv_first_tc): out[blk][n < 64][kk < 64] with GEMM-K index kg = blk*64 + kk.
//   SRK_PACK_FIRST          : kg = (u*k + v)*cin + ci, n = co        -> w[u][v][ci][co]            (cout <= 64, zero beyond)
//   SRK_PACK_FIRST_ROT180T  : kg = (u*k + v)*cout + co, n = ci       -> w[k-1-u][k-1-v][ci][co]    (cin <= 64): the last layer's
//                             data gradient as a first-layer style conv over dY[..., cout]
__device__ __forceinline__ float first_form_value(const float* __restrict__ w, int k, int cin, int cout, int mode, int n, int kg) {
  if (mode == SRK_PACK_FIRST) {
    if (kg >= k * k * cin || n >= cout) return 0.f;
    const int ci = kg % cin, tap = kg / cin;
    return w[(tap * cin + ci) * cout + n];
  }
  if (kg >= k * k * cout || n >= cin) return 0.f;
  const int co = kg % cout, tap = kg / cout;
  const int u = k - 1 - tap / k, v = k - 1 - tap % k;
  return w[((u * k + v) * cin + n) * cout + co];
}

__global__ void pack_first_kernel(const float* __restrict__ w, int k, int cin, int cout, int mode, int blocks, __nv_bfloat16* __restrict__ out) {
  const int total = blocks * 4096;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int kk = e % 64, n = (e / 64) % 64, blk = e / 4096;
    out[e] = __float2bfloat16_rn(first_form_value(w, k, cin, cout, mode, n, blk * 64 + kk));
  }
}

// Batched form: one launch prepares every layer of a model after an optimiser step.  Each job packs one
// HWIO kernel out of the flat fp32 parameter arena; mode SRK_PACK_ROT180T_F32 instead writes the fp32
// [k][k][cout][cin] rotated/transposed kernel that srk_conv_first consumes as the last layer's dgrad.
__global__ void pack_weights_batched_kernel(const float* __restrict__ arena, const srk_pack_job* __restrict__ jobs, int n_jobs,
                                            int64_t total, uint8_t* __restrict__ out_base) {
  extern __shared__ srk_pack_job s_jobs[];
  pdl_wait();
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < n_jobs; i += blockDim.x) s_jobs[i] = jobs[i];
  __syncthreads();
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    int j = 0, hi = n_jobs - 1;  // last job whose first element is <= i (binary search: a model has up to ~50 jobs)
    while (j < hi) {
      const int mid = (j + hi + 1) >> 1;
      if (s_jobs[mid].elem_begin <= i) j = mid;
      else hi = mid - 1;
    }
    const srk_pack_job jb = s_jobs[j];
    const int e = int(i - jb.elem_begin);
    const float* w = arena + jb.src_offset;
    const int k = jb.k, cin = jb.cin, cout = jb.cout;
    if (jb.mode == SRK_PACK_FIRST || jb.mode == SRK_PACK_FIRST_ROT180T) {
      const int kk = e % 64, n = (e / 64) % 64, blk = e / 4096;
      reinterpret_cast<__nv_bfloat16*>(out_base + jb.dst_offset)[e] = __float2bfloat16_rn(first_form_value(w, k, cin, cout, jb.mode, n, blk * 64 + kk));
    } else if (jb.mode == SRK_PACK_ROT180T_F32) {
      // out[u'][v'][co][ci] = w[k-1-u'][k-1-v'][ci][co]
      const int ci = e % cin, co = (e / cin) % cout, tap = e / (cin * cout);
      const int u = k - 1 - tap / k, v = k - 1 - tap % k;
      reinterpret_cast<float*>(out_base + jb.dst_offset)[e] = w[((u * k + v) * cin + ci) * cout + co];
    } else {
      const int kk = e % jb.cinp, n = (e / jb.cinp) % jb.np, tap = e / (jb.cinp * jb.np);
      int u = tap / k, v = tap % k;
      float val = 0.f;
      if (jb.mode == SRK_PACK_FWD) {
        if (n < cout && kk < cin) val = w[((u * k + v) * cin + kk) * cout + n];
      } else {
        u = k - 1 - u;
        v = k - 1 - v;
        if (n < cin && kk < cout) val = w[((u * k + v) * cin + n) * cout + kk];
      }
      reinterpret_cast<__nv_bfloat16*>(out_base + jb.dst_offset)[e] = __float2bfloat16_rn(val);
    }
  }
}

// sum over the arena of mask[i] * w[i]^2, scaled: the l2_regularizer term of the loss
__global__ void __launch_bounds__(256) sumsq_masked_kernel(const float* __restrict__ w, const float* __restrict__ mask, size_t n, float scale,
                                                           float* __restrict__ out) {
  float acc = 0.f;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float v = w[i];
    acc += (mask ? mask[i] : 1.f) * v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += s[i];
    atomicAdd(out, t * scale);
  }
}

// ------------------------------------------------------------------------------------ first layer
struct ConvFirstParams {
  const float* x;
  const float* w;     // HWIO [KS][KS][CIN][64]
  const float* bias;  // [64] or null
  __nv_bfloat16* y;   // FPA 64 ch
  const __nv_bfloat16* relu_mask;
  const srk_panel* panels;
  int FH, FW;      // frame dims
  int Hin, Win;    // panel-local input window (H + KS-1 for VALID, H for SAME)
  int H, W, Wp;    // output FPA geometry
  int64_t rows_valid;
  int po;          // pad offset: KS/2 for SAME, 0 for VALID
  int act;
};

template <int KS, int CIN>
__global__ void __launch_bounds__(128) conv_first_kernel(const ConvFirstParams p) {
  extern __shared__ float s_w[];  // [KS*KS*CIN][64]
  constexpr int kW = KS * KS * CIN * 64;
  for (int i = threadIdx.x; i < kW; i += blockDim.x) s_w[i] = p.w[i];
  __syncthreads();
  const int64_t prow = int64_t(blockIdx.x) * 128 + threadIdx.x;
  if (prow >= p.rows_valid) return;
  const uint32_t pr = uint32_t(prow);
  const uint32_t q = pr / uint32_t(p.Wp);
  const int x = int(pr - q * uint32_t(p.Wp));
  const int n = int(q / uint32_t(p.H + 1));
  const int yy = int(q - uint32_t(n) * uint32_t(p.H + 1));
  const int y = yy - 1;
  uint4* dst = reinterpret_cast<uint4*>(p.y + size_t(prow) * 64);
  if (x >= p.W || yy == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = make_uint4(0, 0, 0, 0);
    return;
  }
  int fn = n, y0 = 0, x0 = 0;
  if (p.panels) {
    const srk_panel e = p.panels[n];
    fn = e.frame;
    y0 = e.y0;
    x0 = e.x0;
  }
  float acc[64];
#pragma unroll
  for (int c = 0; c < 64; ++c) acc[c] = p.bias ? __ldg(p.bias + c) : 0.f;
  const float* frame = p.x + int64_t(fn) * p.FH * p.FW * CIN;
#pragma unroll 1
  for (int u = 0; u < KS; ++u) {
    const int sy = y + u - p.po;
    if (sy < 0 || sy >= p.Hin) continue;
#pragma unroll 1
    for (int v = 0; v < KS; ++v) {
      const int sx = x + v - p.po;
      if (sx < 0 || sx >= p.Win) continue;
      const float* src = frame + (int64_t(y0 + sy) * p.FW + (x0 + sx)) * CIN;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float val = __ldg(src + ci);
        const float4* wr = reinterpret_cast<const float4*>(s_w + ((u * KS + v) * CIN + ci) * 64);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 w4 = wr[j];
          acc[4 * j + 0] = fmaf(val, w4.x, acc[4 * j + 0]);
          acc[4 * j + 1] = fmaf(val, w4.y, acc[4 * j + 1]);
          acc[4 * j + 2] = fmaf(val, w4.z, acc[4 * j + 2]);
          acc[4 * j + 3] = fmaf(val, w4.w, acc[4 * j + 3]);
        }
      }
    }
  }
  if (p.act == SRK_ACT_RELU) {
#pragma unroll
    for (int c = 0; c < 64; ++c) acc[c] = fmaxf(acc[c], 0.f);
  } else if (p.act == SRK_ACT_TANH) {
#pragma unroll
    for (int c = 0; c < 64; ++c) acc[c] = tanhf(acc[c]);
  }
  if (p.relu_mask) {
    const uint4* m = reinterpret_cast<const uint4*>(p.relu_mask + size_t(prow) * 64);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 mv = __ldg(m + j);
      const uint32_t w4[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
        if (!(f.x > 0.f)) acc[j * 8 + e * 2] = 0.f;
        if (!(f.y > 0.f)) acc[j * 8 + e * 2 + 1] = 0.f;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[j * 8 + e * 2], acc[j * 8 + e * 2 + 1]);
      r[e] = *reinterpret_cast<uint32_t*>(&h2);
    }
    dst[j] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

template <int KS, int CIN>
static int launch_conv_first(srk_ctx* h, const ConvFirstParams& p, cudaStream_t s) {
  const int smem = KS * KS * CIN * 64 * 4;
  if (smem > 48 * 1024 && first_use(h, reinterpret_cast<const void*>(&conv_first_kernel<KS, CIN>)))
    SRK_CHECK_CUDA(cudaFuncSetAttribute(conv_first_kernel<KS, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = int((p.rows_valid + 127) / 128);
  conv_first_kernel<KS, CIN><<<grid, 128, smem, s>>>(p);
  SRK_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------ converters
__global__ void fpa_to_nhwc_kernel(const __nv_bfloat16* __restrict__ x, int C, int n_img, int H, int W, float* __restrict__ y) {
  const int64_t total = int64_t(n_img) * H * W * C;
  const int Wp = W + 1;
  const int64_t S = int64_t(H + 1) * Wp;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % C);
    const int64_t pix = i / C;
    const int xx = int(pix % W);
    const int yy = int((pix / W) % H);
    const int64_t n = pix / (int64_t(W) * H);
    y[i] = __bfloat162float(x[(n * S + int64_t(yy + 1) * Wp + xx) * C + c]);
  }
}
// Cp % 8 == 0: one thread converts 8 channels of one pixel row and writes them with a single 16-byte store
__global__ void __launch_bounds__(256) nhwc_to_fpa_vec8_kernel(const float* __restrict__ x, int C, int Cp, int n_img, int H, int W,
                                                               int64_t rows_valid, uint4* __restrict__ y) {
  pdl_wait();
  pdl_launch_dependents();
  const int cpr = Cp / 8;
  const int64_t total = rows_valid * cpr;
  const uint32_t Wp = W + 1, H1 = H + 1;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const uint32_t prow = uint32_t(i / cpr);  // rows_valid < 2^31 (checked by the caller)
    const int c0 = int(i - int64_t(prow) * cpr) * 8;
    const uint32_t q = prow / Wp, xx = prow - q * Wp;
    const uint32_t n = q / H1, yy = q - n * H1;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
    if (xx < uint32_t(W) && yy > 0 && c0 < C) {
      const float* src = x + ((int64_t(n) * H + (yy - 1)) * int64_t(W) + xx) * C;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < C) f[j] = __ldg(src + c0 + j);
    }
    uint4 o;
    __nv_bfloat162 h;
    h = __floats2bfloat162_rn(f[0], f[1]); o.x = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(f[2], f[3]); o.y = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(f[4], f[5]); o.z = *reinterpret_cast<uint32_t*>(&h);
    h = __floats2bfloat162_rn(f[6], f[7]); o.w = *reinterpret_cast<uint32_t*>(&h);
    y[i] = o;
  }
}
__global__ void nhwc_to_fpa_kernel(const float* __restrict__ x, int C, int Cp, int n_img, int H, int W, int64_t rows_valid,
                                   __nv_bfloat16* __restrict__ y) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t total = rows_valid * Cp;
  const int Wp = W + 1;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % Cp);
    const int64_t prow = i / Cp;
    const int64_t q = prow / Wp;
    const int xx = int(prow - q * Wp);
    const int64_t n = q / (H + 1);
    const int yy = int(q - n * (H + 1));
    float v = 0.f;
    if (xx < W && yy > 0 && c < C) v = x[((n * H + (yy - 1)) * int64_t(W) + xx) * C + c];
    y[i] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------ first/last layer wgrad
// Both are tiny contractions (k*k*cin*64 <= 15.5k outputs, reduction over every pixel).  1024 threads =
// 64 channels x 16 pixel lanes; each thread keeps its k*k*c partial sums in registers while the block walks
// its pixel slice, then the 16 pixel lanes are folded through shared memory and one fp32 atomic per
// output and block goes to global (caller zeroes dw/db).
//
// First layer: dw[u][v][ci][co] = sum_{n,y,x} x[n, y+u-po, x+v-po, ci] * dy[n,y,x,co];  db[co] = sum dy.
template <int KS, int CIN>
__global__ void __launch_bounds__(1024) conv_first_wgrad_kernel(const float* __restrict__ x, int n_img, int H, int W,
                                                               const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw,
                                                               float* __restrict__ db) {
  constexpr int kAcc = KS * KS * CIN + 1;  // + bias
  constexpr int po = KS / 2;
  const int co = threadIdx.x & 63, pl = threadIdx.x >> 6;
  const int Wp = W + 1;
  const int64_t S = int64_t(H + 1) * Wp;
  const int64_t npix = int64_t(n_img) * H * W;
  const int64_t per = (npix + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
  float acc[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) acc[i] = 0.f;
  const uint32_t HW = uint32_t(H) * uint32_t(W);
  for (uint32_t pix = uint32_t(p0) + pl; pix < uint32_t(p1); pix += 16) {  // npix < 2^31: 32-bit divisions only
    const uint32_t n = pix / HW, rem = pix - n * HW;
    const int yy = int(rem / uint32_t(W));
    const int xx = int(rem - uint32_t(yy) * uint32_t(W));
    const float g = __bfloat162float(dy[(int64_t(n) * S + int64_t(yy + 1) * Wp + xx) * 64 + co]);
    acc[kAcc - 1] += g;
    const float* img = x + int64_t(n) * H * W * CIN;
#pragma unroll
    for (int u = 0; u < KS; ++u) {
      const int sy = yy + u - po;
#pragma unroll
      for (int v = 0; v < KS; ++v) {
        const int sx = xx + v - po;
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
          const float* src = img + (int64_t(sy) * W + sx) * CIN;
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) acc[(u * KS + v) * CIN + ci] = fmaf(__ldg(src + ci), g, acc[(u * KS + v) * CIN + ci]);
        }
      }
    }
  }
  __shared__ float red[16][64];
#pragma unroll 1
  for (int i = 0; i < kAcc; ++i) {
    red[pl][co] = acc[i];
    __syncthreads();
    if (pl == 0) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) t += red[j][co];
      if (i == kAcc - 1) atomicAdd(db + co, t);
      else atomicAdd(dw + i * 64 + co, t);
    }
    __syncthreads();
  }
}

// Last layer: dw[u][v][ci][co] = sum_p x_fpa[p + (u-1)Wp + (v-1)][ci] * dy[p][co], co < COUT <= 4; db[co] = sum dy.
template <int COUT>
__global__ void __launch_bounds__(1024) conv_last_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ dy,
                                                              int n_img, int H, int W, float* __restrict__ dw, float* __restrict__ db) {
  constexpr int kAcc = 9 * COUT;
  const int ci = threadIdx.x & 63, pl = threadIdx.x >> 6;
  const int Wp = W + 1;
  const int64_t S = int64_t(H + 1) * Wp;
  const int64_t rows_valid = int64_t(n_img) * S;
  const int64_t npix = int64_t(n_img) * H * W;
  const int64_t per = (npix + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = blockIdx.x * per, p1 = min(npix, p0 + per);
  float acc[kAcc], bacc[COUT];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < COUT; ++i) bacc[i] = 0.f;
  const uint32_t HW = uint32_t(H) * uint32_t(W);
  for (uint32_t pix = uint32_t(p0) + pl; pix < uint32_t(p1); pix += 16) {  // npix < 2^31: 32-bit divisions only
    const uint32_t n = pix / HW, rem = pix - n * HW;
    const int yy = int(rem / uint32_t(W));
    const int xx = int(rem - uint32_t(yy) * uint32_t(W));
    float g[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      g[co] = __ldg(dy + size_t(pix) * COUT + co);
      bacc[co] += g[co];
    }
    const int64_t row = int64_t(n) * S + int64_t(yy + 1) * Wp + xx;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int64_t r = row + int64_t(tap / 3 - 1) * Wp + (tap % 3 - 1);
      // FPA pads are zero, rows outside the tensor read as zero (TMA would zero-fill them)
      const float xv = (r >= 0 && r < rows_valid) ? __bfloat162float(x[r * 64 + ci]) : 0.f;
#pragma unroll
      for (int co = 0; co < COUT; ++co) acc[tap * COUT + co] = fmaf(xv, g[co], acc[tap * COUT + co]);
    }
  }
  __shared__ float red[16][64];
#pragma unroll 1
  for (int i = 0; i < kAcc; ++i) {
    red[pl][ci] = acc[i];
    __syncthreads();
    if (pl == 0) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) t += red[j][ci];
      const int tap = i / COUT, co = i % COUT;
      atomicAdd(dw + (tap * 64 + ci) * COUT + co, t);
    }
    __syncthreads();
  }
  if (ci == 0) {  // every ci lane holds the same bias sums: one lane per pixel lane reports
#pragma unroll
    for (int co = 0; co < COUT; ++co) atomicAdd(db + co, bacc[co]);
  }
}

}  // namespace srk

using namespace srk;

extern "C" int srk_pack_conv_weights(srk_handle_t h, const float* w_hwio, int k, int cin, int cout, int mode, int np, int cinp,
                                     void* packed_bf16, srk_stream_t stream) {
  SRK_REQUIRE(h && w_hwio && packed_bf16, "srk_pack_conv_weights: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  if (mode == SRK_PACK_FIRST || mode == SRK_PACK_FIRST_ROT180T) {
    // one job through the batched kernel (jobs array passed by value through a tiny device copy is avoided: use the single-job kernel)
    const int kt = (mode == SRK_PACK_FIRST) ? k * k * cin : k * k * cout;
    const int blocks = ((kt + 15) / 16 * 16 + 63) / 64;
    SRK_REQUIRE((mode == SRK_PACK_FIRST ? cout : cin) == 64, "srk_pack_conv_weights: first-layer forms need 64 output channels");
    pack_first_kernel<<<(blocks * 4096 + 255) / 256, 256, 0, as_stream(stream)>>>(w_hwio, k, cin, cout, mode, blocks,
                                                                                 static_cast<__nv_bfloat16*>(packed_bf16));
    SRK_LAUNCH_CHECK();
    return 0;
  }
  if (mode == SRK_PACK_FWD) SRK_REQUIRE(np >= cout && cinp >= cin, "srk_pack_conv_weights: padded dims too small");
  else SRK_REQUIRE(np >= cin && cinp >= cout, "srk_pack_conv_weights: padded dims too small (dgrad)");
  const int total = k * k * np * cinp;
  pack_weights_kernel<<<(total + 255) / 256, 256, 0, as_stream(stream)>>>(w_hwio, k, cin, cout, mode, np, cinp,
                                                                          static_cast<__nv_bfloat16*>(packed_bf16));
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_conv_first(srk_handle_t h, const float* x, int n_frames, int FH, int FW, int cin, const float* w_hwio,
                              const float* bias, int k, int pad_mode, int act, const srk_panel* panels, int n_img, int H,
                              int W, void* y_fpa, const void* relu_mask_src, srk_stream_t stream) {
  SRK_REQUIRE(h && x && w_hwio && y_fpa, "srk_conv_first: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int halo = (pad_mode == SRK_PAD_VALID) ? k - 1 : 0;
  SRK_REQUIRE(panels || (n_frames == n_img && FH == H + halo && FW == W + halo),
              "srk_conv_first: without panels the frame (%dx%d) must match the output geometry (%dx%d, k=%d)", FH, FW, H, W, k);
  const FpaGeom g = fpa_geom(n_img, H, W);
  SRK_REQUIRE(g.rows_valid < (int64_t(1) << 31), "srk_conv_first: too many rows");
  ConvFirstParams p;
  p.x = x;
  p.w = w_hwio;
  p.bias = bias;
  p.y = static_cast<__nv_bfloat16*>(y_fpa);
  p.relu_mask = static_cast<const __nv_bfloat16*>(relu_mask_src);
  p.panels = panels;
  p.FH = FH;
  p.FW = FW;
  p.Hin = H + halo;
  p.Win = W + halo;
  p.H = H;
  p.W = W;
  p.Wp = g.Wp;
  p.rows_valid = g.rows_valid;
  p.po = (pad_mode == SRK_PAD_VALID) ? 0 : k / 2;
  p.act = act;
  cudaStream_t s = as_stream(stream);
#define SRK_CASE(KS, CIN) \
  if (k == KS && cin == CIN) return launch_conv_first<KS, CIN>(h, p, s);
  SRK_CASE(3, 1) SRK_CASE(3, 3) SRK_CASE(5, 1) SRK_CASE(5, 3) SRK_CASE(9, 1) SRK_CASE(9, 3)
#undef SRK_CASE
  set_error("srk_conv_first: unsupported (k=%d, cin=%d)", k, cin);
  return -1;
}

// Column-halo refresh between neighbouring panels of one frame band (see include/srk.h): every column a panel does not own is
// copied from the panel that owns it.  Only non-owned columns are written and only owned columns are read, so the in-place
// update is race free.  One thread moves 16 bytes.
__global__ void __launch_bounds__(256) fpa_halo_exchange_kernel(__nv_bfloat16* __restrict__ x, int C, const srk_panel* __restrict__ panels,
                                                                int n_img, int H, int W, int max_cols) {
  const int cpr = C / 8;  // 16-byte chunks per pixel row
  const int Wp = W + 1;
  const int64_t S = int64_t(H + 1) * Wp;
  const int64_t total = int64_t(n_img) * H * max_cols * cpr;
  for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total; idx += int64_t(gridDim.x) * blockDim.x) {
    const int ch = int(idx % cpr);
    int64_t r = idx / cpr;
    const int j = int(r % max_cols);
    r /= max_cols;
    const int y = int(r % H), i = int(r / H);
    const srk_panel e = panels[i];
    const int n_left = e.own_x0, n_right = W - e.own_x1;
    int xcol, src_img;
    if (j < n_left) {
      xcol = j;
      src_img = i - 1;
    } else if (j - n_left < n_right) {
      xcol = e.own_x1 + (j - n_left);
      src_img = i + 1;
    } else {
      continue;
    }
    if (src_img < 0 || src_img >= n_img) continue;
    const srk_panel s = panels[src_img];
    if (s.frame != e.frame || s.y0 != e.y0) continue;  // frame edge of the band: keep what the layer computed
    const int sx = e.x0 + xcol - s.x0;
    if (sx < s.own_x0 || sx >= s.own_x1) continue;     // not owned by the direct neighbour (cannot happen for plan_tiles output)
    const uint4* src = reinterpret_cast<const uint4*>(x + (int64_t(src_img) * S + int64_t(y + 1) * Wp + sx) * C) + ch;
    uint4* dst = reinterpret_cast<uint4*>(x + (int64_t(i) * S + int64_t(y + 1) * Wp + xcol) * C) + ch;
    *dst = *src;
  }
}

extern "C" int srk_fpa_halo_exchange(srk_handle_t h, void* x_fpa, int C, const srk_panel* panels, int n_img, int H, int W, int max_cols,
                                     srk_stream_t stream) {
  SRK_REQUIRE(h && x_fpa && panels && C % 8 == 0 && max_cols > 0, "srk_fpa_halo_exchange: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int64_t total = int64_t(n_img) * H * max_cols * (C / 8);
  const int grid = int(std::min<int64_t>((total + 255) / 256, int64_t(h->num_sms) * 8));
  fpa_halo_exchange_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<__nv_bfloat16*>(x_fpa), C, panels, n_img, H, W, max_cols);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_fpa_to_nhwc(srk_handle_t h, const void* x_fpa, int C, int n_img, int H, int W, float* y, srk_stream_t stream) {
  SRK_REQUIRE(h && x_fpa && y, "srk_fpa_to_nhwc: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int64_t total = int64_t(n_img) * H * W * C;
  const int grid = int(std::min<int64_t>((total + 255) / 256, int64_t(h->num_sms) * 16));
  fpa_to_nhwc_kernel<<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x_fpa), C, n_img, H, W, y);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_nhwc_to_fpa_pad(srk_handle_t h, const float* x, int C, int Cp, int n_img, int H, int W, void* y_fpa, srk_stream_t stream) {
  SRK_REQUIRE(h && x && y_fpa && Cp >= C && C > 0, "srk_nhwc_to_fpa_pad: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const FpaGeom g = fpa_geom(n_img, H, W);
  if (Cp % 8 == 0 && g.rows_valid < (int64_t(1) << 31)) {
    const int64_t total8 = g.rows_valid * (Cp / 8);
    const int grid8 = int(std::min<int64_t>((total8 + 255) / 256, int64_t(h->num_sms) * 16));
    SRK_CHECK_CUDA(launch_pdl(nhwc_to_fpa_vec8_kernel, dim3(grid8), dim3(256), 0, as_stream(stream), x, C, Cp, n_img, H, W, g.rows_valid,
                              static_cast<uint4*>(y_fpa)));
    return 0;
  }
  const int64_t total = g.rows_valid * Cp;
  const int grid = int(std::min<int64_t>((total + 255) / 256, int64_t(h->num_sms) * 16));
  SRK_CHECK_CUDA(launch_pdl(nhwc_to_fpa_kernel, dim3(grid), dim3(256), 0, as_stream(stream), x, C, Cp, n_img, H, W, g.rows_valid,
                            static_cast<__nv_bfloat16*>(y_fpa)));
  return 0;
}

extern "C" int srk_nhwc_to_fpa(srk_handle_t h, const float* x, int C, int n_img, int H, int W, void* y_fpa, srk_stream_t stream) {
  if (int rc_dev = check_device(h)) return rc_dev;
  return srk_nhwc_to_fpa_pad(h, x, C, C, n_img, H, W, y_fpa, stream);
}

extern "C" int srk_conv_first_wgrad(srk_handle_t h, const float* x, int n_img, int H, int W, int cin, int k, const void* dy_fpa,
                                    float* dw_hwio, float* dbias, srk_stream_t stream) {
  SRK_REQUIRE(h && x && dy_fpa && dw_hwio && dbias, "srk_conv_first_wgrad: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int64_t npix = int64_t(n_img) * H * W;
  SRK_REQUIRE(npix < (int64_t(1) << 31), "srk_conv_first_wgrad: too many pixels");
  const int grid = int(std::min<int64_t>((npix + 63) / 64, int64_t(h->num_sms)));
  const __nv_bfloat16* dy = static_cast<const __nv_bfloat16*>(dy_fpa);
  cudaStream_t s = as_stream(stream);
#define SRK_CASE(KS, CIN)                                                                           \
  if (k == KS && cin == CIN) {                                                                      \
    conv_first_wgrad_kernel<KS, CIN><<<grid, 1024, 0, s>>>(x, n_img, H, W, dy, dw_hwio, dbias);      \
    SRK_LAUNCH_CHECK();                                                                             \
    return 0;                                                                                       \
  }
  SRK_CASE(3, 1) SRK_CASE(3, 3) SRK_CASE(5, 1) SRK_CASE(5, 3)
#undef SRK_CASE
  set_error("srk_conv_first_wgrad: unsupported (k=%d, cin=%d)", k, cin);
  return -1;
}

extern "C" int srk_conv_last_wgrad(srk_handle_t h, const void* x_fpa, const float* dy, int n_img, int H, int W, int cout,
                                   float* dw_hwio, float* dbias, srk_stream_t stream) {
  SRK_REQUIRE(h && x_fpa && dy && dw_hwio && dbias, "srk_conv_last_wgrad: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int64_t npix = int64_t(n_img) * H * W;
  SRK_REQUIRE(npix < (int64_t(1) << 31), "srk_conv_last_wgrad: too many pixels");
  const int grid = int(std::min<int64_t>((npix + 63) / 64, int64_t(h->num_sms)));
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(x_fpa);
  cudaStream_t s = as_stream(stream);
#define SRK_CASE(COUT)                                                                       \
  if (cout == COUT) {                                                                        \
    conv_last_wgrad_kernel<COUT><<<grid, 1024, 0, s>>>(x, dy, n_img, H, W, dw_hwio, dbias);   \
    SRK_LAUNCH_CHECK();                                                                      \
    return 0;                                                                                \
  }
  SRK_CASE(1) SRK_CASE(2) SRK_CASE(3) SRK_CASE(4)
#undef SRK_CASE
  set_error("srk_conv_last_wgrad: unsupported cout=%d", cout);
  return -1;
}

extern "C" int srk_pack_conv_weights_batched(srk_handle_t h, const float* arena, const srk_pack_job* jobs_device, int n_jobs,
                                             int64_t total_elems, void* out_base, srk_stream_t stream) {
  SRK_REQUIRE(h && arena && jobs_device && out_base && n_jobs > 0 && n_jobs <= 256, "srk_pack_conv_weights_batched: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int grid = int(std::min<int64_t>((total_elems + 255) / 256, int64_t(h->num_sms) * 8));
  SRK_CHECK_CUDA(launch_pdl(pack_weights_batched_kernel, dim3(grid), dim3(256), n_jobs * sizeof(srk_pack_job), as_stream(stream), arena,
                            jobs_device, n_jobs, total_elems, static_cast<uint8_t*>(out_base)));
  return 0;
}

extern "C" int srk_sumsq_masked(srk_handle_t h, const float* w, const float* mask, size_t n, float scale, float* out_accum,
                                srk_stream_t stream) {
  SRK_REQUIRE(h && w && out_accum, "srk_sumsq_masked: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  if (n == 0) return 0;
  const int grid = int(std::min<int64_t>((int64_t(n) + 255) / 256, int64_t(h->num_sms) * 4));
  sumsq_masked_kernel<<<grid, 256, 0, as_stream(stream)>>>(w, mask, n, scale, out_accum);
  SRK_LAUNCH_CHECK();
  return 0;
}

// Bandwidth kernels around the generic GEMM for EnhanceNet's training losses:
//   2x2 max-pool (VGG-19, enet/enet/model_vgg.py:27-36) and its gradient, activation gradients (leaky-ReLU 0.2 / sigmoid of the
//   discriminator, enet/enet/model_enet.py:118-161), the VGG input transform (RGB -> BGR, - mean pixel, model_vgg.py:77-81),
//   per-pixel channel-mean normalisation (model_enet.py:34-41), 16x16 patch extraction for the Gram matrices (:218-246),
//   tf.losses.log_loss (:164-181), and the small axpy / convert / bias-gradient helpers the backward pass needs.
// Grid-stride, coalesced along the channel axis; reductions by warp shuffle + one atomic per warp.
#include <algorithm>

#include "srk_common.cuh"

namespace srk {

static int f2_grid(srk_ctx* h, long long total) { return int(std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)h->num_sms * 16))); }

__device__ __forceinline__ float f2_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- max pool 2x2 stride 2, 'SAME' (windows clipped at the bottom / right edge of odd sizes)
template <typename T>
__global__ void __launch_bounds__(256) maxpool_kernel(const T* __restrict__ x, int n, int H, int W, int C, int Ho, int Wo, T* __restrict__ y) {
  const long long total = (long long)n * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % C), ox = int((i / C) % Wo), oy = int((i / ((long long)C * Wo)) % Ho), img = int(i / ((long long)C * Wo * Ho));
    float m = -3.0e38f;
    for (int u = 0; u < 2; ++u)
      for (int v = 0; v < 2; ++v) {
        const int yy = 2 * oy + u, xx = 2 * ox + v;
        if (yy < H && xx < W) m = fmaxf(m, float(x[(((long long)img * H + yy) * W + xx) * C + c]));
      }
    y[i] = T(m);
  }
}
// gradient: the FIRST maximum of each window (row-major) receives dy, like TF's MaxPoolGrad; every input belongs to one window
template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, int n, int H, int W, int C, int Ho, int Wo,
                                                          T* __restrict__ dx) {
  const long long total = (long long)n * Ho * Wo * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % C), ox = int((i / C) % Wo), oy = int((i / ((long long)C * Wo)) % Ho), img = int(i / ((long long)C * Wo * Ho));
    float m = -3.0e38f;
    int best = -1;
    for (int u = 0; u < 2; ++u)
      for (int v = 0; v < 2; ++v) {
        const int yy = 2 * oy + u, xx = 2 * ox + v;
        if (yy < H && xx < W) {
          const float val = float(x[(((long long)img * H + yy) * W + xx) * C + c]);
          if (val > m) {
            m = val;
            best = u * 2 + v;
          }
        }
      }
    for (int u = 0; u < 2; ++u)
      for (int v = 0; v < 2; ++v) {
        const int yy = 2 * oy + u, xx = 2 * ox + v;
        if (yy < H && xx < W) dx[(((long long)img * H + yy) * W + xx) * C + c] = (u * 2 + v == best) ? dy[i] : T(0.f);
      }
  }
}

// ---- dx = dy * act'(y)   (y = the saved OUTPUT of the activation)
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, long long n, int act, float leaky, T* __restrict__ dx) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = float(dy[i]), o = float(y[i]);
    float d = 1.f;
    if (act == SRK_ACT_RELU) d = o > 0.f ? 1.f : 0.f;
    else if (act == SRK_ACT_LEAKY_RELU) d = o > 0.f ? 1.f : leaky;
    else if (act == SRK_ACT_SIGMOID) d = o * (1.f - o);
    else if (act == SRK_ACT_TANH) d = 1.f - o * o;
    dx[i] = T(g * d);
  }
}

// ---- VGG input: y[p][c] = x[p][2-c] * scale + shift - mean_bgr[c]   (fp32 NHWC in, bf16 or fp32 out); gradient: dx[p][2-c] = dy[p][c] * scale
template <typename T>
__global__ void __launch_bounds__(256) vgg_pre_kernel(const float* __restrict__ x, long long pixels, float scale, float shift, T* __restrict__ y) {
  const float mean[3] = {103.939f, 116.779f, 123.68f};
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pixels * 3; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / 3;
    const int c = int(i - p * 3);
    y[i] = T(x[p * 3 + (2 - c)] * scale + shift - mean[c]);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) vgg_pre_bwd_kernel(const T* __restrict__ dy, long long pixels, float scale, float* __restrict__ dx, int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < pixels * 3; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / 3;
    const int c = int(i - p * 3);
    const float g = float(dy[p * 3 + (2 - c)]) * scale;
    dx[i] = accumulate ? dx[i] + g : g;
  }
}

// ---- normalize: y[p][c] = x[p][c] / (mean_c x[p][:] + eps); one warp per pixel.  bwd: dx_c = dy_c / s - (sum_j dy_j x_j) / (C s^2), s = mean + eps
template <typename T>
__global__ void __launch_bounds__(256) normalize_kernel(const T* __restrict__ x, long long pixels, int C, float eps, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  for (long long p = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; p < pixels; p += ((long long)gridDim.x * blockDim.x) >> 5) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += float(x[p * C + c]);
    s = f2_warp_sum(s) / float(C) + eps;
    for (int c = lane; c < C; c += 32) y[p * C + c] = float(x[p * C + c]) / s;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) normalize_bwd_kernel(const T* __restrict__ x, const float* __restrict__ dy, long long pixels, int C, float eps,
                                                            T* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  for (long long p = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; p < pixels; p += ((long long)gridDim.x * blockDim.x) >> 5) {
    float s = 0.f, dot = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float xv = float(x[p * C + c]);
      s += xv;
      dot += dy[p * C + c] * xv;
    }
    s = f2_warp_sum(s) / float(C) + eps;
    dot = f2_warp_sum(dot);
    const float k = dot / (float(C) * s * s);
    for (int c = lane; c < C; c += 32) dx[p * C + c] = T(dy[p * C + c] / s - k);
  }
}

// ---- 16x16 patches of a [n,H,W,C] fp32 tensor (H, W multiples of 16): patch q = (img, gy, gx), pixel r = py*16 + px
//   x_p [q][r][c]  (the reference's reshape of tf.extract_image_patches, model_enet.py:226-243)   and   x_t [q][c][r] (its transpose)
__global__ void __launch_bounds__(256) patches_kernel(const float* __restrict__ x, int n, int H, int W, int C, __nv_bfloat16* __restrict__ xp,
                                                      __nv_bfloat16* __restrict__ xt) {
  const long long total = (long long)n * H * W * C;
  const int gw = W / 16, gh = H / 16;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % C), xx = int((i / C) % W), yy = int((i / ((long long)C * W)) % H), img = int(i / ((long long)C * W * H));
    const long long q = ((long long)img * gh + yy / 16) * gw + xx / 16;
    const int r = (yy % 16) * 16 + (xx % 16);
    const __nv_bfloat16 v = __float2bfloat16_rn(x[i]);
    xp[(q * 256 + r) * C + c] = v;
    if (xt) xt[(q * C + c) * 256 + r] = v;
  }
}
// gradient back from the transposed patch form: dx[img,yy,xx,c] = dxt[q][c][r]
__global__ void __launch_bounds__(256) patches_bwd_kernel(const float* __restrict__ dxt, int n, int H, int W, int C, float* __restrict__ dx) {
  const long long total = (long long)n * H * W * C;
  const int gw = W / 16, gh = H / 16;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % C), xx = int((i / C) % W), yy = int((i / ((long long)C * W)) % H), img = int(i / ((long long)C * W * H));
    const long long q = ((long long)img * gh + yy / 16) * gw + xx / 16;
    dx[i] = dxt[(q * C + c) * 256 + (yy % 16) * 16 + (xx % 16)];
  }
}

// ---- tf.losses.log_loss(labels = label, predictions = p, eps 1e-7, MEAN): loss += -mean(l log(p+eps) + (1-l) log(1-p+eps)); dp = d loss / d p
__global__ void __launch_bounds__(256) log_loss_kernel(const float* __restrict__ p, long long n, float label, float scale, float* __restrict__ loss,
                                                       float* __restrict__ dp) {
  const float eps = 1e-7f;
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = p[i];
    acc += -(label * logf(v + eps) + (1.f - label) * logf(1.f - v + eps));
    if (dp) dp[i] = -(label / (v + eps) - (1.f - label) / (1.f - v + eps)) * scale / float(n);
  }
  acc = f2_warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && acc != 0.f) atomicAdd(loss, acc * scale / float(n));
}

// ---- y = alpha * x + beta * y (fp32), dtype conversion with scale, column sums (bias gradients)
__global__ void __launch_bounds__(256) axpby_kernel(const float* __restrict__ x, long long n, float alpha, float beta, float* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = alpha * x[i] + (beta != 0.f ? beta * y[i] : 0.f);
}
template <typename S, typename D>
__global__ void __launch_bounds__(256) convert_kernel(const S* __restrict__ x, long long n, float scale, D* __restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = D(float(x[i]) * scale);
}
// out[c] (=|+=) sum_m x[m][c]; one block per 32 columns, rows split over the block's warps, fixed-order combine (deterministic)
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, long long M, int C, float* __restrict__ out, int accumulate) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  float acc = 0.f;
  if (c < C)
    for (long long m = w; m < M; m += 8) acc += float(x[m * C + c]);
  part[w][threadIdx.x & 31] = acc;
  __syncthreads();
  if (w == 0 && c < C) {
    float s = 0.f;
    for (int j = 0; j < 8; ++j) s += part[j][threadIdx.x & 31];
    out[c] = accumulate ? out[c] + s : s;
  }
}

// 3xTF32 operand split: y[r] = [hi | lo | hi] (side 0, the A operand) or [hi | hi | lo] (side 1, the B operand), hi = rna_tf32(x),
// lo = rna_tf32(x - hi); each segment Kp long (zero beyond K).  A.B^T over the 3*Kp-long rows = hi.hi + lo.hi + hi.lo: fp32-level
// accuracy from three tf32 tensor-core products.
__device__ __forceinline__ float rna_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}
__global__ void __launch_bounds__(256) tf32_split_kernel(const float* __restrict__ x, long long R, int K, long long ldx, long long bsx, float* __restrict__ y, int Kp,
                                                         int side) {
  const long long total = R * Kp;
  const long long b = blockIdx.y;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / Kp;
    const int k = int(i - r * Kp);
    const float v = k < K ? x[b * bsx + r * ldx + k] : 0.f;
    const float hi = rna_tf32(v), lo = rna_tf32(v - hi);
    float* row = y + (b * R + r) * 3 * Kp;
    row[k] = hi;
    row[Kp + k] = side ? hi : lo;
    row[2 * Kp + k] = side ? lo : hi;
  }
}

}  // namespace srk

using namespace srk;

extern "C" int srk_tf32_split(srk_handle_t h, const float* x, int batch, long long R, int K, long long ldx, long long stride_x, float* y, int Kp, int side,
                              srk_stream_t stream) {
  SRK_REQUIRE(h && x && y && batch > 0 && R > 0 && K > 0 && Kp >= K && Kp % 4 == 0 && batch <= 65535, "srk_tf32_split: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  tf32_split_kernel<<<dim3(f2_grid(h, R * Kp), batch), 256, 0, as_stream(stream)>>>(x, R, K, ldx, stride_x, y, Kp, side);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_maxpool2x2(srk_handle_t h, const void* x, int dtype, int n, int H, int W, int C, void* y, srk_stream_t stream) {
  SRK_REQUIRE(h && x && y && n > 0 && H > 0 && W > 0 && C > 0, "srk_maxpool2x2: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int grid = f2_grid(h, (long long)n * Ho * Wo * C);
  if (dtype == SRK_DT_BF16) maxpool_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), n, H, W, C, Ho, Wo, static_cast<__nv_bfloat16*>(y));
  else maxpool_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), n, H, W, C, Ho, Wo, static_cast<float*>(y));
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_maxpool2x2_bwd(srk_handle_t h, const void* x, const void* dy, int dtype, int n, int H, int W, int C, void* dx, srk_stream_t stream) {
  SRK_REQUIRE(h && x && dy && dx && n > 0 && H > 0 && W > 0 && C > 0, "srk_maxpool2x2_bwd: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int grid = f2_grid(h, (long long)n * Ho * Wo * C);
  if (dtype == SRK_DT_BF16)
    maxpool_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), n, H, W, C, Ho, Wo,
                                                                          static_cast<__nv_bfloat16*>(dx));
  else maxpool_bwd_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), static_cast<const float*>(dy), n, H, W, C, Ho, Wo, static_cast<float*>(dx));
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_act_bwd(srk_handle_t h, const void* dy, const void* y, int dtype, long long n, int act, float leaky, void* dx, srk_stream_t stream) {
  SRK_REQUIRE(h && dy && y && dx && n > 0, "srk_act_bwd: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int grid = f2_grid(h, n);
  if (dtype == SRK_DT_BF16)
    act_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(y), n, act, leaky,
                                                                      static_cast<__nv_bfloat16*>(dx));
  else act_bwd_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(dy), static_cast<const float*>(y), n, act, leaky, static_cast<float*>(dx));
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_vgg_preprocess(srk_handle_t h, const float* x, long long pixels, float scale, float shift, int out_dtype, void* y, srk_stream_t stream) {
  SRK_REQUIRE(h && x && y && pixels > 0, "srk_vgg_preprocess: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int grid = f2_grid(h, pixels * 3);
  if (out_dtype == SRK_DT_BF16) vgg_pre_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(x, pixels, scale, shift, static_cast<__nv_bfloat16*>(y));
  else vgg_pre_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(x, pixels, scale, shift, static_cast<float*>(y));
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_vgg_preprocess_bwd(srk_handle_t h, const void* dy, int dtype, long long pixels, float scale, float* dx, int accumulate, srk_stream_t stream) {
  SRK_REQUIRE(h && dy && dx && pixels > 0, "srk_vgg_preprocess_bwd: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int grid = f2_grid(h, pixels * 3);
  if (dtype == SRK_DT_BF16) vgg_pre_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(dy), pixels, scale, dx, accumulate);
  else vgg_pre_bwd_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(dy), pixels, scale, dx, accumulate);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_normalize_channels(srk_handle_t h, const void* x, int dtype, long long pixels, int C, float* y, srk_stream_t stream) {
  SRK_REQUIRE(h && x && y && pixels > 0 && C > 0, "srk_normalize_channels: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int grid = f2_grid(h, pixels * 32);
  if (dtype == SRK_DT_BF16) normalize_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), pixels, C, 1e-6f, y);
  else normalize_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), pixels, C, 1e-6f, y);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_normalize_channels_bwd(srk_handle_t h, const void* x, const float* dy, int dtype, long long pixels, int C, void* dx, srk_stream_t stream) {
  SRK_REQUIRE(h && x && dy && dx && pixels > 0 && C > 0, "srk_normalize_channels_bwd: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int grid = f2_grid(h, pixels * 32);
  if (dtype == SRK_DT_BF16)
    normalize_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), dy, pixels, C, 1e-6f, static_cast<__nv_bfloat16*>(dx));
  else normalize_bwd_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), dy, pixels, C, 1e-6f, static_cast<float*>(dx));
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_extract_patches16(srk_handle_t h, const float* x, int n, int H, int W, int C, void* xp_bf16, void* xt_bf16, srk_stream_t stream) {
  SRK_REQUIRE(h && x && xp_bf16 && n > 0 && C > 0 && H >= 16 && W >= 16 && H % 16 == 0 && W % 16 == 0, "srk_extract_patches16: bad argument (H, W multiples of 16)");
  if (int rc_dev = check_device(h)) return rc_dev;
  patches_kernel<<<f2_grid(h, (long long)n * H * W * C), 256, 0, as_stream(stream)>>>(x, n, H, W, C, static_cast<__nv_bfloat16*>(xp_bf16), static_cast<__nv_bfloat16*>(xt_bf16));
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_extract_patches16_bwd(srk_handle_t h, const float* dxt, int n, int H, int W, int C, float* dx, srk_stream_t stream) {
  SRK_REQUIRE(h && dxt && dx && n > 0 && C > 0 && H % 16 == 0 && W % 16 == 0, "srk_extract_patches16_bwd: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  patches_bwd_kernel<<<f2_grid(h, (long long)n * H * W * C), 256, 0, as_stream(stream)>>>(dxt, n, H, W, C, dx);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_log_loss(srk_handle_t h, const float* p, long long n, float label, float scale, float* loss_accum, float* dp, srk_stream_t stream) {
  SRK_REQUIRE(h && p && loss_accum && n > 0, "srk_log_loss: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  log_loss_kernel<<<f2_grid(h, n), 256, 0, as_stream(stream)>>>(p, n, label, scale, loss_accum, dp);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_axpby(srk_handle_t h, const float* x, long long n, float alpha, float beta, float* y, srk_stream_t stream) {
  SRK_REQUIRE(h && x && y && n > 0, "srk_axpby: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  axpby_kernel<<<f2_grid(h, n), 256, 0, as_stream(stream)>>>(x, n, alpha, beta, y);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_convert(srk_handle_t h, const void* x, int src_dtype, long long n, float scale, void* y, int dst_dtype, srk_stream_t stream) {
  SRK_REQUIRE(h && x && y && n > 0, "srk_convert: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int grid = f2_grid(h, n);
  const bool sb = src_dtype == SRK_DT_BF16, db = dst_dtype == SRK_DT_BF16;
  if (sb && db) convert_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), n, scale, static_cast<__nv_bfloat16*>(y));
  else if (sb) convert_kernel<__nv_bfloat16, float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), n, scale, static_cast<float*>(y));
  else if (db) convert_kernel<float, __nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), n, scale, static_cast<__nv_bfloat16*>(y));
  else convert_kernel<float, float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), n, scale, static_cast<float*>(y));
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_colsum(srk_handle_t h, const void* x, int dtype, long long M, int C, float* out, int accumulate, srk_stream_t stream) {
  SRK_REQUIRE(h && x && out && M > 0 && C > 0, "srk_colsum: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int grid = (C + 31) / 32;
  if (dtype == SRK_DT_BF16) colsum_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(x), M, C, out, accumulate);
  else colsum_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(x), M, C, out, accumulate);
  SRK_LAUNCH_CHECK();
  return 0;
}

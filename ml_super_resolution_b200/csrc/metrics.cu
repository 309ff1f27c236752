// Evaluation post-pass on the device: tf.image.psnr / tf.image.ssim per image, RGB -> Y, and the
// saturate_cast(x * 127.5 + 127.5, uint8) hand-off to the PNG encoder.  Bandwidth kernels; fp64 accumulators so that the
// per-image reductions do not depend on the order of the atomics beyond the last bit.
//   vdsr/vdsr/experiment_evaluate.py:57-60   psnr / ssim of [-1,1] images, max_val 2.0
//   espcn/espcn/experiment_test.py:32-55     clip to [0,1], optional rgb_to_yuv Y channel in packed space, max_val 1.0
//   vdsr/vdsr/experiment_resolve.py:65-69    saturate_cast(sr * 127.5 + 127.5, uint8)
#include "srk_common.cuh"

namespace srk {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// acc[n] += sum over this block's slice of image n of (a-b)^2
__global__ void __launch_bounds__(256) sumsq_diff_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t per_img,
                                                         double* __restrict__ acc) {
  const int n = blockIdx.y;
  const float* pa = a + int64_t(n) * per_img;
  const float* pb = b + int64_t(n) * per_img;
  float s = 0.f;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < per_img; i += int64_t(gridDim.x) * blockDim.x) {
    const float d = pa[i] - pb[i];
    s = fmaf(d, d, s);
  }
  s = warp_sum_f(s);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += double(ws[i]);
    atomicAdd(acc + n, t);
  }
}
__global__ void psnr_finalize_kernel(const double* __restrict__ acc, int n_img, double per_img, double max_val, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < n_img) out[n] = float(20.0 * log10(max_val) - 10.0 * log10(acc[n] / per_img));
}

// SSIM (tf.image.ssim, TF 1.8 `_ssim_helper`): 11x11 gaussian (sigma 1.5) VALID window means of x, y, x*y, x^2+y^2 per channel;
// one block = 16x16 window positions of one (image, channel); separable: horizontal pass into shared memory, then vertical.
constexpr int kWin = 11, kTile = 16, kIn = kTile + kWin - 1;  // 26
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ x, const float* __restrict__ y, int H, int W, int C,
                                                   float c1, float c2, double* __restrict__ acc) {
  __shared__ float sx[kIn][kIn + 1], sy[kIn][kIn + 1];
  __shared__ float h0[kIn][kTile], h1[kIn][kTile], h2[kIn][kTile], h3[kIn][kTile];  // row-filtered x, y, x*y, x^2+y^2
  __shared__ float g[kWin];
  __shared__ float ws[8];
  const int n = blockIdx.z / C, c = blockIdx.z % C;
  const int oy0 = blockIdx.y * kTile, ox0 = blockIdx.x * kTile;
  const int OH = H - (kWin - 1), OW = W - (kWin - 1);
  if (threadIdx.x < kWin) {
    float s = 0.f;
    for (int i = 0; i < kWin; ++i) s += expf(-float((i - 5) * (i - 5)) / (2.f * 1.5f * 1.5f));
    g[threadIdx.x] = expf(-float((int(threadIdx.x) - 5) * (int(threadIdx.x) - 5)) / (2.f * 1.5f * 1.5f)) / s;
  }
  const float* px = x + int64_t(n) * H * W * C + c;
  const float* py = y + int64_t(n) * H * W * C + c;
  for (int i = threadIdx.x; i < kIn * kIn; i += 256) {
    const int r = i / kIn, q = i % kIn;
    const int yy = min(oy0 + r, H - 1), xx = min(ox0 + q, W - 1);  // clamped reads only feed window positions outside the image
    sx[r][q] = px[(int64_t(yy) * W + xx) * C];
    sy[r][q] = py[(int64_t(yy) * W + xx) * C];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kIn * kTile; i += 256) {
    const int r = i / kTile, q = i % kTile;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float u = sx[r][q + k], v = sy[r][q + k], w = g[k];
      a0 = fmaf(w, u, a0);
      a1 = fmaf(w, v, a1);
      a2 = fmaf(w, u * v, a2);
      a3 = fmaf(w, fmaf(u, u, v * v), a3);
    }
    h0[r][q] = a0;
    h1[r][q] = a1;
    h2[r][q] = a2;
    h3[r][q] = a3;
  }
  __syncthreads();
  const int r = threadIdx.x / kTile, q = threadIdx.x % kTile;
  float val = 0.f;
  if (oy0 + r < OH && ox0 + q < OW) {
    float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
#pragma unroll
    for (int k = 0; k < kWin; ++k) {
      const float w = g[k];
      m0 = fmaf(w, h0[r + k][q], m0);
      m1 = fmaf(w, h1[r + k][q], m1);
      m2 = fmaf(w, h2[r + k][q], m2);
      m3 = fmaf(w, h3[r + k][q], m3);
    }
    const float num0 = m0 * m1 * 2.f, den0 = m0 * m0 + m1 * m1;
    const float lum = (num0 + c1) / (den0 + c1);
    const float cs = (m2 * 2.f - num0 + c2) / (m3 - den0 + c2);
    val = lum * cs;
  }
  val = warp_sum_f(val);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += double(ws[i]);
    atomicAdd(acc + n, t);
  }
}
__global__ void scale_finalize_kernel(const double* __restrict__ acc, int n_img, double denom, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n < n_img) out[n] = float(acc[n] / denom);
}

__global__ void __launch_bounds__(256) rgb_to_y_kernel(const float* __restrict__ x, int64_t n_pix, float lo, float hi, float scale,
                                                       float bias, float* __restrict__ y) {
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n_pix; i += int64_t(gridDim.x) * blockDim.x) {
    float rgb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] = fminf(fmaxf(fmaf(x[i * 3 + c], scale, bias), lo), hi);
    y[i] = fmaf(rgb[2], 0.114f, fmaf(rgb[1], 0.587f, rgb[0] * 0.299f));
  }
}
__global__ void __launch_bounds__(256) saturate_u8_kernel(const float* __restrict__ x, size_t n, float scale, float bias, uint8_t* __restrict__ y) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const float v = __fadd_rn(__fmul_rn(x[i], scale), bias);
    y[i] = uint8_t(fminf(fmaxf(v, 0.f), 255.f));  // clamp, then truncate: tf.saturate_cast
  }
}

// 8x8 mosaic of the 64 feature maps of one image: channel ch goes to grid cell (ch / 8, ch % 8), saturate-cast like the images
__global__ void __launch_bounds__(256) feature_mosaic_kernel(const float* __restrict__ x, int H, int W, uint8_t* __restrict__ y) {
  const int64_t total = int64_t(64) * H * W;
  const int OW = 8 * W;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int ox = int(i % OW);
    const int oy = int(i / OW);
    const int ch = (oy / H) * 8 + ox / W, yy = oy % H, xx = ox % W;
    const float v = __fadd_rn(__fmul_rn(x[(int64_t(yy) * W + xx) * 64 + ch], 127.5f), 127.5f);
    y[i] = uint8_t(fminf(fmaxf(v, 0.f), 255.f));
  }
}

// ------------------------------------------------------------------------------------ PIL-style uint8 resampling (EnhanceNet input pipeline)
// One pass of Pillow's separable fixed-point resampler (libImaging/Resample.c, 8 bits per channel): out = clip8((2^21 +
// sum_k pixel[lo + k] * coeff[o][k]) >> 22), coefficients and bounds precomputed by the host exactly like precompute_coeffs /
// normalize_coeffs_8bpc.  `along_x` selects the horizontal (over W) or the vertical (over H) pass; x: [n, H, W, C].
__global__ void __launch_bounds__(256) resample_u8_kernel(const uint8_t* __restrict__ x, int n, int H, int W, int C, int out_size, int along_x,
                                                          const int32_t* __restrict__ kk, const int32_t* __restrict__ bounds, int ksize,
                                                          uint8_t* __restrict__ y) {
  const int OH = along_x ? H : out_size, OW = along_x ? out_size : W;
  const int64_t total = int64_t(n) * OH * OW * C;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % C);
    int64_t r = i / C;
    const int ox = int(r % OW);
    r /= OW;
    const int oy = int(r % OH), b = int(r / OH);
    const int o = along_x ? ox : oy;
    const int lo = bounds[2 * o], cnt = bounds[2 * o + 1];
    const int32_t* k = kk + int64_t(o) * ksize;
    int acc = 1 << 21;
    if (along_x) {
      const uint8_t* src = x + ((int64_t(b) * H + oy) * W + lo) * C + c;
      for (int j = 0; j < cnt; ++j) acc += int(src[int64_t(j) * C]) * k[j];
    } else {
      const uint8_t* src = x + ((int64_t(b) * H + lo) * W + ox) * C + c;
      for (int j = 0; j < cnt; ++j) acc += int(src[int64_t(j) * W * C]) * k[j];
    }
    acc >>= 22;
    y[i] = uint8_t(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
  }
}
__global__ void __launch_bounds__(256) crop_u8_kernel(const uint8_t* __restrict__ pool, const srk_pool_image* __restrict__ images,
                                                      const srk_crop* __restrict__ crops, int n, int S, int C, uint8_t* __restrict__ out) {
  const int64_t total = int64_t(n) * S * S * C;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % C);
    int64_t r = i / C;
    const int x = int(r % S);
    r /= S;
    const int y = int(r % S), b = int(r / S);
    const srk_crop cr = crops[b];
    const srk_pool_image im = images[cr.image];
    const int sx = cr.flip ? (cr.x + S - 1 - x) : (cr.x + x);
    out[i] = pool[im.offset + (int64_t(cr.y + y) * im.width + sx) * C + c];
  }
}
__global__ void __launch_bounds__(256) u8_to_pm1_kernel(const uint8_t* __restrict__ x, size_t n, float* __restrict__ y) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    y[i] = __fadd_rn(__fdiv_rn(float(x[i]), 127.5f), -1.f);  // x.astype(float32) / 127.5 - 1.0
}

// numpy's `uint8_image / 127.5 - 1.0` (float64 arithmetic) handed to a float32 placeholder: espcn/espcn/experiment_test.py:159 + :164.
// Half of the 256 values differ by one ulp from the float32 arithmetic of u8_to_pm1_kernel, so this form computes in double too.
__global__ void __launch_bounds__(256) u8_to_pm1_f64_kernel(const uint8_t* __restrict__ x, size_t n, float* __restrict__ y) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    y[i] = float(__dadd_rn(__ddiv_rn(double(x[i]), 127.5), -1.0));
}

static inline int grid1(srk_ctx* h, int64_t items, int block, int per_sm) {
  const int64_t g = (items + block - 1) / block;
  const int64_t cap = int64_t(h->num_sms) * per_sm;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace srk

using namespace srk;

extern "C" int srk_psnr(srk_handle_t h, const float* a, const float* b, int n_img, int64_t numel_per_image, float max_val, double* workspace,
                        float* out, srk_stream_t stream) {
  SRK_REQUIRE(h && a && b && workspace && out && n_img > 0 && numel_per_image > 0, "srk_psnr: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  cudaStream_t s = as_stream(stream);
  SRK_CHECK_CUDA(cudaMemsetAsync(workspace, 0, sizeof(double) * n_img, s));
  const int gx = grid1(h, numel_per_image, 256, 8) / (n_img > 1 ? (n_img > 8 ? 8 : n_img) : 1) + 1;
  sumsq_diff_kernel<<<dim3(gx, n_img), 256, 0, s>>>(a, b, numel_per_image, workspace);
  SRK_LAUNCH_CHECK();
  psnr_finalize_kernel<<<(n_img + 127) / 128, 128, 0, s>>>(workspace, n_img, double(numel_per_image), double(max_val), out);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_ssim(srk_handle_t h, const float* a, const float* b, int n_img, int H, int W, int C, float max_val, double* workspace,
                        float* out, srk_stream_t stream) {
  SRK_REQUIRE(h && a && b && workspace && out && n_img > 0 && C > 0, "srk_ssim: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  SRK_REQUIRE(H >= kWin && W >= kWin, "srk_ssim: image %dx%d smaller than the 11x11 window", H, W);
  SRK_REQUIRE(int64_t(n_img) * C <= 65535, "srk_ssim: n_img * C exceeds the grid limit");
  cudaStream_t s = as_stream(stream);
  SRK_CHECK_CUDA(cudaMemsetAsync(workspace, 0, sizeof(double) * n_img, s));
  const int OH = H - (kWin - 1), OW = W - (kWin - 1);
  const float c1 = (0.01f * max_val) * (0.01f * max_val), c2 = (0.03f * max_val) * (0.03f * max_val);
  ssim_kernel<<<dim3((OW + kTile - 1) / kTile, (OH + kTile - 1) / kTile, n_img * C), 256, 0, s>>>(a, b, H, W, C, c1, c2, workspace);
  SRK_LAUNCH_CHECK();
  scale_finalize_kernel<<<(n_img + 127) / 128, 128, 0, s>>>(workspace, n_img, double(OH) * OW * C, out);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_rgb_to_y(srk_handle_t h, const float* x, int64_t n_pixels, float scale, float bias, float clip_lo, float clip_hi, float* y,
                            srk_stream_t stream) {
  SRK_REQUIRE(h && x && y, "srk_rgb_to_y: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  if (n_pixels == 0) return 0;
  rgb_to_y_kernel<<<grid1(h, n_pixels, 256, 16), 256, 0, as_stream(stream)>>>(x, n_pixels, clip_lo, clip_hi, scale, bias, y);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_saturate_cast_u8(srk_handle_t h, const float* x, size_t n, float scale, float bias, uint8_t* y, srk_stream_t stream) {
  SRK_REQUIRE(h && x && y, "srk_saturate_cast_u8: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  if (n == 0) return 0;
  saturate_u8_kernel<<<grid1(h, int64_t(n), 256, 16), 256, 0, as_stream(stream)>>>(x, n, scale, bias, y);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_feature_mosaic_u8(srk_handle_t h, const float* x, int H, int W, uint8_t* y, srk_stream_t stream) {
  SRK_REQUIRE(h && x && y && H > 0 && W > 0, "srk_feature_mosaic_u8: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  feature_mosaic_kernel<<<grid1(h, int64_t(64) * H * W, 256, 16), 256, 0, as_stream(stream)>>>(x, H, W, y);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_resample_u8(srk_handle_t h, const uint8_t* x, int n, int H, int W, int C, int out_h, int out_w, const int32_t* kx,
                               const int32_t* bx, int ksx, const int32_t* ky, const int32_t* by, int ksy, uint8_t* tmp, uint8_t* y,
                               srk_stream_t stream) {
  SRK_REQUIRE(h && x && y && n > 0 && H > 0 && W > 0 && C > 0 && out_h > 0 && out_w > 0, "srk_resample_u8: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const bool hx = out_w != W, vy = out_h != H;
  SRK_REQUIRE((!hx || (kx && bx && ksx > 0)) && (!vy || (ky && by && ksy > 0)), "srk_resample_u8: missing coefficient table");
  SRK_REQUIRE(!(hx && vy) || tmp, "srk_resample_u8: a two-pass resize needs the [n,H,out_w,C] intermediate");
  cudaStream_t s = as_stream(stream);
  const uint8_t* src = x;
  if (hx) {
    uint8_t* dst = vy ? tmp : y;
    resample_u8_kernel<<<grid1(h, int64_t(n) * H * out_w * C, 256, 16), 256, 0, s>>>(src, n, H, W, C, out_w, 1, kx, bx, ksx, dst);
    SRK_LAUNCH_CHECK();
    src = dst;
  }
  if (vy) {
    resample_u8_kernel<<<grid1(h, int64_t(n) * out_h * out_w * C, 256, 16), 256, 0, s>>>(src, n, H, out_w, C, out_h, 0, ky, by, ksy, y);
    SRK_LAUNCH_CHECK();
  } else if (!hx) {
    SRK_CHECK_CUDA(cudaMemcpyAsync(y, x, size_t(n) * H * W * C, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

extern "C" int srk_crop_u8(srk_handle_t h, const uint8_t* pool, const srk_pool_image* images_device, const srk_crop* crops_device, int n, int S,
                           int C, uint8_t* out, srk_stream_t stream) {
  SRK_REQUIRE(h && pool && images_device && crops_device && out && n > 0 && S > 0 && C > 0, "srk_crop_u8: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  crop_u8_kernel<<<grid1(h, int64_t(n) * S * S * C, 256, 16), 256, 0, as_stream(stream)>>>(pool, images_device, crops_device, n, S, C, out);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_u8_to_pm1_f64(srk_handle_t h, const uint8_t* x, size_t n, float* y, srk_stream_t stream) {
  SRK_REQUIRE(h && x && y, "srk_u8_to_pm1_f64: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  if (n == 0) return 0;
  u8_to_pm1_f64_kernel<<<grid1(h, int64_t(n), 256, 16), 256, 0, as_stream(stream)>>>(x, n, y);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_u8_to_pm1(srk_handle_t h, const uint8_t* x, size_t n, float* y, srk_stream_t stream) {
  SRK_REQUIRE(h && x && y, "srk_u8_to_pm1: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  if (n == 0) return 0;
  u8_to_pm1_kernel<<<grid1(h, int64_t(n), 256, 16), 256, 0, as_stream(stream)>>>(x, n, y);
  SRK_LAUNCH_CHECK();
  return 0;
}

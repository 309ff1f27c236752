// HBM-bound kernels around the conv stack: pixel (un)shuffle, TF1-legacy bicubic, the VDSR degrade
// pre-pass, nearest-neighbour x2 on FPAs, MSE / row-L2-norm losses with their gradients, and the
// TF-Adam / Momentum+clip optimiser steps.  Coalesced, vectorised where the layout allows,
// warp-shuffle reductions, one atomic per CTA.
#include <mutex>

#include "srk_common.cuh"

namespace srk {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// Block-wide sum (blockDim multiple of 32, <= 1024); result valid in thread 0.
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float s[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) s[w] = v;
  __syncthreads();
  if (w == 0) {
    v = (lane < int(blockDim.x + 31) / 32) ? s[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;
}

// ------------------------------------------------------------------------------------ pixel shuffle
// y[n, i*r+dy, j*r+dx, c] = x[n, i, j, (dy*r+dx)*C + c]   (thread per OUTPUT element: coalesced writes)
__global__ void pixel_shuffle_kernel(const float* __restrict__ x, int N, int H, int W, int C, int r, float* __restrict__ y,
                                     int inverse) {
  const int64_t total = int64_t(N) * H * W * C * r * r;
  const int OW = W * r, OH = H * r;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % C);
    const int64_t pix = i / C;
    const int ox = int(pix % OW);
    const int oy = int((pix / OW) % OH);
    const int64_t n = pix / (int64_t(OW) * OH);
    const int ii = oy / r, dy = oy - ii * r, jj = ox / r, dx = ox - jj * r;
    const int64_t packed = ((n * H + ii) * int64_t(W) + jj) * (C * r * r) + (dy * r + dx) * C + c;
    if (inverse) y[packed] = x[i];
    else y[i] = x[packed];
  }
}

// ------------------------------------------------------------------------------------ TF1 bicubic
constexpr int kBicubicTableSize = 1 << 10;
__constant__ float c_bicubic[(kBicubicTableSize + 1) * 2];

static int ensure_bicubic_table(srk_ctx* h) {  // a __constant__ symbol exists once per device: uploaded once per handle
  cudaError_t err = cudaSuccess;
  if (first_use(h, reinterpret_cast<const void*>(&c_bicubic))) {
    // TF-1.x InitCoeffsTable: double arithmetic on a float abscissa, rounded to float on store.
    static float tab[(kBicubicTableSize + 1) * 2];
    static const double A = -0.75;
    for (int i = 0; i <= kBicubicTableSize; ++i) {
      float x = float(i * 1.0 / kBicubicTableSize);
      tab[i * 2] = float(((A + 2) * x - (A + 3)) * x * x + 1);
      x += 1.0;
      tab[i * 2 + 1] = float(((A * x - 5 * A) * x + 8 * A) * x - 4 * A);
    }
    err = cudaMemcpyToSymbol(c_bicubic, tab, sizeof tab);
  }
  SRK_CHECK_CUDA(err);
  return 0;
}

struct BicubicTaps {
  int idx[4];
  float w[4];
};
__device__ __forceinline__ BicubicTaps bicubic_taps(int o, float scale, int limit) {
  BicubicTaps t;
  const float f = __fmul_rn(float(o), scale);
  const int i = int(f);  // trunc, f >= 0
  const float delta = __fsub_rn(f, float(i));
  const int off = __float2int_rn(__fmul_rn(delta, float(kBicubicTableSize)));  // lrintf
  t.w[0] = c_bicubic[off * 2 + 1];
  t.w[1] = c_bicubic[off * 2];
  t.w[2] = c_bicubic[(kBicubicTableSize - off) * 2];
  t.w[3] = c_bicubic[(kBicubicTableSize - off) * 2 + 1];
#pragma unroll
  for (int k = 0; k < 4; ++k) t.idx[k] = min(max(i - 1 + k, 0), limit - 1);
  return t;
}
__device__ __forceinline__ float interp4(float v0, float v1, float v2, float v3, const float* w) {
  // ((v0*w0 + v1*w1) + v2*w2) + v3*w3, every product and sum rounded separately (no FMA contraction)
  float r = __fadd_rn(__fmul_rn(v0, w[0]), __fmul_rn(v1, w[1]));
  r = __fadd_rn(r, __fmul_rn(v2, w[2]));
  return __fadd_rn(r, __fmul_rn(v3, w[3]));
}
__global__ void resize_bicubic_kernel(const float* __restrict__ x, int N, int H, int W, int C, int OH, int OW,
                                      float* __restrict__ y) {
  const float sy = float(H) / float(OH), sx = float(W) / float(OW);
  const int64_t total = int64_t(N) * OH * OW * C;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int c = int(i % C);
    const int64_t pix = i / C;
    const int ox = int(pix % OW);
    const int oy = int((pix / OW) % OH);
    const int64_t n = pix / (int64_t(OW) * OH);
    const BicubicTaps ty = bicubic_taps(oy, sy, H), tx = bicubic_taps(ox, sx, W);
    const float* img = x + n * int64_t(H) * W * C + c;
    float rows[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float* rp = img + int64_t(ty.idx[k]) * W * C;
      rows[k] = interp4(__ldg(rp + int64_t(tx.idx[0]) * C), __ldg(rp + int64_t(tx.idx[1]) * C), __ldg(rp + int64_t(tx.idx[2]) * C),
                        __ldg(rp + int64_t(tx.idx[3]) * C), tx.w);
    }
    y[i] = interp4(rows[0], rows[1], rows[2], rows[3], ty.w);
  }
}

// ------------------------------------------------------------------------------------ VDSR degrade
// One CTA per (image, channel) plane held in shared memory: blur rows, blur columns, bilinear down,
// bilinear up.  Index/fraction tables are computed in fp64 exactly as numpy/skimage do.
__global__ void __launch_bounds__(256) degrade_kernel(const float* __restrict__ hd, int N, int H, int W, int C,
                                                      const float* __restrict__ scales, float* __restrict__ sd) {
  extern __shared__ float sm[];
  float* A = sm;              // [H*W]
  float* B = sm + H * W;      // [H*W]
  __shared__ float kw[32];
  __shared__ int s_lo[2][512], s_hi[2][512];
  __shared__ float s_fr[2][512];
  const int n = blockIdx.x / C, c = blockIdx.x % C;
  const float s = scales[n];
  const int sd_h = int(double(H) / double(s)), sd_w = int(double(W) / double(s));
  const double sigma = fmax(0.0, 0.5 * (double(s) - 1.0));
  const int radius = sigma > 0.0 ? int(4.0 * sigma + 0.5) : 0;
  if (threadIdx.x == 0) {
    double tmp[32], sum = 0.0;
    for (int i = -radius; i <= radius; ++i) {
      tmp[i + radius] = sigma > 0.0 ? exp(-0.5 * double(i) * double(i) / (sigma * sigma)) : 1.0;
      sum += tmp[i + radius];
    }
    for (int i = 0; i <= 2 * radius; ++i) kw[i] = float(tmp[i] / sum);
  }
  const float* src = hd + int64_t(n) * H * W * C + c;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) A[i] = src[int64_t(i) * C];
  __syncthreads();
  // blur along axis 0 (rows index) then axis 1, replicate border -- scipy.ndimage order
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    float acc = 0.f;
    for (int t = -radius; t <= radius; ++t) acc += kw[t + radius] * A[min(max(y + t, 0), H - 1) * W + x];
    B[i] = acc;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    float acc = 0.f;
    for (int t = -radius; t <= radius; ++t) acc += kw[t + radius] * B[y * W + min(max(x + t, 0), W - 1)];
    A[i] = acc;
  }
  // down: [H,W] -> [sd_h, sd_w]
  for (int i = threadIdx.x; i < sd_h + sd_w; i += blockDim.x) {
    const int ax = i < sd_h ? 0 : 1;
    const int o = ax ? i - sd_h : i, in = ax ? W : H, out = ax ? sd_w : sd_h;
    double sc = double(in) / double(out);
    double v = (double(o) + 0.5) * sc - 0.5;
    v = fmin(fmax(v, 0.0), double(in) - 1.0);
    const int lo = int(floor(v));
    s_lo[ax][o] = lo;
    s_hi[ax][o] = min(lo + 1, in - 1);
    s_fr[ax][o] = float(v - double(lo));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < sd_h * sd_w; i += blockDim.x) {
    const int y = i / sd_w, x = i - y * sd_w;
    const float fx = s_fr[1][x], fy = s_fr[0][y];
    const float top = A[s_lo[0][y] * W + s_lo[1][x]] * (1.f - fx) + A[s_lo[0][y] * W + s_hi[1][x]] * fx;
    const float bot = A[s_hi[0][y] * W + s_lo[1][x]] * (1.f - fx) + A[s_hi[0][y] * W + s_hi[1][x]] * fx;
    B[i] = top * (1.f - fy) + bot * fy;
  }
  __syncthreads();
  // up: [sd_h, sd_w] -> [H, W]
  for (int i = threadIdx.x; i < H + W; i += blockDim.x) {
    const int ax = i < H ? 0 : 1;
    const int o = ax ? i - H : i, in = ax ? sd_w : sd_h, out = ax ? W : H;
    double sc = double(in) / double(out);
    double v = (double(o) + 0.5) * sc - 0.5;
    v = fmin(fmax(v, 0.0), double(in) - 1.0);
    const int lo = int(floor(v));
    s_lo[ax][o] = lo;
    s_hi[ax][o] = min(lo + 1, in - 1);
    s_fr[ax][o] = float(v - double(lo));
  }
  __syncthreads();
  float* dst = sd + int64_t(n) * H * W * C + c;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const int y = i / W, x = i - y * W;
    const float fx = s_fr[1][x], fy = s_fr[0][y];
    const float top = B[s_lo[0][y] * sd_w + s_lo[1][x]] * (1.f - fx) + B[s_lo[0][y] * sd_w + s_hi[1][x]] * fx;
    const float bot = B[s_hi[0][y] * sd_w + s_lo[1][x]] * (1.f - fx) + B[s_hi[0][y] * sd_w + s_hi[1][x]] * fx;
    dst[int64_t(i) * C] = top * (1.f - fy) + bot * fy;
  }
}

// ------------------------------------------------------------------------------------ NN x2 on FPAs (64 ch)
// forward: y(n, Y, X) = x(n, Y>>1, X>>1);  thread per (output row, 16-byte chunk)
__global__ void fpa_upsample2_kernel(const uint4* __restrict__ x, int n_img, int H, int W, uint4* __restrict__ y, int64_t rows_out) {
  const int OH = 2 * H, OW = 2 * W, OWp = OW + 1, IWp = W + 1;
  const int64_t IS = int64_t(H + 1) * IWp;
  const int64_t total = rows_out * 8;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int j = int(i & 7);
    const int64_t prow = i >> 3;
    const int64_t q = prow / OWp;
    const int X = int(prow - q * OWp);
    const int64_t n = q / (OH + 1);
    const int YY = int(q - n * (OH + 1));
    uint4 v = make_uint4(0, 0, 0, 0);
    if (X < OW && YY > 0) v = __ldg(x + (n * IS + int64_t(((YY - 1) >> 1) + 1) * IWp + (X >> 1)) * 8 + j);
    y[i] = v;
  }
}
// backward: dx(n, y, x) = sum of the 2x2 block of dy
__global__ void fpa_upsample2_bwd_kernel(const uint4* __restrict__ dy, int n_img, int H, int W, uint4* __restrict__ dx,
                                         int64_t rows_in) {
  const int OW = 2 * W, OWp = OW + 1, IWp = W + 1;
  const int64_t OS = int64_t(2 * H + 1) * OWp;
  const int64_t total = rows_in * 8;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
    const int j = int(i & 7);
    const int64_t prow = i >> 3;
    const int64_t q = prow / IWp;
    const int xx = int(prow - q * IWp);
    const int64_t n = q / (H + 1);
    const int yy = int(q - n * (H + 1));
    uint4 out = make_uint4(0, 0, 0, 0);
    if (xx < W && yy > 0) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const uint4 v = __ldg(dy + (n * OS + int64_t(2 * (yy - 1) + a + 1) * OWp + (2 * xx + b)) * 8 + j);
          const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[e]));
            acc[2 * e] += f.x;
            acc[2 * e + 1] += f.y;
          }
        }
      uint32_t r[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
        r[e] = *reinterpret_cast<uint32_t*>(&h2);
      }
      out = make_uint4(r[0], r[1], r[2], r[3]);
    }
    dx[i] = out;
  }
}

// ------------------------------------------------------------------------------------ losses
__global__ void __launch_bounds__(256) mse_fwd_bwd_kernel(const float* __restrict__ sr, const float* __restrict__ hd, size_t numel,
                                                          float inv_total, float* __restrict__ loss, float* __restrict__ dsr) {
  float acc = 0.f;
  const size_t n4 = numel / 4;
  const float4* a4 = reinterpret_cast<const float4*>(sr);
  const float4* b4 = reinterpret_cast<const float4*>(hd);
  float4* d4 = reinterpret_cast<float4*>(dsr);
  const float g = 2.f * inv_total;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) {
    const float4 a = __ldg(a4 + i), b = __ldg(b4 + i);
    const float4 d = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
    acc += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
    if (dsr) d4[i] = make_float4(d.x * g, d.y * g, d.z * g, d.w * g);
  }
  for (size_t i = n4 * 4 + blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < numel; i += size_t(gridDim.x) * blockDim.x) {
    const float d = sr[i] - hd[i];
    acc += d * d;
    if (dsr) dsr[i] = d * g;
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc * inv_total);
}

// one CTA per row of `cols` consecutive elements
__global__ void __launch_bounds__(256) l2norm_rows_kernel(const float* __restrict__ sr, const float* __restrict__ hd, int rows, int cols,
                                                          float* __restrict__ loss, float* __restrict__ dsr, int sr_act) {
  const int r = blockIdx.x;
  const float* a = sr + int64_t(r) * cols;
  const float* b = hd + int64_t(r) * cols;
  float acc = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    const float d = a[i] - b[i];
    acc += d * d;
  }
  acc = block_sum(acc);
  __shared__ float s_nrm;
  if (threadIdx.x == 0) {
    s_nrm = sqrtf(acc);
    atomicAdd(loss, s_nrm / float(rows));
  }
  __syncthreads();
  if (dsr) {
    const float inv = s_nrm > 0.f ? 1.f / (s_nrm * float(rows)) : 0.f;
    for (int i = threadIdx.x; i < cols; i += blockDim.x) {
      float g = (a[i] - b[i]) * inv;
      if (sr_act == SRK_ACT_TANH) g *= 1.f - a[i] * a[i];  // sr = tanh(pre): hand back dloss/dpre
      dsr[int64_t(r) * cols + i] = g;
    }
  }
}

// ------------------------------------------------------------------------------------ optimisers
__global__ void adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            size_t n, float lr_t, const float* __restrict__ lr_t_dev, float b1, float b2, float eps, float wd,
                            const float* __restrict__ mask) {
  if (lr_t_dev) lr_t = __ldg(lr_t_dev);
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    adam_update(w, m, v, i, g[i], lr_t, b1, b2, eps, wd, mask);
}
__global__ void momentum_clip_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ acc, size_t n, float lr,
                                     float mom, float cap, float wd, const float* __restrict__ mask) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    float gi = g[i];
    const float wi = w[i];
    if (mask) gi = fmaf(wd * mask[i], wi, gi);
    gi = fminf(fmaxf(gi, -cap), cap);
    const float a = mom * acc[i] + gi;
    acc[i] = a;
    w[i] = wi - lr * a;
  }
}

static inline int grid_for(srk_ctx* h, int64_t work_items, int block) {
  int64_t g = (work_items + block - 1) / block;
  const int64_t cap = int64_t(h->num_sms) * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return int(g);
}

}  // namespace srk

using namespace srk;

extern "C" int srk_pixel_shuffle(srk_handle_t h, const float* x, int N, int H, int W, int C, int r, float* y, srk_stream_t stream) {
  SRK_REQUIRE(h && x && y && r >= 1, "srk_pixel_shuffle: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int64_t total = int64_t(N) * H * W * C * r * r;
  if (total == 0) return 0;
  pixel_shuffle_kernel<<<grid_for(h, total, 256), 256, 0, as_stream(stream)>>>(x, N, H, W, C, r, y, 0);
  SRK_LAUNCH_CHECK();
  return 0;
}
extern "C" int srk_pixel_unshuffle(srk_handle_t h, const float* x, int N, int H, int W, int C, int r, float* y, srk_stream_t stream) {
  // x: [N, H*r, W*r, C] -> y: [N, H, W, C*r*r]
  SRK_REQUIRE(h && x && y && r >= 1, "srk_pixel_unshuffle: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int64_t total = int64_t(N) * H * W * C * r * r;
  if (total == 0) return 0;
  pixel_shuffle_kernel<<<grid_for(h, total, 256), 256, 0, as_stream(stream)>>>(x, N, H, W, C, r, y, 1);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_resize_bicubic_tf1(srk_handle_t h, const float* x, int N, int H, int W, int C, int OH, int OW, float* y,
                                      srk_stream_t stream) {
  SRK_REQUIRE(h && x && y && OH > 0 && OW > 0, "srk_resize_bicubic_tf1: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  if (int rc = ensure_bicubic_table(h)) return rc;
  const int64_t total = int64_t(N) * OH * OW * C;
  resize_bicubic_kernel<<<grid_for(h, total, 256), 256, 0, as_stream(stream)>>>(x, N, H, W, C, OH, OW, y);
  SRK_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------ device-resident input pipeline
// One thread per output element: random crop + horizontal flip of a uint8 image that lives in a device pool, cast to
// float the way skimage.util.img_as_float32 does (x / 255 in fp32), and optionally the [-1,1] mapping x*2-1 with the two
// roundings numpy performs (no FMA contraction).
__global__ void __launch_bounds__(256) crop_flip_u8_kernel(const uint8_t* __restrict__ pool, const srk_pool_image* __restrict__ images,
                                                           const srk_crop* __restrict__ crops, int n, int S, int C,
                                                           float* __restrict__ out01, float* __restrict__ out_pm1) {
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
    const float v = __fdiv_rn(float(pool[im.offset + (int64_t(cr.y + y) * im.width + sx) * C + c]), 255.f);
    if (out01) out01[i] = v;
    if (out_pm1) out_pm1[i] = __fadd_rn(__fmul_rn(v, 2.f), -1.f);
  }
}
__global__ void __launch_bounds__(256) affine_kernel(const float* __restrict__ x, size_t n, float a, float b, float* __restrict__ y) {
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    y[i] = __fadd_rn(__fmul_rn(x[i], a), b);
}

extern "C" int srk_crop_flip_u8(srk_handle_t h, const uint8_t* pool, const srk_pool_image* images_device, const srk_crop* crops_device,
                                int n, int S, int C, float* out01, float* out_pm1, srk_stream_t stream) {
  SRK_REQUIRE(h && pool && images_device && crops_device && (out01 || out_pm1) && n > 0 && S > 0 && C > 0, "srk_crop_flip_u8: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const int64_t total = int64_t(n) * S * S * C;
  crop_flip_u8_kernel<<<grid_for(h, total, 256), 256, 0, as_stream(stream)>>>(pool, images_device, crops_device, n, S, C, out01, out_pm1);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_affine_f32(srk_handle_t h, const float* x, size_t n, float a, float b, float* y, srk_stream_t stream) {
  SRK_REQUIRE(h && x && y, "srk_affine_f32: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  if (n == 0) return 0;
  affine_kernel<<<grid_for(h, int64_t(n), 256), 256, 0, as_stream(stream)>>>(x, n, a, b, y);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_degrade_gauss_bilinear(srk_handle_t h, const float* hd, int N, int H, int W, int C, const float* scale_per_sample,
                                          float* sd, srk_stream_t stream) {
  SRK_REQUIRE(h && hd && sd && scale_per_sample, "srk_degrade_gauss_bilinear: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  SRK_REQUIRE(H <= 512 && W <= 512, "srk_degrade_gauss_bilinear: patch %dx%d exceeds 512 (per-patch kernel)", H, W);
  const int smem = 2 * H * W * 4;
  SRK_REQUIRE(smem <= h->smem_optin - 16 * 1024, "srk_degrade_gauss_bilinear: patch %dx%d does not fit shared memory", H, W);
  if (smem > 48 * 1024 && first_use(h, reinterpret_cast<const void*>(&degrade_kernel)))  // raised once, to the device's limit
    SRK_CHECK_CUDA(cudaFuncSetAttribute(degrade_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_optin - 16 * 1024));
  degrade_kernel<<<N * C, 256, smem, as_stream(stream)>>>(hd, N, H, W, C, scale_per_sample, sd);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_fpa_upsample2(srk_handle_t h, const void* x_fpa, int n_img, int H, int W, void* y_fpa, srk_stream_t stream) {
  SRK_REQUIRE(h && x_fpa && y_fpa, "srk_fpa_upsample2: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const FpaGeom go = fpa_geom(n_img, 2 * H, 2 * W);
  fpa_upsample2_kernel<<<grid_for(h, go.rows_valid * 8, 256), 256, 0, as_stream(stream)>>>(
      static_cast<const uint4*>(x_fpa), n_img, H, W, static_cast<uint4*>(y_fpa), go.rows_valid);
  SRK_LAUNCH_CHECK();
  return 0;
}
extern "C" int srk_fpa_upsample2_bwd(srk_handle_t h, const void* dy_fpa, int n_img, int H, int W, void* dx_fpa, srk_stream_t stream) {
  SRK_REQUIRE(h && dy_fpa && dx_fpa, "srk_fpa_upsample2_bwd: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  const FpaGeom gi = fpa_geom(n_img, H, W);
  fpa_upsample2_bwd_kernel<<<grid_for(h, gi.rows_valid * 8, 256), 256, 0, as_stream(stream)>>>(
      static_cast<const uint4*>(dy_fpa), n_img, H, W, static_cast<uint4*>(dx_fpa), gi.rows_valid);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_mse_fwd_bwd(srk_handle_t h, const float* sr, const float* hd, size_t numel, double numel_total, float* loss_accum,
                               float* dsr, srk_stream_t stream) {
  SRK_REQUIRE(h && sr && hd && loss_accum, "srk_mse_fwd_bwd: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  SRK_REQUIRE((reinterpret_cast<uintptr_t>(sr) | reinterpret_cast<uintptr_t>(hd) | reinterpret_cast<uintptr_t>(dsr)) % 16 == 0,
              "srk_mse_fwd_bwd: pointers must be 16-byte aligned");
  if (numel == 0) return 0;
  mse_fwd_bwd_kernel<<<grid_for(h, int64_t(numel / 4 + 1), 256), 256, 0, as_stream(stream)>>>(sr, hd, numel, float(1.0 / numel_total),
                                                                                           loss_accum, dsr);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_l2norm_rows_mean_fwd_bwd(srk_handle_t h, const float* sr, const float* hd, int rows, int cols, float* loss_accum,
                                            float* dsr, int sr_act, srk_stream_t stream) {
  SRK_REQUIRE(h && sr && hd && loss_accum && rows > 0 && cols > 0, "srk_l2norm_rows_mean_fwd_bwd: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  SRK_REQUIRE(sr_act == SRK_ACT_NONE || sr_act == SRK_ACT_TANH, "srk_l2norm_rows_mean_fwd_bwd: sr_act must be NONE or TANH");
  l2norm_rows_kernel<<<rows, 256, 0, as_stream(stream)>>>(sr, hd, rows, cols, loss_accum, dsr, sr_act);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_adam_step(srk_handle_t h, float* w, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                             float eps, int64_t t, float weight_decay, const float* decay_mask, srk_stream_t stream) {
  SRK_REQUIRE(h && w && g && m && v && t >= 1, "srk_adam_step: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  if (n == 0) return 0;
  const double lr_t = double(lr) * sqrt(1.0 - pow(double(beta2), double(t))) / (1.0 - pow(double(beta1), double(t)));
  adam_kernel<<<grid_for(h, int64_t(n), 256), 256, 0, as_stream(stream)>>>(w, g, m, v, n, float(lr_t), nullptr, beta1, beta2, eps,
                                                                         weight_decay, decay_mask);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_momentum_clip_step(srk_handle_t h, float* w, const float* g, float* accum, size_t n, float lr, float momentum,
                                      float gradient_cap, float weight_decay, const float* decay_mask, srk_stream_t stream) {
  SRK_REQUIRE(h && w && g && accum && lr > 0.f, "srk_momentum_clip_step: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  if (n == 0) return 0;
  momentum_clip_kernel<<<grid_for(h, int64_t(n), 256), 256, 0, as_stream(stream)>>>(w, g, accum, n, lr, momentum, gradient_cap / lr,
                                                                                   weight_decay, decay_mask);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_adam_step_dev(srk_handle_t h, float* w, const float* g, float* m, float* v, size_t n, const float* lr_t_device,
                                 float beta1, float beta2, float eps, float weight_decay, const float* decay_mask, srk_stream_t stream) {
  SRK_REQUIRE(h && w && g && m && v && lr_t_device, "srk_adam_step_dev: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  if (n == 0) return 0;
  adam_kernel<<<grid_for(h, int64_t(n), 256), 256, 0, as_stream(stream)>>>(w, g, m, v, n, 0.f, lr_t_device, beta1, beta2, eps,
                                                                         weight_decay, decay_mask);
  SRK_LAUNCH_CHECK();
  return 0;
}

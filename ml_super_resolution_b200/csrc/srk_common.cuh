// Shared host-side plumbing for libsrk: handle, error reporting, tensor-map encoding, FPA geometry.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <array>
#include <cstring>
#include <map>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/srk.h"

// Key of one cached tensor map: everything cuTensorMapEncodeTiled is given.
struct srk_tmap_key {
  const void* ptr;
  uint64_t rows;
  uint32_t cols, box_rows, box_cols, elem_bytes;
  uint64_t row_stride = 0, batch = 0, batch_stride = 0;  // (3-D maps of the generic GEMM; 0 for the packed 2-D maps)
  bool operator==(const srk_tmap_key& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && box_rows == o.box_rows && box_cols == o.box_cols && elem_bytes == o.elem_bytes &&
           row_stride == o.row_stride && batch == o.batch && batch_stride == o.batch_stride;
  }
};
struct srk_tmap_key_hash {
  size_t operator()(const srk_tmap_key& k) const {
    uint64_t h = reinterpret_cast<uint64_t>(k.ptr) * 0x9E3779B97F4A7C15ull;
    h ^= (k.rows + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= ((uint64_t(k.cols) << 40) ^ (uint64_t(k.box_rows) << 20) ^ (uint64_t(k.box_cols) << 4) ^ k.elem_bytes) * 0xC2B2AE3D27D4EB4Full;
    h ^= (k.row_stride * 0x9E3779B97F4A7C15ull) ^ (k.batch << 17) ^ (k.batch_stride * 0xD6E8FEB86659FD93ull);
    return size_t(h);
  }
};

// One handle per device (include/srk.h): everything CUDA keeps PER DEVICE -- kernel attributes, __constant__ uploads, encoded
// tensor maps -- is remembered here, never in process-wide statics, so that a second handle on another GPU of the same process
// sets its own.  Not thread-safe per handle.
struct srk_ctx {
  int device;
  int num_sms;
  int smem_optin;
  std::unordered_set<const void*> once;  // kernels whose attributes are set / tables that are uploaded on this device
  std::unordered_map<srk_tmap_key, CUtensorMap, srk_tmap_key_hash> tmaps;
  std::map<std::array<uint64_t, 6>, CUtensorMap> tmaps_fpa;  // 4-D FPA maps (make_tensor_map_fpa4)
  int conv_form = 0;         // SRK_CONV_FORM_*: kernel form of the plain 3x3 64->64 layers (srk_set_conv_form)
  void* comm = nullptr;      // ncclComm_t once srk_comm_init has run (collective.cu)
  int comm_world = 1;
  void* peer = nullptr;      // PeerState once srk_peer_alloc has run (peer_reduce.cu)
  void* host_pipe = nullptr; // copy streams + events of srk_espcn_forward_host (espcn_fused.cu)
};

namespace srk {

void set_error(const char* fmt, ...);

#define SRK_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t e_ = (expr);                                                          \
    if (e_ != cudaSuccess) {                                                          \
      ::srk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return -2;                                                                      \
    }                                                                                 \
  } while (0)

#define SRK_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::srk::set_error(__VA_ARGS__);  \
      return -1;                      \
    }                                 \
  } while (0)

#define SRK_LAUNCH_CHECK()                                   \
  do {                                                       \
    cudaError_t e_ = cudaGetLastError();                     \
    if (e_ != cudaSuccess) {                                 \
      ::srk::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return -3;                                             \
    }                                                        \
  } while (0)

// FPA geometry (see include/srk.h).
struct FpaGeom {
  int n_img, H, W, Wp, S;
  int64_t rows_valid;  // n_img * S
  int64_t rows_alloc;  // multiple of 128, >= rows_valid + Wp + 1
};
inline FpaGeom fpa_geom(int n_img, int H, int W) {
  FpaGeom g;
  g.n_img = n_img;
  g.H = H;
  g.W = W;
  g.Wp = W + 1;
  g.S = (H + 1) * (W + 1);
  g.rows_valid = int64_t(n_img) * g.S;
  g.rows_alloc = ((g.rows_valid + g.Wp + 1 + 127) / 128) * 128;
  return g;
}

// 2-D bf16 row-major tensor map [rows][cols], box {cols, box_rows}, swizzle chosen from the row bytes
// (128 B -> SW128, 64 B -> SW64, 32 B -> SW32).  Returns 0 on success.  The encoded map is cached in the handle (a model
// replays the same few buffers every step, and cuTensorMapEncodeTiled costs microseconds of host time per call).
int make_tensor_map_2d(srk_ctx* h, CUtensorMap* out, const void* gptr, uint64_t rows, uint32_t cols, uint32_t box_rows);

// 3-D row-major tensor map [batch][rows][cols] of 2- or 4-byte elements with explicit strides (bytes, multiples of 16), box
// {128 bytes of columns, box_rows, 1}, SWIZZLE_128B, zero fill out of bounds: the operand tiles of the generic GEMM.
int make_tensor_map_3d(srk_ctx* h, CUtensorMap* out, const void* gptr, uint32_t elem_bytes, uint64_t rows, uint64_t cols, uint64_t batch,
                       uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_rows);

// 4-D bf16 view of an FPA of 64 channels: dims {64 channels, Wp pixels, H+1 image rows (row 0 = the zero row), n_img}, box
// {64, box_x, 1, box_n}, SWIZZLE_128B, zero fill out of bounds.  A box is box_n * box_x consecutive 128-byte rows in shared
// memory: one image row of box_n images side by side (conv_strip.cu).
int make_tensor_map_fpa4(srk_ctx* h, CUtensorMap* out, const void* gptr, int n_img, int H, int W, uint32_t box_x, uint32_t box_n);

// true exactly once per (handle, key): guards per-device one-time work such as cudaFuncSetAttribute or a __constant__ upload
inline bool first_use(srk_ctx* h, const void* key) { return h->once.insert(key).second; }

// Every compute entry point runs on the handle's device; a caller that switched the current device gets an error, not a
// launch on the wrong GPU.
int check_device(srk_ctx* h);

inline cudaStream_t as_stream(srk_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch: the kernel may become resident while its predecessor in the stream is still
// draining (its CTAs start on an SM the moment the predecessor's CTA there exits, instead of after the whole grid
// has finished plus a launch latency).  Every kernel launched this way calls pdl_wait() before its first access to
// global memory and pdl_launch_dependents() right after, so data dependencies stay exactly those of the stream.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// releases what srk_espcn_forward_host created on the handle (espcn_fused.cu)
void host_pipe_destroy(srk_ctx* h);

// Column-strip form of the 3x3 64->64 FPA convolution (conv_strip.cu): srk_conv_tc routes wide frames there.
bool conv_strip_applicable(srk_ctx* h, int n_img, int H, int W);
int launch_conv_strip(srk_ctx* h, const void* x_fpa, const void* w_packed, const float* bias, int act, int n_img, int H, int W, void* y_fpa,
                      const void* mask_src, cudaStream_t stream);

#ifdef __CUDACC__
// One element of tf.train.AdamOptimizer's update with the bias-corrected rate lr_t (TF's epsilon-hat form), optional masked l2
// decay folded into the gradient first: shared by srk_adam_step(_dev) and the fused exchange + Adam kernel, with explicit
// roundings so that both produce the same bits.
__device__ __forceinline__ void adam_update(float* w, float* m, float* v, size_t i, float gi, float lr_t, float b1, float b2, float eps, float wd,
                                            const float* mask) {
  const float wi = w[i];
  if (mask) gi = __fmaf_rn(__fmul_rn(wd, mask[i]), wi, gi);
  const float mi = __fmaf_rn(b1, m[i], __fmul_rn(1.f - b1, gi));
  const float vi = __fmaf_rn(b2, v[i], __fmul_rn(__fmul_rn(1.f - b2, gi), gi));
  m[i] = mi;
  v[i] = vi;
  w[i] = __fsub_rn(wi, __fdiv_rn(__fmul_rn(lr_t, mi), __fadd_rn(sqrtf(vi), eps)));
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

}  // namespace srk

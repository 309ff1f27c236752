// Shared host-side plumbing for libsrk: handle, error reporting, tensor-map encoding, FPA geometry.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <utility>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/srk.h"

struct srk_ctx {
  int device;
  int num_sms;
  int smem_optin;
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
// (128 B -> SW128, 64 B -> SW64, 32 B -> SW32).  Returns 0 on success.
int make_tensor_map_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint32_t cols, uint32_t box_rows);

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

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

}  // namespace srk

// Handle management, error text and tensor-map encoding for libsrk (include/srk.h).
#include <cstring>
#include <mutex>

#include "srk_common.cuh"

namespace srk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int check_device(srk_ctx* h) {
  SRK_REQUIRE(h != nullptr, "libsrk: null handle");
  int cur = -1;
  SRK_CHECK_CUDA(cudaGetDevice(&cur));
  SRK_REQUIRE(cur == h->device, "libsrk handle belongs to device %d but the current device is %d (one handle per device)", h->device, cur);
  return 0;
}

int make_tensor_map_2d(srk_ctx* h, CUtensorMap* out, const void* gptr, uint64_t rows, uint32_t cols, uint32_t box_rows) {
  const srk_tmap_key key{gptr, rows, cols, box_rows, cols, 2};
  auto it = h->tmaps.find(key);
  if (it != h->tmaps.end()) {
    *out = it->second;
    return 0;
  }
  if (h->tmaps.size() > 4096) h->tmaps.clear();  // bounded: a long-lived handle that sees ever new buffers starts over
  EncodeTiledFn enc = encode_fn();
  SRK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  const uint32_t row_bytes = cols * 2;
  CUtensorMapSwizzle sw;
  if (row_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  else if (row_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (row_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else {
    set_error("tensor map: unsupported row bytes %u", row_bytes);
    return -1;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_bytes};
  cuuint32_t box[2] = {cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (ptr %p rows %llu cols %u)", int(r), gptr,
              (unsigned long long)rows, cols);
  h->tmaps.emplace(key, *out);
  return 0;
}

int make_tensor_map_3d(srk_ctx* h, CUtensorMap* out, const void* gptr, uint32_t elem_bytes, uint64_t rows, uint64_t cols, uint64_t batch,
                       uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_rows) {
  SRK_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "tensor map: element size %u", elem_bytes);
  SRK_REQUIRE(row_stride_bytes % 16 == 0 && (batch <= 1 || batch_stride_bytes % 16 == 0) && reinterpret_cast<uintptr_t>(gptr) % 16 == 0,
              "tensor map: base and strides must be multiples of 16 bytes (row stride %llu, batch stride %llu)", (unsigned long long)row_stride_bytes,
              (unsigned long long)batch_stride_bytes);
  const uint32_t box_cols = 128 / elem_bytes;
  srk_tmap_key key{gptr, rows, uint32_t(cols), box_rows, box_cols, elem_bytes};
  key.row_stride = row_stride_bytes;
  key.batch = batch;
  key.batch_stride = batch_stride_bytes;
  auto it = h->tmaps.find(key);
  if (it != h->tmaps.end()) {
    *out = it->second;
    return 0;
  }
  if (h->tmaps.size() > 4096) h->tmaps.clear();
  EncodeTiledFn enc = encode_fn();
  SRK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  if (batch < 1) batch = 1;
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {row_stride_bytes, batch > 1 ? batch_stride_bytes : row_stride_bytes * rows};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(gptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed with CUresult %d (ptr %p rows %llu cols %llu batch %llu)", int(r), gptr,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)batch);
  h->tmaps.emplace(key, *out);
  return 0;
}

int make_tensor_map_fpa4(srk_ctx* h, CUtensorMap* out, const void* gptr, int n_img, int H, int W, uint32_t box_x, uint32_t box_n) {
  const std::array<uint64_t, 6> key{reinterpret_cast<uint64_t>(gptr), uint64_t(n_img), uint64_t(H), uint64_t(W), box_x, box_n};
  auto it = h->tmaps_fpa.find(key);
  if (it != h->tmaps_fpa.end()) {
    *out = it->second;
    return 0;
  }
  if (h->tmaps_fpa.size() > 4096) h->tmaps_fpa.clear();
  EncodeTiledFn enc = encode_fn();
  SRK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  SRK_REQUIRE(box_x >= 1 && box_x <= 256 && box_n >= 1 && box_n <= 256 && box_x * box_n <= 128, "tensor map: FPA box %u x %u", box_x, box_n);
  const uint64_t Wp = uint64_t(W) + 1, rows = uint64_t(H) + 1;
  cuuint64_t dims[4] = {64, Wp, rows, uint64_t(n_img)};
  cuuint64_t strides[3] = {128, Wp * 128, rows * Wp * 128};
  cuuint32_t box[4] = {64, box_x, 1, box_n};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(gptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SRK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (FPA 4-D) failed with CUresult %d (ptr %p n %d H %d W %d box %u x %u)", int(r), gptr, n_img, H, W,
              box_x, box_n);
  h->tmaps_fpa.emplace(key, *out);
  return 0;
}

}  // namespace srk

extern "C" {

int srk_version(void) { return 100; }

const char* srk_last_error(void) { return srk::g_err; }

int srk_create(int device, srk_handle_t* out) {
  SRK_REQUIRE(out != nullptr, "srk_create: out is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  SRK_REQUIRE(e == cudaSuccess && count > 0, "srk_create: no CUDA device (%s); libsrk has no CPU fallback",
              cudaGetErrorString(e));
  SRK_REQUIRE(device >= 0 && device < count, "srk_create: device %d out of range [0,%d)", device, count);
  cudaDeviceProp prop;
  SRK_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  SRK_REQUIRE(prop.major == 10, "srk_create: device %d is sm_%d%d; libsrk is built for sm_100a only", device, prop.major,
              prop.minor);
  srk_ctx* c = new srk_ctx;
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->smem_optin = int(prop.sharedMemPerBlockOptin);
  *out = c;
  return 0;
}

int srk_destroy(srk_handle_t h) {
  if (!h) return 0;
  srk_peer_close(h);              // unmaps the peers' exchange regions, frees the local one (peer_reduce.cu)
  srk_comm_destroy(h);            // ncclCommDestroy (collective.cu)
  srk::host_pipe_destroy(h);      // copy streams + events of srk_espcn_forward_host (espcn_fused.cu)
  delete h;
  return 0;
}

int srk_num_sms(srk_handle_t h) { return h ? h->num_sms : -1; }

int srk_set_conv_form(srk_handle_t h, int form) {
  SRK_REQUIRE(h && form >= SRK_CONV_FORM_AUTO && form <= SRK_CONV_FORM_STRIP, "srk_set_conv_form: bad argument (form %d)", form);
  h->conv_form = form;
  return 0;
}

int srk_memcpy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows, int to_device,
                       srk_stream_t stream) {
  SRK_REQUIRE(dst && src && width_bytes <= dst_pitch && width_bytes <= src_pitch, "srk_memcpy2d_async: bad argument");
  if (rows == 0 || width_bytes == 0) return 0;
  SRK_CHECK_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost,
                                   srk::as_stream(stream)));
  return 0;
}

int64_t srk_fpa_rows(int n_img, int H, int W) { return srk::fpa_geom(n_img, H, W).rows_alloc; }

}  // extern "C"

// Hardware probe (test infrastructure, not product): pins down the tcgen05 shared-memory
// descriptor semantics and the TMA swizzle pattern the conv kernels rely on, using exact
// small-integer bf16 data so every comparison is bit-exact.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu -lcuda
//   run  : ./umma_probe   (prints one PASS/FAIL line per experiment)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <cuda_bf16.h>
#include "../sm100_ptx.cuh"

using namespace srk;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int MAX_OPS = 64;
struct MmaOp {
  uint32_t a_off, b_off;   // byte offsets from the 1024-aligned smem base
  uint64_t a_hi, b_hi;     // descriptor templates (everything but the start address)
  uint32_t idesc, accum, tmem_col, pad;
};
struct ProbeParams {
  const uint8_t* img;  // smem image
  uint32_t img_bytes;
  uint32_t n_ops;
  uint32_t ncols;  // TMEM columns to dump (multiple of 32)
  float* out;      // [128][ncols]
  MmaOp ops[MAX_OPS];
};

__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sbase = smem_u32(smem);
  for (uint32_t i = threadIdx.x * 16; i < p.img_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(p.img + i);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc<256>(smem_u32(&tmem_base_s));
  fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the async (UMMA) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < p.n_ops; ++i) {
      const MmaOp& o = p.ops[i];
      umma_bf16(tmem + o.tmem_col, umma_desc(o.a_hi, sbase + o.a_off), umma_desc(o.b_hi, sbase + o.b_off),
                o.idesc, o.accum);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const uint32_t warp = threadIdx.x >> 5;
  for (uint32_t c = 0; c < p.ncols; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32b_x32(tmem + ((warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) p.out[threadIdx.x * p.ncols + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<256>(tmem);
}

// TMA probe: load one box into smem at dst_off, dump smem; optionally store smem box back to global.
struct TmaParams {
  uint32_t dst_off, box_bytes, dump_bytes;
  int c0, c1;
  uint8_t* dump;
  int do_store, s0, s1;
};
__global__ void __launch_bounds__(128, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out,
                 const __grid_constant__ TmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  for (uint32_t i = threadIdx.x; i < p.dump_bytes; i += blockDim.x) smem[i] = 0xEE;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(smem_u32(&bar), p.box_bytes);
    tma_load_2d(smem_u32(smem) + p.dst_off, &map_in, p.c0, p.c1, smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  for (uint32_t i = threadIdx.x; i < p.dump_bytes; i += blockDim.x) p.dump[i] = smem[i];
  __syncthreads();
  if (p.do_store && threadIdx.x == 0) {
    tma_store_2d(&map_out, p.s0, p.s1, smem_u32(smem) + p.dst_off);
    tma_store_commit();
    tma_store_wait_all<0>();
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
  return reinterpret_cast<EncodeTiledFn>(fn);
}
static CUtensorMap make_map_2d(void* gptr, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows,
                               CUtensorMapSwizzle sw) {
  static EncodeTiledFn enc = get_encode();
  CUtensorMap m;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, gptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed %d\n", int(r)); exit(2); }
  return m;
}

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return uint16_t(u >> 16);  // exact for the small integers used here
}
static float bf2f(uint16_t h) {
  uint32_t u = uint32_t(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
struct Mat {  // row-major [rows][cols] small integers
  int rows, cols;
  std::vector<float> v;
  Mat(int r, int c, uint32_t seed) : rows(r), cols(c), v(size_t(r) * c) {
    uint32_t s = seed * 2654435761u + 12345u;
    for (auto& x : v) {
      s = s * 1664525u + 1013904223u;
      x = float(int((s >> 24) % 5) - 2);
    }
  }
  float at(int r, int c) const { return v[size_t(r) * cols + c]; }
};
// byte offset of element (r,c) of a [rows][C] bf16 matrix stored with the 128B/64B swizzle; region
// base is 1024-aligned so the pattern phase equals absolute-address bits.
static uint32_t off_sw128(int r, int c) { return r * 128 + ((((c >> 3) ^ (r & 7)) & 7) << 4) + (c & 7) * 2; }
static uint32_t off_sw64(int r, int c) { return r * 64 + ((((c >> 3) ^ ((r >> 1) & 3)) & 3) << 4) + (c & 7) * 2; }
static void put(std::vector<uint8_t>& img, uint32_t off, float f) {
  uint16_t h = f2bf(f);
  memcpy(&img[off], &h, 2);
}
static void place_sw128(std::vector<uint8_t>& img, uint32_t base, const Mat& m) {
  for (int r = 0; r < m.rows; ++r)
    for (int c = 0; c < 64; ++c) put(img, base + off_sw128(r, c), m.at(r, c));
}
static void place_sw64(std::vector<uint8_t>& img, uint32_t base, const Mat& m) {
  for (int r = 0; r < m.rows; ++r)
    for (int c = 0; c < 32; ++c) put(img, base + off_sw64(r, c), m.at(r, c));
}

static int g_fail = 0;
static std::vector<float> run_probe(const std::vector<uint8_t>& img, const std::vector<MmaOp>& ops, int ncols) {
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  uint8_t* dimg;
  float* dout;
  CK(cudaMalloc(&dimg, img.size()));
  CK(cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dout, 128 * ncols * 4));
  CK(cudaMemset(dout, 0xFF, 128 * ncols * 4));
  p.img = dimg;
  p.img_bytes = uint32_t(img.size());
  p.n_ops = uint32_t(ops.size());
  p.ncols = ncols;
  p.out = dout;
  for (size_t i = 0; i < ops.size(); ++i) p.ops[i] = ops[i];
  size_t smem = img.size() + 1024;
  CK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  umma_probe_kernel<<<1, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> out(size_t(128) * ncols, -12345.f);
  if (e != cudaSuccess) {
    printf("  kernel error: %s\n", cudaGetErrorString(e));
    exit(3);  // sticky error: stop, later experiments would all fail
  }
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  cudaFree(dimg);
  cudaFree(dout);
  return out;
}
static void report(const char* name, const std::vector<float>& got, const std::vector<float>& want, int rows,
                   int cols, int ld) {
  int bad = 0, first = -1;
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c)
      if (got[size_t(r) * ld + c] != want[size_t(r) * cols + c]) {
        if (first < 0) first = r * cols + c;
        ++bad;
      }
  printf("%-58s %s", name, bad ? "FAIL" : "PASS");
  if (bad)
    printf("  (%d/%d wrong, first at r=%d c=%d got %g want %g)", bad, rows * cols, first / cols, first % cols,
           got[size_t(first / cols) * ld + first % cols], want[first]);
  printf("\n");
  fflush(stdout);
  if (bad) ++g_fail;
}

// D[m][n] = sum_k A[a_row0 + m][k] * B[n][k]   (both K-major)
static std::vector<float> ref_kmajor(const Mat& A, int a_row0, const Mat& B, int M, int N, int K) {
  std::vector<float> d(size_t(M) * N, 0.f);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += A.at(a_row0 + m, k) * B.at(n, k);
      d[size_t(m) * N + n] = s;
    }
  return d;
}

int main() {
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);

  // ---------------------------------------------------------------- E1/E2/E3: K-major SW128, row shifts
  {
    const int AROWS = 512;
    Mat A(AROWS, 64, 1), B(64, 64, 2);
    const uint32_t A_OFF = 0, B_OFF = AROWS * 128;
    std::vector<uint8_t> img(B_OFF + 64 * 128, 0);
    place_sw128(img, A_OFF, A);
    place_sw128(img, B_OFF, B);
    const int shifts[] = {0, 1, 2, 3, 5, 8, 42, 43, 129, 257};
    for (int bo_mode = 0; bo_mode < 2; ++bo_mode)
      for (int s : shifts) {
        if (bo_mode == 1 && (s % 8) == 0) continue;
        std::vector<MmaOp> ops;
        for (int k = 0; k < 4; ++k) {
          MmaOp o{};
          o.a_off = A_OFF + s * 128 + k * 32;
          o.b_off = B_OFF + k * 32;
          o.a_hi = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128, bo_mode ? (s & 7) : 0);
          o.b_hi = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
          o.idesc = umma_idesc_bf16(128, 64, 0, 0);
          o.accum = k > 0;
          ops.push_back(o);
        }
        auto got = run_probe(img, ops, 64);
        auto want = ref_kmajor(A, s, B, 128, 64, 64);
        char name[128];
        snprintf(name, sizeof name, "E%d K-major SW128 M128 N64 K64 rowshift=%d base_offset=%d", bo_mode ? 3 : 2, s,
                 bo_mode ? (s & 7) : 0);
        report(name, got, want, 128, 64, 64);
      }
    // N = 32 and N = 16 (rows of B beyond N ignored)
    for (int N : {32, 16}) {
      std::vector<MmaOp> ops;
      for (int k = 0; k < 4; ++k) {
        MmaOp o{};
        o.a_off = A_OFF + 3 * 128 + k * 32;
        o.b_off = B_OFF + k * 32;
        o.a_hi = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
        o.b_hi = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
        o.idesc = umma_idesc_bf16(128, N, 0, 0);
        o.accum = k > 0;
        ops.push_back(o);
      }
      auto got = run_probe(img, ops, 64);
      auto want = ref_kmajor(A, 3, B, 128, N, 64);
      char name[128];
      snprintf(name, sizeof name, "E4 K-major SW128 M128 N%d K64 rowshift=3", N);
      report(name, got, want, 128, N, 64);
    }
  }
  // ---------------------------------------------------------------- E5: K-major SW64 (32-channel rows)
  {
    const int AROWS = 512;
    Mat A(AROWS, 32, 3), B(32, 32, 4);
    const uint32_t A_OFF = 0, B_OFF = AROWS * 64;
    std::vector<uint8_t> img(B_OFF + 32 * 64, 0);
    place_sw64(img, A_OFF, A);
    place_sw64(img, B_OFF, B);
    for (int s : {0, 1, 2, 3, 43, 130}) {
      std::vector<MmaOp> ops;
      for (int k = 0; k < 2; ++k) {
        MmaOp o{};
        o.a_off = A_OFF + s * 64 + k * 32;
        o.b_off = B_OFF + k * 32;
        o.a_hi = umma_desc_hi(0, 512, UMMA_LAYOUT_SW64);
        o.b_hi = umma_desc_hi(0, 512, UMMA_LAYOUT_SW64);
        o.idesc = umma_idesc_bf16(128, 32, 0, 0);
        o.accum = k > 0;
        ops.push_back(o);
      }
      auto got = run_probe(img, ops, 32);
      auto want = ref_kmajor(A, s, B, 128, 32, 32);
      char name[128];
      snprintf(name, sizeof name, "E5 K-major SW64 M128 N32 K32 rowshift=%d", s);
      report(name, got, want, 128, 32, 32);
    }
  }
  // ---------------------------------------------------------------- E6: MN-major SW128 (wgrad form)
  // X image [pixel rows][64 ci], dY image [pixel rows][64 co].  D[m][n] = sum_k X[s + k + (m>=64)*d][m%64] * dY[k][n]
  {
    const int XROWS = 512, KTOT = 128;
    Mat X(XROWS, 64, 5), dY(KTOT, 64, 6);
    const uint32_t X_OFF = 0, Y_OFF = XROWS * 128;
    std::vector<uint8_t> img(Y_OFF + KTOT * 128, 0);
    place_sw128(img, X_OFF, X);
    place_sw128(img, Y_OFF, dY);
    struct Cfg { int s, d, swap; };
    // negative d: second atom BEFORE the first one, LBO encoded modulo 2^18 bytes (14-bit field of 16 B units)
    const Cfg cfgs[] = {{0, 8, 0}, {0, 1, 0}, {1, 1, 0}, {3, 42, 0}, {43, 84, 0}, {0, 8, 1}, {3, 42, 1},
                        {50, -1, 0}, {50, -8, 0}, {300, -257, 0}, {131, -42, 0}};
    for (const Cfg& c : cfgs) {
      std::vector<MmaOp> ops;
      for (int k = 0; k < KTOT / 16; ++k) {
        MmaOp o{};
        o.a_off = X_OFF + (c.s + 16 * k) * 128;
        o.b_off = Y_OFF + (16 * k) * 128;
        uint32_t lbo = uint32_t(c.d * 128) & 0x3FFFFu, sbo = 1024;
        o.a_hi = c.swap ? umma_desc_hi(sbo, lbo, UMMA_LAYOUT_SW128) : umma_desc_hi(lbo, sbo, UMMA_LAYOUT_SW128);
        o.b_hi = c.swap ? umma_desc_hi(1024, 1024, UMMA_LAYOUT_SW128) : umma_desc_hi(1024, 1024, UMMA_LAYOUT_SW128);
        o.idesc = umma_idesc_bf16(128, 64, 1, 1);
        o.accum = k > 0;
        ops.push_back(o);
      }
      auto got = run_probe(img, ops, 64);
      std::vector<float> want(128 * 64, 0.f);
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          float sum = 0;
          for (int k = 0; k < KTOT; ++k) sum += X.at(c.s + k + (m >= 64 ? c.d : 0), m % 64) * dY.at(k, n);
          want[m * 64 + n] = sum;
        }
      char name[128];
      snprintf(name, sizeof name, "E6 MN-major SW128 M128(2 atoms) N64 K128 s=%d d=%d %s", c.s, c.d,
               c.swap ? "LBO/SBO swapped" : "LBO=atom SBO=kgroup");
      report(name, got, want, 128, 64, 64);
    }
    // M = 64 single atom (TMEM layout for M=64 dumped raw for inspection: lanes 0..127)
    {
      std::vector<MmaOp> ops;
      for (int k = 0; k < KTOT / 16; ++k) {
        MmaOp o{};
        o.a_off = X_OFF + (16 * k) * 128;
        o.b_off = Y_OFF + (16 * k) * 128;
        o.a_hi = umma_desc_hi(1024, 1024, UMMA_LAYOUT_SW128);
        o.b_hi = umma_desc_hi(1024, 1024, UMMA_LAYOUT_SW128);
        o.idesc = umma_idesc_bf16(64, 64, 1, 1);
        o.accum = k > 0;
        ops.push_back(o);
      }
      auto got = run_probe(img, ops, 64);
      // hypothesis: row m lives in lane m (lanes 0..63)
      std::vector<float> want(64 * 64, 0.f);
      for (int m = 0; m < 64; ++m)
        for (int n = 0; n < 64; ++n) {
          float sum = 0;
          for (int k = 0; k < KTOT; ++k) sum += X.at(k, m) * dY.at(k, n);
          want[m * 64 + n] = sum;
        }
      report("E7 MN-major SW128 M64 N64 K128 (row m -> lane m?)", got, want, 64, 64, 64);
      // alternative hypothesis: row m -> lane (m%16) + 32*(m/16)
      std::vector<float> got2(64 * 64);
      for (int m = 0; m < 64; ++m)
        for (int n = 0; n < 64; ++n) got2[m * 64 + n] = got[((m % 16) + 32 * (m / 16)) * 64 + n];
      report("E7b same, row m -> lane (m%16)+32*(m/16)?", got2, want, 64, 64, 64);
    }
  }
  // ---------------------------------------------------------------- E8: TMA load swizzle + OOB zero fill, TMA store
  {
    const int ROWS = 300;
    Mat G(ROWS, 64, 7);
    std::vector<uint16_t> hg(size_t(ROWS) * 64);
    for (int r = 0; r < ROWS; ++r)
      for (int c = 0; c < 64; ++c) hg[size_t(r) * 64 + c] = f2bf(G.at(r, c));
    uint16_t *dg, *dgo;
    CK(cudaMalloc(&dg, hg.size() * 2));
    CK(cudaMalloc(&dgo, hg.size() * 2));
    CK(cudaMemcpy(dg, hg.data(), hg.size() * 2, cudaMemcpyHostToDevice));
    uint8_t* ddump;
    const uint32_t DUMP = 128 * 128 + 2048;
    CK(cudaMalloc(&ddump, DUMP));
    CUtensorMap mi = make_map_2d(dg, 64, ROWS, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    CUtensorMap mo = make_map_2d(dgo, 64, ROWS, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    CK(cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(DUMP + 1024)));
    const int row0s[] = {0, -3, 200};
    for (int row0 : row0s) {
      CK(cudaMemset(dgo, 0x11, hg.size() * 2));
      TmaParams tp{1024, 128 * 128, DUMP, 0, row0, ddump, 1, 0, row0 < 0 ? 5 : row0};
      tma_probe_kernel<<<1, 128, DUMP + 1024>>>(mi, mo, tp);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  tma kernel error: %s\n", cudaGetErrorString(e)); exit(3); }
      std::vector<uint8_t> dump(DUMP);
      CK(cudaMemcpy(dump.data(), ddump, DUMP, cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 64; ++c) {
          int gr = row0 + r;
          float want = (gr >= 0 && gr < ROWS) ? G.at(gr, c) : 0.f;
          uint16_t h;
          memcpy(&h, &dump[1024 + off_sw128(r, c)], 2);
          if (bf2f(h) != want) ++bad;
        }
      char name[128];
      snprintf(name, sizeof name, "E8 TMA load SW128 box{64,128} row0=%d (swizzle model + OOB zero)", row0);
      printf("%-58s %s\n", name, bad ? "FAIL" : "PASS");
      if (bad) { ++g_fail; printf("  %d wrong\n", bad); }
      // store check: smem box -> global rows [s1, s1+128) clipped
      std::vector<uint16_t> ho(hg.size());
      CK(cudaMemcpy(ho.data(), dgo, ho.size() * 2, cudaMemcpyDeviceToHost));
      int s1 = tp.s1, bad2 = 0;
      for (int r = 0; r < ROWS; ++r)
        for (int c = 0; c < 64; ++c) {
          uint16_t want = 0x1111;
          if (r >= s1 && r < s1 + 128) {
            int gr = row0 + (r - s1);
            want = f2bf((gr >= 0 && gr < ROWS) ? G.at(gr, c) : 0.f);
          }
          if (ho[size_t(r) * 64 + c] != want) ++bad2;
        }
      snprintf(name, sizeof name, "E9 TMA store SW128 box{64,128} to row %d (clipped at end)", s1);
      printf("%-58s %s\n", name, bad2 ? "FAIL" : "PASS");
      if (bad2) { ++g_fail; printf("  %d wrong\n", bad2); }
    }
    // SW64 map over a 32-column matrix
    {
      Mat G2(ROWS, 32, 8);
      std::vector<uint16_t> h2(size_t(ROWS) * 32);
      for (int r = 0; r < ROWS; ++r)
        for (int c = 0; c < 32; ++c) h2[size_t(r) * 32 + c] = f2bf(G2.at(r, c));
      uint16_t* d2;
      CK(cudaMalloc(&d2, h2.size() * 2));
      CK(cudaMemcpy(d2, h2.data(), h2.size() * 2, cudaMemcpyHostToDevice));
      CUtensorMap m2 = make_map_2d(d2, 32, ROWS, 32, 128, CU_TENSOR_MAP_SWIZZLE_64B);
      TmaParams tp{1024, 128 * 64, DUMP, 0, 7, ddump, 0, 0, 0};
      tma_probe_kernel<<<1, 128, DUMP + 1024>>>(m2, m2, tp);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("  tma kernel error: %s\n", cudaGetErrorString(e)); exit(3); }
      std::vector<uint8_t> dump(DUMP);
      CK(cudaMemcpy(dump.data(), ddump, DUMP, cudaMemcpyDeviceToHost));
      int bad = 0;
      for (int r = 0; r < 128; ++r)
        for (int c = 0; c < 32; ++c) {
          uint16_t h;
          memcpy(&h, &dump[1024 + off_sw64(r, c)], 2);
          if (bf2f(h) != G2.at(7 + r, c)) ++bad;
        }
      printf("%-58s %s\n", "E10 TMA load SW64 box{32,128} row0=7", bad ? "FAIL" : "PASS");
      if (bad) { ++g_fail; printf("  %d wrong\n", bad); }
    }
  }
  printf("probe done: %d failing experiments\n", g_fail);
  return 0;
}

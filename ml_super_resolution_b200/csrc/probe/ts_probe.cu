// Hardware probe (test infrastructure, not product): tcgen05.mma with the A operand in TENSOR MEMORY (TS mode),
// which the fused ESPCN kernel is built on, plus the epilogue-side rates that bound it.
//   1. layout : A written by tcgen05.st.32x32b (thread i <-> lane i = row i, bf16 pairs packed in consecutive
//               32-bit columns), B = [N][K] bf16 K-major SW128 / SW64 in shared memory; result compared with a CPU GEMM
//   2. rate   : cycles per TS-mode MMA (M=128, K=16) for N = 16..256
//   3. rates of tanh.approx.f32 / tanh.approx.bf16x2 / ex2, tcgen05.ld, tcgen05.st and shfl with 4..16 busy warps
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o ts_probe ts_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../sm100_ptx.cuh"
using namespace srk;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ 1. layout
// A [128][KT] bf16 (global, row-major), B image = pre-swizzled smem bytes.  D [128][N] fp32 out.
template <int N, int KT, int SW /*128 or 64*/>
__global__ void __launch_bounds__(128, 1) ts_layout_kernel(const __nv_bfloat16* A, const uint8_t* Bimg, int b_bytes, float* D) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x * 16; i < b_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(Bimg + i);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&tslot));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t a_tmem = tmem + 256;  // A operand at columns [256, 256 + KT/2)
  // every thread writes its row of A: KT bf16 = KT/2 columns
  for (int c = 0; c < KT / 2; c += 8) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const uint32_t*>(A + threadIdx.x * KT + 2 * (c + j));
    tmem_st_32x32b_x8(a_tmem + ((warp * 32) << 16) + c, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    constexpr int kRowBytes = SW;  // one swizzle row = SW bytes of K
    constexpr uint64_t hi = umma_desc_hi(0, 8 * kRowBytes, SW == 128 ? UMMA_LAYOUT_SW128 : UMMA_LAYOUT_SW64);
    constexpr int kPerRow = kRowBytes / 2;  // K elements per swizzle row
    for (int k = 0; k < KT / 16; ++k) {
      const int blk = (k * 16) / kPerRow, within = (k * 16) % kPerRow;
      const uint32_t b_addr = smem_u32(smem) + blk * (N * kRowBytes) + within * 2;
      umma_bf16_ts(tmem, a_tmem + k * 8, umma_desc(hi, b_addr), idesc, k != 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 16) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(tmem + ((warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[threadIdx.x * N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int N, int KT, int SW>
void run_layout() {
  std::vector<__nv_bfloat16> A(128 * KT), B(N * KT);
  std::vector<float> Af(128 * KT), Bf(N * KT);
  srand(7);
  for (int i = 0; i < 128 * KT; ++i) { float v = float((rand() % 17) - 8); A[i] = __float2bfloat16(v); Af[i] = v; }
  for (int i = 0; i < N * KT; ++i) { float v = float((rand() % 9) - 4); B[i] = __float2bfloat16(v); Bf[i] = v; }
  // B image: K blocks of (SW/2) elements; block b = [N rows][SW bytes], 16-byte chunks XOR-swizzled with the row
  const int kPerRow = SW / 2, nblk = KT / kPerRow, chunks = SW / 16;
  std::vector<uint8_t> img(size_t(nblk) * N * SW, 0);
  for (int b = 0; b < nblk; ++b)
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < chunks; ++c) {
        const int sw = (SW == 128) ? (n & 7) : ((n >> 1) & 3);
        uint8_t* dst = img.data() + size_t(b) * N * SW + size_t(n) * SW + ((c ^ sw) << 4);
        memcpy(dst, &B[n * KT + b * kPerRow + c * 8], 16);
      }
  __nv_bfloat16* dA; uint8_t* dB; float* dD;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, img.size())); CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, img.data(), img.size(), cudaMemcpyHostToDevice));
  const int smem = int(img.size()) + 2048;
  CK(cudaFuncSetAttribute(ts_layout_kernel<N, KT, SW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  ts_layout_kernel<N, KT, SW><<<1, 128, smem>>>(dA, dB, int(img.size()), dD);
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float ref = 0;
      for (int k = 0; k < KT; ++k) ref += Af[m * KT + k] * Bf[n * KT + k];
      maxerr = fmax(maxerr, fabs(ref - D[m * N + n]));
    }
  printf("TS layout  M=128 N=%3d K=%3d SW%d : max |err| = %g  %s\n", N, KT, SW, maxerr, maxerr == 0 ? "PASS" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

// ------------------------------------------------------------------------------------------------ 2. TS rate
template <int N>
__global__ void __launch_bounds__(128, 1) ts_rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&tslot));
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
    constexpr uint64_t hi = umma_desc_hi(0, 1024, UMMA_LAYOUT_SW128);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t b = smem_u32(smem) + ((i >> 2) % 3) * (N * 128) * 0 + (i & 3) * 32;
      umma_bf16_ts(tmem + (i & 1) * 0, tmem + 256 + (i & 3) * 8 + ((i >> 2) % 3) * 32, umma_desc(hi, b), idesc, 1);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}
template <int N>
void run_rate() {
  const int grid = 148, iters = 4096;
  long long* d; CK(cudaMalloc(&d, grid * 8));
  const int smem = 66 * 1024;
  CK(cudaFuncSetAttribute(ts_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  for (int r = 0; r < 2; ++r) { ts_rate_kernel<N><<<grid, 128, smem>>>(iters, d); CK(cudaDeviceSynchronize()); }
  std::vector<long long> h(grid);
  CK(cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost));
  long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
  printf("TS rate    M=128 N=%3d K=16 : %.1f cycles/MMA -> %.0f MAC/cyc/SM\n", N, double(mx) / iters, 128.0 * N * 16 / (double(mx) / iters));
  cudaFree(d);
}

// ------------------------------------------------------------------------------------------------ 3. epilogue-side rates
// mode 0 tanh.approx.f32, 1 tanh.approx.bf16x2, 2 ex2.approx.f32, 3 shfl, 4 tcgen05.ld x16, 5 tcgen05.ld x32, 6 tcgen05.st x16,
// 7 FFMA polynomial tanh (degree-9 odd, clamp)
template <int MODE>
__global__ void __launch_bounds__(512, 1) unit_rate_kernel(int iters, long long* out, float* sink) {
  __shared__ uint32_t tslot;
  if (threadIdx.x < 32) tmem_alloc<512>(smem_u32(&tslot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot, warp = threadIdx.x >> 5;
  const uint32_t taddr = tmem + (((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  float x[8];
  uint32_t u[8];
  for (int j = 0; j < 8; ++j) { x[j] = 0.001f * float(threadIdx.x + j); u[j] = 0x3C003C00u + threadIdx.x + j; }
  uint32_t r16[16], r32[32];
  for (int j = 0; j < 16; ++j) r16[j] = j;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[j]));
    } else if (MODE == 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u[j]));
    } else if (MODE == 2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
    } else if (MODE == 3) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = __shfl_sync(0xffffffffu, x[j], (threadIdx.x + 1) & 31);
    } else if (MODE == 4) {
      tmem_ld_32x32b_x16(taddr, r16);
      tmem_ld_wait();
    } else if (MODE == 5) {
      tmem_ld_32x32b_x32(taddr, r32);
      tmem_ld_wait();
    } else if (MODE == 6) {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
                   "r"(r16[0]), "r"(r16[1]), "r"(r16[2]), "r"(r16[3]), "r"(r16[4]), "r"(r16[5]), "r"(r16[6]), "r"(r16[7]), "r"(r16[8]), "r"(r16[9]),
                   "r"(r16[10]), "r"(r16[11]), "r"(r16[12]), "r"(r16[13]), "r"(r16[14]), "r"(r16[15])
                   : "memory");
      tmem_st_wait();
    } else if (MODE == 7) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float c = fminf(fmaxf(x[j], -3.f), 3.f), s = c * c;
        float p = fmaf(s, -2.1e-5f, 6.4e-4f);
        p = fmaf(p, s, -8.3e-3f);
        p = fmaf(p, s, 6.2e-2f);
        p = fmaf(p, s, -0.31f);
        p = fmaf(p, s, 1.f);
        x[j] = p * c + 1e-3f;
      }
    }
  }
  long long t1 = clock64();
  float acc = 0;
  for (int j = 0; j < 8; ++j) acc += x[j] + __uint_as_float(u[j]);
  if (MODE == 4) for (int j = 0; j < 16; ++j) acc += __uint_as_float(r16[j]);
  if (MODE == 5) for (int j = 0; j < 32; ++j) acc += __uint_as_float(r32[j]);
  if (acc == 12345.678f) sink[0] = acc;
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}
template <int MODE>
void run_unit(const char* name, int warps, int per_iter) {
  const int grid = 148, iters = 2048;
  long long* d; float* s;
  CK(cudaMalloc(&d, grid * 16 * 8)); CK(cudaMalloc(&s, 4));
  for (int r = 0; r < 2; ++r) { unit_rate_kernel<MODE><<<grid, warps * 32>>>(iters, d, s); CK(cudaDeviceSynchronize()); }
  std::vector<long long> h(grid * 16);
  CK(cudaMemcpy(h.data(), d, grid * 16 * 8, cudaMemcpyDeviceToHost));
  long long mx = 0; for (int b = 0; b < grid; ++b) for (int w = 0; w < warps; ++w) mx = h[b * 16 + w] > mx ? h[b * 16 + w] : mx;
  const double cyc = double(mx) / iters;
  printf("unit rate  %-28s warps=%2d : %.1f cycles per iteration per warp; %.2f warp-instructions/cycle/SM\n", name, warps, cyc, warps * per_iter / cyc);
  cudaFree(d); cudaFree(s);
}

int main() {
  run_layout<64, 64, 128>();
  run_layout<96, 64, 128>();
  run_layout<48, 32, 64>();
  run_layout<64, 32, 64>();
  run_layout<96, 32, 64>();
  run_rate<16>(); run_rate<32>(); run_rate<48>(); run_rate<64>(); run_rate<96>(); run_rate<128>(); run_rate<192>(); run_rate<256>();
  for (int w : {4, 8, 16}) {
    run_unit<0>("tanh.approx.f32 x8", w, 8);
    run_unit<1>("tanh.approx.bf16x2 x8", w, 8);
    run_unit<2>("ex2.approx.f32 x8", w, 8);
    run_unit<7>("poly tanh (6 FMA + clamp) x8", w, 8);
    run_unit<3>("shfl.sync x8", w, 8);
    run_unit<4>("tcgen05.ld 32x32b.x16 + wait", w, 1);
    run_unit<5>("tcgen05.ld 32x32b.x32 + wait", w, 1);
    run_unit<6>("tcgen05.st 32x32b.x16 + wait", w, 1);
  }
  return 0;
}

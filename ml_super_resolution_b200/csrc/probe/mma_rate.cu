// Microbenchmark (test infrastructure): tcgen05.mma issue rate for M=128 SS-mode bf16 as a function of N
// with all operands resident in shared memory -- decides whether the N=64 conv mainloop is bound by
// tensor math (32 cyc/MMA) or by the shared-memory operand fetch.
#include <cstdio>
#include <cstdlib>
#include "../sm100_ptx.cuh"
using namespace srk;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

template <int N, int KMAJOR_LAYOUT>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int a_stride_rows, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sb = smem_u32(smem);
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
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
    const uint32_t a_base = sb, b_base = sb + 96 * 1024;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t a = a_base + ((i >> 2) % 9) * a_stride_rows * 128 + (i & 3) * 32;
      const uint32_t b = b_base + (i & 3) * 32;
      umma_bf16(tmem + (i & 1) * 256 * 0, umma_desc(hi, a), umma_desc(hi, b), idesc, 1);
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
void run(int grid, int iters, int stride) {
  long long* d;
  CK(cudaMalloc(&d, grid * 8));
  const int smem = 161 * 1024 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel<N, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  rate_kernel<N, 0><<<grid, 128, smem>>>(iters, stride, d);
  CK(cudaDeviceSynchronize());
  rate_kernel<N, 0><<<grid, 128, smem>>>(iters, stride, d);
  CK(cudaDeviceSynchronize());
  long long* h = (long long*)malloc(grid * 8);
  CK(cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost));
  long long mx = 0, mn = 1LL << 60;
  for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
  printf("M=128 N=%3d K=16 SS  grid=%3d  a_stride_rows=%3d : %.1f cycles/MMA (min CTA %.1f)  -> %.0f MAC/cyc/SM\n", N, grid, stride,
         double(mx) / iters, double(mn) / iters, 128.0 * N * 16 / (double(mx) / iters));
  cudaFree(d);
  free(h);
}

int main() {
  for (int grid : {148}) {
    run<64>(grid, 4096, 1);
    run<64>(grid, 4096, 43);
    run<32>(grid, 4096, 43);
    run<16>(grid, 4096, 43);
    run<96>(grid, 4096, 43);
    run<128>(grid, 4096, 43);
    run<192>(grid, 4096, 43);
    run<256>(grid, 4096, 43);
  }
  return 0;
}

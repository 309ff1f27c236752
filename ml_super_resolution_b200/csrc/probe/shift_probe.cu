// Probe (test infrastructure): semantics and cost of tcgen05.shift.down -- which lanes move, in which direction,
// how many columns one instruction covers, whether the shift crosses the 32-lane quadrants, and cycles per shift.
#include <cstdio>
#include <cstdlib>
#include "../sm100_ptx.cuh"
using namespace srk;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2);} } while (0)

__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
               "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_shift_down(uint32_t taddr) {
  asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory");
}

// out[lane][16 cols] after: fill value = lane*100 + col; `nshift` shifts at column offset `col_off`, lane base `lane_base`
__global__ void __launch_bounds__(128, 1) shift_kernel(int nshift, int col_off, int lane_base, int timing_iters, uint32_t* out, long long* cyc) {
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc<64>(smem_u32(&tslot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  const uint32_t taddr = tmem + (uint32_t(warp * 32) << 16);
  uint32_t v[16];
  for (int c = 0; c < 16; ++c) v[c] = (warp * 32 + lane) * 100 + c;
  tmem_st_x16(taddr, v);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    for (int i = 0; i < nshift; ++i) tmem_shift_down(tmem + col_off + (uint32_t(lane_base) << 16));
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    if (timing_iters > 0) {
      long long t0 = clock64();
      for (int i = 0; i < timing_iters; ++i) tmem_shift_down(tmem + 32 + (uint32_t(0) << 16));
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 1);
      cyc[0] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t u[16];
  tmem_ld_32x32b_x16(taddr, u);
  tmem_ld_wait();
  for (int c = 0; c < 16; ++c) out[threadIdx.x * 16 + c] = u[c];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

int main() {
  uint32_t* d;
  long long* dc;
  CK(cudaMalloc(&d, 128 * 16 * 4));
  CK(cudaMalloc(&dc, 8));
  uint32_t h[128 * 16];
  struct { int n, col, lane, it; } cases[] = {{1, 0, 0, 0}, {2, 0, 0, 0}, {1, 8, 0, 0}, {1, 0, 32, 0}, {1, 4, 0, 0}, {1, 0, 0, 1000}};
  for (auto cs : cases) {
    CK(cudaMemset(dc, 0, 8));
    shift_kernel<<<1, 128>>>(cs.n, cs.col, cs.lane, cs.it, d, dc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost));
    long long c = 0;
    CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
    printf("--- %d x shift.down at col %d lane %d%s\n", cs.n, cs.col, cs.lane, cs.it ? " (+timing)" : "");
    // summarise: for each column, which source lane each lane now holds
    for (int col = 0; col < 16; ++col) {
      int moved = 0, first_bad = -1;
      int delta_hist[7] = {0};
      for (int l = 0; l < 128; ++l) {
        const uint32_t x = h[l * 16 + col];
        const int src_lane = int(x / 100), src_col = int(x % 100);
        const int dl = src_lane - l;
        if (src_col != col && first_bad < 0) first_bad = l;
        if (dl != 0) ++moved;
        if (dl >= -3 && dl <= 3) ++delta_hist[dl + 3];
      }
      printf("  col %2d: moved lanes %3d  src-lane delta hist[-3..3] = %d %d %d %d %d %d %d%s\n", col, moved, delta_hist[0], delta_hist[1],
             delta_hist[2], delta_hist[3], delta_hist[4], delta_hist[5], delta_hist[6], first_bad >= 0 ? "  (column mixing!)" : "");
    }
    printf("  lanes 0,1,2,31,32,33,63,64,95,96,126,127 of col 0 hold source lanes:");
    for (int l : {0, 1, 2, 31, 32, 33, 63, 64, 95, 96, 126, 127}) printf(" %u", h[l * 16] / 100);
    printf("\n");
    if (cs.it) printf("  %d shifts: %lld cycles -> %.1f cycles/shift\n", cs.it, c, double(c) / cs.it);
  }
  return 0;
}

// The data-parallel exchange step FUSED with the optimiser, over NVLink peer memory (absent in the reference: one
// P100, vdsr/README.md:14 -- the call site it extends is the `minimize` of vdsr/vdsr/model_vdsr.py:146-148):
//
//     g_sum = sum over ranks of g_r  (fixed rank order)   +   Adam(w, g_sum)           in ONE kernel per rank, no NCCL call.
//
// Every rank owns an exchange region (cudaMalloc'ed here, exported as a cudaIpcMemHandle and opened by its peers): two gradient
// staging buffers and a flag array.  A block of the kernel (the same block index covers the same slice of the flat parameter
// arena on every rank)
//   A. copies its slice of the local gradient into the local staging buffer `epoch & 1`,
//   B. tells every peer "my slice of epoch e is staged" (one 4-byte store into the peer's flag array, after a system-scope
//      fence) and waits for the same word from every peer -- a per-slice barrier between GPUs, no grid-wide or host sync,
//   C. reads the slice from every rank's staging buffer through the peer mappings (NVLink loads, volatile: not cached in L1),
//      adds them in rank order -- the same order on every rank, so the replicas stay bit-identical -- writes the sum to the
//      local gradient arena and applies the Adam update of srk_adam_step_dev to the local replica.
// The staging buffers alternate per launch: a buffer is rewritten two launches later, when every peer has provably finished
// reading it (it signalled the epoch in between, which it does only after its previous launch has completed in stream order).
// The epoch counters live in device memory, so the kernel replays inside a CUDA graph.  Waits are bounded (a trap, not a hang).
#include <algorithm>

#include "srk_common.cuh"

namespace srk {

constexpr int kPeerMaxWorld = 8;
constexpr int kPeerBlocks = 128;
constexpr int kPeerThreads = 512;

struct PeerState {
  void* region = nullptr;        // local: [staging 0][staging 1][flags kPeerBlocks x kPeerMaxWorld][epoch kPeerBlocks]
  size_t elems = 0;              // floats per staging buffer (multiple of 4)
  void* peer[kPeerMaxWorld] = {};  // peers' regions (peer[rank] == region)
  int rank = 0, world = 1;
  bool opened = false;
};

struct PeerParams {
  const float* stage[kPeerMaxWorld];  // every rank's region base (staging buffers at +0 and +elems floats)
  unsigned int* flags[kPeerMaxWorld]; // every rank's flag array
  unsigned int* epoch;                // local, one word per block
  float* w;
  float* g;
  float* m;
  float* v;
  const float* lr_t;
  const float* mask;
  size_t n, elems;
  float b1, b2, eps, wd;
  int rank, world;
};

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_adam_kernel(const PeerParams p) {
  __shared__ unsigned int s_epoch;
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) s_epoch = p.epoch[b] + 1u;
  __syncthreads();
  const unsigned int epoch = s_epoch;
  const size_t n4 = p.n / 4;  // float4 groups; the tail (n % 4 elements) belongs to the last block
  const size_t i0 = n4 * b / gridDim.x, i1 = n4 * (b + 1) / gridDim.x;
  const size_t boff = (epoch & 1u) ? p.elems : 0;
  // ---- A: stage my slice
  {
    const float4* src = reinterpret_cast<const float4*>(p.g);
    float4* dst = reinterpret_cast<float4*>(const_cast<float*>(p.stage[p.rank]) + boff);
    for (size_t i = i0 + tid; i < i1; i += kPeerThreads) dst[i] = src[i];
    if (b == int(gridDim.x) - 1)
      for (size_t i = n4 * 4 + tid; i < p.n; i += kPeerThreads) const_cast<float*>(p.stage[p.rank])[boff + i] = p.g[i];
  }
  __threadfence_system();
  __syncthreads();
  // ---- B: per-slice barrier between the GPUs
  if (tid < p.world && tid != p.rank) {
    volatile unsigned int* theirs = p.flags[tid] + b * kPeerMaxWorld + p.rank;
    *theirs = epoch;
    volatile unsigned int* mine = p.flags[p.rank] + b * kPeerMaxWorld + tid;
    unsigned int spins = 0;
    while (int(*mine - epoch) < 0) {
      __nanosleep(40);
      if (++spins > (1u << 23)) __trap();
    }
  }
  __threadfence_system();
  __syncthreads();
  // ---- C: sum in rank order, Adam
  const float lr_t = __ldg(p.lr_t);
  auto adam1 = [&](size_t i, float gi) {
    p.g[i] = gi;
    adam_update(p.w, p.m, p.v, i, gi, lr_t, p.b1, p.b2, p.eps, p.wd, p.mask);
  };
  for (size_t i = i0 + tid; i < i1; i += kPeerThreads) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < p.world; ++r) {
      const float* q = p.stage[r] + boff + 4 * i;
      float4 t;
      asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "l"(q));
      s.x += t.x, s.y += t.y, s.z += t.z, s.w += t.w;
    }
    adam1(4 * i, s.x);
    adam1(4 * i + 1, s.y);
    adam1(4 * i + 2, s.z);
    adam1(4 * i + 3, s.w);
  }
  if (b == int(gridDim.x) - 1)
    for (size_t i = n4 * 4 + tid; i < p.n; i += kPeerThreads) {
      float s = 0.f;
      for (int r = 0; r < p.world; ++r) {
        float t;
        asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(t) : "l"(p.stage[r] + boff + i));
        s += t;
      }
      adam1(i, s);
    }
  if (tid == 0) p.epoch[b] = epoch;
}

static PeerState* peer_of(srk_ctx* h) { return static_cast<PeerState*>(h->peer); }
static size_t peer_region_bytes(size_t elems) { return 2 * elems * sizeof(float) + (kPeerBlocks * kPeerMaxWorld + kPeerBlocks) * sizeof(unsigned int); }

}  // namespace srk

using namespace srk;

extern "C" int srk_peer_alloc(srk_handle_t h, size_t grad_elems, void* ipc_handle_out64) {
  SRK_REQUIRE(h && grad_elems > 0 && ipc_handle_out64, "srk_peer_alloc: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI passes IPC handles as 64 bytes");
  if (h->peer) {
    if (int rc = srk_peer_close(h)) return rc;
  }
  PeerState* st = new PeerState();
  st->elems = (grad_elems + 3) / 4 * 4;
  const size_t bytes = peer_region_bytes(st->elems);
  cudaError_t e = cudaMalloc(&st->region, bytes);
  if (e != cudaSuccess) {
    delete st;
    set_error("srk_peer_alloc: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    return -2;
  }
  SRK_CHECK_CUDA(cudaMemset(st->region, 0, bytes));
  cudaIpcMemHandle_t hd;
  e = cudaIpcGetMemHandle(&hd, st->region);
  if (e != cudaSuccess) {
    cudaFree(st->region);
    delete st;
    set_error("srk_peer_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return -2;
  }
  memcpy(ipc_handle_out64, &hd, 64);
  h->peer = st;
  return 0;
}

extern "C" int srk_peer_open(srk_handle_t h, int rank, int world, const void* handles) {
  SRK_REQUIRE(h && h->peer && handles && world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "srk_peer_open: bad argument (world %d, rank %d)", world, rank);
  if (int rc_dev = check_device(h)) return rc_dev;
  PeerState* st = peer_of(h);
  SRK_REQUIRE(!st->opened, "srk_peer_open: already open");
  st->rank = rank;
  st->world = world;
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      st->peer[r] = st->region;
      continue;
    }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, static_cast<const char*>(handles) + 64 * r, 64);
    cudaError_t e = cudaIpcOpenMemHandle(&st->peer[r], hd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("srk_peer_open: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
      return -2;
    }
  }
  st->opened = true;
  return 0;
}

extern "C" int srk_allreduce_adam_step_dev(srk_handle_t h, float* w, float* g, float* m, float* v, size_t n, const float* lr_t_device, float beta1,
                                           float beta2, float eps, float weight_decay, const float* decay_mask, srk_stream_t stream) {
  SRK_REQUIRE(h && w && g && m && v && lr_t_device, "srk_allreduce_adam_step_dev: bad argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  PeerState* st = peer_of(h);
  SRK_REQUIRE(st && st->opened, "srk_allreduce_adam_step_dev: srk_peer_alloc / srk_peer_open have not been called on this handle");
  SRK_REQUIRE(n <= st->elems, "srk_allreduce_adam_step_dev: %zu elements, the exchange region holds %zu", n, st->elems);
  SRK_REQUIRE(reinterpret_cast<uintptr_t>(g) % 16 == 0, "srk_allreduce_adam_step_dev: gradient arena must be 16-byte aligned");
  PeerParams p{};
  for (int r = 0; r < st->world; ++r) {
    p.stage[r] = static_cast<const float*>(st->peer[r]);
    p.flags[r] = reinterpret_cast<unsigned int*>(static_cast<char*>(st->peer[r]) + 2 * st->elems * sizeof(float));
  }
  p.epoch = p.flags[st->rank] + kPeerBlocks * kPeerMaxWorld;
  p.w = w;
  p.g = g;
  p.m = m;
  p.v = v;
  p.lr_t = lr_t_device;
  p.mask = decay_mask;
  p.n = n;
  p.elems = st->elems;
  p.b1 = beta1;
  p.b2 = beta2;
  p.eps = eps;
  p.wd = decay_mask ? weight_decay : 0.f;
  p.rank = st->rank;
  p.world = st->world;
  const int grid = std::min(kPeerBlocks, h->num_sms);  // every rank must launch the SAME grid: the slices are matched by block index
  peer_allreduce_adam_kernel<<<grid, kPeerThreads, 0, as_stream(stream)>>>(p);
  SRK_LAUNCH_CHECK();
  return 0;
}

extern "C" int srk_peer_close(srk_handle_t h) {
  SRK_REQUIRE(h != nullptr, "srk_peer_close: null handle");
  PeerState* st = peer_of(h);
  if (!st) return 0;
  cudaDeviceSynchronize();
  for (int r = 0; r < st->world; ++r)
    if (st->opened && r != st->rank && st->peer[r]) cudaIpcCloseMemHandle(st->peer[r]);
  if (st->region) cudaFree(st->region);
  delete st;
  h->peer = nullptr;
  return 0;
}

// Instantiations of espcn_fused_kernel for 1-channel frames (espcn_fused.cuh): scaling factors 2, 3, 4, shuffled and packed output.
#include "espcn_fused.cuh"

namespace srk {

int launch_espcn_fused_c1(srk_ctx* h, EspcnFusedParams& p, int r, bool shuffle, const void* w1p, const void* w2p, const void* w3p, cudaStream_t stream) {
#define SRK_CASE(RR)                                                                                   \
  if (r == RR) return shuffle ? launch_espcn_fused<1, RR, true>(h, p, w1p, w2p, w3p, stream)         \
                              : launch_espcn_fused<1, RR, false>(h, p, w1p, w2p, w3p, stream);
  SRK_CASE(2)
  SRK_CASE(3)
  SRK_CASE(4)
#undef SRK_CASE
  set_error("srk_espcn_forward: unsupported scaling factor %d", r);
  return -1;
}

}  // namespace srk

#ifdef SRK_TRACE
extern "C" int srk_debug_ef_trace_read(unsigned long long* host_dst) {  // development builds only, not part of include/srk.h
  return int(cudaMemcpyFromSymbol(host_dst, srk::g_ef_trace, sizeof(srk::g_ef_trace)));
}
#endif

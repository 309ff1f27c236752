// C-ABI entry of the fused ESPCN forward (kernel: espcn_fused.cuh; instantiations: espcn_fused_c1.cu, espcn_fused_c3.cu).
#include "espcn_fused.cuh"

using namespace srk;

extern "C" int srk_espcn_forward(srk_handle_t h, const srk_espcn_net* net, const float* lr, int n, int H, int W, int y_begin, int y_end,
                                 int shuffle, int out_kind, void* out, srk_stream_t stream) {
  SRK_REQUIRE(h && net && lr && out, "srk_espcn_forward: null argument");
  if (int rc_dev = check_device(h)) return rc_dev;
  SRK_REQUIRE(net->w1_packed && net->w2_packed && net->w3_packed && net->b1 && net->b2 && net->b3, "srk_espcn_forward: incomplete srk_espcn_net");
  SRK_REQUIRE(n > 0 && H > 0 && W > 0 && y_begin >= 0 && y_end <= H, "srk_espcn_forward: bad geometry n=%d H=%d W=%d rows [%d,%d)", n, H, W, y_begin, y_end);
  SRK_REQUIRE(out_kind == SRK_OUT_F32 || out_kind == SRK_OUT_U8, "srk_espcn_forward: out_kind %d", out_kind);
  const int C = net->channels, r = net->scaling_factor;
  SRK_REQUIRE((C == 1 || C == 3) && r >= 2 && r <= 4, "srk_espcn_forward: unsupported (channels=%d, scaling_factor=%d)", C, r);
  SRK_REQUIRE(int64_t(n) * H * W * C * r * r < (int64_t(1) << 40), "srk_espcn_forward: output too large");
  if (y_end <= y_begin) return 0;
  // A CTA follows at most kEfMaxSegs strip segments: frames are processed in chunks of at most 62 * num_sms strips per launch
  // (one launch for anything but tens of thousands of tiny frames).
  const int strips = (W + kEfStripW - 1) / kEfStripW;
  SRK_REQUIRE(strips <= (kEfMaxSegs - 2) * h->num_sms, "srk_espcn_forward: frame width %d too large", W);
  const int chunk = std::max(1, (kEfMaxSegs - 2) * h->num_sms / strips);
  const size_t out_elem = out_kind == SRK_OUT_U8 ? 1 : 4;
  cudaStream_t s = as_stream(stream);
  for (int n0 = 0; n0 < n; n0 += chunk) {
    const int nc = std::min(chunk, n - n0);
    EspcnFusedParams p{};
    p.lr = lr + size_t(n0) * H * W * C;
    p.b1 = net->b1;
    p.b2 = net->b2;
    p.b3 = net->b3;
    p.out = static_cast<uint8_t*>(out) + size_t(n0) * H * W * C * r * r * out_elem;
    p.n = nc;
    p.H = H;
    p.W = W;
    p.y0 = y_begin;
    p.hb = y_end - y_begin;
    p.strips = strips;
    p.units = (long long)nc * strips * p.hb;
    p.out_kind = out_kind;
    SRK_REQUIRE(p.units < (1ll << 31), "srk_espcn_forward: too many strip rows for one launch");
    const int rc = C == 1 ? launch_espcn_fused_c1(h, p, r, shuffle != 0, net->w1_packed, net->w2_packed, net->w3_packed, s)
                          : launch_espcn_fused_c3(h, p, r, shuffle != 0, net->w1_packed, net->w2_packed, net->w3_packed, s);
    if (rc) return rc;
  }
  return 0;
}

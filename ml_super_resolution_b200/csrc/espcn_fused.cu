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


// ---------------------------------------------------------------------------------------------------------------------------------
// Host array in, host array out: what `session.run(sr, feed_dict={lr: frames})` does in the reference
// (espcn/espcn/experiment_test.py:164-184), with the copies hidden -- every frame is cut into row bands, and the host->device copy
// of band k+1, the fused kernel on band k and the device->host copy of band k-1 run on three streams.  The whole band loop lives
// here, not in Python: at ~100 us of interpreter time per band the host was a co-bottleneck of a 1.6 ms call.
// ---------------------------------------------------------------------------------------------------------------------------------
namespace srk {
struct HostPipe {
  cudaStream_t s_in = nullptr, s_out = nullptr;
  static constexpr int kEvents = 64;
  cudaEvent_t ev[kEvents] = {};
  int next = 0;
  cudaEvent_t get() {
    cudaEvent_t e = ev[next];
    next = (next + 1) % kEvents;
    return e;
  }
};
void host_pipe_destroy(srk_ctx* h) {
  HostPipe* hp = static_cast<HostPipe*>(h->host_pipe);
  if (!hp) return;
  cudaStreamSynchronize(hp->s_in);
  cudaStreamSynchronize(hp->s_out);
  for (int i = 0; i < HostPipe::kEvents; ++i) cudaEventDestroy(hp->ev[i]);
  cudaStreamDestroy(hp->s_in);
  cudaStreamDestroy(hp->s_out);
  delete hp;
  h->host_pipe = nullptr;
}
}  // namespace srk

extern "C" int srk_espcn_forward_host(srk_handle_t h, const srk_espcn_net* net, const void* lr_host, int lr_is_u8, int n, int H, int W, int shuffle,
                                      int out_kind, void* out_host, float* lr_dev, uint8_t* lr_u8_dev, void* out_dev, int band_rows,
                                      srk_stream_t stream) {
  SRK_REQUIRE(h && net && lr_host && out_host && lr_dev && out_dev && (!lr_is_u8 || lr_u8_dev), "srk_espcn_forward_host: null argument");
  SRK_REQUIRE(n > 0 && H > 0 && W > 0 && band_rows > 0, "srk_espcn_forward_host: bad geometry n=%d H=%d W=%d band_rows=%d", n, H, W, band_rows);
  if (int rc_dev = check_device(h)) return rc_dev;
  if (!h->host_pipe) {
    HostPipe* hp = new HostPipe();
    SRK_CHECK_CUDA(cudaStreamCreateWithFlags(&hp->s_in, cudaStreamNonBlocking));
    SRK_CHECK_CUDA(cudaStreamCreateWithFlags(&hp->s_out, cudaStreamNonBlocking));
    for (int i = 0; i < HostPipe::kEvents; ++i) SRK_CHECK_CUDA(cudaEventCreateWithFlags(&hp->ev[i], cudaEventDisableTiming));
    h->host_pipe = hp;
  }
  HostPipe* hp = static_cast<HostPipe*>(h->host_pipe);
  cudaStream_t s_c = as_stream(stream);
  const int C = net->channels, r = net->scaling_factor;
  const size_t in_row = size_t(W) * C;                                                  // samples per LR row
  const size_t out_elem = out_kind == SRK_OUT_U8 ? 1 : 4;
  const size_t out_row = shuffle ? size_t(W) * r * C * out_elem : size_t(W) * C * r * r * out_elem;  // bytes per OUTPUT row
  const int rows_per_lr = shuffle ? r : 1;
  // the side streams start after whatever the caller queued on its stream (weights re-packed, buffers released)
  cudaEvent_t e0 = hp->get();
  SRK_CHECK_CUDA(cudaEventRecord(e0, s_c));
  SRK_CHECK_CUDA(cudaStreamWaitEvent(hp->s_in, e0, 0));
  SRK_CHECK_CUDA(cudaStreamWaitEvent(hp->s_out, e0, 0));
  for (int f = 0; f < n; ++f) {
    int have = 0;  // input rows of frame f already on the device
    for (int y0 = 0; y0 < H; y0 += band_rows) {
      const int y1 = std::min(H, y0 + band_rows), need = std::min(H, y1 + 4);  // (4 LR rows of receptive field below the band)
      if (need > have) {
        const size_t off = (size_t(f) * H + have) * in_row, cnt = size_t(need - have) * in_row;
        if (lr_is_u8) {
          SRK_CHECK_CUDA(cudaMemcpyAsync(lr_u8_dev + off, static_cast<const uint8_t*>(lr_host) + off, cnt, cudaMemcpyHostToDevice, hp->s_in));
          if (int rc = srk_u8_to_pm1_f64(h, lr_u8_dev + off, cnt, lr_dev + off, reinterpret_cast<srk_stream_t>(hp->s_in))) return rc;
        } else {
          SRK_CHECK_CUDA(cudaMemcpyAsync(lr_dev + off, static_cast<const float*>(lr_host) + off, cnt * 4, cudaMemcpyHostToDevice, hp->s_in));
        }
        have = need;
        cudaEvent_t e_in = hp->get();
        SRK_CHECK_CUDA(cudaEventRecord(e_in, hp->s_in));
        SRK_CHECK_CUDA(cudaStreamWaitEvent(s_c, e_in, 0));
      }
      if (int rc = srk_espcn_forward(h, net, lr_dev + size_t(f) * H * in_row, 1, H, W, y0, y1, shuffle, out_kind,
                                     static_cast<uint8_t*>(out_dev) + size_t(f) * H * rows_per_lr * out_row, stream))
        return rc;
      cudaEvent_t e_c = hp->get();
      SRK_CHECK_CUDA(cudaEventRecord(e_c, s_c));
      SRK_CHECK_CUDA(cudaStreamWaitEvent(hp->s_out, e_c, 0));
      const size_t ooff = (size_t(f) * H + y0) * rows_per_lr * out_row, obytes = size_t(y1 - y0) * rows_per_lr * out_row;
      SRK_CHECK_CUDA(cudaMemcpyAsync(static_cast<uint8_t*>(out_host) + ooff, static_cast<const uint8_t*>(out_dev) + ooff, obytes, cudaMemcpyDeviceToHost,
                                     hp->s_out));
    }
  }
  // the caller's stream continues after the last copy; the call itself returns when the result is on the host
  cudaEvent_t e1 = hp->get();
  SRK_CHECK_CUDA(cudaEventRecord(e1, hp->s_out));
  SRK_CHECK_CUDA(cudaStreamWaitEvent(s_c, e1, 0));
  SRK_CHECK_CUDA(cudaStreamSynchronize(hp->s_out));
  return 0;
}

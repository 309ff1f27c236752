/* srk.h -- C ABI of the B200-native super-resolution conv hot path (libsrk.so).
 *
 * Drop-in boundary for the convolutional hot path of imironhead/ml_super_resolution.  The
 * reference has no native ABI: its seam is the TensorFlow-1.8 op layer that the model builders
 * call (tf.layers.conv2d, tf.image.resize_bicubic, tf.losses.mean_squared_error,
 * tf.train.AdamOptimizer, ...).  Each entry point below replaces one of those op call sites; the
 * comment on each cites the reference file:line (relative to the reference root) it stands in for.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  All tensor pointers are DEVICE pointers unless the
 *     name ends in `_host`.  The caller owns every buffer; the library owns only its handle.
 *   - Every call is asynchronous on the supplied cudaStream_t (passed as void*), performs no
 *     host<->device synchronisation and no allocation, and is CUDA-graph capturable.
 *   - Return value: 0 = OK, negative = error; text via srk_last_error() (thread-local).
 *   - There is no CPU fallback: every compute entry point fails if no sm_100 device is present.
 *
 * Activation layout ("FPA" = flat padded activation), the HBM format between conv layers:
 *   bf16 [rows][C] (C = 64 or 32), one row per pixel.  n_img images of H x W are stored with
 *   pitch Wp = W+1 and image stride S = (H+1)*Wp; pixel (n,y,x) lives at row n*S + (y+1)*Wp + x.
 *   Row y = -1 of every image and column x = W of every row are ZERO, so the 3x3 tap (dy,dx) of
 *   pixel row p is simply row p + dy*Wp + dx: a convolution is 9 shifted GEMMs over one flat
 *   matrix.  srk_fpa_rows() gives the allocation size in rows (multiple of 128).
 */
#ifndef SRK_H_
#define SRK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct srk_ctx* srk_handle_t;
typedef void* srk_stream_t; /* cudaStream_t */

enum { SRK_ACT_NONE = 0, SRK_ACT_RELU = 1, SRK_ACT_TANH = 2, SRK_ACT_LEAKY_RELU = 3, SRK_ACT_SIGMOID = 4 };
enum { SRK_DT_BF16 = 0, SRK_DT_TF32 = 1, SRK_DT_F32 = 1 }; /* storage: bf16, or fp32 (read by the tensor core as tf32) */
enum { SRK_PAD_SAME = 0, SRK_PAD_VALID = 1 };
enum { SRK_PACK_FWD = 0, SRK_PACK_DGRAD = 1, SRK_PACK_ROT180T_F32 = 2, SRK_PACK_FIRST = 3, SRK_PACK_FIRST_ROT180T = 4 };

/* One entry per FPA image when a large frame is processed as column/row panels (tiled inference,
 * SURVEY 8e): the panel's top-left corner in the frame and the rectangle (panel-local, half-open)
 * of output pixels this panel owns. */
typedef struct {
  int32_t frame, y0, x0;
  int32_t own_y0, own_y1, own_x0, own_x1;
  int32_t reserved;
} srk_panel;

/* ---- handle / errors ------------------------------------------------------------------------- */
int srk_version(void);
const char* srk_last_error(void);
int srk_create(int device, srk_handle_t* out);
int srk_destroy(srk_handle_t h);
int srk_num_sms(srk_handle_t h);

/* Kernel form of the plain 3x3 64->64 srk_conv_tc layers on this handle (sticky until changed):
 *   SRK_CONV_FORM_AUTO  (default) column strips when at least 80 % of a 126-lane tile carries pixels (rows of 100+ pixels cut
 *                       into strips, or several narrow images side by side: 3 x 42 lanes for 41-pixel patches), else the flat stream
 *   SRK_CONV_FORM_FLAT / SRK_CONV_FORM_STRIP  force one form.
 * A model that cuts a frame into panels sets the form from the FRAME width before it runs its layers, so that tiled and
 * un-tiled runs of one frame use the same arithmetic and stay bit-identical (the reference has no tiling: its
 * experiment_resolve.py:60-69 feeds whole frames to tf.layers.conv2d). */
enum { SRK_CONV_FORM_AUTO = 0, SRK_CONV_FORM_FLAT = 1, SRK_CONV_FORM_STRIP = 2 };
int srk_set_conv_form(srk_handle_t h, int form);

/* Strided host <-> device copy of a rectangle (rows x width_bytes) on a stream: a rank of tile-sharded inference moves only its
 * region of the frame, as one DMA transfer (cudaMemcpy2DAsync; page-locked host memory keeps it asynchronous). */
int srk_memcpy2d_async(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes, size_t rows, int to_device,
                       srk_stream_t stream);

/* rows (multiple of 128) an FPA buffer for n_img images of H x W must hold */
int64_t srk_fpa_rows(int n_img, int H, int W);

/* ---- weight layout prep ----------------------------------------------------------------------
 * TF HWIO fp32 kernel [k,k,cin,cout] (tf.layers.conv2d `kernel:0`, vdsr/vdsr/model_vdsr.py:62-70)
 * -> bf16 [k*k][np][cinp] K-major GEMM-B blocks (np, cinp = cout, cin rounded up; zero padded).
 * mode SRK_PACK_DGRAD packs the transposed, 180-degree rotated kernel used by the data-gradient
 * convolution (dX = dY * rot180(W)^T), i.e. block[tap'][ci][co] = w[k-1-u][k-1-v][ci][co]. */
int srk_pack_conv_weights(srk_handle_t h, const float* w_hwio, int k, int cin, int cout, int mode,
                          int np, int cinp, void* packed_bf16, srk_stream_t stream);

/* Batched form of srk_pack_conv_weights: one launch re-packs every layer of a model after an optimiser
 * step.  jobs_device is a device array; src_offset is in floats from `arena`, dst_offset in bytes from
 * `out_base`, elem_begin the exclusive prefix sum of output elements (jobs sorted by it).  Mode
 * SRK_PACK_ROT180T_F32 writes the fp32 [k][k][cout][cin] rotated+transposed kernel (the last layer's
 * data-gradient kernel in the form srk_conv_first consumes). */
typedef struct {
  int64_t src_offset, dst_offset, elem_begin;
  int32_t k, cin, cout, mode, np, cinp;
} srk_pack_job;
int srk_pack_conv_weights_batched(srk_handle_t h, const float* arena, const srk_pack_job* jobs_device, int n_jobs,
                                  int64_t total_elems, void* out_base, srk_stream_t stream);

/* ---- conv layers ------------------------------------------------------------------------------
 * First layer of every model: small-Cin conv from an fp32 NHWC frame straight into an FPA.
 *   replaces tf.layers.conv2d(sd_images, 64, 3, 'same', relu)  vdsr/vdsr/model_vdsr.py:62 (i=0)
 *            tf.layers.conv2d(lr_source, 64, 5, 'same', tanh)  espcn/espcn/model_espcn.py:30-38
 *            convolution2d(lo_images, 64, 9, 'VALID', relu)    srcnn/srcnn.py:100-108
 *            tf.layers.conv2d(sd_images, 64, 3, 'same', relu)  enet/enet/model_enet.py:63-70
 * x: fp32 [n_frames, FH, FW, cin] ; w_hwio fp32 [k,k,cin,64] ; bias fp32 [64].
 * y: FPA bf16 64ch of (n_img, H, W).  Without panels n_img == n_frames and (H,W) is the output
 * size (FH-k+1 for VALID).  With panels each FPA image n reads the frame window at panels[n].
 * relu_mask_src (optional FPA, same geometry): multiplies the result by (mask_src > 0); this is
 * how the last layer's data gradient is produced (conv of dY[...,cout<=4] with the packed
 * rot180 kernel, masked by ReLU' of the saved activation). */
int srk_conv_first(srk_handle_t h, const float* x, int n_frames, int FH, int FW, int cin,
                   const float* w_hwio, const float* bias, int k, int pad_mode, int act,
                   const srk_panel* panels, int n_img, int H, int W, void* y_fpa,
                   const void* relu_mask_src, srk_stream_t stream);

/* Tensor-core form of srk_conv_first (same call sites): the im2col rows of each 128-pixel tile are gathered
 * from the fp32 frame into shared memory (bf16) and multiplied by the packed kernel with tcgen05.mma.
 * w_packed: srk_pack_conv_weights(..., mode = SRK_PACK_FIRST, ...) -> bf16 [ceil(k*k*cin/64)][64][64]; or mode
 * SRK_PACK_FIRST_ROT180T of a [k,k,64,cout] kernel, which makes this call the LAST layer's data gradient
 * (x = dY fp32 [.., cout], mask_src = saved activation, mask_kind = SRK_ACT_RELU). */
int srk_conv_first_tc(srk_handle_t h, const float* x, int n_frames, int FH, int FW, int cin, const void* w_packed,
                      const float* bias, int k, int pad_mode, int act, const srk_panel* panels, int n_img, int H, int W,
                      void* y_fpa, const void* mask_src, int mask_kind, srk_stream_t stream);

/* Tensor-core conv between FPA buffers (tcgen05 implicit GEMM, 9 shifted GEMMs, fp32 TMEM accum):
 *   y = act(conv_kxk(x) + bias) [* act'(mask_src)] [then relu(y + addend)]
 *   replaces tf.layers.conv2d(t, 64, 3, 'same', relu)   vdsr/vdsr/model_vdsr.py:62-70 (i>=1)
 *            tf.layers.conv2d(t, 32, 3, 'same', tanh)   espcn/espcn/model_espcn.py:40-48
 *            residual_block convs (3x3 relu, 1x1)       enet/enet/model_enet.py:13-31
 *   and, with SRK_PACK_DGRAD weights + mask_src, the data gradient of the same layers.
 * x: FPA bf16 [rows, cin_p] (cin_p 64 or 32); y: FPA bf16 [rows, cout_p] (64 or 32).
 * k is 3 or 1.  mask_kind: SRK_ACT_RELU -> (mask_src>0), SRK_ACT_TANH -> (1-mask_src^2).
 * addend (optional FPA, cout_p channels): relu_after_add = 0 -> y + addend; 1 -> relu(y + addend) (residual block
 * output, enet/enet/model_enet.py:29-31); 2 -> (conv + addend) * act'(mask_src): the data gradient through that
 * junction, where the skip path's gradient joins before the producing layer's activation mask.
 * The plain 3x3 64->64 layer (no mask, no addend, act none | relu) has two kernel forms: the flat-stream one (any geometry)
 * and the column-strip one (csrc/conv_strip.cu: no lane shift in the epilogue, ~25 % faster on wide frames); they add the
 * nine taps in different fp32 orders, so their outputs differ in the last bf16 bit.  srk_set_conv_form chooses. */
int srk_conv_tc(srk_handle_t h, const void* x_fpa, int cin_p, const void* w_packed, const float* bias,
                int k, int cout_p, int act, int n_img, int H, int W, void* y_fpa,
                const void* mask_src, int mask_kind, const void* addend_fpa, int relu_after_add,
                srk_stream_t stream);

/* A CHAIN of plain 3x3 64->64 layers of one geometry in ONE persistent launch (column-strip form, csrc/conv_strip.cu): layer l
 * reads x_fpa[l] and writes y_fpa[l] (normally x_fpa[l+1] == y_fpa[l]); between layers the CTAs meet at a grid barrier instead
 * of paying a launch each -- at 64 patches of 41x41 a launch lives 12.7 us of which ~5 us is work.
 *   replaces the 18 middle tf.layers.conv2d(.., 64, 3, 'same', relu) of vdsr/vdsr/model_vdsr.py:64-83 (forward chain) and
 *   their data gradients (tf.gradients of the same lines: SRK_PACK_DGRAD weights, mask_src[l] = the saved activation whose
 *   ReLU' masks layer l's output, bias null, act none).
 * Arrays of n_layers (<= 20) HOST pointers to device buffers; bias / mask_src may be null (or hold null entries); act[l] is
 * SRK_ACT_NONE | SRK_ACT_RELU.  sync_word: 16 bytes of device memory for the grid barriers (the call zeroes them on the stream);
 * may be null when n_layers == 1. */
int srk_conv_tc_chain(srk_handle_t h, int n_layers, const void* const* x_fpa, const void* const* w_packed, const float* const* bias,
                      const int* act, void* const* y_fpa, const void* const* mask_src, int n_img, int H, int W, void* sync_word,
                      srk_stream_t stream);

/* Last layer: tensor-core conv from an FPA into an fp32 NHWC frame, fused with the global
 * residual add and (ESPCN) the depth_to_space pixel shuffle:
 *   out[n, Y*r+dy, X*r+dx, c] = act(conv(x)[n,Y,X,(dy*r+dx)*C + c] + bias) + addend[same index]
 *   replaces tf.layers.conv2d(t, 3, 3, 'same') + `sd_images + tensors` vdsr/vdsr/model_vdsr.py:85-104
 *            f3 conv + host un-pack  espcn/espcn/model_espcn.py:54-62, experiment_test.py:173-177
 *            last conv + `bq_images + tensors`  enet/enet/model_enet.py:102-113
 * cout = true output channels (<= 32); r = 1 (no shuffle) or the ESPCN scaling factor.
 * out / addend: fp32 [n_frames, FH*r, FW*r, cout/(r*r)].  Without panels FH=H, FW=W. */
int srk_conv_tc_last(srk_handle_t h, const void* x_fpa, int cin_p, const void* w_packed, const float* bias,
                     int k, int cout, int cout_p, int act, int n_img, int H, int W,
                     const srk_panel* panels, int n_frames, int FH, int FW, int shuffle_r,
                     const float* addend, float* out, srk_stream_t stream);

/* ---- model-level fast path: the whole ESPCN test graph in ONE persistent kernel (csrc/espcn_fused.cu) -------------------
 *   replaces, for one `session.run(sr_results)` of the constant-weight test graph,
 *     tf.nn.conv2d(.., f1) + bias_add + tanh ; tf.nn.conv2d(.., f2) + bias_add + tanh ; tf.nn.conv2d(.., f3) + bias_add
 *                                                                   espcn/espcn/model_espcn.py:117-134
 *   and (shuffle != 0) the host un-pack np.split / reshape / concatenate of espcn/espcn/experiment_test.py:173-177,
 *   and (out_kind == SRK_OUT_U8) the uint8 conversion of the written image, :179-184 (tf.saturate_cast semantics).
 * The two intermediate activations never leave the SM (tensor memory); HBM sees the LR frame once and the result once.
 * net: device pointers to the packed kernels the layer-by-layer path uses as well --
 *   w1_packed srk_pack_conv_weights(f1 [5,5,C,64],  SRK_PACK_FIRST)              bf16 [ceil(25C/64)][64][64]
 *   w2_packed srk_pack_conv_weights(f2 [3,3,64,32], SRK_PACK_FWD, np 32, cinp 64) bf16 [9][32][64]
 *   w3_packed srk_pack_conv_weights(f3 [3,3,32,C*r^2], SRK_PACK_FWD, np = C*r^2 rounded up to 16, cinp 32)
 *   b1 [64], b2 [32], b3 [C*r^2] fp32.
 * lr: fp32 NHWC [n,H,W,C] (C = 1 or 3).  Rows [y_begin, y_end) of every frame are produced (a rank's row band in tiled
 * multi-GPU inference; the 4 LR rows above / below the band are read from `lr`, which therefore addresses whole frames);
 * out addresses whole frames too: fp32 or uint8 [n, H*r, W*r, C] (shuffle) or [n, H, W, C*r^2] (packed, shuffle = 0). */
enum { SRK_OUT_F32 = 0, SRK_OUT_U8 = 1 };
typedef struct {
  const void* w1_packed;
  const void* w2_packed;
  const void* w3_packed;
  const float* b1;
  const float* b2;
  const float* b3;
  int32_t channels, scaling_factor;
} srk_espcn_net;
int srk_espcn_forward(srk_handle_t h, const srk_espcn_net* net, const float* lr, int n, int H, int W, int y_begin,
                      int y_end, int shuffle, int out_kind, void* out, srk_stream_t stream);

/* The same, HOST array in, HOST array out -- `session.run(model['sr_results'], feed_dict={lr_sources: frames})` of
 * espcn/espcn/experiment_test.py:164-169 plus the un-pack / uint8 conversion of :173-184 -- with the copies hidden: every frame
 * is cut into bands of `band_rows` LR rows, and the host->device copy of band k+1, the fused kernel on band k and the device->host
 * copy of band k-1 run on three streams; the call returns when the last band is in `out_host`.  lr_host / out_host should be
 * page-locked (pageable memory works, without overlap).  lr_is_u8 != 0: `lr_host` is the decoded uint8 image; it crosses PCIe at
 * one byte per sample and `image / 127.5 - 1.0` (:159) runs on the device (srk_u8_to_pm1_f64) into lr_dev.  lr_dev (fp32
 * [n,H,W,C]), lr_u8_dev (uint8, only for lr_is_u8) and out_dev (the result's shape and type) are device scratch owned by the
 * caller.  `stream`: the compute stream (weights must be ready on it). */
int srk_espcn_forward_host(srk_handle_t h, const srk_espcn_net* net, const void* lr_host, int lr_is_u8, int n, int H, int W, int shuffle,
                           int out_kind, void* out_host, float* lr_dev, uint8_t* lr_u8_dev, void* out_dev, int band_rows,
                           srk_stream_t stream);

/* ---- generic tensor-core GEMM and the layers built on it (csrc/gemm_tc.cu, csrc/f2_ops.cu) ------------------------------
 * D[b][M][N] = act(A[b][M][K] x B[b][N][K]^T + bias[N]): both operands row-major with K contiguous (leading dimensions lda / ldb
 * and batch strides in ELEMENTS; rows must start on 16-byte boundaries), in_dtype SRK_DT_BF16 or SRK_DT_TF32 (fp32 storage,
 * tcgen05 kind::tf32), fp32 accumulation, D bf16 or fp32.  act: SRK_ACT_* (leaky = slope of SRK_ACT_LEAKY_RELU; tanh is tanhf).
 * With srk_im2col / srk_col2im / srk_transpose this is
 *   tf.layers.conv2d(.., 3, strides 1|2, 'same', leaky_relu), tf.layers.dense       enet/enet/model_enet.py:129-159 (discriminator)
 *   tf.nn.conv2d + bias_add + relu                                                  enet/enet/model_vgg.py:11-24 (VGG-19)
 *   tf.matmul(x, x, transpose_a=True) per 16x16 patch                               enet/enet/model_enet.py:248-249 (Gram matrices)
 * and their gradients (what `minimize` adds, :325-335), and -- with SRK_DT_TF32 -- the fp32-storage form of every convolution of
 * the four models (north_star's tf32 variant). */
int srk_gemm_tc(srk_handle_t h, const void* A, const void* B, void* D, int M, int N, int K, int batch, long long lda,
                long long ldb, long long ldd, long long stride_a, long long stride_b, long long stride_d, const float* bias,
                int act, float leaky, int in_dtype, int out_dtype, srk_stream_t stream);
/* TF output size of a k x k / stride window: 'SAME' ceil(in / stride), 'VALID' (in - k) / stride + 1. */
int srk_conv_out_size(int in, int k, int stride, int pad_mode);
/* col[m][(u*k+v)*C + c] = x[n, oy*s+u-pt, ox*s+v-pl, c] (zero outside; TF 'SAME' pads pad_total/2 before and the remainder AFTER:
 * asymmetric for stride 2), rows Kp >= k*k*C long (zero beyond); transposed != 0 writes colT[kk][m] with row length Mp instead
 * (the K-major operand of the weight gradient).  x, col: bf16 or fp32 NHWC. */
int srk_im2col(srk_handle_t h, const void* x, int dtype, int n, int H, int W, int C, int k, int stride, int pad_mode, void* col,
               int Kp, long long Mp, int transposed, srk_stream_t stream);
/* dx[n,y,x,c] = sum of the dcol entries that srk_im2col filled from it (gather form, deterministic): the data gradient. */
int srk_col2im(srk_handle_t h, const void* dcol, int dtype, int n, int H, int W, int C, int k, int stride, int pad_mode, int Kp,
               void* dx, srk_stream_t stream);
/* y[b][c][r] = x[b][r][c]. */
int srk_transpose(srk_handle_t h, const void* x, int dtype, int batch, int R, int Ccols, long long ldx, long long stride_x, void* y,
                  long long ldy, long long stride_y, srk_stream_t stream);
/* tf.nn.max_pool(ksize 2, strides 2, 'SAME') and its gradient (first maximum of a window receives it).  model_vgg.py:27-36 */
int srk_maxpool2x2(srk_handle_t h, const void* x, int dtype, int n, int H, int W, int C, void* y, srk_stream_t stream);
int srk_maxpool2x2_bwd(srk_handle_t h, const void* x, const void* dy, int dtype, int n, int H, int W, int C, void* dx, srk_stream_t stream);
/* dx = dy * act'(y), y = the saved activation OUTPUT (relu, leaky-relu, sigmoid, tanh). */
int srk_act_bwd(srk_handle_t h, const void* dy, const void* y, int dtype, long long n, int act, float leaky, void* dx, srk_stream_t stream);
/* VGG input: y[p][c] = x[p][2-c] * scale + shift - (103.939, 116.779, 123.68)[c]   (tf.reverse + mean pixel, model_vgg.py:77-81; the
 * caller's scale / shift = 127.5 / 127.5 map [-1,1] images to [0,255], model_enet.py:291-292); gradient w.r.t. x (optionally added). */
int srk_vgg_preprocess(srk_handle_t h, const float* x, long long pixels, float scale, float shift, int out_dtype, void* y, srk_stream_t stream);
int srk_vgg_preprocess_bwd(srk_handle_t h, const void* dy, int dtype, long long pixels, float scale, float* dx, int accumulate, srk_stream_t stream);
/* normalize(): y = x / (reduce_mean(x, axis=-1, keepdims) + 1e-6), fp32 out, and its gradient.  model_enet.py:34-41 */
int srk_normalize_channels(srk_handle_t h, const void* x, int dtype, long long pixels, int C, float* y, srk_stream_t stream);
int srk_normalize_channels_bwd(srk_handle_t h, const void* x, const float* dy, int dtype, long long pixels, int C, void* dx, srk_stream_t stream);
/* tf.extract_image_patches(16x16, stride 16, VALID) + reshape [-1, 256, C] (model_enet.py:226-243) as bf16: xp [q][256][C] and its
 * transpose xt [q][C][256] (the K-major operand of the Gram GEMM); gradient from d(xt) back to the NHWC tensor. */
int srk_extract_patches16(srk_handle_t h, const float* x, int n, int H, int W, int C, void* xp_bf16, void* xt_bf16, srk_stream_t stream);
int srk_extract_patches16_bwd(srk_handle_t h, const float* dxt, int n, int H, int W, int C, float* dx, srk_stream_t stream);
/* tf.losses.log_loss(labels = label * ones, predictions = p, MEAN): *loss_accum += scale * loss; dp = scale * d loss / d p.  :164-181 */
int srk_log_loss(srk_handle_t h, const float* p, long long n, float label, float scale, float* loss_accum, float* dp, srk_stream_t stream);
/* y = alpha * x + beta * y ;  y = convert(x * scale) ;  out[c] (=|+=) sum_m x[m][c] (bias gradients). */
int srk_axpby(srk_handle_t h, const float* x, long long n, float alpha, float beta, float* y, srk_stream_t stream);
int srk_convert(srk_handle_t h, const void* x, int src_dtype, long long n, float scale, void* y, int dst_dtype, srk_stream_t stream);
int srk_colsum(srk_handle_t h, const void* x, int dtype, long long M, int C, float* out, int accumulate, srk_stream_t stream);

/* 3xTF32 operand split for fp32-level accuracy on the tf32 tensor cores: row r of y = [hi | lo | hi] (side 0: the A operand) or
 * [hi | hi | lo] (side 1: the B operand), hi = round-to-nearest tf32(x), lo = tf32(x - hi), segments Kp (multiple of 4) long.
 * srk_gemm_tc over the 3*Kp-long rows then yields hi.hi + lo.hi + hi.lo. */
int srk_tf32_split(srk_handle_t h, const float* x, int batch, long long R, int K, long long ldx, long long stride_x, float* y, int Kp,
                   int side, srk_stream_t stream);

/* ---- data-parallel exchange step: NCCL sum all-reduce of the flat fp32 gradient arena over NVLink ----------
 * The reference trains on one GPU (vdsr/README.md:14); these calls extend the `minimize` of vdsr/vdsr/model_vdsr.py:146-148
 * (gradients are summed across ranks before the optimiser applies them; every rank then runs the identical update).
 * NCCL is loaded at run time (libnccl.so.2 of the host process).  srk_comm_unique_id: rank 0 creates the 128-byte
 * ncclUniqueId, the host distributes it (any side channel), every rank calls srk_comm_init (collective, blocking).
 * srk_allreduce_grads is asynchronous on `stream` and may be captured into a CUDA graph. */
int srk_comm_unique_id(void* id128);
int srk_comm_init(srk_handle_t h, const void* id128, int rank, int world);
int srk_allreduce_grads(srk_handle_t h, float* flat, size_t n, srk_stream_t stream);
int srk_comm_destroy(srk_handle_t h);

/* The same exchange step FUSED with the optimiser over NVLink peer memory (csrc/peer_reduce.cu): one kernel per rank stages its
 * gradient, meets its peers at per-slice flags in each other's memory, sums all ranks' gradients in rank order through the peer
 * mappings and applies Adam -- no NCCL call, no separate pass over the summed gradient, replicas bit-identical.
 *   srk_peer_alloc   allocates this rank's exchange region for `grad_elems` floats and returns its cudaIpcMemHandle (64 bytes)
 *   srk_peer_open    after the ranks have all-gathered those handles (world x 64 bytes, rank-major): maps the peers' regions
 *   srk_allreduce_adam_step_dev   g <- sum_r g_r, then the update of srk_adam_step_dev with it (same arguments); every rank
 *                    must call it once per step with the same n; capturable in a CUDA graph (epochs live on the device)
 *   srk_peer_close   unmaps and frees
 * One process per GPU on one node, at most 8 ranks (extends `minimize`, vdsr/vdsr/model_vdsr.py:146-148). */
int srk_peer_alloc(srk_handle_t h, size_t grad_elems, void* ipc_handle_out64);
int srk_peer_open(srk_handle_t h, int rank, int world, const void* handles);
int srk_allreduce_adam_step_dev(srk_handle_t h, float* w, float* g, float* m, float* v, size_t n, const float* lr_t_device, float beta1,
                                float beta2, float eps, float weight_decay, const float* decay_mask, srk_stream_t stream);
int srk_peer_close(srk_handle_t h);

/* Weight gradient of a 3x3 64->64 layer on tensor cores: dW[u,v,ci,co] = sum_p x[p+(u-1)*Wp+(v-1)][ci]
 * * dy[p][co], dbias[co] = sum_p dy[p][co].  Split over CTAs along the pixel axis; every CTA stores its
 * partial [9*64*64+64] block into `workspace` and a second kernel sums the partials in a fixed order
 * (deterministic).  accumulate != 0 adds to dw/dbias instead of overwriting.
 *   replaces the wgrad/bgrad ops autodiff adds for  vdsr/vdsr/model_vdsr.py:146-148 (minimize). */
size_t srk_conv_wgrad_tc_workspace_bytes(srk_handle_t h, int n_img, int H, int W);
int srk_conv_wgrad_tc(srk_handle_t h, const void* x_fpa, const void* dy_fpa, int n_img, int H, int W,
                      float* dw_hwio, float* dbias, int accumulate, void* workspace, size_t workspace_bytes,
                      srk_stream_t stream);
/* Deferred form: call srk_conv_wgrad_tc with dw_hwio = NULL for each of n_layers layers, each with its own
 * workspace slice `workspace_base + l * layer_stride_bytes`, then fold all of them with ONE launch.
 * dsts_device: device array of n_layers destinations.  ci_n / co_n < 64 keep only the leading input / output
 * channels and write the dense [9][ci_n][co_n] kernel gradient: this is how the first layer (x = the input frame
 * zero-padded to 64 channels by srk_nhwc_to_fpa_pad, ci_n = C) and the last layer (dy zero-padded, co_n = C)
 * share the tensor-core kernel. */
typedef struct {
  float* dw;
  float* db;
  int32_t ci_n, co_n;
} srk_wgrad_dst;
/* All layers of one geometry in ONE weight-gradient launch + ONE reduce launch: x_fpas / dy_fpas are HOST arrays of n_layers
 * device pointers (layer l: dW_l = X_l^T dY_l), workspace slices `layer_stride_bytes` apart (size them with
 * srk_conv_wgrad_tc_workspace_bytes), destinations as for srk_wgrad_reduce_many.  Needs every layer's dY kept until the
 * call (the training step stores them instead of ping-ponging two buffers). */
int srk_conv_wgrad_tc_batched(srk_handle_t h, const void* const* x_fpas, const void* const* dy_fpas, int n_layers,
                              int n_img, int H, int W, void* workspace, size_t layer_stride_bytes,
                              const srk_wgrad_dst* dsts_device, int accumulate, srk_stream_t stream);
int srk_wgrad_reduce_many(srk_handle_t h, const void* workspace_base, size_t layer_stride_bytes, int n_layers,
                          int n_img, int H, int W, const srk_wgrad_dst* dsts_device, int accumulate,
                          srk_stream_t stream);

/* Weight gradient of the first layer (x fp32 NHWC cin<=4, dy FPA 64ch) -> dw [k,k,cin,64], db[64]. */
int srk_conv_first_wgrad(srk_handle_t h, const float* x, int n_img, int H, int W, int cin, int k,
                         const void* dy_fpa, float* dw_hwio, float* dbias, srk_stream_t stream);
/* Weight gradient of the last layer (x FPA 64ch, dy fp32 NHWC cout<=4) -> dw [3,3,64,cout], db[cout]. */
int srk_conv_last_wgrad(srk_handle_t h, const void* x_fpa, const float* dy, int n_img, int H, int W,
                        int cout, float* dw_hwio, float* dbias, srk_stream_t stream);

/* ---- bandwidth kernels ------------------------------------------------------------------------ */
/* Standalone depth_to_space: [N,H,W,C*r*r] -> [N,H*r,W*r,C], packed channel (dy*r+dx)*C+c.
 * espcn/espcn/experiment_test.py:173-177 (host np.split/reshape/concatenate). fp32. */
int srk_pixel_shuffle(srk_handle_t h, const float* x, int N, int H, int W, int C, int r, float* y,
                      srk_stream_t stream);
/* Inverse packing, espcn/espcn/dataset.py:140-156. */
int srk_pixel_unshuffle(srk_handle_t h, const float* x, int N, int H, int W, int C, int r, float* y,
                        srk_stream_t stream);
/* TF1-legacy bicubic (A=-0.75, 1024-entry table, no half-pixel centres): srcnn/srcnn.py:89-93. */
int srk_resize_bicubic_tf1(srk_handle_t h, const float* x, int N, int H, int W, int C, int OH, int OW,
                           float* y, srk_stream_t stream);
/* Device-resident input pipeline: the training images stay in HBM as one uint8 pool; a batch is cut
 * out of it by the same steps the reference's python generator performs per sample (vdsr/vdsr/dataset.py:95-116): random
 * crop [y:y+S, x:x+S], optional horizontal flip, img_as_float32 (x/255), then -- after the degrade below -- x*2-1.
 * out01 (optional) receives the [0,1] patch that srk_degrade_gauss_bilinear consumes, out_pm1 (optional) the [-1,1] one. */
typedef struct {
  int64_t offset;        /* byte offset of pixel (0,0,0) in the pool */
  int32_t height, width; /* channels = C of the call */
} srk_pool_image;
typedef struct {
  int32_t image, y, x, flip;
} srk_crop;
int srk_crop_flip_u8(srk_handle_t h, const uint8_t* pool, const srk_pool_image* images_device, const srk_crop* crops_device,
                     int n, int S, int C, float* out01, float* out_pm1, srk_stream_t stream);
/* y = x*a + b with two fp32 roundings (numpy semantics of `sd_image * 2.0 - 1.0`, vdsr/vdsr/dataset.py:113-115); y may be x. */
int srk_affine_f32(srk_handle_t h, const float* x, size_t n, float a, float b, float* y, srk_stream_t stream);

/* EnhanceNet's input pipeline on the device (enet/enet/datasets.py:100-125): 128x128 uint8 crop, `scipy.misc.imresize` = Pillow's
 * separable fixed-point resampler (25 % bilinear with antialiasing, then 400 % bicubic), and x.astype(float32)/127.5 - 1.
 * srk_resample_u8: one or two passes (horizontal first, uint8 intermediate `tmp` [n,H,out_w,C]); kx/ky = int32 coefficients
 * [out][ks] scaled by 2^22, bx/by = int32 (first input index, count) per output, both computed by the host exactly like
 * Resample.c precompute_coeffs / normalize_coeffs_8bpc (ml_super_resolution_b200/enet/datasets.py). */
int srk_crop_u8(srk_handle_t h, const uint8_t* pool, const srk_pool_image* images_device, const srk_crop* crops_device,
                int n, int S, int C, uint8_t* out, srk_stream_t stream);
int srk_resample_u8(srk_handle_t h, const uint8_t* x, int n, int H, int W, int C, int out_h, int out_w,
                    const int32_t* kx, const int32_t* bx, int ksx, const int32_t* ky, const int32_t* by, int ksy,
                    uint8_t* tmp, uint8_t* y, srk_stream_t stream);
int srk_u8_to_pm1(srk_handle_t h, const uint8_t* x, size_t n, float* y, srk_stream_t stream);
/* uint8 -> float32 as numpy computes `image / 127.5 - 1.0` in float64 before the feed casts it to float32
 * (espcn/espcn/experiment_test.py:159,164; vdsr/vdsr/experiment_resolve.py likewise): the raw image crosses PCIe at one byte per
 * sample and is normalised on the device, bit-identical to the host arithmetic (srk_u8_to_pm1 is the float32 form of
 * enet/enet/datasets.py:114-116: one ulp apart for half of the 256 values). */
int srk_u8_to_pm1_f64(srk_handle_t h, const uint8_t* x, size_t n, float* y, srk_stream_t stream);

/* VDSR degrade pre-pass: gaussian(sigma=0.5(s-1), replicate) -> bilinear down to int(H/s) x int(W/s)
 * -> bilinear up (half-pixel, edge clamp), per-sample scale: vdsr/vdsr/dataset.py:13-38. */
int srk_degrade_gauss_bilinear(srk_handle_t h, const float* hd, int N, int H, int W, int C,
                               const float* scale_per_sample, float* sd, srk_stream_t stream);
/* Nearest-neighbour x2 upsample of an FPA (enet/enet/model_enet.py:78-80) and its backward (2x2 sum). */
int srk_fpa_upsample2(srk_handle_t h, const void* x_fpa, int n_img, int H, int W, void* y_fpa,
                      srk_stream_t stream);
int srk_fpa_upsample2_bwd(srk_handle_t h, const void* dy_fpa, int n_img, int H, int W, void* dx_fpa,
                          srk_stream_t stream);
/* MSE (tf.losses.mean_squared_error MEAN, vdsr/vdsr/model_vdsr.py:120-123): *loss_accum += sum((sr-hd)^2)/numel_total
 * and dsr = 2(sr-hd)/numel_total.  numel_total lets a data-parallel rank scale by the GLOBAL element count. */
int srk_mse_fwd_bwd(srk_handle_t h, const float* sr, const float* hd, size_t numel, double numel_total,
                    float* loss_accum, float* dsr, srk_stream_t stream);
/* SRCNN loss (srcnn/srcnn.py:142-144): mean over rows of ||reshape(sr-hd,[rows,cols])||_2, and its gradient.
 * sr_act = SRK_ACT_TANH: sr is the tanh output of the reconstruction layer (srcnn/srcnn.py:128) and dsr receives the
 * gradient w.r.t. its pre-activation, dloss/dsr * (1 - sr^2); SRK_ACT_NONE: plain dloss/dsr. */
int srk_l2norm_rows_mean_fwd_bwd(srk_handle_t h, const float* sr, const float* hd, int rows, int cols,
                                 float* loss_accum, float* dsr, int sr_act, srk_stream_t stream);
/* tf.train.AdamOptimizer step over a flat fp32 arena (vdsr/vdsr/model_vdsr.py:145-148; A.8 epsilon-hat
 * form); t = 1-based step.  decay_mask (optional, n floats): g += weight_decay * decay_mask[i] * w[i]
 * (the l2_regularizer term of model_vdsr.py:34,125: kernels 1, biases 0). */
int srk_adam_step(srk_handle_t h, float* w, const float* g, float* m, float* v, size_t n, float lr,
                  float beta1, float beta2, float eps, int64_t t, float weight_decay,
                  const float* decay_mask, srk_stream_t stream);
/* Same step with the bias-corrected rate lr_t = lr*sqrt(1-b2^t)/(1-b1^t) read from DEVICE memory, so a
 * captured CUDA graph of the whole training step can be replayed while lr and t advance (the reference
 * feeds the learning rate every step: vdsr/vdsr/experiment_train.py:130,137). */
int srk_adam_step_dev(srk_handle_t h, float* w, const float* g, float* m, float* v, size_t n,
                      const float* lr_t_device, float beta1, float beta2, float eps, float weight_decay,
                      const float* decay_mask, srk_stream_t stream);
/* *out_accum += scale * sum_i mask[i]*w[i]^2 : the l2_regularizer part of the VDSR loss
 * (vdsr/vdsr/model_vdsr.py:34,125: scale = 0.5 * 1e-4, mask = 1 on kernels, 0 on biases). */
int srk_sumsq_masked(srk_handle_t h, const float* w, const float* mask, size_t n, float scale, float* out_accum,
                     srk_stream_t stream);
/* Momentum(0.9) with gradient clip +-cap/lr (vdsr/vdsr/model_vdsr.py:158-184). */
int srk_momentum_clip_step(srk_handle_t h, float* w, const float* g, float* accum, size_t n, float lr,
                           float momentum, float gradient_cap, float weight_decay, const float* decay_mask,
                           srk_stream_t stream);
/* Tiled full-frame inference without recomputing a column halo: the FPA images are panels of one frame
 * (srk_panel, ordered left to right within a band) that overlap by a few columns; after every layer each column a panel
 * does not own ([0, own_x0) and [own_x1, W)) is refreshed in place from the neighbouring panel that owns it, so the next
 * 3x3 layer sees exact neighbours at the panel seam.  max_cols >= the largest number of non-owned columns of any panel. */
int srk_fpa_halo_exchange(srk_handle_t h, void* x_fpa, int C, const srk_panel* panels, int n_img, int H, int W,
                          int max_cols, srk_stream_t stream);
/* ---- evaluation post-pass ---------------------------------------------------------------------
 * tf.image.psnr(a, b, max_val) per image over H,W,C: 20 log10(max_val) - 10 log10(mean((a-b)^2))
 *   vdsr/vdsr/experiment_evaluate.py:57-58 (max_val 2.0), espcn/espcn/experiment_test.py:52 (1.0).
 * workspace: n_img doubles of device scratch; out: n_img floats. */
int srk_psnr(srk_handle_t h, const float* a, const float* b, int n_img, int64_t numel_per_image, float max_val,
             double* workspace, float* out, srk_stream_t stream);
/* tf.image.ssim(a, b, max_val) per image: 11x11 gaussian window (sigma 1.5), VALID, k1 = 0.01, k2 = 0.03, mean over window
 * positions and channels (TF 1.8 _ssim_helper: luminance * contrast-structure).  a, b fp32 NHWC [n_img,H,W,C], H,W >= 11.
 *   vdsr/vdsr/experiment_evaluate.py:59-60, espcn/espcn/experiment_test.py:53. */
int srk_ssim(srk_handle_t h, const float* a, const float* b, int n_img, int H, int W, int C, float max_val,
             double* workspace, float* out, srk_stream_t stream);
/* Y of tf.image.rgb_to_yuv applied to clip(x*scale+bias, clip_lo, clip_hi): espcn/espcn/experiment_test.py:32-48 remaps
 * [-1,1] -> [0,1] (scale 0.5, bias 0.5), clips, and scores the Y channel.  x: [n_pixels,3], y: [n_pixels]. */
int srk_rgb_to_y(srk_handle_t h, const float* x, int64_t n_pixels, float scale, float bias, float clip_lo, float clip_hi,
                 float* y, srk_stream_t stream);
/* tf.saturate_cast(x*scale + bias, uint8) (clamp to [0,255], truncate): vdsr/vdsr/experiment_resolve.py:65-67. */
int srk_saturate_cast_u8(srk_handle_t h, const float* x, size_t n, float scale, float bias, uint8_t* y, srk_stream_t stream);

/* 8x8 mosaic of a 64-channel feature map x fp32 [H,W,64] -> uint8 [8H, 8W]: channel ch at grid cell (ch/8, ch%8),
 * saturate_cast(x*127.5+127.5) (vdsr/vdsr/experiment_feature_map_visualize.py:80-110). */
int srk_feature_mosaic_u8(srk_handle_t h, const float* x, int H, int W, uint8_t* y, srk_stream_t stream);

/* FPA (bf16, C ch) <-> fp32 NHWC [n_img,H,W,C] converters (feature-map taps `conv.N:0`, tests). */
int srk_fpa_to_nhwc(srk_handle_t h, const void* x_fpa, int C, int n_img, int H, int W, float* y,
                    srk_stream_t stream);
int srk_nhwc_to_fpa(srk_handle_t h, const float* x, int C, int n_img, int H, int W, void* y_fpa,
                    srk_stream_t stream);
/* Same with the channel count padded with zeros to Cp (e.g. 3 -> 64), so a 3-channel frame or gradient can feed
 * the 64-channel tensor-core kernels. */
int srk_nhwc_to_fpa_pad(srk_handle_t h, const float* x, int C, int Cp, int n_img, int H, int W, void* y_fpa,
                        srk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SRK_H_ */

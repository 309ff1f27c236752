"""EnhanceNet's training losses on the B200 -- drop-in for the second half of enet/enet/model_enet.py:

  build_discriminator   5 x [3x3 s1 leaky ; 3x3 s2 leaky] (32..512 ch), flatten, dense 1024 leaky, dense 1 sigmoid      :118-161
  generator_loss / discriminator_loss   tf.losses.log_loss against ones / zeros                                         :164-181
  perceptual_loss       0.2 MSE(norm(pool2)) + 0.02 MSE(norm(pool5)) of VGG-19 features                                  :184-205
  texture_matching_loss Gram matrices of 16x16 patches of norm(block{1,2,3}_conv1), weights 3e-7 / 1e-6 / 1e-6          :208-256
  model_vgg.build_vgg19_model   16 conv3x3+ReLU, 5 max-pools, BGR - mean pixel input                                     model_vgg.py:11-99

Every convolution / dense layer / Gram product runs on the generic tcgen05 GEMM (csrc/gemm_tc.cu) through `nn.py`; pooling,
normalisation, patch extraction, log-loss and the input transform are the bandwidth kernels of csrc/f2_ops.cu.  Each function
returns the loss AND its gradient with respect to the generated image (what `g_trainer.minimize(g_losses, var_list=g_vars)`
back-propagates into the generator, :336-341); `Discriminator.loss_and_grads` also returns the gradients of its own weights
(`d_trainer`, :343-347).  The two loss networks keep fp32 activations and multiply on the tf32 tensor cores (bf16 activations
put the gradient ~14 % off the fp64 reference through ReLU-mask / arg-max flips, see Vgg19).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .. import _ffi, nn, ops
from .._ffi import check
from ..params import ParamArena

VGG_LAYERS = ["block1_conv1", "block1_conv2", "block1_pool", "block2_conv1", "block2_conv2", "block2_pool", "block3_conv1", "block3_conv2",
              "block3_conv3", "block3_conv4", "block3_pool", "block4_conv1", "block4_conv2", "block4_conv3", "block4_conv4", "block4_pool",
              "block5_conv1", "block5_conv2", "block5_conv3", "block5_conv4", "block5_pool"]
VGG_CHANNELS = {"block1": 64, "block2": 128, "block3": 256, "block4": 512, "block5": 512}
TEXTURE_LAYERS = [("block1_conv1", 3e-7), ("block2_conv1", 1e-6), ("block3_conv1", 1e-6)]


def vgg19_random_weights(seed=0) -> "OrderedDict[str, np.ndarray]":
    """He-initialised stand-in for keras' vgg19 weights (there is no network here), keyed like the reference's npz: the
    arrays `<layer>_W_1:0` [3,3,cin,cout] and `<layer>_b_1:0` (model_vgg.py:44-60)."""
    rng = np.random.default_rng(seed)
    out, cin = OrderedDict(), 3
    for name in VGG_LAYERS:
        if name.endswith("pool"):
            continue
        cout = VGG_CHANNELS[name[:6]]
        out[f"{name}_W_1:0"] = (rng.standard_normal((3, 3, cin, cout)) * np.sqrt(2.0 / (9 * cin))).astype(np.float32)
        out[f"{name}_b_1:0"] = (0.01 * rng.standard_normal(cout)).astype(np.float32)
        cin = cout
    return out


def load_vgg_weights(weights_path: str):
    """model_vgg.load_vgg_weights: the keras vgg19 weights as an `.npz` with `<layer>_W_1:0` / `<layer>_b_1:0` arrays."""
    with np.load(weights_path) as data:
        return OrderedDict((k, np.asarray(data[k], np.float32)) for k in data.files)


class Vgg19:
    """model_vgg.build_vgg19_model with constant weights: features of a [-1,1] RGB image batch and the gradient of a weighted set
    of feature taps back to the image."""

    def __init__(self, weights: dict, device="cuda", upto="block5_pool", dtype=torch.float32, precise=True):
        """dtype: storage of the activations; precise: 3xTF32 products.  The default is fp32 + 3xTF32: with coarser arithmetic
        the ReLU masks and pooling arg-maxima of 16 + 5 layers flip (pre-activations within rounding distance of zero / of each
        other) and the gradient reaching the image drifts from the fp64 reference -- measured relative L2 error of
        d(perceptual loss)/d(image): bf16 14 %, one tf32 product 4.9 %, 3xTF32 5e-5 (tests/test_enet_losses_gpu.py)."""
        self.device, self.dtype = device, dtype
        self.names = VGG_LAYERS[: VGG_LAYERS.index(upto) + 1]
        self.layers = {}
        for i, name in enumerate(self.names):
            if name.endswith("pool"):
                continue
            w = torch.from_numpy(np.ascontiguousarray(weights[f"{name}_W_1:0"], np.float32)).to(device)
            b = torch.from_numpy(np.ascontiguousarray(weights[f"{name}_b_1:0"], np.float32)).to(device)
            self.layers[name] = nn.Conv(w, b, 1, "SAME", "relu", dtype=dtype, in_dtype=torch.float32 if i == 0 else dtype, precise=precise)

    def forward(self, images: torch.Tensor) -> dict:
        """images fp32 NHWC in [-1,1] -> {'input': ..., layer name: activation} (hd_vgg_input = images * 127.5 + 127.5, :291-292)."""
        n, H, W, _ = images.shape
        x = torch.empty((n, H, W, 3), dtype=torch.float32, device=images.device)
        check(_ffi.lib().srk_vgg_preprocess(ops.handle(), ops._ptr(images), n * H * W, 127.5, 127.5, nn.DT_F32, ops._ptr(x), ops._stream()), "srk_vgg_preprocess")
        acts = {"input": x}
        t = x
        for name in self.names:
            t = nn.maxpool2x2(t) if name.endswith("pool") else self.layers[name].forward(t)
            acts[name] = t
        return acts

    def backward(self, acts: dict, tap_grads: dict, d_images: torch.Tensor, accumulate=True) -> torch.Tensor:
        """tap_grads: {layer name: d loss / d activation (bf16, NHWC)}; adds (or writes) d loss / d images into `d_images`."""
        last = max(self.names.index(k) for k in tap_grads)
        g = None
        for i in range(last, -1, -1):
            name = self.names[i]
            if name in tap_grads:
                tg = tap_grads[name]
                if g is None:
                    g = tg
                else:  # two gradients meet at a tap: add in fp32
                    s = nn.convert(g, torch.float32)
                    nn.axpby(nn.convert(tg, torch.float32), s, 1.0, 1.0)
                    g = nn.convert(s, self.dtype)
            x_in = acts[self.names[i - 1]] if i > 0 else acts["input"]
            if name.endswith("pool"):
                g = nn.maxpool2x2_bwd(x_in, g)
            else:
                g = self.layers[name].backward(x_in, acts[name], g, True)
        n, H, W, _ = d_images.shape
        check(_ffi.lib().srk_vgg_preprocess_bwd(ops.handle(), ops._ptr_any(g), nn._dt(g), n * H * W, 127.5, ops._ptr(d_images), int(accumulate), ops._stream()),
              "srk_vgg_preprocess_bwd")
        return d_images


def normalize(x: torch.Tensor) -> torch.Tensor:
    """model_enet.normalize: x / (mean over channels + 1e-6), fp32."""
    C = x.shape[-1]
    y = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    check(_ffi.lib().srk_normalize_channels(ops.handle(), ops._ptr_any(x), nn._dt(x), x.numel() // C, C, ops._ptr(y), ops._stream()), "srk_normalize_channels")
    return y


def normalize_bwd(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    C = x.shape[-1]
    dx = torch.empty_like(x)
    check(_ffi.lib().srk_normalize_channels_bwd(ops.handle(), ops._ptr_any(x), ops._ptr(dy), nn._dt(x), x.numel() // C, C, ops._ptr_any(dx), ops._stream()),
          "srk_normalize_channels_bwd")
    return dx


def _mse(a: torch.Tensor, b: torch.Tensor, weight: float, loss: torch.Tensor, want_grad=True):
    """loss += weight * mean((a-b)^2); returns weight * d/da (fp32) -- srk_mse_fwd_bwd scaled by `weight`."""
    tmp = torch.zeros(1, dtype=torch.float32, device=a.device)
    da = torch.empty_like(a) if want_grad else None
    ops.mse_fwd_bwd(a, b, tmp, da)
    nn.axpby(tmp, loss, weight, 1.0)
    if want_grad and weight != 1.0:
        nn.axpby(da, da, weight, 0.0)
    return da


def perceptual_loss(vgg: Vgg19, sr_acts: dict, hd_acts: dict, loss: torch.Tensor) -> dict:
    """model_enet.perceptual_loss (:184-205): loss += 0.2 MSE(norm pool2) + 0.02 MSE(norm pool5); returns the tap gradients."""
    taps = {}
    for name, wgt in (("block2_pool", 0.2), ("block5_pool", 0.02)):
        s, h = normalize(sr_acts[name]), normalize(hd_acts[name])
        ds = _mse(s, h, wgt, loss)
        taps[name] = normalize_bwd(sr_acts[name], ds)
    return taps


def texture_matching_loss(sr_acts: dict, hd_acts: dict, loss: torch.Tensor) -> dict:
    """model_enet.texture_matching_loss (:208-256): per layer, normalise, cut into 16x16 patches, Gram = X^T X per patch on the
    tensor cores, MSE between the Gram matrices; returns the tap gradients."""
    taps = {}
    for name, wgt in TEXTURE_LAYERS:
        xs, xh = sr_acts[name], hd_acts[name]
        n, H, W, C = xs.shape
        q = n * (H // 16) * (W // 16)
        grams, keep = [], None
        for x, is_sr in ((xs, True), (xh, False)):
            nx = normalize(x)
            xp = torch.empty((q, 256, C), dtype=torch.bfloat16, device=x.device)
            xt = torch.empty((q, C, 256), dtype=torch.bfloat16, device=x.device)
            check(_ffi.lib().srk_extract_patches16(ops.handle(), ops._ptr(nx), n, H, W, C, ops._ptr_any(xp), ops._ptr_any(xt), ops._stream()),
                  "srk_extract_patches16")
            grams.append(nn.gemm(xt, xt, out_dtype=torch.float32))           # [q, C, C] = X^T X
            if is_sr:
                keep = xp
        dG = _mse(grams[0], grams[1], wgt, loss)                               # symmetric: d/dX^T = 2 dG X^T
        # d(X^T)[C, 256] = (2 dG)[C, C] x X^T[C, 256]: as a K-major GEMM the B operand [N = 256][K = C] is the patch matrix itself
        dxt = nn.gemm(nn.convert(dG, torch.bfloat16, 2.0), keep, out_dtype=torch.float32)
        dn = torch.empty((n, H, W, C), dtype=torch.float32, device=xs.device)
        check(_ffi.lib().srk_extract_patches16_bwd(ops.handle(), ops._ptr(dxt), n, H, W, C, ops._ptr(dn), ops._stream()), "srk_extract_patches16_bwd")
        taps[name] = normalize_bwd(xs, dn)
    return taps


class Discriminator:
    """model_enet.build_discriminator (:118-161) with its own parameter arena (TF variable names `d_/conv2d[_i]/kernel:0`, ...,
    `d_/dense[_1]/kernel:0`); input images NHWC fp32 in [-1,1] of the size the dense layer was built for."""

    def __init__(self, image_size: int, params=None, scope_name="d_", device="cuda", seed=0, dtype=torch.float32, precise=True):
        self.scope, self.device, self.size, self.dtype = scope_name, device, image_size, dtype
        if params is None:
            params = discriminator_params(image_size, seed, scope_name)
        self.names = list(params.keys())
        self.arena = ParamArena(OrderedDict((k, np.asarray(v, np.float32)) for k, v in params.items()), device, decay_suffix=None)
        self.arena.enable_training()
        a = self.arena
        self.convs = []
        for i in range(10):
            kname, bname = self.names[2 * i], self.names[2 * i + 1]
            self.convs.append((nn.Conv(a.view(kname), a.view(bname), 1 + (i % 2), "SAME", "leaky_relu", 0.2, dtype=dtype, precise=precise), kname, bname))
        self.dense1 = (nn.Dense(a.view(self.names[20]), a.view(self.names[21]), "leaky_relu", 0.2, dtype=dtype, precise=precise), self.names[20], self.names[21])
        self.dense2 = (nn.Dense(a.view(self.names[22]), a.view(self.names[23]), "sigmoid", dtype=dtype, precise=precise), self.names[22], self.names[23])
        self.step = 0

    def repack(self):
        for layer, _, _ in self.convs + [self.dense1, self.dense2]:
            layer.pack()

    def forward(self, images: torch.Tensor):
        """-> (probabilities fp32 [n, 1], saved activations)."""
        acts = [images if images.dtype == self.dtype else nn.convert(images, self.dtype)]
        for layer, _, _ in self.convs:
            acts.append(layer.forward(acts[-1]))
        flat = acts[-1].reshape(images.shape[0], -1)
        h = self.dense1[0].forward(flat)
        p = self.dense2[0].forward(h, out_dtype=torch.float32)
        return p, (acts, flat, h, p)

    def backward(self, saved, dp: torch.Tensor, grads: dict | None, need_dx: bool):
        """dp = d loss / d p.  Fills `grads` (weight gradients, TF names) when given; returns d loss / d images (fp32) if asked."""
        acts, flat, h, p = saved
        dh = self.dense2[0].backward(h, p, dp, grads, self.dense2[1:])
        dflat = self.dense1[0].backward(flat, h, dh, grads, self.dense1[1:])
        g = dflat.reshape(acts[-1].shape)
        for i in range(9, -1, -1):
            layer, kname, bname = self.convs[i]
            g = layer.backward(acts[i], acts[i + 1], g, need_dx or i > 0, grads, (kname, bname))
        return (g if g.dtype == torch.float32 else nn.convert(g, torch.float32)) if need_dx else None

    def generator_loss(self, sr_images: torch.Tensor, loss: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
        """model_enet.generator_loss (:164-169): loss += scale * log_loss(ones, D(sr)); returns scale * d/d(sr)."""
        p, saved = self.forward(sr_images)
        dp = torch.empty_like(p)
        check(_ffi.lib().srk_log_loss(ops.handle(), ops._ptr(p), p.numel(), 1.0, float(scale), ops._ptr(loss), ops._ptr(dp), ops._stream()), "srk_log_loss")
        return self.backward(saved, dp, None, True)

    def discriminator_loss_and_grads(self, sr_images: torch.Tensor, hd_images: torch.Tensor, loss: torch.Tensor):
        """model_enet.discriminator_loss (:172-181): log_loss(zeros, D(sr)) + log_loss(ones, D(hd)); the weight gradients of both
        terms are accumulated into the arena's gradient buffer (`arena.g`)."""
        self.arena.g.zero_()
        for images, label in ((sr_images, 0.0), (hd_images, 1.0)):
            p, saved = self.forward(images)
            dp = torch.empty_like(p)
            check(_ffi.lib().srk_log_loss(ops.handle(), ops._ptr(p), p.numel(), label, 1.0, ops._ptr(loss), ops._ptr(dp), ops._stream()), "srk_log_loss")
            grads = {}
            self.backward(saved, dp, grads, False)
            for k, g in grads.items():
                nn.axpby(g.reshape(-1), self.arena.view(k, "g").reshape(-1), 1.0, 1.0)

    def adam_step(self, learning_rate=1e-4):
        """d_trainer: tf.train.AdamOptimizer(0.0001).minimize(a_loss, var_list=d_vars) (:343-347)."""
        a = self.arena
        self.step += 1
        ops.adam_step(a.w, a.g, a.m, a.v, learning_rate, self.step)
        self.repack()


def discriminator_params(image_size: int, seed=0, scope_name="d_") -> "OrderedDict[str, np.ndarray]":
    """truncated_normal(stddev 0.02) kernels, zero biases, TF auto-names in creation order (model_enet.py:122-159)."""
    from ..initializers import tf_conv_name, truncated_normal
    rng = np.random.default_rng(seed)
    out, cin, size = OrderedDict(), 3, image_size
    idx = 0
    for i in range(5):
        f = 2 ** (i + 5)
        for stride in (1, 2):
            out[f"{scope_name}/{tf_conv_name(idx)}/kernel:0"] = truncated_normal(rng, (3, 3, cin, f), 0.02)
            out[f"{scope_name}/{tf_conv_name(idx)}/bias:0"] = np.zeros(f, np.float32)
            cin, idx = f, idx + 1
            if stride == 2:
                size = -(-size // 2)
    flat = size * size * cin
    out[f"{scope_name}/dense/kernel:0"] = truncated_normal(rng, (flat, 1024), 0.02)
    out[f"{scope_name}/dense/bias:0"] = np.zeros(1024, np.float32)
    out[f"{scope_name}/dense_1/kernel:0"] = truncated_normal(rng, (1024, 1), 0.02)
    out[f"{scope_name}/dense_1/bias:0"] = np.zeros(1, np.float32)
    return out

"""TEST INFRASTRUCTURE (not product): CPU restatement (torch-CPU fp64 + autograd) of EnhanceNet's training losses --
enet/enet/model_enet.py:118-261 (discriminator, log losses, perceptual loss, texture-matching loss, normalize) and
enet/enet/model_vgg.py:11-99 (VGG-19 feature extractor) -- the checker of ml_super_resolution_b200/enet/losses.py.
PARITY UNPINNED: the arithmetic lives in TensorFlow 1.8, which cannot run here; each function follows the reference line it
cites and uses torch's conv2d / max_pool2d / matmul as the primitive (TF 'SAME' padding restated explicitly: the extra pixel of
an odd pad goes after)."""
import numpy as np
import torch
import torch.nn.functional as F

VGG_LAYERS = ["block1_conv1", "block1_conv2", "block1_pool", "block2_conv1", "block2_conv2", "block2_pool", "block3_conv1", "block3_conv2",
              "block3_conv3", "block3_conv4", "block3_pool", "block4_conv1", "block4_conv2", "block4_conv3", "block4_conv4", "block4_pool",
              "block5_conv1", "block5_conv2", "block5_conv3", "block5_conv4", "block5_pool"]


def _t(a):
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.asarray(a)).double()


def conv_same(x, w, b, stride):
    """tf.layers.conv2d / tf.nn.conv2d, NHWC, 'same': out = ceil(in / s), pad_total = max((out-1) s + k - in, 0), before = total // 2."""
    k = w.shape[0]

    def pads(n):
        o = -(-n // stride)
        t = max((o - 1) * stride + k - n, 0)
        return t // 2, t - t // 2

    (pt, pb), (pl, pr) = pads(x.shape[1]), pads(x.shape[2])
    y = F.conv2d(F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb)), _t(w).permute(3, 2, 0, 1), _t(b), stride=stride)
    return y.permute(0, 2, 3, 1)


def vgg19(images_pm1, weights):
    """model_vgg.build_vgg19_model (:63-99) of hd_vgg_input = images * 127.5 + 127.5 (model_enet.py:291-292)."""
    x = images_pm1 * 127.5 + 127.5
    t = torch.flip(x, dims=[-1]) - torch.tensor([103.939, 116.779, 123.68], dtype=x.dtype)
    model = {}
    for name in VGG_LAYERS:
        if name.endswith("pool"):
            t = F.max_pool2d(t.permute(0, 3, 1, 2), 2, 2, ceil_mode=True).permute(0, 2, 3, 1)  # 'SAME' 2x2/2
        else:
            t = F.relu(conv_same(t, weights[f"{name}_W_1:0"], weights[f"{name}_b_1:0"], 1))
        model[name] = t
    return model


def normalize(t):
    """model_enet.py:34-41."""
    return t / (t.mean(dim=-1, keepdim=True) + 0.000001)


def perceptual_loss(sr_vgg, hd_vgg):
    """model_enet.py:184-205."""
    l2 = ((normalize(sr_vgg["block2_pool"]) - normalize(hd_vgg["block2_pool"])) ** 2).mean()
    l5 = ((normalize(sr_vgg["block5_pool"]) - normalize(hd_vgg["block5_pool"])) ** 2).mean()
    return 0.2 * l2 + 0.02 * l5


def _patches(t):
    """tf.extract_image_patches(16x16, stride 16, VALID) + reshape [-1, h*w//256, 256, c] (model_enet.py:226-243)."""
    n, h, w, c = t.shape
    p = t.reshape(n, h // 16, 16, w // 16, 16, c).permute(0, 1, 3, 2, 4, 5)  # [n, gy, gx, py, px, c]
    return p.reshape(n, (h // 16) * (w // 16), 256, c)


def texture_matching_loss(sr_vgg, hd_vgg):
    """model_enet.py:208-256."""
    loss = 0
    for name, weight in (("block1_conv1", 3e-7), ("block2_conv1", 1e-6), ("block3_conv1", 1e-6)):
        s, h = _patches(normalize(sr_vgg[name])), _patches(normalize(hd_vgg[name]))
        gs, gh = s.transpose(-1, -2) @ s, h.transpose(-1, -2) @ h
        loss = loss + weight * ((gs - gh) ** 2).mean()
    return loss


def discriminator(images, p, scope="d_"):
    """model_enet.py:118-161."""
    tf_conv_name = lambda i: "conv2d" if i == 0 else f"conv2d_{i}"  # noqa: E731  (TF auto-naming in creation order)
    t, idx = images, 0
    for _ in range(5):
        for stride in (1, 2):
            t = F.leaky_relu(conv_same(t, p[f"{scope}/{tf_conv_name(idx)}/kernel:0"], p[f"{scope}/{tf_conv_name(idx)}/bias:0"], stride), 0.2)
            idx += 1
    t = t.reshape(t.shape[0], -1)
    t = F.leaky_relu(t @ _t(p[f"{scope}/dense/kernel:0"]) + _t(p[f"{scope}/dense/bias:0"]), 0.2)
    return torch.sigmoid(t @ _t(p[f"{scope}/dense_1/kernel:0"]) + _t(p[f"{scope}/dense_1/bias:0"]))


def log_loss(labels, predictions, eps=1e-7):
    """tf.losses.log_loss, Reduction.MEAN."""
    return (-(labels * torch.log(predictions + eps) + (1 - labels) * torch.log(1 - predictions + eps))).mean()


def enet_losses(sr, hd, vgg_weights, d_params, pat_model="pat"):
    """The loss section of build_enet (model_enet.py:288-322) for given sr / hd images: returns the losses, d(g_losses)/d(sr), and
    the gradients of a_loss with respect to the discriminator's variables."""
    sr = _t(sr).clone().requires_grad_(True)
    hd = _t(hd)
    dp = {k: _t(v).clone().requires_grad_(True) for k, v in d_params.items()}
    sr_vgg, hd_vgg = vgg19(sr, vgg_weights), vgg19(hd, vgg_weights)
    out = {}
    g_losses = p_loss = perceptual_loss(sr_vgg, hd_vgg)
    out["p_loss"] = float(p_loss.detach())
    if "a" in pat_model:
        fake, real = discriminator(sr, dp), discriminator(hd, dp)
        a_loss = log_loss(torch.zeros_like(fake), fake) + log_loss(torch.ones_like(real), real)
        g_loss = log_loss(torch.ones_like(fake), fake)
        g_losses = g_losses + (g_loss * 2.0 if "t" in pat_model else g_loss)
        out["a_loss"], out["g_loss"] = float(a_loss.detach()), float(g_loss.detach())
        d_grads = torch.autograd.grad(a_loss, list(dp.values()), retain_graph=True)
        out["d_grads"] = {k: g.numpy() for k, g in zip(dp.keys(), d_grads)}
    if "t" in pat_model:
        t_loss = texture_matching_loss(sr_vgg, hd_vgg)
        g_losses = g_losses + t_loss
        out["t_loss"] = float(t_loss.detach())
    out["g_loss_all"] = float(g_losses.detach())
    out["dsr"] = torch.autograd.grad(g_losses, sr)[0].numpy()
    return out

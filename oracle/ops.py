"""Op-level CPU oracle (test infrastructure; see oracle/__init__.py).

Every function restates one TensorFlow-1.8 / scikit-image / numpy op the reference calls on
the hot path; the docstring cites the reference call site (paths relative to /root/reference)
and the SURVEY.md appendix that spells out the semantics.  numpy for index/byte work, torch
CPU (fp64 by default) for the dense contractions.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------
# conv2d  (tf.layers.conv2d / tf.nn.conv2d+bias_add / tf.contrib.layers.convolution2d)
# ------------------------------------------------------------------------------------------


def _act(y: torch.Tensor, act: str | None) -> torch.Tensor:
    if act in (None, "none", "linear"):
        return y
    if act == "relu":
        return torch.relu(y)
    if act == "tanh":
        return torch.tanh(y)
    raise ValueError(f"unknown activation {act!r}")


def conv2d_nhwc_t(x: torch.Tensor, w_hwio: torch.Tensor, b: torch.Tensor | None, padding: str = "SAME",
                  act: str | None = None) -> torch.Tensor:
    """Differentiable torch version of `conv2d_nhwc` (same semantics, tensors in/out).

    Follows vdsr/vdsr/model_vdsr.py:62-70,85-93; espcn/espcn/model_espcn.py:30-62,117-134;
    srcnn/srcnn.py:100-130; enet/enet/model_enet.py:13-29,63-70: NHWC input,
    HWIO kernel, stride 1, cross-correlation (no flip), bias always present,
    padding 'SAME' (odd k: symmetric zero pad k//2) or 'VALID'.
    """
    kh, kw, cin, cout = w_hwio.shape
    assert x.shape[-1] == cin, (x.shape, w_hwio.shape)
    pad = padding.upper()
    if pad == "SAME":
        assert kh % 2 == 1 and kw % 2 == 1, "SAME restated for odd kernels / stride 1 only"
        p = (kh // 2, kw // 2)
    elif pad == "VALID":
        p = (0, 0)
    else:
        raise ValueError(padding)
    y = F.conv2d(x.permute(0, 3, 1, 2).contiguous(), w_hwio.permute(3, 2, 0, 1).contiguous(), b, stride=1, padding=p)
    return _act(y.permute(0, 2, 3, 1), act)


def conv2d_nhwc(x, w_hwio, b=None, padding="SAME", act=None, dtype=np.float64) -> np.ndarray:
    """numpy in / numpy out wrapper around `conv2d_nhwc_t` computed in `dtype` (fp64 default)."""
    td = torch.float64 if dtype == np.float64 else torch.float32
    xt = torch.as_tensor(np.ascontiguousarray(x)).to(td)
    wt = torch.as_tensor(np.ascontiguousarray(w_hwio)).to(td)
    bt = None if b is None else torch.as_tensor(np.ascontiguousarray(b)).to(td)
    return conv2d_nhwc_t(xt, wt, bt, padding, act).numpy()


def conv2d_nhwc_im2col(x, w_hwio, b=None, padding="SAME", act=None) -> np.ndarray:
    """Independent restatement of A.1 as an explicit im2col GEMM in numpy fp64.

    A[M = N*Ho*Wo, K = kh*kw*Cin] with K index (u*kw+v)*Cin+ci times HWIO flattened [K, Cout]
    (no transpose) -- exactly the implicit-GEMM view the CUDA kernels use.  Used to
    cross-validate the torch-based oracle; small shapes only.
    """
    x = np.asarray(x, np.float64)
    w = np.asarray(w_hwio, np.float64)
    kh, kw, cin, cout = w.shape
    n, h, wd, _ = x.shape
    if padding.upper() == "SAME":
        ph, pw = kh // 2, kw // 2
        xp = np.zeros((n, h + 2 * ph, wd + 2 * pw, cin))
        xp[:, ph:ph + h, pw:pw + wd] = x
        ho, wo = h, wd
    else:
        xp = x
        ho, wo = h - kh + 1, wd - kw + 1
    cols = np.empty((n, ho, wo, kh * kw * cin))
    for u in range(kh):
        for v in range(kw):
            cols[..., (u * kw + v) * cin:(u * kw + v + 1) * cin] = xp[:, u:u + ho, v:v + wo, :]
    y = cols.reshape(-1, kh * kw * cin) @ w.reshape(kh * kw * cin, cout)
    if b is not None:
        y = y + np.asarray(b, np.float64)
    y = y.reshape(n, ho, wo, cout)
    if act == "relu":
        y = np.maximum(y, 0)
    elif act == "tanh":
        y = np.tanh(y)
    return y


def conv2d_backward(x, w_hwio, b, dy, padding="SAME", act=None, dtype=np.float64):
    """dL/dx, dL/dw, dL/db for y = act(conv(x,w)+b) given dL/dy (autodiff of A.1/A.2).

    This is what `optimizer.minimize` builds implicitly (vdsr/vdsr/model_vdsr.py:146-148,
    espcn/espcn/model_espcn.py:87-89, srcnn/srcnn.py:155-157): dgrad, wgrad, bgrad with the
    activation mask (ReLU: y>0, tanh: 1-y^2).
    """
    td = torch.float64 if dtype == np.float64 else torch.float32
    xt = torch.as_tensor(np.ascontiguousarray(x)).to(td).requires_grad_(True)
    wt = torch.as_tensor(np.ascontiguousarray(w_hwio)).to(td).requires_grad_(True)
    bt = torch.as_tensor(np.ascontiguousarray(b)).to(td).requires_grad_(True)
    y = conv2d_nhwc_t(xt, wt, bt, padding, act)
    y.backward(torch.as_tensor(np.ascontiguousarray(dy)).to(td))
    return xt.grad.numpy(), wt.grad.numpy(), bt.grad.numpy()


# ------------------------------------------------------------------------------------------
# ESPCN sub-pixel packing (pixel shuffle) -- index-only, bit exact
# ------------------------------------------------------------------------------------------


def pixel_shuffle(packed: np.ndarray, r: int) -> np.ndarray:
    """[..., h, w, r*r*C] -> [..., h*r, w*r, C] with packed channel k = (dy*r+dx)*C + c.

    Closed form of the host un-pack at espcn/espcn/experiment_test.py:173-177:
    out[i*r+dy, j*r+dx, c] = packed[i, j, (dy*r+dx)*C + c].  Equals tf.depth_to_space(NHWC),
    NOT torch.pixel_shuffle's c*r^2+dy*r+dx ordering.
    """
    *lead, h, w, k = packed.shape
    c = k // (r * r)
    assert c * r * r == k
    t = packed.reshape(*lead, h, w, r, r, c)  # [.., i, j, dy, dx, c]
    nl = len(lead)
    t = np.moveaxis(t, nl + 2, nl + 1)  # [.., i, dy, j, dx, c]
    return np.ascontiguousarray(t).reshape(*lead, h * r, w * r, c)


def pixel_unshuffle(hr: np.ndarray, r: int) -> np.ndarray:
    """Inverse of `pixel_shuffle`: the training-label packing at espcn/espcn/dataset.py:140-156
    and espcn/espcn/experiment_test.py:91-96: packed[i,j,(dy*r+dx)*C+c] = hr[i*r+dy, j*r+dx, c]."""
    *lead, hh, ww, c = hr.shape
    h, w = hh // r, ww // r
    assert h * r == hh and w * r == ww
    t = hr.reshape(*lead, h, r, w, r, c)  # [.., i, dy, j, dx, c]
    nl = len(lead)
    t = np.moveaxis(t, nl + 1, nl + 2)  # [.., i, j, dy, dx, c]
    return np.ascontiguousarray(t).reshape(*lead, h, w, r * r * c)


def pixel_shuffle_reference_literal(packed_hw_k: np.ndarray, r: int) -> np.ndarray:
    """The reference's own numpy sequence, restated call for call (np.split / reshape /
    concatenate) for a single [h, w, r*r*3] image -- espcn/espcn/experiment_test.py:173-177.
    Exists only to pin `pixel_shuffle` against the reference's algorithm."""
    lrh, lrw, _ = packed_hw_k.shape
    patches = np.split(packed_hw_k, lrw, axis=1)
    patches = [np.reshape(p, [lrh * r, r, 3]) for p in patches]
    return np.concatenate(patches, axis=1)


def pixel_unshuffle_reference_literal(hr_hw_c: np.ndarray, r: int) -> np.ndarray:
    """espcn/espcn/experiment_test.py:91-96 (same as dataset.py:140-156), call for call."""
    h, w, _ = hr_hw_c.shape
    patches = np.split(hr_hw_c, w // r, axis=1)
    patches = [np.reshape(im, [h // r, 1, -1]) for im in patches]
    return np.concatenate(patches, axis=1)


# ------------------------------------------------------------------------------------------
# TF1 legacy resize ops
# ------------------------------------------------------------------------------------------

_BICUBIC_TABLE_SIZE = 1 << 10


def _bicubic_table() -> np.ndarray:
    """TF-1.x `InitCoeffsTable` (A = -0.75): double arithmetic on a float abscissa, stored as fp32;
    T[2j], T[2j+1] for j = 0..1024."""
    a = -0.75
    t = np.zeros((_BICUBIC_TABLE_SIZE + 1) * 2, np.float32)
    for j in range(_BICUBIC_TABLE_SIZE + 1):
        x = float(np.float32(j * 1.0 / _BICUBIC_TABLE_SIZE))
        t[2 * j] = np.float32(((a + 2) * x - (a + 3)) * x * x + 1)
        x = float(np.float32(x + 1.0))
        t[2 * j + 1] = np.float32(((a * x - 5 * a) * x + 8 * a) * x - 4 * a)
    return t


BICUBIC_TABLE = _bicubic_table()


def bicubic_taps_tf1(out_size: int, in_size: int):
    """Per-output-index tap indices [out,4] (int64) and fp32 weights [out,4] of TF1's legacy
    `tf.image.resize_bicubic(align_corners=False)` -- srcnn/srcnn.py:89-93, SURVEY A.4."""
    scale = np.float32(in_size) / np.float32(out_size)
    idx = np.zeros((out_size, 4), np.int64)
    wts = np.zeros((out_size, 4), np.float32)
    for o in range(out_size):
        f = np.float32(o) * scale
        i = int(f)  # trunc; f >= 0
        delta = np.float32(f - np.float32(i))
        t = int(np.rint(np.float32(delta * np.float32(_BICUBIC_TABLE_SIZE))))  # lrintf
        wts[o] = (BICUBIC_TABLE[2 * t + 1], BICUBIC_TABLE[2 * t],
                  BICUBIC_TABLE[2 * (_BICUBIC_TABLE_SIZE - t)], BICUBIC_TABLE[2 * (_BICUBIC_TABLE_SIZE - t) + 1])
        for k in range(4):
            idx[o, k] = min(max(i - 1 + k, 0), in_size - 1)
    return idx, wts


def resize_bicubic_tf1(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """TF1 legacy bicubic, NHWC fp32: for each output pixel, 4 rows each reduced over 4 x-taps,
    then 4 y-taps, all fp32 in the order w0*v0+w1*v1+w2*v2+w3*v3."""
    x = np.asarray(x, np.float32)
    n, h, w, c = x.shape
    yi, yw = bicubic_taps_tf1(out_h, h)
    xi, xw = bicubic_taps_tf1(out_w, w)
    out = np.zeros((n, out_h, out_w, c), np.float32)
    for oy in range(out_h):
        rows = []
        for k in range(4):
            src = x[:, yi[oy, k]]  # [n, w, c]
            g = src[:, xi, :]  # [n, out_w, 4, c]
            xwb = xw[None, :, :, None]
            r = g[:, :, 0] * xwb[:, :, 0] + g[:, :, 1] * xwb[:, :, 1]
            r = r + g[:, :, 2] * xwb[:, :, 2]
            r = r + g[:, :, 3] * xwb[:, :, 3]
            rows.append(r.astype(np.float32))
        acc = rows[0] * yw[oy, 0] + rows[1] * yw[oy, 1]
        acc = acc + rows[2] * yw[oy, 2]
        acc = acc + rows[3] * yw[oy, 3]
        out[:, oy] = acc
    return out


def resize_nearest_tf1(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """TF1 legacy `tf.image.resize_nearest_neighbor` -- enet/enet/model_enet.py:78-80, SURVEY A.9:
    src = min(floor(dst * in/out), in-1) with an fp32 scale."""
    n, h, w, c = x.shape
    sy = np.float32(h) / np.float32(out_h)
    sx = np.float32(w) / np.float32(out_w)
    yi = np.minimum(np.floor(np.arange(out_h, dtype=np.float32) * sy).astype(np.int64), h - 1)
    xi = np.minimum(np.floor(np.arange(out_w, dtype=np.float32) * sx).astype(np.int64), w - 1)
    return x[:, yi][:, :, xi]


# ------------------------------------------------------------------------------------------
# VDSR / ESPCN degrade pre-pass (skimage gaussian + skimage resize order=1)
# ------------------------------------------------------------------------------------------


def gaussian_kernel1d(sigma: float) -> np.ndarray:
    """scipy.ndimage / skimage.filters.gaussian taps: radius int(4*sigma+0.5), exp(-x^2/2s^2)
    normalised, fp64."""
    if sigma <= 0:
        return np.ones(1)
    radius = int(4.0 * sigma + 0.5)
    xs = np.arange(-radius, radius + 1, dtype=np.float64)
    k = np.exp(-0.5 * xs * xs / (sigma * sigma))
    return k / k.sum()


def gaussian_blur_nearest(img: np.ndarray, sigma: float) -> np.ndarray:
    """`skimage.filters.gaussian(img[H,W,C], sigma, mode='nearest')` == separable blur over H and
    W only (channels untouched), replicate border, fp64.  Correlation order as scipy.ndimage:
    axis 0 then axis 1."""
    img = np.asarray(img, np.float64)
    k = gaussian_kernel1d(sigma)
    r = len(k) // 2
    if r == 0:
        return img.copy()
    out = img
    for axis in (0, 1):
        n = out.shape[axis]
        idx = np.clip(np.arange(-r, n + r), 0, n - 1)
        padded = np.take(out, idx, axis=axis)
        acc = np.zeros_like(out)
        for t in range(2 * r + 1):
            sl = [slice(None)] * out.ndim
            sl[axis] = slice(t, t + n)
            acc = acc + k[t] * padded[tuple(sl)]
        out = acc
    return out


def bilinear_taps_halfpixel(out_size: int, in_size: int):
    """skimage.transform.resize(order=1, mode='edge') source coordinates:
    src = (dst+0.5)*(in/out)-0.5 clamped to [0,in-1]; lo=floor, hi=min(lo+1,in-1), frac."""
    scale = in_size / out_size
    src = (np.arange(out_size, dtype=np.float64) + 0.5) * scale - 0.5
    src = np.clip(src, 0.0, in_size - 1.0)
    lo = np.floor(src).astype(np.int64)
    hi = np.minimum(lo + 1, in_size - 1)
    return lo, hi, src - lo


def resize_bilinear_edge(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """[H,W,C] fp64 bilinear, lerp along x then along y (vdsr/vdsr/dataset.py:32-36)."""
    img = np.asarray(img, np.float64)
    ylo, yhi, yf = bilinear_taps_halfpixel(out_h, img.shape[0])
    xlo, xhi, xf = bilinear_taps_halfpixel(out_w, img.shape[1])
    xf = xf[None, :, None]
    top = img[ylo][:, xlo] * (1.0 - xf) + img[ylo][:, xhi] * xf
    bot = img[yhi][:, xlo] * (1.0 - xf) + img[yhi][:, xhi] * xf
    yf = yf[:, None, None]
    return top * (1.0 - yf) + bot * yf


def hd_image_to_sd_image(hd_image: np.ndarray, scaling_factor: float) -> np.ndarray:
    """vdsr/vdsr/dataset.py:13-38: blur sigma=max(0,0.5(s-1)), bilinear down to int(H/s) x int(W/s),
    bilinear back up; fp64 throughout (skimage converts to float64)."""
    hd_h, hd_w, _ = hd_image.shape
    sd_h = int(hd_h / scaling_factor)
    sd_w = int(hd_w / scaling_factor)
    sigma = max(0.0, 0.5 * (scaling_factor - 1.0))
    bl = gaussian_blur_nearest(hd_image, sigma)
    sd = resize_bilinear_edge(bl, sd_h, sd_w)
    return resize_bilinear_edge(sd, hd_h, hd_w)


def vdsr_image_batches(images, scaling_factors, image_size, batch_size, rng):
    """vdsr/vdsr/dataset.py:41-128 `image_batches` with the directory replaced by its decoded uint8 images (the order of
    `images` stands for `tf.gfile.ListDirectory`): same random-number call sequence on `rng` (a numpy RandomState standing for
    the global `np.random`), same numpy / skimage arithmetic (img_as_float32 = x/255 in fp32; the degrade runs in fp64 and the
    stacked batch is float64 like the reference's).  Endless generator of (sd_images, hd_images)."""
    if scaling_factors is None or len(scaling_factors) <= 0:
        scaling_factors = [2.0, 3.0, 4.0]
    if any([s <= 1 for s in scaling_factors]):
        raise Exception('invalide scaling factors')
    order = list(range(len(images)))

    def image_indices():
        while True:
            rng.shuffle(order)
            for i in order:
                yield i

    gen = image_indices()
    sd_images, hd_images = [], []
    while True:
        hd_image = images[next(gen)]
        h, w, c = hd_image.shape
        if h < image_size or w < image_size or c != 3:
            continue
        x = rng.randint(w - image_size)
        y = rng.randint(h - image_size)
        hd_image = hd_image[y:y + image_size, x:x + image_size, :]
        if 1 == rng.choice([0, 1]):
            hd_image = hd_image[:, ::-1, :]
        hd_image = np.divide(hd_image, 255, dtype=np.float32)  # skimage.util.img_as_float32 of a uint8 image
        scaling_factor = rng.choice(scaling_factors)
        sd_image = hd_image_to_sd_image(hd_image, scaling_factor)
        sd_image = sd_image * 2.0 - 1.0
        hd_image = hd_image * 2.0 - 1.0
        sd_images.append(sd_image)
        hd_images.append(hd_image)
        if len(sd_images) == batch_size:
            yield np.stack(sd_images, axis=0), np.stack(hd_images, axis=0)
            sd_images, hd_images = [], []


def espcn_lr_from_hr(hr_image: np.ndarray, r: int) -> np.ndarray:
    """espcn/espcn/experiment_test.py:76-87: gaussian sigma=0.5(r-1) then stride-r decimation
    at offset r//2.  `hr_image` already in [-1,1]."""
    sigma = max(0.0, 0.5 * (r - 1.0))
    bl = gaussian_blur_nearest(hr_image, sigma)
    off = r // 2
    return bl[off::r, off::r]


# ------------------------------------------------------------------------------------------
# losses / metrics
# ------------------------------------------------------------------------------------------


def mse_mean(a, b) -> float:
    """tf.losses.mean_squared_error(reduction=MEAN) = sum((a-b)^2)/numel;
    vdsr/vdsr/model_vdsr.py:120-123, espcn/espcn/model_espcn.py:76-77."""
    d = np.asarray(a, np.float64) - np.asarray(b, np.float64)
    return float((d * d).sum() / d.size)


def l2norm_rows_mean(sr, hd, cols: int):
    """srcnn/srcnn.py:142-144: d = reshape(sr-hd, [-1, cols]); loss = mean_rows ||d_row||_2.
    Returns (loss, dloss/dsr)."""
    d = (np.asarray(sr, np.float64) - np.asarray(hd, np.float64)).reshape(-1, cols)
    nrm = np.sqrt((d * d).sum(axis=1))
    loss = float(nrm.mean())
    grad = d / (np.maximum(nrm, 1e-300)[:, None] * d.shape[0])
    return loss, grad.reshape(np.shape(sr))


def psnr(a, b, max_val: float) -> np.ndarray:
    """tf.image.psnr per image over H,W,C (vdsr/vdsr/experiment_train.py:80, max_val 2.0)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    mse = ((a - b) ** 2).reshape(a.shape[0], -1).mean(axis=1)
    return 20.0 * math.log10(max_val) - 10.0 * np.log10(mse)


def ssim_tf(a, b, max_val: float) -> np.ndarray:
    """tf.image.ssim per image (TF 1.8 image_ops_impl._ssim_helper / _ssim_per_channel): 11x11 gaussian window with
    sigma 1.5 (softmax-normalised outer product), VALID depthwise filtering of x, y, x*y and x^2+y^2, k1 = 0.01, k2 = 0.03,
    luminance * contrast-structure averaged over window positions, then over channels.
    vdsr/vdsr/experiment_evaluate.py:59-60 (max_val 2.0); espcn/espcn/experiment_test.py:53 (1.0)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    size, sigma = 11, 1.5
    coords = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
    g = np.exp(-(coords ** 2) / (2.0 * sigma * sigma))
    g /= g.sum()

    def reducer(t):  # [N,H,W,C] -> VALID separable gaussian
        t = np.lib.stride_tricks.sliding_window_view(t, size, axis=1) @ g      # [N,H-10,W,C]
        return np.lib.stride_tricks.sliding_window_view(t, size, axis=2) @ g   # [N,H-10,W-10,C]

    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    mean0, mean1 = reducer(a), reducer(b)
    num0 = mean0 * mean1 * 2.0
    den0 = mean0 ** 2 + mean1 ** 2
    luminance = (num0 + c1) / (den0 + c1)
    num1 = reducer(a * b) * 2.0
    den1 = reducer(a ** 2 + b ** 2)
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    return (luminance * cs).mean(axis=(1, 2)).mean(axis=-1)


def rgb_to_yuv_y(rgb) -> np.ndarray:
    """First channel of tf.image.rgb_to_yuv (kernel column [0.299, 0.587, 0.114]); espcn/espcn/experiment_test.py:45-49."""
    rgb = np.asarray(rgb, np.float64)
    return rgb[..., 0] * 0.299 + rgb[..., 1] * 0.587 + rgb[..., 2] * 0.114


def espcn_scores(sr_packed, hr_packed, scaling_factor: int, score_space: str = "rgb"):
    """espcn/espcn/experiment_test.py:30-53: remap [-1,1] -> [0,1], clip, optionally reinterpret the packed tensor
    [N,h,w,3r^2] as [N,h,w*r^2,3] and keep Y, then psnr / ssim with max_val 1.0.  Returns (psnrs, ssims)."""
    sr = np.clip(np.asarray(sr_packed, np.float64) * 0.5 + 0.5, 0.0, 1.0)
    hr = np.clip(np.asarray(hr_packed, np.float64) * 0.5 + 0.5, 0.0, 1.0)
    if score_space == "y":
        n, h, w, _ = hr.shape
        w2 = w * scaling_factor ** 2
        sr = rgb_to_yuv_y(sr.reshape(-1, h, w2, 3))[..., None]
        hr = rgb_to_yuv_y(hr.reshape(-1, h, w2, 3))[..., None]
    return psnr(hr, sr, 1.0), ssim_tf(hr, sr, 1.0)


def saturate_cast_u8(x, scale=127.5, bias=127.5) -> np.ndarray:
    """tf.saturate_cast(x * 127.5 + 127.5, tf.uint8) in fp32 (vdsr/vdsr/experiment_resolve.py:65-67): clamp, then truncate."""
    v = np.asarray(x, np.float32) * np.float32(scale) + np.float32(bias)
    return np.clip(v, 0.0, 255.0).astype(np.uint8)


# ------------------------------------------------------------------------------------------
# PIL / scipy.misc.imresize on uint8 (EnhanceNet's input pipeline, enet/enet/datasets.py:112-113)
# ------------------------------------------------------------------------------------------
# `scipy.misc.imresize(arr, percent, interp)` of a uint8 RGB array is `PIL.Image.fromarray(arr).resize(size, resample)` with
# size = (arr.shape[:2][::-1] * percent / 100).astype(int), default interp 'bilinear'.  Pillow resamples separably, first
# horizontally then vertically, with fixed-point coefficients (libImaging/Resample.c): PINNED against the Pillow installed in
# this environment by tests/test_oracle_cpu.py.
PIL_PRECISION_BITS = 32 - 8 - 2


def _pil_bilinear(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _pil_bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


PIL_FILTERS = {"bilinear": (_pil_bilinear, 1.0), "bicubic": (_pil_bicubic, 2.0)}


def pil_resample_coeffs(in_size: int, out_size: int, interp: str):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc -> (ksize, bounds int32 [out,2] = (xmin, count), kk int32 [out,ksize])."""
    filt, fsupport = PIL_FILTERS[interp]
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = fsupport * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        bounds[xx] = (xmin, xmax)
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << PIL_PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PIL_PRECISION_BITS))
    return ksize, bounds, kk


def pil_resize_u8(img: np.ndarray, out_h: int, out_w: int, interp: str) -> np.ndarray:
    """PIL.Image.resize of a uint8 HWC image: horizontal pass to a uint8 intermediate, then vertical pass; each output is
    clip8((2^21 + sum pixel * coeff) >> 22)."""
    img = np.asarray(img, np.uint8)
    h, w, c = img.shape

    def one_pass(src, in_size, out_size, axis):
        if in_size == out_size:
            return src
        _, bounds, kk = pil_resample_coeffs(in_size, out_size, interp)
        src = np.moveaxis(src, axis, 0).astype(np.int64)
        out = np.empty((out_size,) + src.shape[1:], np.uint8)
        for xx in range(out_size):
            x0, n = bounds[xx]
            acc = (1 << (PIL_PRECISION_BITS - 1)) + np.tensordot(kk[xx, :n].astype(np.int64), src[x0:x0 + n], axes=(0, 0))
            out[xx] = np.clip(acc >> PIL_PRECISION_BITS, 0, 255).astype(np.uint8)
        return np.moveaxis(out, 0, axis)

    tmp = one_pass(img, w, out_w, 1)
    return one_pass(tmp, h, out_h, 0)


def imresize_percent(arr: np.ndarray, percent: int, interp: str = "bilinear") -> np.ndarray:
    """scipy.misc.imresize(arr, percent, interp) for a uint8 HWC array (SciPy <= 1.2: PIL based)."""
    h, w = arr.shape[:2]
    ow, oh = (np.array([w, h]) * percent / 100.0).astype(int)
    return pil_resize_u8(arr, int(oh), int(ow), interp)


def enet_batch(images, crops):
    """enet/enet/datasets.py:100-125 for given crop origins [(image, y, x), ...]: 128x128 crop, sd = imresize(hd, 25),
    bq = imresize(sd, 400, 'bicubic'), all three mapped by x.astype(float32) / 127.5 - 1.0.  -> (sd, bq, hd) float32."""
    sd_images, bq_images, hd_images = [], [], []
    for (i, y, x) in crops:
        hd = images[i][y:y + 128, x:x + 128, :]
        sd = imresize_percent(hd, 25)
        bq = imresize_percent(sd, 400, "bicubic")
        sd_images.append(sd.astype(np.float32) / 127.5 - 1.0)
        bq_images.append(bq.astype(np.float32) / 127.5 - 1.0)
        hd_images.append(hd.astype(np.float32) / 127.5 - 1.0)
    return np.stack(sd_images), np.stack(bq_images), np.stack(hd_images)


def enet_image_batches(images, batch_size, rng):
    """enet/enet/datasets.py:78-127 `image_batches` with the directory replaced by its decoded uint8 images: shuffle per epoch,
    per sample x = randint(128) then y = randint(128), 128x128 crop; endless generator of (sd, bq, hd)."""
    order = list(range(len(images)))

    def indices():
        while True:
            rng.shuffle(order)
            for i in order:
                yield i

    gen = indices()
    while True:
        crops = []
        for _ in range(batch_size):
            i = next(gen)
            x = rng.randint(128)
            y = rng.randint(128)
            crops.append((i, y, x))
        yield enet_batch(images, crops)


def feature_mosaic_u8(feature_map) -> np.ndarray:
    """vdsr/vdsr/experiment_feature_map_visualize.py:80-110 `encode_feature_map` before encode_png: split the 64 channels,
    concat 8 per row along the width, rows along the height, saturate_cast(x*127.5+127.5)."""
    t = np.asarray(feature_map, np.float32)[0]
    maps = np.split(t, 64, axis=-1)
    rows = [np.concatenate(maps[i:i + 8], axis=1) for i in range(0, 64, 8)]
    return saturate_cast_u8(np.concatenate(rows, axis=0))


# ------------------------------------------------------------------------------------------
# optimisers (TF formulas)
# ------------------------------------------------------------------------------------------


def adam_tf(w, g, m, v, t: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8, dtype=np.float32):
    """One tf.train.AdamOptimizer step; `t` is the 1-based step after increment.
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; w -= lr_t*m/(sqrt(v)+eps)  (epsilon-hat form).
    vdsr/vdsr/model_vdsr.py:145-148; espcn/espcn/model_espcn.py:87-89; srcnn/srcnn.py:155-156."""
    w = np.asarray(w, dtype)
    g = np.asarray(g, dtype)
    m = np.asarray(m, dtype)
    v = np.asarray(v, dtype)
    lr_t = dtype(lr) * dtype(math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t))
    m = dtype(beta1) * m + dtype(1.0 - beta1) * g
    v = dtype(beta2) * v + dtype(1.0 - beta2) * g * g
    w = w - lr_t * m / (np.sqrt(v) + dtype(eps))
    return w.astype(dtype), m.astype(dtype), v.astype(dtype)


def momentum_clip_tf(w, g, accum, lr: float, momentum=0.9, gradient_cap=0.01, dtype=np.float32):
    """vdsr/vdsr/model_vdsr.py:158-184: g <- clip(g, +-cap/lr); a <- mom*a + g; w <- w - lr*a."""
    cap = dtype(gradient_cap) / dtype(lr)
    g = np.clip(np.asarray(g, dtype), -cap, cap)
    accum = dtype(momentum) * np.asarray(accum, dtype) + g
    w = np.asarray(w, dtype) - dtype(lr) * accum
    return w.astype(dtype), accum.astype(dtype)


def stepwise_lr(lr0: float, factor: float, step: int, decay_steps: int) -> float:
    """vdsr/vdsr/experiment_train.py:130; espcn/espcn/experiment_train.py:101-107."""
    return lr0 * (factor ** (step // decay_steps))


# ------------------------------------------------------------------------------------------
# bf16 helpers shared by tests (round-to-nearest-even like cvt.rn.bf16.f32)
# ------------------------------------------------------------------------------------------


def bf16_round(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 -> fp32 with round-to-nearest-even (what the kernels store)."""
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()

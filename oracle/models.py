"""Model-level CPU oracle (test infrastructure; see oracle/__init__.py).

The four reference graphs restated as pure functions over a `{tf_variable_name: ndarray}`
parameter dict (HWIO kernels), computed with torch CPU in fp64 (parity) or fp32 (timing).
Backward passes come from torch autograd over the same graph, which is what
`optimizer.minimize(loss)` builds in the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from .ops import conv2d_nhwc_t

# ------------------------------------------------------------------------------------------
# initialisers
# ------------------------------------------------------------------------------------------


def xavier_uniform(rng: np.random.Generator, kh, kw, cin, cout) -> np.ndarray:
    """tf.contrib.layers.xavier_initializer(): U(+-sqrt(6/(fan_in+fan_out))), fan = k*k*C
    (vdsr/vdsr/model_vdsr.py:27)."""
    lim = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
    return rng.uniform(-lim, lim, size=(kh, kw, cin, cout)).astype(np.float32)


def truncated_normal(rng: np.random.Generator, shape, stddev) -> np.ndarray:
    """tf.truncated_normal_initializer: resample beyond 2 sigma (espcn model_espcn.py:21,
    enet model_enet.py:11,47, srcnn.py:84)."""
    out = rng.standard_normal(size=shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * stddev).astype(np.float32)


def _tf_conv_name(i: int) -> str:
    return "conv2d" if i == 0 else f"conv2d_{i}"


def vdsr_init(seed=42, num_layers=20, channels=3, bias_scale=0.0) -> dict:
    """Variables `conv2d/kernel:0 ... conv2d_19/bias:0` in creation order."""
    rng = np.random.default_rng(seed)
    p = {}
    for i in range(num_layers):
        cin = channels if i == 0 else 64
        cout = channels if i == num_layers - 1 else 64
        p[f"{_tf_conv_name(i)}/kernel:0"] = xavier_uniform(rng, 3, 3, cin, cout)
        p[f"{_tf_conv_name(i)}/bias:0"] = (bias_scale * rng.standard_normal(cout)).astype(np.float32)
    return p


def espcn_init(seed=42, scaling_factor=3, channels=3, stddev=0.02, bias_scale=0.0) -> dict:
    rng = np.random.default_rng(seed)
    shapes = {"f1": (5, 5, channels, 64), "f2": (3, 3, 64, 32), "f3": (3, 3, 32, channels * scaling_factor ** 2)}
    p = {}
    for name, s in shapes.items():
        p[f"{name}/kernel:0"] = truncated_normal(rng, s, stddev)
        p[f"{name}/bias:0"] = (bias_scale * rng.standard_normal(s[3])).astype(np.float32)
    return p


def srcnn_init(seed=42, channels=3, f=(9, 1, 5), n=(64, 32), stddev=0.001, bias_scale=0.0) -> dict:
    rng = np.random.default_rng(seed)
    shapes = {"patch_extraction": (f[0], f[0], channels, n[0]), "non_linear_mapping": (f[1], f[1], n[0], n[1]),
              "reconstruction": (f[2], f[2], n[1], channels)}
    p = {}
    for name, s in shapes.items():
        p[f"{name}/weights:0"] = truncated_normal(rng, s, stddev)
        p[f"{name}/biases:0"] = (bias_scale * rng.standard_normal(s[3])).astype(np.float32)
    return p


ENET_G_LAYERS = ([(3, 3, 64)] + [(3, 64, 64), (1, 64, 64)] * 10 + [(3, 64, 64)] * 3 + [(3, 64, 3)])


def enet_g_init(seed=42, stddev=0.02, bias_scale=0.0) -> dict:
    """25 convs `g_/conv2d ... g_/conv2d_24` (enet/enet/model_enet.py:60,275)."""
    rng = np.random.default_rng(seed)
    p = {}
    for i, (k, cin, cout) in enumerate(ENET_G_LAYERS):
        p[f"g_/{_tf_conv_name(i)}/kernel:0"] = truncated_normal(rng, (k, k, cin, cout), stddev)
        p[f"g_/{_tf_conv_name(i)}/bias:0"] = (bias_scale * rng.standard_normal(cout)).astype(np.float32)
    return p


# ------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------


def _td(dtype):
    return torch.float64 if dtype in (np.float64, "float64", torch.float64) else torch.float32


def _to_t(params: dict, dtype, requires_grad=False) -> dict:
    td = _td(dtype)
    return {k: torch.as_tensor(np.ascontiguousarray(v)).to(td).requires_grad_(requires_grad) for k, v in params.items()}


def _t(x, dtype):
    return torch.as_tensor(np.ascontiguousarray(x)).to(_td(dtype))


# ------------------------------------------------------------------------------------------
# VDSR  (vdsr/vdsr/model_vdsr.py:6-192)
# ------------------------------------------------------------------------------------------


def vdsr_forward_t(p: dict, sd: torch.Tensor, num_layers=20, taps: dict | None = None) -> torch.Tensor:
    """19x[conv3x3->64, ReLU] (ReLU applied twice in the reference -- idempotent, :68,:74) then
    conv3x3->C linear, sr = sd + residual (:104)."""
    t = sd
    for i in range(num_layers - 1):
        n = _tf_conv_name(i)
        t = conv2d_nhwc_t(t, p[f"{n}/kernel:0"], p[f"{n}/bias:0"], "SAME", "relu")
        if taps is not None:
            taps[f"conv.{i + 1}"] = t
            taps[f"relu.{i + 1}"] = t
    n = _tf_conv_name(num_layers - 1)
    res = conv2d_nhwc_t(t, p[f"{n}/kernel:0"], p[f"{n}/bias:0"], "SAME", None)
    if taps is not None:
        taps[f"conv.{num_layers}"] = res
    return sd + res


def vdsr_forward(params: dict, sd: np.ndarray, num_layers=20, dtype=np.float64) -> dict:
    taps = {}
    with torch.no_grad():
        sr = vdsr_forward_t(_to_t(params, dtype), _t(sd, dtype), num_layers, taps)
    out = {k: v.numpy() for k, v in taps.items()}
    out["sd_images"] = np.asarray(sd)
    out["sr_images"] = sr.numpy()
    return out


def vdsr_loss_and_grads(params: dict, sd, hd, num_layers=20, weight_decay=1e-4, dtype=np.float64):
    """loss = MSE_mean(hd, sr) + sum_kernels wd * 0.5*sum(w^2)  (:120-125);
    returns (loss, mse, {var: grad}, sr)."""
    p = _to_t(params, dtype, requires_grad=True)
    sdt, hdt = _t(sd, dtype), _t(hd, dtype)
    sr = vdsr_forward_t(p, sdt, num_layers)
    mse = ((sr - hdt) ** 2).mean()
    reg = sum(weight_decay * 0.5 * (v ** 2).sum() for k, v in p.items() if k.endswith("kernel:0"))
    loss = mse + reg
    loss.backward()
    return float(loss.detach()), float(mse.detach()), {k: v.grad.numpy() for k, v in p.items()}, sr.detach().numpy()


# ------------------------------------------------------------------------------------------
# ESPCN  (espcn/espcn/model_espcn.py:6-147)
# ------------------------------------------------------------------------------------------


def espcn_forward_t(p: dict, lr: torch.Tensor) -> torch.Tensor:
    t = conv2d_nhwc_t(lr, p["f1/kernel:0"], p["f1/bias:0"], "SAME", "tanh")
    t = conv2d_nhwc_t(t, p["f2/kernel:0"], p["f2/bias:0"], "SAME", "tanh")
    return conv2d_nhwc_t(t, p["f3/kernel:0"], p["f3/bias:0"], "SAME", None)


def espcn_forward(params: dict, lr: np.ndarray, dtype=np.float64) -> np.ndarray:
    """Packed (un-shuffled) `sr_result` [N,h,w,C*r^2], as build_test_model returns (:143-147)."""
    with torch.no_grad():
        return espcn_forward_t(_to_t(params, dtype), _t(lr, dtype)).numpy()


def espcn_loss_and_grads(params: dict, lr, hr_packed, dtype=np.float64):
    """MSE in packed space (:76-77)."""
    p = _to_t(params, dtype, requires_grad=True)
    sr = espcn_forward_t(p, _t(lr, dtype))
    loss = ((sr - _t(hr_packed, dtype)) ** 2).mean()
    loss.backward()
    return float(loss.detach()), {k: v.grad.numpy() for k, v in p.items()}, sr.detach().numpy()


# ------------------------------------------------------------------------------------------
# SRCNN  (srcnn/srcnn.py:81-166) -- convs + crop + loss; the bicubic pre-pass is ops.resize_bicubic_tf1
# ------------------------------------------------------------------------------------------


def srcnn_forward_t(p: dict, lo: torch.Tensor) -> torch.Tensor:
    t = conv2d_nhwc_t(lo, p["patch_extraction/weights:0"], p["patch_extraction/biases:0"], "VALID", "relu")
    t = conv2d_nhwc_t(t, p["non_linear_mapping/weights:0"], p["non_linear_mapping/biases:0"], "VALID", "relu")
    return conv2d_nhwc_t(t, p["reconstruction/weights:0"], p["reconstruction/biases:0"], "VALID", "tanh")


def srcnn_forward(params: dict, lo: np.ndarray, dtype=np.float64) -> np.ndarray:
    with torch.no_grad():
        return srcnn_forward_t(_to_t(params, dtype), _t(lo, dtype)).numpy()


def srcnn_loss_and_grads(params: dict, lo, hi, dtype=np.float64):
    """hi cropped by the VALID border; loss = mean over rows of ||reshape(sr-hi,[-1,bb^2])||_2
    (:132-144)."""
    p = _to_t(params, dtype, requires_grad=True)
    sr = srcnn_forward_t(p, _t(lo, dtype))
    bb = sr.shape[1]
    side = (hi.shape[1] - bb) // 2
    hit = _t(hi, dtype)[:, side:side + bb, side:side + bb, :]
    d = (sr - hit).reshape(-1, bb * bb)
    loss = torch.linalg.vector_norm(d, ord=2, dim=1).mean()
    loss.backward()
    return float(loss.detach()), {k: v.grad.numpy() for k, v in p.items()}, sr.detach().numpy()


# ------------------------------------------------------------------------------------------
# EnhanceNet generator  (enet/enet/model_enet.py:8-31,44-115)
# ------------------------------------------------------------------------------------------


def _nn_up2_t(t: torch.Tensor) -> torch.Tensor:
    """resize_nearest_neighbor to exactly 2x: src = dst >> 1."""
    return t.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)


def enet_generator_forward_t(p: dict, sd: torch.Tensor, bq: torch.Tensor) -> torch.Tensor:
    def conv(i, t, act):
        n = f"g_/{_tf_conv_name(i)}"
        return conv2d_nhwc_t(t, p[f"{n}/kernel:0"], p[f"{n}/bias:0"], "SAME", act)

    t = conv(0, sd, "relu")
    i = 1
    for _ in range(10):
        x = conv(i, t, "relu")
        x = conv(i + 1, x, None)
        t = torch.relu(t + x)
        i += 2
    for _ in range(2):
        t = conv(i, _nn_up2_t(t), "relu")
        i += 1
    t = conv(i, t, "relu")
    t = conv(i + 1, t, None)
    return bq + t


def enet_generator_forward(params: dict, sd, bq, dtype=np.float64) -> np.ndarray:
    with torch.no_grad():
        return enet_generator_forward_t(_to_t(params, dtype), _t(sd, dtype), _t(bq, dtype)).numpy()


def enet_generator_grads(params: dict, sd, bq, dsr, dtype=np.float64):
    """Backward of the generator for a supplied upstream gradient d(sr) (BASELINE cfg5)."""
    p = _to_t(params, dtype, requires_grad=True)
    sr = enet_generator_forward_t(p, _t(sd, dtype), _t(bq, dtype))
    sr.backward(_t(dsr, dtype))
    return {k: v.grad.numpy() for k, v in p.items()}, sr.detach().numpy()


# ------------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------------


def synthetic_images(seed: int, n: int, h: int, w: int, c: int) -> np.ndarray:
    """Smooth-ish content in [-1,1]: clip(0.5 + 0.25*boxlowpass5(N(0,1))*k, 0, 1)*2-1."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((n, h + 4, w + 4, c))
    acc = np.zeros((n, h, w, c))
    for dy in range(5):
        for dx in range(5):
            acc += z[:, dy:dy + h, dx:dx + w]
    acc /= 5.0  # box sum / sqrt(25) keeps unit variance
    img = np.clip(0.5 + 0.25 * acc, 0.0, 1.0)
    return (img * 2.0 - 1.0).astype(np.float32)

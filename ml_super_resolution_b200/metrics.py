"""Evaluation post-pass on the device: the metrics and image hand-off the reference's evaluate / test /
resolve drivers compute with tf.image.psnr, tf.image.ssim, tf.image.rgb_to_yuv and tf.saturate_cast
(vdsr/vdsr/experiment_evaluate.py:57-60, espcn/espcn/experiment_test.py:30-53, vdsr/vdsr/experiment_resolve.py:65-69)."""
from __future__ import annotations

import torch

from . import _ffi, ops
from ._ffi import check


def _ws(n, device):
    return torch.empty(n, dtype=torch.float64, device=device)


def psnr(a: torch.Tensor, b: torch.Tensor, max_val: float) -> torch.Tensor:
    """tf.image.psnr: fp32 [N] for NHWC fp32 inputs."""
    a, b = ops._f32(a), ops._f32(b)
    n = a.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=a.device)
    check(_ffi.lib().srk_psnr(ops.handle(), ops._ptr(a), ops._ptr(b), n, a.numel() // n, float(max_val), ops._ptr(_ws(n, a.device)),
                              ops._ptr(out), ops._stream()), "srk_psnr")
    return out


def ssim(a: torch.Tensor, b: torch.Tensor, max_val: float) -> torch.Tensor:
    """tf.image.ssim: fp32 [N] for NHWC fp32 inputs (H, W >= 11)."""
    a, b = ops._f32(a), ops._f32(b)
    n, h, w, c = a.shape
    out = torch.empty(n, dtype=torch.float32, device=a.device)
    check(_ffi.lib().srk_ssim(ops.handle(), ops._ptr(a), ops._ptr(b), n, h, w, c, float(max_val), ops._ptr(_ws(n, a.device)),
                              ops._ptr(out), ops._stream()), "srk_ssim")
    return out


def rgb_to_y(x: torch.Tensor, scale=1.0, bias=0.0, clip=(-3.0e38, 3.0e38)) -> torch.Tensor:
    """Y channel of tf.image.rgb_to_yuv(clip(x * scale + bias)): [..., 3] -> [..., 1]."""
    x = ops._f32(x)
    assert x.shape[-1] == 3
    y = torch.empty(x.shape[:-1] + (1,), dtype=torch.float32, device=x.device)
    check(_ffi.lib().srk_rgb_to_y(ops.handle(), ops._ptr(x), x.numel() // 3, float(scale), float(bias), float(clip[0]), float(clip[1]),
                                  ops._ptr(y), ops._stream()), "srk_rgb_to_y")
    return y


def espcn_scores(sr_packed: torch.Tensor, hr_packed: torch.Tensor, scaling_factor: int, score_space: str = "y"):
    """espcn/espcn/experiment_test.py:30-53 -> (psnrs, ssims), max_val 1.0, in RGB or Y space of the PACKED tensors."""
    n, h, w, c = hr_packed.shape
    if score_space == "y":
        w2 = w * scaling_factor ** 2
        sr = rgb_to_y(sr_packed.reshape(-1, h, w2, 3), 0.5, 0.5, (0.0, 1.0))
        hr = rgb_to_y(hr_packed.reshape(-1, h, w2, 3), 0.5, 0.5, (0.0, 1.0))
    else:
        sr = torch.clamp(ops._f32(sr_packed) * 0.5 + 0.5, 0.0, 1.0)
        hr = torch.clamp(ops._f32(hr_packed) * 0.5 + 0.5, 0.0, 1.0)
    return psnr(hr, sr, 1.0), ssim(hr, sr, 1.0)


def saturate_cast_u8(x: torch.Tensor, scale: float = 127.5, bias: float = 127.5) -> torch.Tensor:
    """tf.saturate_cast(x * 127.5 + 127.5, tf.uint8): the image handed to the PNG encoder (vdsr/vdsr/experiment_resolve.py:65-67)."""
    x = ops._f32(x)
    y = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    check(_ffi.lib().srk_saturate_cast_u8(ops.handle(), ops._ptr(x), x.numel(), float(scale), float(bias), ops._ptr(y), ops._stream()),
          "srk_saturate_cast_u8")
    return y


def feature_mosaic_u8(feature_map: torch.Tensor) -> torch.Tensor:
    """`encode_feature_map` of vdsr/vdsr/experiment_feature_map_visualize.py:80-110 up to the PNG encoder: fp32 [1,H,W,64] (a
    `conv.N` / `relu.N` tap of VdsrNet.forward) -> uint8 [8H, 8W, 1], the 64 maps in an 8x8 grid."""
    x = ops._f32(feature_map)
    assert x.shape[0] == 1 and x.shape[3] == 64
    _, h, w, _ = x.shape
    y = torch.empty((8 * h, 8 * w, 1), dtype=torch.uint8, device=x.device)
    check(_ffi.lib().srk_feature_mosaic_u8(ops.handle(), ops._ptr(x), h, w, ops._ptr(y), ops._stream()), "srk_feature_mosaic_u8")
    return y

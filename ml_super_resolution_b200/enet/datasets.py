"""Device-resident EnhanceNet input pipeline -- drop-in for `image_batches` of enet/enet/datasets.py:78-127.

The reference crops a 128x128 patch per sample on the host, shrinks it with `scipy.misc.imresize(hd, 25)` (Pillow bilinear,
antialiased), blows it back up with `scipy.misc.imresize(sd, 400, 'bicubic')` (Pillow bicubic) and maps all three uint8 images
to [-1,1].  Here the decoded images live in one uint8 pool in HBM (`vdsr.dataset.DevicePool`) and a batch is
    srk_crop_u8 -> srk_resample_u8 (128->32 bilinear) -> srk_resample_u8 (32->128 bicubic) -> 3 x srk_u8_to_pm1
with Pillow's fixed-point coefficient tables computed once on the host (the arithmetic of libImaging/Resample.c
precompute_coeffs / normalize_coeffs_8bpc): results are bit-identical to Pillow.  The host keeps the reference's random-number
call sequence (`shuffle` per epoch; per sample `randint(128)` for x, then for y) on one numpy RandomState.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _ffi, ops
from .._ffi import check
from ..vdsr.dataset import DevicePool

_PRECISION_BITS = 32 - 8 - 2


def _bilinear(x):
    x = abs(x)
    return 1.0 - x if x < 1.0 else 0.0


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


_FILTERS = {"bilinear": (_bilinear, 1.0), "bicubic": (_bicubic, 2.0)}


def resample_tables(in_size: int, out_size: int, interp: str):
    """Pillow's coefficient tables for one axis: (ksize, bounds int32 [out,2] = (first input index, count), kk int32 [out,ksize])."""
    filt, fsupport = _FILTERS[interp]
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
        cnt = min(int(center + support + 0.5), in_size) - xmin
        w = [filt((x + xmin - center + 0.5) * ss) for x in range(cnt)]
        ww = 0.0
        for v in w:
            ww += v
        if ww != 0.0:
            w = [v / ww for v in w]
        bounds[xx] = (xmin, cnt)
        for x, v in enumerate(w):
            kk[xx, x] = int(-0.5 + v * (1 << _PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PRECISION_BITS))
    return ksize, bounds, kk


class _Resizer:
    """uint8 NHWC resize H x W -> out_h x out_w with Pillow's arithmetic; tables live on the device."""

    def __init__(self, H, W, out_h, out_w, interp, device):
        self.shape = (H, W, out_h, out_w)
        self.ksx, bx, kx = resample_tables(W, out_w, interp)
        self.ksy, by, ky = resample_tables(H, out_h, interp)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        self.kx, self.bx, self.ky, self.by = up(kx), up(bx), up(ky), up(by)

    def __call__(self, x: torch.Tensor, tmp: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        H, W, oh, ow = self.shape
        n, c = x.shape[0], x.shape[3]
        check(_ffi.lib().srk_resample_u8(ops.handle(), ops._ptr(x), n, H, W, c, oh, ow, ops._ptr(self.kx), ops._ptr(self.bx), self.ksx,
                                         ops._ptr(self.ky), ops._ptr(self.by), self.ksy, ops._ptr(tmp), ops._ptr(out), ops._stream()),
              "srk_resample_u8")
        return out


def draw_crops(shapes, batch_size, rng):
    """The reference's draws (enet/enet/datasets.py:82-110): endless generator of int32 [B,4] = (image, y, x, flip=0)."""
    order = list(range(len(shapes)))

    def indices():
        while True:
            rng.shuffle(order)
            for i in order:
                yield i

    it = indices()
    while True:
        crops = []
        for _ in range(batch_size):
            i = next(it)
            x = rng.randint(128)
            y = rng.randint(128)
            crops.append((i, y, x, 0))
        yield np.asarray(crops, np.int32)


class EnetBatcher:
    def __init__(self, images, batch_size=32, device="cuda"):
        self.pool = images if isinstance(images, DevicePool) else DevicePool(images, device)
        assert all(h >= 255 and w >= 255 for (h, w, _) in self.pool.shapes), "the reference crops 128x128 at offsets up to 127"
        self.B, dev, C = batch_size, self.pool.device, self.pool.C
        u8 = lambda *s: torch.empty(s, dtype=torch.uint8, device=dev)  # noqa: E731
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)  # noqa: E731
        self.hd_u8, self.sd_u8, self.bq_u8 = u8(batch_size, 128, 128, C), u8(batch_size, 32, 32, C), u8(batch_size, 128, 128, C)
        self.tmp_dn, self.tmp_up = u8(batch_size, 128, 32, C), u8(batch_size, 32, 128, C)
        self.out = [{"sd": f32(batch_size, 32, 32, C), "bq": f32(batch_size, 128, 128, C), "hd": f32(batch_size, 128, 128, C)} for _ in range(2)]
        self.down = _Resizer(128, 128, 32, 32, "bilinear", dev)   # scipy.misc.imresize(hd, 25): default interp is bilinear
        self.up = _Resizer(32, 32, 128, 128, "bicubic", dev)      # scipy.misc.imresize(sd, 400, 'bicubic')

    def make(self, crops: np.ndarray, slot: int = 0):
        lib, h, st = _ffi.lib(), ops.handle(), ops._stream()
        crops_d = torch.from_numpy(np.ascontiguousarray(crops, np.int32)).to(self.pool.device, non_blocking=True)
        check(lib.srk_crop_u8(h, ops._ptr(self.pool.pool), ops._ptr(self.pool.table), ops._ptr(crops_d), self.B, 128, self.pool.C,
                              ops._ptr(self.hd_u8), st), "srk_crop_u8")
        self.down(self.hd_u8, self.tmp_dn, self.sd_u8)
        self.up(self.sd_u8, self.tmp_up, self.bq_u8)
        o = self.out[slot]
        for src, key in ((self.sd_u8, "sd"), (self.bq_u8, "bq"), (self.hd_u8, "hd")):
            check(lib.srk_u8_to_pm1(h, ops._ptr(src), src.numel(), ops._ptr(o[key]), st), "srk_u8_to_pm1")
        return o["sd"], o["bq"], o["hd"]


def image_batches(images, scale_factor=4, batch_size=32, seed=None, rng=None, device="cuda"):
    """`image_batches(source_dir_path, scale_factor, batch_size)` of the reference with the directory replaced by its decoded
    uint8 images (the reference's body ignores scale_factor and always uses 128 -> 32 -> 128).  Yields device tensors
    (sd_images [B,32,32,3], bq_images [B,128,128,3], hd_images [B,128,128,3]) in [-1,1]; two output sets alternate so the
    previous batch stays valid while the next one is produced."""
    batcher = EnetBatcher(images, batch_size, device)
    rng = rng if rng is not None else np.random.RandomState(seed)
    slot = 0
    for crops in draw_crops(batcher.pool.shapes, batch_size, rng):
        yield batcher.make(crops, slot)
        slot ^= 1

"""Device-resident VDSR input pipeline -- drop-in for `image_batches` of vdsr/vdsr/dataset.py:41-128.

The reference's python generator reads one image per sample, crops / flips it with numpy, blurs and resizes it twice with
skimage (milliseconds per patch) and stacks a batch between two `session.run`s.  Here the decoded uint8 images are
uploaded ONCE into a single pool in HBM; a batch is then three small launches on a side stream
  srk_crop_flip_u8            random crop + flip + img_as_float32 (+ the [-1,1] target)      reference :98-107,115
  srk_degrade_gauss_bilinear  gaussian + bilinear down/up at the sample's scale               reference :13-38,110-111
  srk_affine_f32              sd * 2 - 1                                                      reference :114
and is prefetched while the previous step trains.  The host keeps the reference's random-number call sequence
(`shuffle` of the image order per epoch, then per sample `randint(w - S)`, `randint(h - S)`, `choice([0, 1])`,
`choice(scaling_factors)` on one `numpy.random.RandomState`), so a given seed yields the reference's batches.
Image decoding / directory listing (`skimage.io.imread`, `tf.gfile`) stay outside: the caller passes decoded arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _ffi, ops
from .._ffi import check


class _PoolImage(C.Structure):
    _fields_ = [("offset", C.c_int64), ("height", C.c_int32), ("width", C.c_int32)]


class _Crop(C.Structure):
    _fields_ = [("image", C.c_int32), ("y", C.c_int32), ("x", C.c_int32), ("flip", C.c_int32)]


class DevicePool:
    """uint8 HWC images concatenated in one device buffer + the table srk_crop_flip_u8 indexes."""

    def __init__(self, images, device="cuda"):
        self.shapes = [tuple(im.shape) for im in images]
        assert all(len(s) == 3 for s in self.shapes)
        self.C = self.shapes[0][2]
        table = (_PoolImage * len(images))()
        off = 0
        for i, im in enumerate(images):
            assert im.dtype == np.uint8 and im.shape[2] == self.C
            table[i] = _PoolImage(off, im.shape[0], im.shape[1])
            off += im.size
        host = np.empty(off, np.uint8)
        for t, im in zip(table, images):
            host[t.offset:t.offset + im.size] = np.ascontiguousarray(im).ravel()
        self.pool = torch.from_numpy(host).to(device)
        self.table = torch.from_numpy(np.frombuffer(bytes(table), dtype=np.uint8).copy()).to(device)
        self.device = device


def draw_samples(shapes, scaling_factors, image_size, batch_size, rng):
    """The reference's per-sample random draws (vdsr/vdsr/dataset.py:54-63,85-110), in its call order, as an endless
    generator of (crops int32 [B,4] = image, y, x, flip ; scales float32 [B])."""
    if scaling_factors is None or len(scaling_factors) <= 0:
        scaling_factors = [2.0, 3.0, 4.0]
    if any(s <= 1 for s in scaling_factors):
        raise Exception("invalide scaling factors")
    order = list(range(len(shapes)))

    def indices():
        while True:
            rng.shuffle(order)
            for i in order:
                yield i

    it = indices()
    crops, scales = [], []
    while True:
        i = next(it)
        h, w, c = shapes[i]
        if h < image_size or w < image_size or c != 3:
            continue
        x = rng.randint(w - image_size)
        y = rng.randint(h - image_size)
        flip = int(rng.choice([0, 1]))
        scale = float(rng.choice(scaling_factors))
        crops.append((i, y, x, flip))
        scales.append(scale)
        if len(crops) == batch_size:
            yield np.asarray(crops, np.int32), np.asarray(scales, np.float32)
            crops, scales = [], []


def make_batch(pool: DevicePool, crops: np.ndarray, scales: np.ndarray, image_size: int, out=None):
    """One batch on the current stream: returns (sd, hd) fp32 [B,S,S,3] in [-1,1] on the device."""
    B, S, Cc = len(crops), image_size, pool.C
    dev = pool.device
    if out is None:
        out = {k: torch.empty((B, S, S, Cc), dtype=torch.float32, device=dev) for k in ("hd01", "hd", "sd")}
    crops_d = torch.from_numpy(np.ascontiguousarray(crops, np.int32)).to(dev, non_blocking=True)
    scales_d = torch.from_numpy(np.ascontiguousarray(scales, np.float32)).to(dev, non_blocking=True)
    lib, h, st = _ffi.lib(), ops.handle(), ops._stream()
    check(lib.srk_crop_flip_u8(h, ops._ptr(pool.pool), ops._ptr(pool.table), ops._ptr(crops_d), B, S, Cc, ops._ptr(out["hd01"]),
                               ops._ptr(out["hd"]), st), "srk_crop_flip_u8")
    ops.degrade_gauss_bilinear(out["hd01"], scales_d, out=out["sd"])
    check(lib.srk_affine_f32(h, ops._ptr(out["sd"]), out["sd"].numel(), 2.0, -1.0, ops._ptr(out["sd"]), st), "srk_affine_f32")
    return out["sd"], out["hd"]


def image_batches(images, scaling_factors, image_size, batch_size, seed=None, rng=None, device="cuda", prefetch=True):
    """`image_batches(source_dir_path, scaling_factors, image_size, batch_size)` of the reference with the directory replaced
    by its decoded uint8 images.  Yields (sd_images, hd_images) device tensors; with `prefetch` the next batch is produced on
    a side stream while the caller consumes the current one (two buffer sets alternate)."""
    pool = images if isinstance(images, DevicePool) else DevicePool(images, device)
    rng = rng if rng is not None else np.random.RandomState(seed)
    draws = draw_samples(pool.shapes, scaling_factors, image_size, batch_size, rng)
    if not prefetch:
        for crops, scales in draws:
            yield make_batch(pool, crops, scales, image_size)
        return
    side = torch.cuda.Stream()
    bufs = [{k: torch.empty((batch_size, image_size, image_size, pool.C), dtype=torch.float32, device=pool.device) for k in ("hd01", "hd", "sd")}
            for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def produce(slot):
        crops, scales = next(draws)
        with torch.cuda.stream(side):
            side.wait_event(consumed[slot])  # the step that read this buffer set has been enqueued and finished
            make_batch(pool, crops, scales, image_size, out=bufs[slot])
            ready[slot].record(side)

    produce(0)
    slot = 0
    while True:
        produce(slot ^ 1)
        torch.cuda.current_stream().wait_event(ready[slot])
        yield bufs[slot]["sd"], bufs[slot]["hd"]
        consumed[slot].record(torch.cuda.current_stream())  # recorded when the consumer asks for the next batch
        slot ^= 1


def hd_image_to_sd_image(hd_image, scaling_factor):
    """vdsr/vdsr/dataset.py:13-38, same call: `hd_image` float [H,W,C] in [0,1] -> the blurred, bilinearly down- and up-scaled
    `sd_image` [H,W,C] (numpy in, numpy out); the arithmetic runs in srk_degrade_gauss_bilinear on the current device."""
    hd = np.ascontiguousarray(hd_image, dtype=np.float32)
    assert hd.ndim == 3
    x = torch.from_numpy(hd[None]).cuda()
    scales = torch.full((1,), float(scaling_factor), dtype=torch.float32, device=x.device)
    return ops.degrade_gauss_bilinear(x, scales)[0].cpu().numpy()

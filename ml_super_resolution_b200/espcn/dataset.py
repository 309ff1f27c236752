"""Reader / writer of the reference's ESPCN training patches -- the TensorFlow-free counterpart of `decode_patch_pair`,
`build_image_batch_iterator` and `write_patch` in espcn/espcn/dataset.py:10-75,160-195.

`build_image_batch_iterator(dir_path, batch_size, upscaling_factor)` yields the batches `EspcnNet.train_step` consumes:
lr [B,h,w,3] and hr already packed to [B,h,w,3*r^2] by `extract_image_patches` (reference :78-157, an offline numpy tool that
is not part of the hot path and is not re-implemented here).  With `device=` the batches are uploaded through pinned staging
buffers on a side stream while the previous step trains.
"""
from __future__ import annotations

import numpy as np
import torch

from ..io.tfrecord import decode_patch_pair as _decode, patch_batches, write_patch  # noqa: F401  (write_patch re-exported)


def decode_patch_pair(scaling_factor=3):
    """Reference signature: returns the decoder for one serialized record."""
    return lambda record: _decode(record, scaling_factor)


def build_image_batch_iterator(dir_path, batch_size=32, upscaling_factor=3, seed=None, device=None):
    """Endless iterator over shuffled `*.tfrecord` patch pairs.  device=None: numpy batches (feed them like the reference's
    placeholders); device='cuda': device tensors, double-buffered through pinned memory on a copy stream."""
    it = patch_batches(dir_path, batch_size, upscaling_factor, seed)
    if device is None:
        return it
    return _device_batches(it, device)


def _device_batches(it, device):
    side = torch.cuda.Stream()
    slots = [None, None]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(slot):
        lr, hr = next(it)
        if slots[slot] is None or slots[slot]["lr_h"].shape != lr.shape or slots[slot]["hr_h"].shape != hr.shape:
            slots[slot] = {"lr_h": torch.empty(lr.shape, dtype=torch.float32).pin_memory(), "hr_h": torch.empty(hr.shape, dtype=torch.float32).pin_memory(),
                           "lr": torch.empty(lr.shape, dtype=torch.float32, device=device), "hr": torch.empty(hr.shape, dtype=torch.float32, device=device)}
        s = slots[slot]
        consumed[slot].synchronize()  # the pinned staging buffer is rewritten by the host below
        s["lr_h"].copy_(torch.from_numpy(np.ascontiguousarray(lr)))
        s["hr_h"].copy_(torch.from_numpy(np.ascontiguousarray(hr)))
        with torch.cuda.stream(side):
            side.wait_event(consumed[slot])
            s["lr"].copy_(s["lr_h"], non_blocking=True)
            s["hr"].copy_(s["hr_h"], non_blocking=True)
            ready[slot].record(side)

    upload(0)
    slot = 0
    while True:
        upload(slot ^ 1)
        torch.cuda.current_stream().wait_event(ready[slot])
        yield slots[slot]["lr"], slots[slot]["hr"]
        consumed[slot].record(torch.cuda.current_stream())
        slot ^= 1

"""Flat fp32 parameter arena keyed by the reference's TensorFlow variable names.

All trainable variables of a model live in ONE contiguous fp32 device buffer (plus same-shaped
gradient / Adam-moment / decay-mask buffers) so the optimiser step, the gradient all-reduce and the
weight re-packing are each a single launch over the arena.  Checkpoints are `.npz` files keyed by the
TF variable names in HWIO layout (`conv2d/kernel:0`, `f1/bias:0`, ...), the layout
`tf.train.Saver`-restored variables have in the reference (checkpoint row).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch


class ParamArena:
    def __init__(self, params: "OrderedDict[str, np.ndarray]", device="cuda", decay_suffix: str | None = "kernel:0"):
        self.names = list(params.keys())
        self.shapes = {k: tuple(v.shape) for k, v in params.items()}
        self.offsets = {}
        off = 0
        for k, v in params.items():
            self.offsets[k] = off
            off += ((v.size + 3) // 4) * 4  # keep every variable 16-byte aligned
        self.size = off
        host = np.zeros(off, np.float32)
        mask = np.zeros(off, np.float32)
        for k, v in params.items():
            o = self.offsets[k]
            host[o:o + v.size] = np.asarray(v, np.float32).ravel()
            if decay_suffix is not None and k.endswith(decay_suffix):
                mask[o:o + v.size] = 1.0
        self.w = torch.from_numpy(host).to(device)
        self.decay_mask = torch.from_numpy(mask).to(device)
        self.g = None
        self.m = None
        self.v = None

    def enable_training(self) -> None:
        if self.g is None:
            self.g = torch.zeros_like(self.w)
            self.m = torch.zeros_like(self.w)
            self.v = torch.zeros_like(self.w)

    def view(self, name: str, which: str = "w") -> torch.Tensor:
        buf = getattr(self, which)
        o, shp = self.offsets[name], self.shapes[name]
        n = int(np.prod(shp))
        return buf[o:o + n].view(shp)

    def to_numpy(self, which: str = "w") -> "OrderedDict[str, np.ndarray]":
        host = getattr(self, which).detach().cpu().numpy()
        out = OrderedDict()
        for k in self.names:
            o, shp = self.offsets[k], self.shapes[k]
            out[k] = host[o:o + int(np.prod(shp))].reshape(shp).copy()
        return out

    def load_numpy(self, params: dict) -> None:
        host = self.w.detach().cpu().numpy().copy()
        for k in self.names:
            v = np.asarray(params[k], np.float32)
            assert tuple(v.shape) == self.shapes[k], (k, v.shape, self.shapes[k])
            o = self.offsets[k]
            host[o:o + v.size] = v.ravel()
        self.w.copy_(torch.from_numpy(host))

    def save(self, path: str, **extra) -> None:
        np.savez(path, **self.to_numpy(), **extra)


def load_params(ckpt_path: str) -> "OrderedDict[str, np.ndarray]":
    """Weights of a checkpoint as {TF tensor name ('f1/kernel:0', ...): ndarray}: either this package's `.npz` (ParamArena.save)
    or a TensorFlow Saver V2 checkpoint prefix (`model.ckpt-20000` with its `.index` / `.data-*` files), read without
    TensorFlow by io/tf_checkpoint.py."""
    import os
    if ckpt_path.endswith(".npz") or (os.path.exists(ckpt_path) and not os.path.exists(ckpt_path + ".index")):
        with np.load(ckpt_path) as z:
            return OrderedDict((k, z[k]) for k in z.files)
    if os.path.exists(ckpt_path + ".index"):
        from .io.tf_checkpoint import load_checkpoint, to_params
        return to_params(load_checkpoint(ckpt_path))
    raise FileNotFoundError(f"{ckpt_path}: neither an .npz file nor a TensorFlow checkpoint prefix")

"""SRCNN 9-1-5 on the B200 conv hot path -- drop-in for srcnn/srcnn.py of the reference.

The reference is a single script that reads a global `FLAGS` (srcnn/srcnn.py:8-23) and builds, in one graph,
the JPEG reader, the bicubic degrade, the three VALID convolutions, the crop, the loss and Adam
(`build_srcnn`, :81-166).  Here the same flag names live on `FLAGS`, `sanity_check()` is the same arithmetic
(:28-43, python-2 integer division), and `build_srcnn(hi_images)` returns the same dict keys
(`step, loss, trainer, hd_images, sd_images, sr_images`) as session handles.  The file reader / queue runners are
out of scope: `hi_images` is fed like any other placeholder.

  lo = resize_bicubic(resize_bicubic(hi, S/r), S)   srk_resize_bicubic_tf1 x2        reference :89-93
  patch_extraction     9x9 VALID C->64 ReLU         srk_conv_first_tc               reference :100-108
  non_linear_mapping   1x1 VALID 64->32 ReLU        srk_conv_tc (k=1, N=32)         reference :111-119
  reconstruction       5x5 VALID 32->C tanh         srk_conv_tc_last (k=5, crop)    reference :122-130
  loss                 mean_rows ||sr - hi||_2      srk_l2norm_rows_mean_fwd_bwd    reference :142-144
"""
from __future__ import annotations

from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import torch

from .. import ops
from ..initializers import srcnn_params
from ..params import ParamArena
from ..session import Handle

FLAGS = SimpleNamespace(ckpt_dir_path="./ckpts/", logs_dir_path="./logs/", training_images_path=None, sr_source_path=None,
                        sr_target_path=None, train=False, batch_size=64, upscaling_factor=3, crop_image_size=256, crop_image_side=6,
                        srcnn_fsub=33, srcnn_f1=9, srcnn_f2=1, srcnn_f3=5, srcnn_n1=64, srcnn_n2=32)


def sanity_check(flags=FLAGS):
    """srcnn/srcnn.py:28-43 (py2 `/` on ints is floor division)."""
    smaller_output_size = flags.srcnn_fsub - flags.srcnn_f1 - flags.srcnn_f2 - flags.srcnn_f3 + 3
    boundary = (flags.srcnn_fsub - smaller_output_size) // 2
    crop_size = (flags.crop_image_size - boundary * 2) // smaller_output_size
    flags.crop_image_side = boundary
    flags.crop_image_size = crop_size * smaller_output_size + boundary * 2
    if not flags.train:
        flags.batch_size = 1


class SrcnnNet:
    NAMES = ("patch_extraction", "non_linear_mapping", "reconstruction")

    def __init__(self, params=None, channels=3, device="cuda", seed=0, flags=FLAGS):
        f = (flags.srcnn_f1, flags.srcnn_f2, flags.srcnn_f3)
        assert f == (9, 1, 5) and (flags.srcnn_n1, flags.srcnn_n2) == (64, 32), "kernels are instantiated for the 9-1-5 / 64-32 SRCNN"
        if params is None:
            params = srcnn_params(seed, channels, f, (64, 32))
        order = OrderedDict()
        for n in self.NAMES:
            order[f"{n}/weights:0"] = np.asarray(params[f"{n}/weights:0"], np.float32)
            order[f"{n}/biases:0"] = np.asarray(params[f"{n}/biases:0"], np.float32)
        self.C = order["patch_extraction/weights:0"].shape[2]
        self.device = device
        self.r = flags.upscaling_factor
        self.arena = ParamArena(order, device, decay_suffix=None)
        a = self.arena
        plan = ops.PackPlan(device)
        self._i1 = plan.add(a.offsets["patch_extraction/weights:0"], 9, self.C, 64, ops.PACK_FIRST)
        self._i2 = plan.add(a.offsets["non_linear_mapping/weights:0"], 1, 64, 32, ops.PACK_FWD, 32, 64)
        self._i3 = plan.add(a.offsets["reconstruction/weights:0"], 5, 32, self.C, ops.PACK_FWD, 16, 32)
        plan.finalize()
        self.plan = plan
        self.bias3 = torch.zeros(16, dtype=torch.float32, device=device)
        self.repack()
        self._panels = {}

    def repack(self):
        self.plan.run(self.arena.w)
        self.bias3[: self.C].copy_(self.arena.view("reconstruction/biases:0"))

    def degrade(self, hi: torch.Tensor) -> torch.Tensor:
        """In-graph bicubic down then up, TF1 legacy kernel (reference :89-93; sizes are py2 integer divisions)."""
        n, h, w, c = hi.shape
        lo = ops.resize_bicubic_tf1(hi, h // self.r, w // self.r)
        return ops.resize_bicubic_tf1(lo, h, w)

    def forward(self, lo: torch.Tensor) -> torch.Tensor:
        """lo fp32 [N,S,S,C] -> sr [N,S-12,S-12,C] (three VALID convolutions)."""
        n, H, W, C = lo.shape
        a = self.arena
        t1 = ops.conv_first_tc(lo, self.plan.views[self._i1], a.view("patch_extraction/biases:0"), 9, "VALID", "relu")
        t2 = ops.conv_tc(t1, self.plan.views[self._i2], a.view("non_linear_mapping/biases:0"), 1, "relu")
        h2, w2 = t2.H, t2.W
        key = (n, h2, w2)
        if key not in self._panels:  # VALID 5x5 == SAME over the FPA, keeping only outputs 2 px inside the border
            self._panels[key] = ops.make_panels([(i, -2, -2, 2, h2 - 2, 2, w2 - 2) for i in range(n)], self.device)
        return ops.conv_tc_last(t2, self.plan.views[self._i3], self.bias3, 5, C, "tanh", panels=self._panels[key],
                                frame_shape=(n, h2 - 4, w2 - 4))

    def loss(self, sr: torch.Tensor, hi: torch.Tensor):
        """mean over rows of ||reshape(sr - crop(hi), [-1, bb^2])||_2 (reference :132-144); returns (loss, dloss/dsr)."""
        bb = sr.shape[1]
        side = (hi.shape[1] - bb) // 2
        hic = hi[:, side:side + bb, side:side + bb, :].contiguous()
        acc = torch.zeros(1, device=self.device)
        dsr = torch.empty_like(sr)
        ops.l2norm_rows_mean_fwd_bwd(sr, hic, bb * bb, acc, dsr)
        return acc, dsr


class _SrcnnGraph:
    def __init__(self, net, hi_ph):
        self.net, self.hi_ph = net, hi_ph

    def execute(self, keys, feeds):
        net = self.net
        if "trainer" in keys:
            raise NotImplementedError("SRCNN training is a 'next' row (DESIGN.md section 7); forward, degrade and loss run on the GPU path")
        x = feeds[self.hi_ph]
        hi = (x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))).to(net.device).contiguous()
        lo = net.degrade(hi)
        sr = net.forward(lo)
        side = (hi.shape[1] - sr.shape[1]) // 2
        crop = lambda t: t[:, side:side + sr.shape[1], side:side + sr.shape[1], :]
        out = {"step": 0, "sr_images": sr.cpu().numpy(), "hd_images": crop(hi).cpu().numpy(), "sd_images": crop(lo).cpu().numpy()}
        if "loss" in keys:
            out["loss"] = float(net.loss(sr, hi)[0])
        return out


def build_srcnn(hi_images=None, params=None, channels=3, device="cuda", seed=0, flags=FLAGS):
    """srcnn/srcnn.py:81 `build_srcnn()`; `hi_images` replaces the in-graph dataset reader (:86)."""
    net = SrcnnNet(params, channels, device, seed, flags)
    g = _SrcnnGraph(net, hi_images)
    return {k: Handle(g, k) for k in ("step", "loss", "trainer", "hd_images", "sd_images", "sr_images")}

"""EnhanceNet generator on the B200 conv hot path -- drop-in for `build_generator` / `residual_block` of
enet/enet/model_enet.py (the discriminator / VGG / texture losses of `build_enet` are outside the hot path,
SURVEY section 2.1 #10-11).

  conv2d            3x3 3->64 ReLU                      srk_conv_first_tc                    reference :63-70
  10 x residual     3x3 64->64 ReLU ; 1x1 64->64 ;      srk_conv_tc ; srk_conv_tc(k=1) with  reference :8-31
                    ReLU(x + y)                         the block residual + ReLU fused in the epilogue
  2 x upsample      NN x2 ; 3x3 64->64 ReLU             srk_fpa_upsample2 ; srk_conv_tc      reference :77-89
  conv2d_23         3x3 64->64 ReLU                     srk_conv_tc                          reference :92-99
  conv2d_24         3x3 64->3, + bq_images              srk_conv_tc_last (fused addend)      reference :102-113
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .. import ops
from ..initializers import ENET_G_LAYERS, enet_g_params, tf_conv_name
from ..params import ParamArena
from ..session import Handle
from ..tiling import MAX_PANEL_W


class EnetGenerator:
    def __init__(self, params=None, scope_name="g_", device="cuda", seed=0):
        self.scope = scope_name
        if params is None:
            params = enet_g_params(seed, scope_name)
        order = OrderedDict()
        for i in range(len(ENET_G_LAYERS)):
            for suf in ("kernel:0", "bias:0"):
                k = f"{scope_name}/{tf_conv_name(i)}/{suf}"
                order[k] = np.asarray(params[k], np.float32)
        self.device = device
        self.arena = ParamArena(order, device, decay_suffix=None)
        a = self.arena
        plan = ops.PackPlan(device)
        self._idx = {}
        for i, (k, cin, cout) in enumerate(ENET_G_LAYERS):
            off = a.offsets[self._k(i)]
            if i == 0:
                self._idx[i] = plan.add(off, 3, cin, 64, ops.PACK_FIRST)
            elif cout == 64:
                self._idx[i] = plan.add(off, k, 64, 64, ops.PACK_FWD, 64, 64)
            else:
                self._idx[i] = plan.add(off, 3, 64, cout, ops.PACK_FWD, 16, 64)
        plan.finalize()
        self.plan = plan
        self.bias_last = torch.zeros(16, dtype=torch.float32, device=device)
        self.repack()

    def _k(self, i):
        return f"{self.scope}/{tf_conv_name(i)}/kernel:0"

    def _b(self, i):
        return f"{self.scope}/{tf_conv_name(i)}/bias:0"

    def repack(self):
        self.plan.run(self.arena.w)
        self.bias_last[:3].copy_(self.arena.view(self._b(len(ENET_G_LAYERS) - 1)))

    def forward(self, sd: torch.Tensor, bq: torch.Tensor) -> torch.Tensor:
        """sd fp32 [N,h,w,3], bq fp32 [N,4h,4w,3] -> sr = bq + generator(sd)."""
        n, h, w, _ = sd.shape
        assert 4 * w <= MAX_PANEL_W, "EnhanceNet frames wider than 63 LR px need tiling (13-px LR halo) -- not wired up yet"
        a, W = self.arena, self.plan.views
        t = ops.conv_first_tc(sd, W[self._idx[0]], a.view(self._b(0)), 3, "SAME", "relu")
        i = 1
        for _ in range(10):
            x = ops.conv_tc(t, W[self._idx[i]], a.view(self._b(i)), 3, "relu")
            t = ops.conv_tc(x, W[self._idx[i + 1]], a.view(self._b(i + 1)), 1, None, addend=t, relu_after_add=True)
            i += 2
        for _ in range(2):
            t = ops.conv_tc(ops.fpa_upsample2(t), W[self._idx[i]], a.view(self._b(i)), 3, "relu")
            i += 1
        t = ops.conv_tc(t, W[self._idx[i]], a.view(self._b(i)), 3, "relu")
        return ops.conv_tc_last(t, W[self._idx[i + 1]], self.bias_last, 3, 3, None, addend=bq)


class _EnetGraph:
    def __init__(self, net, sd_ph, bq_ph):
        self.net, self.sd_ph, self.bq_ph = net, sd_ph, bq_ph

    def execute(self, keys, feeds):
        dev = self.net.device
        to = lambda x: (x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))).to(dev).contiguous()
        return {"sr_images": self.net.forward(to(feeds[self.sd_ph]), to(feeds[self.bq_ph])).cpu().numpy()}


def build_generator(sd_images, bq_images, hd_images=None, scope_name="g_", params=None, device="cuda", seed=0):
    """enet/enet/model_enet.py:44 `build_generator(sd_images, bq_images, hd_images, scope_name)` -> sr_images handle."""
    net = EnetGenerator(params, scope_name, device, seed)
    return Handle(_EnetGraph(net, sd_images, bq_images), "sr_images")

"""Small driver for ncu: a few launches of the 64->64 conv_tc forward (and optionally dgrad / wgrad) on one panel group."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_super_resolution_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
form = sys.argv[3] if len(sys.argv) > 3 else "auto"  # auto | flat | strip: kernel form of the plain 64->64 layer
n, h, w = (4, 540, 252) if len(sys.argv) <= 2 else tuple(int(v) for v in sys.argv[2].split("x"))
g = torch.Generator(device="cuda").manual_seed(0)
x = ops.fpa_empty(n, h, w, 64)
x.data.normal_(generator=g)
y = ops.fpa_empty(n, h, w, 64)
wp = ops.pack_conv_weights(torch.randn((3, 3, 64, 64), device="cuda", generator=g) / 24)
b = torch.zeros(64, device="cuda")
dw = torch.zeros((3, 3, 64, 64), device="cuda")
db = torch.zeros(64, device="cuda")
for _ in range(4):
    if which == "fwd":
        with ops.conv_form(form):
            ops.conv_tc(x, wp, b, 3, "relu", out=y)
    elif which == "dgrad":
        ops.conv_tc(x, wp, None, 3, None, out=y, mask_src=x, mask_kind="relu")
    else:
        ops.conv_wgrad_tc(x, y, dw, db)
torch.cuda.synchronize()
print("ok")

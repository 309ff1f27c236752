"""Randomised shape sweep of the tensor-core conv kernels against torch's conv2d on the same bf16-rounded operands
(development tool; run on the GPU box).  Exercises tile counts from 1 to hundreds per CTA, narrow / wide panels and the
barrier round-robins of conv_tc_kernel, conv_gather_tc_kernel and wgrad_tc_kernel."""
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, "/root/repo")
from ml_super_resolution_b200 import ops  # noqa: E402

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n_cases = int(sys.argv[2]) if len(sys.argv) > 2 else 60
g = torch.Generator(device="cuda").manual_seed(1)
worst = {}


def fpa_from(x_nhwc):
    return ops.fpa_from_nhwc(x_nhwc, x_nhwc.shape[-1])


def ref_conv(x_nhwc, w_hwio, b, k):
    xr = x_nhwc.to(torch.bfloat16).float().permute(0, 3, 1, 2)
    wr = w_hwio.to(torch.bfloat16).float().permute(3, 2, 0, 1)
    return (F.conv2d(xr, wr, b, padding=k // 2)).permute(0, 2, 3, 1)


for case in range(n_cases):
    kind = case % 5
    n = int(rng.integers(1, 6))
    H = int(rng.integers(3, 140)) if case % 7 else int(rng.integers(500, 1500))
    W = int(rng.integers(3, 254))
    if kind == 0:  # 64 -> 64 3x3 relu
        cin, cout, k, act = 64, 64, 3, "relu"
    elif kind == 1:  # 64 -> 32 3x3 tanh
        cin, cout, k, act = 64, 32, 3, "tanh"
    elif kind == 2:  # 64 -> 64 1x1
        cin, cout, k, act = 64, 64, 1, None
    elif kind == 3:  # last layer 64 -> 3 with residual
        cin, cout, k, act = 64, 3, 3, None
    else:  # first layer 3 -> 64 (gather kernel) + wgrad of a 64 -> 64 layer
        cin, cout, k, act = 3, 64, 3, "relu"
    x = torch.randn((n, H, W, cin), device="cuda", generator=g)
    w = torch.randn((k, k, cin, cout), device="cuda", generator=g) * (0.5 / (k * np.sqrt(cin)))
    b = torch.randn(cout, device="cuda", generator=g) * 0.1
    ref = ref_conv(x, w, b, k)
    if act == "relu":
        ref = torch.relu(ref)
    elif act == "tanh":
        ref = torch.tanh(ref)
    if kind in (0, 1, 2):
        np_ = 64 if cout == 64 else 32
        got = ops.fpa_to_nhwc(ops.conv_tc(fpa_from(x), ops.pack_conv_weights(w, ops.PACK_FWD, np_, 64), ops.pad_bias(b, np_), k, act))[..., :cout]
        tol = 2e-2
    elif kind == 3:
        add = torch.randn((n, H, W, cout), device="cuda", generator=g)
        got = ops.conv_tc_last(fpa_from(x), ops.pack_conv_weights(w, ops.PACK_FWD, 16, 64), ops.pad_bias(b, 16), k, cout, None, addend=add)
        ref = ref + add
        tol = 2e-3
    else:
        got = ops.fpa_to_nhwc(ops.conv_first_tc(x, ops.pack_first_weights(w), b, k, "SAME", act))
        tol = 2e-2
        # wgrad: dW = X^T dY over the same geometry
        xa = torch.randn((n, H, W, 64), device="cuda", generator=g)
        dy = torch.randn((n, H, W, 64), device="cuda", generator=g) * 0.1
        dw = torch.empty((3, 3, 64, 64), device="cuda")
        db = torch.empty(64, device="cuda")
        ops.conv_wgrad_tc(fpa_from(xa), fpa_from(dy), dw, db)
        xr = xa.to(torch.bfloat16).float().permute(3, 0, 1, 2)          # [ci, n, H, W] as a batch of ci "images"
        dr = dy.to(torch.bfloat16).float().permute(3, 0, 1, 2)          # [co, n, H, W] as co filters
        rw = F.conv2d(xr, dr, padding=1).permute(2, 3, 0, 1)            # [3, 3, ci, co]
        e = float((dw - rw).abs().max() / (rw.abs().max() + 1e-9))
        worst["wgrad"] = max(worst.get("wgrad", 0.0), e)
        assert e < 5e-3, ("wgrad", n, H, W, e)
        assert float((db - dr.sum(dim=(1, 2, 3))).abs().max()) < 1e-2 * max(1.0, float(dr.abs().sum(dim=(1, 2, 3)).max()) * 1e-2)
    e = float((got - ref).abs().max())
    name = ["c64", "c32", "1x1", "last", "first"][kind]
    worst[name] = max(worst.get(name, 0.0), e)
    assert e <= tol * max(1.0, float(ref.abs().max())), (name, n, H, W, e)
torch.cuda.synchronize()
print("stress OK:", n_cases, "cases; worst abs errors", {k: round(v, 5) for k, v in worst.items()})

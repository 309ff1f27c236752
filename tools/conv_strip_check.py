"""Development check of the column-strip 64->64 convolution (csrc/conv_strip.cu) on the GPU box: against the flat-stream kernel
(conv_tc.cu, selected with SRK_NO_STRIP=1) and an fp32 torch reference, then timing of both on 4K panels."""
import os
import sys

import torch

sys.path.insert(0, "/root/repo")
from ml_super_resolution_b200 import ops  # noqa: E402


def run(x, wp, b, act, strip):
    if strip:
        os.environ.pop("SRK_NO_STRIP", None)
    else:
        os.environ["SRK_NO_STRIP"] = "1"
    y = ops.fpa_empty(x.n_img, x.H, x.W, 64)
    y.data.fill_(float("nan"))
    ops.conv_tc(x, wp, b, 3, act, out=y)
    return y


ok = True
g = torch.Generator(device="cuda").manual_seed(0)
for (n, h, w) in [(1, 9, 126), (2, 20, 130), (1, 33, 242), (3, 17, 253), (2, 64, 400), (1, 300, 1000)]:
    xin = torch.randn((n, h, w, 64), device="cuda", generator=g)
    wt = torch.randn((3, 3, 64, 64), device="cuda", generator=g) / 24
    b = torch.randn(64, device="cuda", generator=g) * 0.1
    x = ops.fpa_from_nhwc(xin)
    wp = ops.pack_conv_weights(wt)
    for act in ("relu", None):
        ys = run(x, wp, b, act, True)
        a = ops.fpa_to_nhwc(ys)
        xb = xin.to(torch.bfloat16).float().permute(0, 3, 1, 2)
        ref = torch.nn.functional.conv2d(xb, wt.to(torch.bfloat16).float().permute(3, 2, 0, 1), b, padding=1).permute(0, 2, 3, 1)
        if act == "relu":
            ref = ref.relu()
        e_s = (a - ref).abs().max().item()
        nv = n * (h + 1) * (w + 1)
        raw_s = ys.data[:nv].float().view(n, h + 1, w + 1, 64)
        pads_ok = bool(torch.isfinite(raw_s).all()) and bool((raw_s[:, 0] == 0).all()) and bool((raw_s[:, :, w] == 0).all())
        msg = ""
        if w <= 254:
            yo = run(x, wp, b, act, False)
            o = ops.fpa_to_nhwc(yo)
            msg = f"flat {(o - ref).abs().max().item():.3e}  strip-vs-flat {(a - o).abs().max().item():.3e}"
        good = e_s <= 3e-2 and pads_ok
        ok &= good
        print(f"{(n, h, w)} act={act}: strip max|err| {e_s:.3e}  {msg}  pads {pads_ok}  {'OK' if good else 'FAIL'}", flush=True)
os.environ.pop("SRK_NO_STRIP", None)
for shape in [(16, 270, 242), (16, 2160, 242), (1, 2160, 3840)]:
    n, h, w = shape
    x = ops.fpa_empty(n, h, w, 64)
    x.data.normal_(generator=g)
    wp = ops.pack_conv_weights(torch.randn((3, 3, 64, 64), device="cuda", generator=g) / 24)
    b = torch.zeros(64, device="cuda")
    for strip in (True, False):
        if not strip and w > 254:
            continue
        y = run(x, wp, b, "relu", strip)
        for _ in range(3):
            ops.conv_tc(x, wp, b, 3, "relu", out=y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            ops.conv_tc(x, wp, b, 3, "relu", out=y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{shape} {'strip' if strip else 'flat '}: {ms:.3f} ms  {2 * 576 * 64 * n * h * w / ms / 1e9:.0f} TFLOP/s", flush=True)
sys.exit(0 if ok else 1)

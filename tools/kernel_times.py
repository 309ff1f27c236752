"""Per-kernel CUDA-event timings on the BASELINE shapes (development aid; bench.py is the contract)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ml_super_resolution_b200 import ops  # noqa: E402

FLUSH = None


def timeit(fn, iters=20, warm=3, flush=True):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    res = {}
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    # ---- 64->64 3x3 layer on the VDSR training shape and on one inference panel group
    for name, (n, h, w) in {"train_64x41x41": (64, 41, 41), "train_64x128x128": (64, 128, 128), "panel_18x2160x252": (18, 2160, 252),
                            "panel_4x540x252": (4, 540, 252)}.items():
        x = ops.fpa_empty(n, h, w, 64)
        x.data.normal_(generator=g)
        y = ops.fpa_empty(n, h, w, 64)
        w_hwio = torch.randn((3, 3, 64, 64), device=dev, generator=g) / 24
        wp = ops.pack_conv_weights(w_hwio)
        b = torch.zeros(64, device=dev)
        rows = n * (h + 1) * (w + 1)
        flops = 2.0 * n * h * w * 576 * 64
        med, mn = timeit(lambda: ops.conv_tc(x, wp, b, 3, "relu", out=y))
        res[f"conv_tc_fwd/{name}"] = dict(ms=med, ms_min=mn, tflops=flops / med / 1e9, gbs=rows * 256 / med / 1e6)
        med, mn = timeit(lambda: ops.conv_tc(x, wp, None, 3, None, out=y, mask_src=x, mask_kind="relu"))
        res[f"conv_tc_dgrad/{name}"] = dict(ms=med, ms_min=mn, tflops=flops / med / 1e9)
        if n * h * w < 3e6:
            dw = torch.zeros((3, 3, 64, 64), device=dev)
            db = torch.zeros(64, device=dev)
            med, mn = timeit(lambda: ops.conv_wgrad_tc(x, y, dw, db))
            res[f"wgrad_tc/{name}"] = dict(ms=med, ms_min=mn, tflops=flops / med / 1e9)
        del x, y
    # ---- ESPCN 1080p LR Y, per layer
    from ml_super_resolution_b200.espcn.model_espcn import EspcnNet
    for C in (1, 3):
        net = EspcnNet(None, 3, C)
        lr = torch.rand((1, 1080, 1920, C), device=dev, generator=g) * 2 - 1
        out = torch.empty((1, 3240, 5760, C), device=dev)
        med, mn = timeit(lambda: net.forward(lr, True, out=out))
        res[f"espcn_1080p_C{C}/total"] = dict(ms=med, ms_min=mn, out_mpix_s=3240 * 5760 / med / 1e3)
        from ml_super_resolution_b200.tiling import plan_tiles
        Ht, Wt, tiles = plan_tiles(1, 1080, 1920, 4)
        panels = ops.make_panels([t.as_tuple() for t in tiles])
        t1, t2 = net._get_bufs(len(tiles), Ht, Wt)
        a = net.arena
        med, mn = timeit(lambda: ops.conv_first_tc(lr, net.plan.views[net._i1], a.view("f1/bias:0"), 5, "SAME", "tanh", panels=panels,
                                                    panel_hw=(Ht, Wt), out=t1))
        res[f"espcn_1080p_C{C}/f1_conv_first"] = dict(ms=med, ms_min=mn)
        med, mn = timeit(lambda: ops.conv_tc(t1, net.plan.views[net._i2], a.view("f2/bias:0"), 3, "tanh", out=t2))
        res[f"espcn_1080p_C{C}/f2_conv_tc"] = dict(ms=med, ms_min=mn)
        med, mn = timeit(lambda: ops.conv_tc_last(t2, net.plan.views[net._i3], net.bias3, 3, net.cout3, None, shuffle_r=3, panels=panels,
                                                   frame_shape=(1, 1080, 1920), out=out))
        res[f"espcn_1080p_C{C}/f3_conv_tc_last"] = dict(ms=med, ms_min=mn)
        del net, out
    # ---- VDSR: training step and 4K inference
    from ml_super_resolution_b200.vdsr.model_vdsr import VdsrNet
    net = VdsrNet(None, 20, 3)
    sd = torch.rand((64, 41, 41, 3), device=dev, generator=g) * 2 - 1
    hd = torch.rand((64, 41, 41, 3), device=dev, generator=g) * 2 - 1
    med, mn = timeit(lambda: net.train_step(sd, hd, 1e-4), iters=10, flush=False)
    res["vdsr_train_step_64x41x41"] = dict(ms=med, ms_min=mn, patches_s=64 / med * 1e3)
    med, mn = timeit(lambda: net.forward_backward(sd, hd), iters=10, flush=False)
    res["vdsr_fwd_bwd_64x41x41"] = dict(ms=med, ms_min=mn)
    med, mn = timeit(lambda: net.forward(sd), iters=10, flush=False)
    res["vdsr_fwd_64x41x41"] = dict(ms=med, ms_min=mn)
    frame = torch.rand((1, 2160, 3840, 3), device=dev, generator=g) * 2 - 1
    out = torch.empty_like(frame)
    med, mn = timeit(lambda: net.forward(frame, out=out), iters=5, warm=2, flush=False)
    res["vdsr_infer_4k"] = dict(ms=med, ms_min=mn, mpix_s=2160 * 3840 / med / 1e3)
    for tr, gb in ((540, 64 << 20),):
        net.tile_group_bytes = gb
        med, mn = timeit(lambda: net.forward(frame, out=out, tile_rows=tr), iters=5, warm=2, flush=False)
        res[f"vdsr_infer_4k_tilerows{tr}_group{gb >> 20}MB"] = dict(ms=med, ms_min=mn, mpix_s=2160 * 3840 / med / 1e3)
    for k, v in res.items():
        print(k, json.dumps({a: round(b, 4) for a, b in v.items()}))


if __name__ == "__main__":
    main()

"""Development check of the column-strip 64->64 convolution (csrc/conv_strip.cu) on the GPU box: against the flat-stream kernel
(conv_tc.cu) and an fp32 torch reference -- wide frames, several narrow images per tile, the ReLU' mask form, layer chains -- then
timing of both forms."""
import sys

import torch

sys.path.insert(0, "/root/repo")
from ml_super_resolution_b200 import ops  # noqa: E402


def run(x, wp, b, act, form, mask=None):
    y = ops.fpa_empty(x.n_img, x.H, x.W, 64)
    y.data.fill_(float("nan"))
    with ops.conv_form(form):
        ops.conv_tc(x, wp, b, 3, act, out=y, mask_src=mask, mask_kind="relu" if mask is not None else None)
    return y


ok = True
g = torch.Generator(device="cuda").manual_seed(0)
for (n, h, w) in [(1, 9, 126), (2, 20, 130), (1, 33, 242), (3, 17, 253), (2, 64, 400), (8, 41, 41), (64, 41, 41), (7, 20, 24), (5, 3, 62), (1, 1, 1)]:
    xin = torch.randn((n, h, w, 64), device="cuda", generator=g)
    wt = torch.randn((3, 3, 64, 64), device="cuda", generator=g) / 24
    b = torch.randn(64, device="cuda", generator=g) * 0.1
    msk = ops.fpa_from_nhwc(torch.randn((n, h, w, 64), device="cuda", generator=g))
    x = ops.fpa_from_nhwc(xin)
    wp = ops.pack_conv_weights(wt)
    for act, mask in (("relu", None), (None, None), (None, msk)):
        ys = run(x, wp, b if mask is None else None, act, "strip", mask)
        a = ops.fpa_to_nhwc(ys)
        xb = xin.to(torch.bfloat16).float().permute(0, 3, 1, 2)
        ref = torch.nn.functional.conv2d(xb, wt.to(torch.bfloat16).float().permute(3, 2, 0, 1), b if mask is None else None, padding=1).permute(0, 2, 3, 1)
        if act == "relu":
            ref = ref.relu()
        if mask is not None:
            ref = ref * (ops.fpa_to_nhwc(msk) > 0)
        e_s = (a - ref).abs().max().item()
        nv = n * (h + 1) * (w + 1)
        raw_s = ys.data[:nv].float().view(n, h + 1, w + 1, 64)
        pads_ok = bool(torch.isfinite(raw_s).all()) and bool((raw_s[:, 0] == 0).all()) and bool((raw_s[:, :, w] == 0).all())
        msg = ""
        if w <= 254:
            o = ops.fpa_to_nhwc(run(x, wp, b if mask is None else None, act, "flat", mask))
            msg = f"flat {(o - ref).abs().max().item():.3e}  strip-vs-flat {(a - o).abs().max().item():.3e}"
        good = e_s <= 3e-2 and pads_ok
        ok &= good
        print(f"{(n, h, w)} act={act} mask={mask is not None}: strip max|err| {e_s:.3e}  {msg}  pads {pads_ok}  {'OK' if good else 'FAIL'}", flush=True)

# chains: n layers in one launch == the same layers launched one by one (bit for bit), forward (bias + relu) and masked
for (n, h, w, L) in [(64, 41, 41, 6), (2, 50, 300, 5), (16, 41, 41, 18)]:
    ws = [ops.pack_conv_weights(torch.randn((3, 3, 64, 64), device="cuda", generator=g) / 24) for _ in range(L)]
    bs = [torch.randn(64, device="cuda", generator=g) * 0.1 for _ in range(L)]
    x0 = ops.fpa_from_nhwc(torch.randn((n, h, w, 64), device="cuda", generator=g))
    masks = [ops.fpa_from_nhwc(torch.randn((n, h, w, 64), device="cuda", generator=g)) for _ in range(L)]
    for use_mask in (False, True):
        seq = [x0]
        for l in range(L):
            seq.append(run(seq[-1], ws[l], None if use_mask else bs[l], None if use_mask else "relu", "strip", masks[l] if use_mask else None))
        bufs = [ops.fpa_empty(n, h, w, 64) for _ in range(L)]
        for bq in bufs:
            bq.data.fill_(float("nan"))
        chain = ops.ConvChain([x0] + bufs[:-1], ws, [None] * L if use_mask else bs, [None] * L if use_mask else ["relu"] * L, bufs, masks if use_mask else None)
        for _ in range(3):  # (repeat: the barrier word is re-armed by every call)
            chain.run()
        torch.cuda.synchronize()
        nv = n * (h + 1) * (w + 1)
        same = all(torch.equal(bufs[l].data[:nv], seq[l + 1].data[:nv]) for l in range(L))
        ok &= same
        print(f"chain {(n, h, w)} x{L} mask={use_mask}: equal to layer-by-layer {same}  {'OK' if same else 'FAIL'}", flush=True)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for shape in [(16, 2160, 242), (1, 2160, 3840), (64, 41, 41)]:
    n, h, w = shape
    x = ops.fpa_empty(n, h, w, 64)
    x.data.normal_(generator=g)
    wp = ops.pack_conv_weights(torch.randn((3, 3, 64, 64), device="cuda", generator=g) / 24)
    b = torch.zeros(64, device="cuda")
    y = ops.fpa_empty(n, h, w, 64)
    for form in ("strip", "flat"):
        if form == "flat" and w > 254:
            continue

        def one():
            with ops.conv_form(form):
                ops.conv_tc(x, wp, b, 3, "relu", out=y)
        ms = timeit(one)
        print(f"{shape} {form:5s}: {ms * 1000:.1f} us  {2 * 576 * 64 * n * h * w / ms / 1e9:.0f} TFLOP/s", flush=True)
# the training shape: 18 layers one by one vs one chain
n, h, w, L = 64, 41, 41, 18
ws = [ops.pack_conv_weights(torch.randn((3, 3, 64, 64), device="cuda", generator=g) / 24) for _ in range(L)]
bs = [torch.zeros(64, device="cuda") for _ in range(L)]
bufs = [ops.fpa_empty(n, h, w, 64) for _ in range(L + 1)]
bufs[0].data.normal_(generator=g)
chain = ops.ConvChain(bufs[:-1], ws, bs, ["relu"] * L, bufs[1:])
for form in ("flat", "strip"):
    def seq():
        with ops.conv_form(form):
            for l in range(L):
                ops.conv_tc(bufs[l], ws[l], bs[l], 3, "relu", out=bufs[l + 1])
    print(f"18 layers at {(n, h, w)} one by one, {form}: {timeit(seq) * 1000:.1f} us", flush=True)
print(f"18 layers at {(n, h, w)} as one chain: {timeit(chain.run) * 1000:.1f} us", flush=True)
sys.exit(0 if ok else 1)

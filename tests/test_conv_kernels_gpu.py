"""Parity of the CUDA conv kernels (through the C ABI) against the CPU oracle.  Tolerances:
north_star allows max-abs 2e-2 on [0,1] images in bf16 == 4e-2 on the reference's [-1,1] range;
kernel-level checks here are much tighter (bf16 store rounding: 2^-8 relative)."""
import numpy as np
import pytest
import torch

from oracle import ops as O

pytestmark = pytest.mark.gpu


def _rng(seed):
    return np.random.default_rng(seed)


def _bf(x):
    return O.bf16_round(np.asarray(x, np.float32))


def _close_bf16(got, ref, rel=2.0 ** -7, abs_=2e-3):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    err = np.abs(got - ref)
    tol = rel * np.abs(ref) + abs_
    bad = err > tol
    assert not bad.any(), f"{bad.sum()} / {bad.size} mismatches, max err {err.max():.4g} (ref max {np.abs(ref).max():.4g})"


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("k,cin,pad,act,shape", [
    (3, 3, "SAME", "relu", (2, 9, 11)),
    (3, 1, "SAME", "relu", (1, 16, 7)),
    (5, 1, "SAME", "tanh", (2, 12, 13)),
    (5, 3, "SAME", "tanh", (1, 17, 17)),
    (9, 3, "VALID", "relu", (2, 21, 23)),
    (9, 1, "VALID", "relu", (3, 33, 33)),
    (3, 3, "SAME", "relu", (4, 41, 41)),
])
def test_conv_first(srk_ops, k, cin, pad, act, shape):
    n, h, w = shape
    r = _rng(k * 100 + cin)
    x = r.uniform(-1, 1, (n, h, w, cin)).astype(np.float32)
    wt = (r.standard_normal((k, k, cin, 64)) * 0.2).astype(np.float32)
    b = (r.standard_normal(64) * 0.1).astype(np.float32)
    y = srk_ops.conv_first(_dev(x), _dev(wt), _dev(b), pad, act)
    got = srk_ops.fpa_to_nhwc(y).cpu().numpy()
    ref = O.conv2d_nhwc(x, wt, b, pad, act)
    assert got.shape == ref.shape
    _close_bf16(got, ref)
    # pad rows / columns of the FPA must be exactly zero
    raw = y.data.float().cpu().numpy()
    Wp, S = y.W + 1, (y.H + 1) * (y.W + 1)
    rows = np.arange(y.n_img * S)
    is_pad = ((rows % Wp) == y.W) | (((rows // Wp) % (y.H + 1)) == 0)
    assert np.all(raw[: y.n_img * S][is_pad] == 0)


@pytest.mark.parametrize("k,cin,pad,act,shape", [
    (3, 3, "SAME", "relu", (2, 9, 11)),
    (3, 1, "SAME", None, (1, 16, 7)),
    (5, 1, "SAME", "tanh", (2, 12, 13)),
    (5, 3, "SAME", "tanh", (1, 17, 17)),
    (9, 3, "VALID", "relu", (2, 21, 23)),
    (9, 1, "VALID", "relu", (3, 33, 33)),
    (3, 3, "SAME", "relu", (64, 41, 41)),
    (5, 1, "SAME", "tanh", (1, 40, 250)),
])
def test_conv_first_tc(srk_ops, k, cin, pad, act, shape):
    """Tensor-core first layer: inputs and weights are rounded to bf16 by the kernel, so the oracle gets the same."""
    n, h, w = shape
    r = _rng(k * 100 + cin + 7)
    x = r.uniform(-1, 1, (n, h, w, cin)).astype(np.float32)
    wt = (r.standard_normal((k, k, cin, 64)) * 0.2).astype(np.float32)
    b = (r.standard_normal(64) * 0.1).astype(np.float32)
    wp = srk_ops.pack_first_weights(_dev(wt))
    y = srk_ops.conv_first_tc(_dev(x), wp, _dev(b), k, pad, act)
    got = srk_ops.fpa_to_nhwc(y).cpu().numpy()
    ref = O.conv2d_nhwc(_bf(x), _bf(wt), b, pad, act)
    assert got.shape == ref.shape
    _close_bf16(got, ref, abs_=3e-3)
    # and against the un-rounded fp64 oracle within the model-level bf16 budget
    assert np.abs(got - O.conv2d_nhwc(x, wt, b, pad, act)).max() <= 4e-2
    raw = y.data.float().cpu().numpy()
    Wp, S = y.W + 1, (y.H + 1) * (y.W + 1)
    rows = np.arange(y.n_img * S)
    is_pad = ((rows % Wp) == y.W) | (((rows // Wp) % (y.H + 1)) == 0)
    assert np.all(raw[: y.n_img * S][is_pad] == 0)


def test_last_layer_dgrad_via_first_tc(srk_ops):
    """dX = conv(dY[...,3], rot180(W)^T) * relu'(saved): the first-layer kernel with SRK_PACK_FIRST_ROT180T weights."""
    r = _rng(11)
    n, h, w = 4, 19, 21
    dy = (r.standard_normal((n, h, w, 3)) * 1e-3).astype(np.float32)
    wt = (r.standard_normal((3, 3, 64, 3)) / 24).astype(np.float32)
    saved = _bf(r.standard_normal((n, h, w, 64)))
    wp = srk_ops.pack_first_weights(_dev(wt), srk_ops.PACK_FIRST_ROT180T)
    dx = srk_ops.conv_first_tc(_dev(dy), wp, None, 3, "SAME", None, mask_src=srk_ops.fpa_from_nhwc(_dev(saved)), mask_kind="relu")
    gx, _, _ = O.conv2d_backward(np.zeros((n, h, w, 64)), _bf(wt), np.zeros(3), _bf(dy), "SAME", None)
    _close_bf16(srk_ops.fpa_to_nhwc(dx).cpu().numpy(), gx * (saved > 0), abs_=1e-6)


@pytest.mark.parametrize("cin,cout,k,act,shape", [
    (64, 64, 3, "relu", (2, 7, 9)),
    (64, 64, 3, "relu", (64, 41, 41)),
    (64, 64, 3, None, (1, 5, 130)),     # Wp > 127: two chunks of look-behind
    (64, 64, 3, "relu", (1, 300, 254)),  # widest supported panel
    (64, 32, 3, "tanh", (2, 17, 17)),
    (64, 64, 1, None, (3, 32, 32)),
    (64, 32, 1, "relu", (2, 25, 25)),
    (32, 64, 3, None, (2, 17, 17)),
])
def test_conv_tc_fpa(srk_ops, cin, cout, k, act, shape):
    n, h, w = shape
    r = _rng(cin + cout + k + h)
    x = _bf(r.uniform(-1, 1, (n, h, w, cin)))
    wt = _bf(r.standard_normal((k, k, cin, cout)) * (1.0 / np.sqrt(k * k * cin)))
    b = (r.standard_normal(cout) * 0.1).astype(np.float32)
    xf = srk_ops.fpa_from_nhwc(_dev(x))
    wp = srk_ops.pack_conv_weights(_dev(wt))
    y = srk_ops.conv_tc(xf, wp, _dev(b), k, act)
    got = srk_ops.fpa_to_nhwc(y).cpu().numpy()
    ref = O.conv2d_nhwc(x, wt, b, "SAME", act)
    _close_bf16(got, ref)
    raw = y.data.float().cpu().numpy()
    Wp, S = y.W + 1, (y.H + 1) * (y.W + 1)
    rows = np.arange(y.n_img * S)
    is_pad = ((rows % Wp) == y.W) | (((rows // Wp) % (y.H + 1)) == 0)
    assert np.all(raw[: y.n_img * S][is_pad] == 0)


def test_conv_tc_dgrad_mask_and_addend(srk_ops):
    """dgrad form: rot180/transposed weights + ReLU' mask; and the ENet block form relu(x + conv1x1)."""
    r = _rng(7)
    n, h, w = 3, 19, 23
    dy = _bf(r.standard_normal((n, h, w, 64)))
    wt = _bf(r.standard_normal((3, 3, 64, 64)) / 24.0)
    saved = _bf(r.standard_normal((n, h, w, 64)))
    dyf = srk_ops.fpa_from_nhwc(_dev(dy))
    sf = srk_ops.fpa_from_nhwc(_dev(saved))
    wd = srk_ops.pack_conv_weights(_dev(wt), srk_ops.PACK_DGRAD)
    dx = srk_ops.conv_tc(dyf, wd, None, 3, None, mask_src=sf, mask_kind="relu")
    got = srk_ops.fpa_to_nhwc(dx).cpu().numpy()
    # oracle: dL/dx of y = conv(x, w) given dy, times relu'(saved)
    gx, _, _ = O.conv2d_backward(np.zeros((n, h, w, 64)), wt, np.zeros(64), dy, "SAME", None)
    _close_bf16(got, gx * (saved > 0))
    # tanh mask
    dx2 = srk_ops.conv_tc(dyf, wd, None, 3, None, mask_src=sf, mask_kind="tanh")
    _close_bf16(srk_ops.fpa_to_nhwc(dx2).cpu().numpy(), gx * (1 - saved.astype(np.float64) ** 2), rel=2.0 ** -6)
    # relu(x + conv1x1(t) + b)
    w1 = _bf(r.standard_normal((1, 1, 64, 64)) / 8.0)
    b1 = (r.standard_normal(64) * 0.1).astype(np.float32)
    y = srk_ops.conv_tc(dyf, srk_ops.pack_conv_weights(_dev(w1)), _dev(b1), 1, None, addend=sf, relu_after_add=True)
    ref = np.maximum(O.conv2d_nhwc(dy, w1, b1, "SAME", None) + saved, 0)
    _close_bf16(srk_ops.fpa_to_nhwc(y).cpu().numpy(), ref)


@pytest.mark.parametrize("cin,cout,r,shape,act", [
    (64, 3, 1, (2, 13, 15), None),
    (64, 1, 1, (1, 9, 140), None),
    (32, 27, 3, (2, 10, 12), None),
    (32, 9, 3, (1, 24, 31), None),
    (32, 12, 2, (1, 8, 8), None),
])
def test_conv_tc_last(srk_ops, cin, cout, r, shape, act):
    n, h, w = shape
    g = _rng(cin + cout + r)
    x = _bf(g.uniform(-1, 1, (n, h, w, cin)))
    wt = _bf(g.standard_normal((3, 3, cin, cout)) / np.sqrt(9 * cin))
    b = (g.standard_normal(cout) * 0.1).astype(np.float32)
    c = cout // (r * r)
    addend = g.uniform(-1, 1, (n, h * r, w * r, c)).astype(np.float32)
    xf = srk_ops.fpa_from_nhwc(_dev(x))
    wp = srk_ops.pack_conv_weights(_dev(wt))
    bp = srk_ops.pad_bias(_dev(b), wp.shape[1])
    out = srk_ops.conv_tc_last(xf, wp, bp, 3, cout, act, addend=_dev(addend), shuffle_r=r).cpu().numpy()
    ref = O.pixel_shuffle(O.conv2d_nhwc(x, wt, b, "SAME", act), r) + addend
    assert out.shape == ref.shape
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=2e-5)


def test_conv_tc_last_valid_5x5_via_panel_crop(srk_ops):
    """SRCNN reconstruction: 5x5 VALID 32->3 tanh == SAME over the FPA, cropped by 2 px."""
    g = _rng(55)
    n, h, w = 2, 25, 25
    x = _bf(g.uniform(0, 1, (n, h, w, 32)))
    wt = _bf(g.standard_normal((5, 5, 32, 3)) / np.sqrt(25 * 32))
    b = (g.standard_normal(3) * 0.1).astype(np.float32)
    xf = srk_ops.fpa_from_nhwc(_dev(x))
    wp = srk_ops.pack_conv_weights(_dev(wt))
    panels = srk_ops.make_panels([(i, -2, -2, 2, h - 2, 2, w - 2) for i in range(n)])
    out = srk_ops.conv_tc_last(xf, wp, srk_ops.pad_bias(_dev(b), 16), 5, 3, "tanh", panels=panels,
                               frame_shape=(n, h - 4, w - 4)).cpu().numpy()
    ref = O.conv2d_nhwc(x, wt, b, "VALID", "tanh")
    np.testing.assert_allclose(out, ref, rtol=2e-3, atol=2e-3)  # tanh.approx


@pytest.mark.parametrize("shape", [(2, 7, 9), (64, 41, 41), (1, 20, 200), (3, 33, 120)])
def test_wgrad_tc(srk_ops, shape):
    n, h, w = shape
    g = _rng(h * w)
    x = _bf(g.uniform(-1, 1, (n, h, w, 64)))
    dy = _bf(g.standard_normal((n, h, w, 64)) * 0.1)
    xf = srk_ops.fpa_from_nhwc(_dev(x))
    dyf = srk_ops.fpa_from_nhwc(_dev(dy))
    dw = torch.zeros((3, 3, 64, 64), device="cuda")
    db = torch.zeros(64, device="cuda")
    srk_ops.conv_wgrad_tc(xf, dyf, dw, db)
    _, gw, gb = O.conv2d_backward(x, np.zeros((3, 3, 64, 64)), np.zeros(64), dy, "SAME", None)
    scale = np.abs(gw).max()
    np.testing.assert_allclose(dw.cpu().numpy(), gw, rtol=1e-3, atol=1e-4 * scale)
    np.testing.assert_allclose(db.cpu().numpy(), gb, rtol=1e-3, atol=1e-4 * np.abs(gb).max())


def test_first_and_last_layer_wgrad(srk_ops):
    g = _rng(3)
    n, h, w = 3, 11, 13
    x = g.uniform(-1, 1, (n, h, w, 3)).astype(np.float32)
    dy = _bf(g.standard_normal((n, h, w, 64)) * 0.1)
    dw = torch.zeros((3, 3, 3, 64), device="cuda")
    db = torch.zeros(64, device="cuda")
    srk_ops.conv_first_wgrad(_dev(x), srk_ops.fpa_from_nhwc(_dev(dy)), 3, dw, db)
    _, gw, gb = O.conv2d_backward(x, np.zeros((3, 3, 3, 64)), np.zeros(64), dy, "SAME", None)
    np.testing.assert_allclose(dw.cpu().numpy(), gw, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(db.cpu().numpy(), gb, rtol=1e-4, atol=1e-5)
    x2 = _bf(g.uniform(-1, 1, (n, h, w, 64)))
    dy2 = (g.standard_normal((n, h, w, 3)) * 0.1).astype(np.float32)
    dw2 = torch.zeros((3, 3, 64, 3), device="cuda")
    db2 = torch.zeros(3, device="cuda")
    srk_ops.conv_last_wgrad(srk_ops.fpa_from_nhwc(_dev(x2)), _dev(dy2), dw2, db2)
    _, gw2, gb2 = O.conv2d_backward(x2, np.zeros((3, 3, 64, 3)), np.zeros(3), dy2, "SAME", None)
    np.testing.assert_allclose(dw2.cpu().numpy(), gw2, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(db2.cpu().numpy(), gb2, rtol=1e-4, atol=1e-5)


def test_random_shape_sweep(srk_ops):
    """25 random geometries (1..5 images, 3..1500 rows, 3..253 columns) through the 64->64, 64->32, 1x1, last-layer, first-layer
    and weight-gradient kernels against torch conv2d on the same bf16-rounded operands: covers 1-tile CTAs, panels at
    the width limit and the barrier round-robins (tools/stress_conv.py)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "stress_conv.py"), "7", "25"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "stress OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_edge_geometries_and_error_behaviour(srk_ops):
    """Degenerate geometries the flat-stream kernels must still get right (1x1 and 1-row images, the 254-px width limit, one
    image smaller than a tile), the residual-junction data gradient, and the error contract: every refusal is an SrkError with
    text from srk_last_error (the analogue of TF raising at session.run), never a crash or a silent fallback."""
    from ml_super_resolution_b200._ffi import SrkError
    import torch.nn.functional as F
    g = torch.Generator(device="cuda").manual_seed(5)
    w = torch.randn((3, 3, 64, 64), device="cuda", generator=g) * 0.04
    b = torch.randn(64, device="cuda", generator=g) * 0.1
    wp = srk_ops.pack_conv_weights(w)
    for (n, H, W) in [(1, 1, 1), (3, 1, 7), (2, 5, 1), (1, 2, 254), (1, 300, 254)]:
        x = torch.randn((n, H, W, 64), device="cuda", generator=g)
        got = srk_ops.fpa_to_nhwc(srk_ops.conv_tc(srk_ops.fpa_from_nhwc(x), wp, b, 3, "relu"))
        ref = torch.relu(F.conv2d(x.to(torch.bfloat16).float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float().permute(3, 2, 0, 1), b, padding=1))
        assert float((got - ref.permute(0, 2, 3, 1)).abs().max()) <= 2e-2 * max(1.0, float(ref.abs().max())), (n, H, W)
    # (conv + skip gradient) * relu'(saved activation): EnhanceNet's residual junction
    x = torch.randn((2, 9, 11, 64), device="cuda", generator=g)
    skip = torch.randn((2, 9, 11, 64), device="cuda", generator=g)
    saved = torch.randn((2, 9, 11, 64), device="cuda", generator=g)
    got = srk_ops.fpa_to_nhwc(srk_ops.conv_tc(srk_ops.fpa_from_nhwc(x), wp, None, 3, None, mask_src=srk_ops.fpa_from_nhwc(saved), mask_kind="relu",
                                              addend=srk_ops.fpa_from_nhwc(skip), relu_after_add=2))
    conv = F.conv2d(x.to(torch.bfloat16).float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float().permute(3, 2, 0, 1), None, padding=1).permute(0, 2, 3, 1)
    ref = (conv + skip.to(torch.bfloat16).float()) * (saved.to(torch.bfloat16).float() > 0)
    assert float((got - ref).abs().max()) <= 3e-2
    # refusals
    wide = srk_ops.fpa_empty(1, 4, 300, 64)
    with srk_ops.conv_form("flat"):  # the flat-stream form keeps 2*Wp + 128 rows in shared memory: 254 px is its limit
        with pytest.raises(SrkError, match="too large|column panels"):
            srk_ops.conv_tc(wide, wp, b, 3, "relu")
    srk_ops.conv_tc(wide, wp, b, 3, "relu")  # (any width in the column-strip form, which is what wide frames get by default)
    srk_ops.conv_tc(wide, wp, None, 3, None, mask_src=wide, mask_kind="relu")  # (so does the masked data-gradient form)
    with pytest.raises(SrkError, match="too large|column panels"):  # forms without a strip variant (residual junction) still refuse
        srk_ops.conv_tc(wide, wp, None, 3, None, addend=wide, relu_after_add=True)
    with pytest.raises(SrkError, match="unsupported"):
        srk_ops.conv_tc(srk_ops.fpa_empty(1, 4, 4, 64), torch.zeros((49, 64, 64), dtype=torch.bfloat16, device="cuda"), b, 7, "relu")
    with pytest.raises(SrkError, match="smaller than the 11x11"):
        from ml_super_resolution_b200 import metrics as M
        M.ssim(torch.zeros((1, 8, 8, 3), device="cuda"), torch.zeros((1, 8, 8, 3), device="cuda"), 2.0)


@pytest.mark.parametrize("shape,act", [((1, 9, 126), "relu"), ((2, 20, 130), None), ((3, 17, 253), "relu"), ((2, 40, 400), "relu"), ((5, 3, 24), "relu"),
                                       ((1, 1, 1), None)])
def test_conv_strip_form(srk_ops, shape, act):
    """The column-strip form of the plain 3x3 64->64 layer (csrc/conv_strip.cu; vdsr/vdsr/model_vdsr.py:64-83) against the oracle and
    against the flat-stream form: same bf16 gate, a complete FPA (zero pad row / column written), ragged strips (widths that are
    not multiples of 126), images narrower than a strip and smaller than the receptive field, several images per launch."""
    n, h, w = shape
    r = _rng(h * 1000 + w)
    x = _bf(r.uniform(-1, 1, (n, h, w, 64)))
    wt = _bf(r.standard_normal((3, 3, 64, 64)) / 24)
    b = (r.standard_normal(64) * 0.1).astype(np.float32)
    xf, wp, bd = srk_ops.fpa_from_nhwc(_dev(x)), srk_ops.pack_conv_weights(_dev(wt)), _dev(b)
    ref = O.conv2d_nhwc(x, wt, b, "SAME", act)
    outs = {}
    for form in ("strip", "flat") if w <= 254 else ("strip",):  # (254 px is the flat-stream form's limit)
        y = srk_ops.fpa_empty(n, h, w, 64)
        y.data.fill_(float("nan"))
        with srk_ops.conv_form(form):
            srk_ops.conv_tc(xf, wp, bd, 3, act, out=y)
        outs[form] = srk_ops.fpa_to_nhwc(y).cpu().numpy()
        _close_bf16(outs[form], ref)
        raw = y.data.float().cpu().numpy()[: n * (h + 1) * (w + 1)].reshape(n, h + 1, w + 1, 64)
        assert np.all(raw[:, 0] == 0) and np.all(raw[:, :, w] == 0), form
    # the two forms add the nine taps in different fp32 orders: equal to one bf16 rounding step
    if "flat" in outs:
        assert np.abs(outs["strip"] - outs["flat"]).max() <= 2.0 ** -7 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("shape", [(8, 41, 41), (7, 20, 24), (5, 3, 62), (2, 20, 130)])
def test_conv_strip_side_by_side_and_mask(srk_ops, shape):
    """Column-strip form with several narrow images side by side in one tile (3 x 42 lanes for VDSR's 41-pixel training patches,
    vdsr/vdsr/experiment_train.py:17-26; a last group that is only partly filled) and its data-gradient form: SRK_PACK_DGRAD-style
    call with the ReLU' mask of the saved activation (tf.gradients through vdsr/vdsr/model_vdsr.py:64-83)."""
    n, h, w = shape
    r = _rng(n * 100 + w)
    x = _bf(r.uniform(-1, 1, (n, h, w, 64)))
    wt = _bf(r.standard_normal((3, 3, 64, 64)) / 24)
    saved = _bf(r.standard_normal((n, h, w, 64)))
    xf, wp, mf = srk_ops.fpa_from_nhwc(_dev(x)), srk_ops.pack_conv_weights(_dev(wt)), srk_ops.fpa_from_nhwc(_dev(saved))
    ref = O.conv2d_nhwc(x, wt, None, "SAME", None) * (saved > 0)
    outs = {}
    for form in ("strip", "flat"):
        y = srk_ops.fpa_empty(n, h, w, 64)
        y.data.fill_(float("nan"))
        with srk_ops.conv_form(form):
            srk_ops.conv_tc(xf, wp, None, 3, None, out=y, mask_src=mf, mask_kind="relu")
        outs[form] = srk_ops.fpa_to_nhwc(y).cpu().numpy()
        _close_bf16(outs[form], ref)
        raw = y.data.float().cpu().numpy()[: n * (h + 1) * (w + 1)].reshape(n, h + 1, w + 1, 64)
        assert np.all(raw[:, 0] == 0) and np.all(raw[:, :, w] == 0), form
    assert np.abs(outs["strip"] - outs["flat"]).max() <= 2.0 ** -7 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("shape,layers", [((16, 41, 41), 6), ((2, 50, 300), 4)])
def test_conv_chain_equals_layer_by_layer(srk_ops, shape, layers):
    """srk_conv_tc_chain: n layers in one persistent launch (grid barrier between layers, two alternating parts for small layers)
    == the same layers launched one by one, bit for bit -- forward (bias + ReLU) and masked data-gradient form; repeated calls
    re-arm the barrier words."""
    n, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    ws = [srk_ops.pack_conv_weights(torch.randn((3, 3, 64, 64), device="cuda", generator=g) / 24) for _ in range(layers)]
    bs = [torch.randn(64, device="cuda", generator=g) * 0.1 for _ in range(layers)]
    x0 = srk_ops.fpa_from_nhwc(torch.randn((n, h, w, 64), device="cuda", generator=g))
    masks = [srk_ops.fpa_from_nhwc(torch.randn((n, h, w, 64), device="cuda", generator=g)) for _ in range(layers)]
    nv = n * (h + 1) * (w + 1)
    for use_mask in (False, True):
        seq = [x0]
        with srk_ops.conv_form("strip"):
            for l in range(layers):
                seq.append(srk_ops.conv_tc(seq[-1], ws[l], None if use_mask else bs[l], 3, None if use_mask else "relu",
                                           mask_src=masks[l] if use_mask else None, mask_kind="relu" if use_mask else None))
        bufs = [srk_ops.fpa_empty(n, h, w, 64) for _ in range(layers)]
        chain = srk_ops.ConvChain([x0] + bufs[:-1], ws, [None] * layers if use_mask else bs, [None] * layers if use_mask else ["relu"] * layers, bufs,
                                  masks if use_mask else None)
        for _ in range(2):
            chain.run()
        for l in range(layers):
            assert torch.equal(bufs[l].data[:nv], seq[l + 1].data[:nv]), (use_mask, l)

"""CPU suite: the oracle against the committed golden vectors and against independent restatements."""
import os

import numpy as np
import pytest

from oracle import models as OM
from oracle import ops as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def test_pixel_shuffle_matches_reference_numpy_path():
    # golden produced by the reference's own np.split/reshape/concatenate sequence
    assert np.array_equal(O.pixel_shuffle(GOLD["ps_packed"], 3), GOLD["ps_shuffled_ref_literal"])
    assert np.array_equal(O.pixel_unshuffle(GOLD["pu_hr"], 2), GOLD["pu_packed_ref_literal_r2"])
    x = np.random.default_rng(0).standard_normal((2, 3, 4, 5, 48))
    assert np.array_equal(O.pixel_unshuffle(O.pixel_shuffle(x, 4), 4), x)  # round trip, ragged leading dims


def test_pixel_shuffle_is_not_torch_order_without_permute():
    import torch
    p = np.random.default_rng(1).standard_normal((1, 4, 5, 27)).astype(np.float32)
    ours = O.pixel_shuffle(p, 3)
    t = torch.from_numpy(p).permute(0, 3, 1, 2)
    naive = torch.nn.functional.pixel_shuffle(t, 3).permute(0, 2, 3, 1).numpy()
    assert not np.array_equal(ours, naive)
    t2 = torch.from_numpy(p.reshape(1, 4, 5, 9, 3).transpose(0, 1, 2, 4, 3).reshape(1, 4, 5, 27)).permute(0, 3, 1, 2)
    assert np.array_equal(ours, torch.nn.functional.pixel_shuffle(t2, 3).permute(0, 2, 3, 1).numpy())


@pytest.mark.parametrize("name,pad,act", [("conv_a", "SAME", "relu"), ("conv_b", "VALID", "relu"), ("conv_c", "SAME", "tanh")])
def test_conv_golden_and_im2col(name, pad, act):
    x, w, b, y = (GOLD[f"{name}_{s}"] for s in "xwby")
    assert np.abs(O.conv2d_nhwc(x, w, b, pad, act) - y).max() < 1e-12
    assert np.abs(O.conv2d_nhwc_im2col(x, w, b, pad, act) - y).max() < 1e-12


def test_conv_backward_matches_finite_differences():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((1, 5, 6, 2))
    w = rng.standard_normal((3, 3, 2, 3)) * 0.3
    b = rng.standard_normal(3) * 0.1
    dy = rng.standard_normal((1, 5, 6, 3))
    gx, gw, gb = O.conv2d_backward(x, w, b, dy, "SAME", "tanh")
    f = lambda xx, ww, bb: float((O.conv2d_nhwc(xx, ww, bb, "SAME", "tanh") * dy).sum())
    eps = 1e-6
    for idx in [(0, 2, 3, 1), (0, 0, 0, 0), (0, 4, 5, 1)]:
        xp, xm = x.copy(), x.copy()
        xp[idx] += eps
        xm[idx] -= eps
        assert abs((f(xp, w, b) - f(xm, w, b)) / (2 * eps) - gx[idx]) < 1e-6
    for idx in [(0, 0, 0, 0), (1, 2, 1, 2), (2, 2, 0, 1)]:
        wp, wm = w.copy(), w.copy()
        wp[idx] += eps
        wm[idx] -= eps
        assert abs((f(x, wp, b) - f(x, wm, b)) / (2 * eps) - gw[idx]) < 1e-6
    bp, bm = b.copy(), b.copy()
    bp[1] += eps
    bm[1] -= eps
    assert abs((f(x, w, bp) - f(x, w, bm)) / (2 * eps) - gb[1]) < 1e-6


def test_bicubic_tf1_golden_and_properties():
    x = GOLD["bicubic_x"]
    assert np.array_equal(O.resize_bicubic_tf1(x, 36, 45), GOLD["bicubic_up"])
    assert np.array_equal(O.resize_bicubic_tf1(x, 4, 5), x[:, ::3, ::3])  # integer-factor downscale is pure decimation
    idx, w = O.bicubic_taps_tf1(243, 81)
    assert np.allclose(w.sum(axis=1), 1.0, atol=3e-7)  # partition of unity
    assert np.allclose(w[1], [-0.111111, 0.796649, 0.369936, -0.055475], atol=2e-6)  # t = 341
    assert np.array_equal(w[1], GOLD["bicubic_w341"])
    assert idx.min() == 0 and idx.max() == 80
    # constant image stays constant
    c = np.full((1, 7, 7, 1), 0.37, np.float32)
    assert np.allclose(O.resize_bicubic_tf1(c, 21, 21), 0.37, atol=1e-6)


def test_degrade_golden_and_independent_restatements():
    hd = GOLD["degrade_hd"]
    for s in (2, 3, 4):
        assert np.abs(O.hd_image_to_sd_image(hd, s) - GOLD[f"degrade_s{s}"]).max() < 1e-14
    ndi = pytest.importorskip("scipy.ndimage")
    for s in (2, 3, 4):
        sig = 0.5 * (s - 1)
        assert np.abs(ndi.gaussian_filter(hd, [sig, sig, 0], mode="nearest", truncate=4.0) - O.gaussian_blur_nearest(hd, sig)).max() < 1e-12
    cv2 = pytest.importorskip("cv2")
    assert np.abs(cv2.resize(hd, (13, 13), interpolation=cv2.INTER_LINEAR) - O.resize_bilinear_edge(hd, 13, 13)).max() < 1e-9
    assert O.hd_image_to_sd_image(hd, 3).shape == hd.shape  # int(41/3) = 13 -> back to 41


def test_adam_and_momentum_closed_forms():
    w, g = GOLD["adam_w"], GOLD["adam_g"]
    w1, m1, v1 = O.adam_tf(w, g, np.zeros_like(w), np.zeros_like(w), 1, 0.01)
    assert np.array_equal(w1, GOLD["adam_w1"]) and np.array_equal(m1, GOLD["adam_m1"]) and np.array_equal(v1, GOLD["adam_v1"])
    # first step closed form: w - lr * g/(|g| + eps*sqrt(1-b2)/... ) ~= w - lr*sign(g)
    assert np.allclose(w1, w - 0.01 * np.sign(g), atol=1e-6)
    w2, a2 = O.momentum_clip_tf(np.zeros(3, np.float32), np.array([10.0, -10.0, 0.05], np.float32), np.zeros(3, np.float32), 0.1)
    assert np.allclose(a2, [0.1, -0.1, 0.05]) and np.allclose(w2, [-0.01, 0.01, -0.005])
    assert O.stepwise_lr(5e-5, 0.9, 2560 * 3 + 5, 2560) == pytest.approx(5e-5 * 0.9 ** 3)


def test_models_golden_and_losses():
    pv = {k[len("vdsr_p/"):]: GOLD[k] for k in GOLD.files if k.startswith("vdsr_p/")}
    loss, mse, grads, sr = OM.vdsr_loss_and_grads(pv, GOLD["vdsr_sd"], GOLD["vdsr_hd"], num_layers=4)
    assert np.abs(sr - GOLD["vdsr_sr"]).max() < 1e-12
    assert abs(loss - float(GOLD["vdsr_loss"])) < 1e-12
    reg = sum(1e-4 * 0.5 * (v.astype(np.float64) ** 2).sum() for k, v in pv.items() if k.endswith("kernel:0"))
    assert abs(loss - (O.mse_mean(sr, GOLD["vdsr_hd"]) + reg)) < 1e-12
    for k, g in grads.items():
        assert np.allclose(g, GOLD["vdsr_g/" + k], rtol=1e-5, atol=1e-8)
    pe = {k[len("espcn_p/"):]: GOLD[k] for k in GOLD.files if k.startswith("espcn_p/")}
    assert np.abs(OM.espcn_forward(pe, GOLD["espcn_lr"]) - GOLD["espcn_packed"]).max() < 1e-12
    # SRCNN loss restatement vs its closed form; ENet generator shape
    l, g = O.l2norm_rows_mean(np.ones((2, 3, 3, 1)), np.zeros((2, 3, 3, 1)), 9)
    assert l == pytest.approx(3.0) and np.allclose(g, 1.0 / (3.0 * 2))
    pg = OM.enet_g_init(seed=1)
    assert len(pg) == 50 and OM.enet_generator_forward(pg, np.zeros((1, 4, 4, 3)), np.zeros((1, 16, 16, 3))).shape == (1, 16, 16, 3)
    assert O.psnr(np.zeros((1, 2, 2, 1)), np.full((1, 2, 2, 1), 0.2), 2.0)[0] == pytest.approx(20.0)


def test_ssim_restatement_properties():
    """tf.image.ssim restatement: identical images score 1, the score is symmetric, and a uniform image pair reduces to
    the closed-form luminance term (2 m1 m2 + c1) / (m1^2 + m2^2 + c1)."""
    x = OM.synthetic_images(5, 2, 24, 31, 3)
    y = np.roll(x, 3, axis=2)
    assert np.allclose(O.ssim_tf(x, x, 2.0), 1.0)
    assert np.allclose(O.ssim_tf(x, y, 2.0), O.ssim_tf(y, x, 2.0))
    a, b = np.full((1, 12, 12, 1), 0.25), np.full((1, 12, 12, 1), 0.5)
    c1 = (0.01 * 1.0) ** 2
    assert np.allclose(O.ssim_tf(a, b, 1.0), (2 * 0.25 * 0.5 + c1) / (0.25 ** 2 + 0.5 ** 2 + c1))
    assert np.array_equal(O.saturate_cast_u8(np.array([-2.0, -1.0, 0.0, 1.0, 2.0], np.float32)), np.array([0, 0, 127, 255, 255], np.uint8))


def test_ssim_restatement_against_scipy_filters():
    """Independent check of the SSIM restatement: the window statistics computed with scipy.ndimage.correlate1d (interior =
    VALID region) reproduce it to fp64 round-off."""
    ndi = pytest.importorskip("scipy.ndimage")
    x = OM.synthetic_images(6, 1, 40, 37, 2).astype(np.float64)
    y = np.clip(x + 0.1 * np.random.default_rng(0).standard_normal(x.shape), -1, 1)
    coords = np.arange(11) - 5.0
    g = np.exp(-coords ** 2 / (2 * 1.5 ** 2))
    g /= g.sum()

    def filt(t):
        t = ndi.correlate1d(t, g, axis=1, mode="constant")
        t = ndi.correlate1d(t, g, axis=2, mode="constant")
        return t[:, 5:-5, 5:-5, :]

    c1, c2 = (0.01 * 2.0) ** 2, (0.03 * 2.0) ** 2
    m0, m1 = filt(x), filt(y)
    lum = (2 * m0 * m1 + c1) / (m0 ** 2 + m1 ** 2 + c1)
    cs = (2 * filt(x * y) - 2 * m0 * m1 + c2) / (filt(x * x + y * y) - m0 ** 2 - m1 ** 2 + c2)
    ref = (lum * cs).mean(axis=(1, 2)).mean(axis=-1)
    assert np.allclose(O.ssim_tf(x, y, 2.0), ref, rtol=0, atol=1e-12)


def test_pil_resize_restatement_is_pinned_against_pillow():
    """EnhanceNet's input pipeline resizes uint8 patches with scipy.misc.imresize = PIL.Image.resize
    (enet/enet/datasets.py:112-113).  The restatement of Pillow's fixed-point separable resampler must match the Pillow
    installed here BIT FOR BIT: the 128->32 bilinear (antialiased) and 32->128 bicubic cases the reference uses, plus odd sizes."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(11)
    cases = [((128, 128), (32, 32), "bilinear"), ((32, 32), (128, 128), "bicubic"), ((57, 91), (23, 40), "bilinear"),
             ((23, 40), (57, 91), "bicubic"), ((64, 48), (17, 48), "bicubic"), ((40, 40), (40, 13), "bilinear")]
    res = {"bilinear": Image.BILINEAR, "bicubic": Image.BICUBIC}
    for (h, w), (oh, ow), interp in cases:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((ow, oh), resample=res[interp]))
        got = O.pil_resize_u8(img, oh, ow, interp)
        assert np.array_equal(got, ref), (h, w, oh, ow, interp, int(np.abs(got.astype(int) - ref.astype(int)).max()))
    hd = rng.integers(0, 256, (128, 128, 3), dtype=np.uint8)
    sd = np.asarray(Image.fromarray(hd).resize((32, 32), resample=Image.BILINEAR))
    bq = np.asarray(Image.fromarray(sd).resize((128, 128), resample=Image.BICUBIC))
    s, b, h_ = O.enet_batch([hd], [(0, 0, 0)])
    assert np.array_equal(s[0], sd.astype(np.float32) / 127.5 - 1.0) and np.array_equal(b[0], bq.astype(np.float32) / 127.5 - 1.0)
    assert np.array_equal(h_[0], hd.astype(np.float32) / 127.5 - 1.0)


def test_independent_library_cross_checks():
    """More of the restatement pinned against independent implementations that ARE in this image (TensorFlow is not):
    nearest-neighbour resize vs torch (TF1 legacy floor(dst * in / out) == torch 'nearest' for the integer factors the models use,
    enet/enet/model_enet.py:78-80); the Keys cubic-convolution kernel with A = -0.75 in closed form at t = 0 and 1/2 and torch's
    own bicubic kernel (also A = -0.75) at half-pixel taps; TF's Adam against torch.optim.Adam with epsilon -> 0, where the two
    epsilon conventions coincide (moment and bias-correction arithmetic, vdsr/vdsr/model_vdsr.py:145-148); MSE against torch."""
    import torch
    rng = np.random.default_rng(11)
    x = rng.standard_normal((2, 5, 7, 3)).astype(np.float32)
    for f in (2, 4):
        ours = O.resize_nearest_tf1(x, 5 * f, 7 * f)
        ref = torch.nn.functional.interpolate(torch.from_numpy(x).permute(0, 3, 1, 2), scale_factor=f, mode="nearest").permute(0, 2, 3, 1).numpy()
        assert np.array_equal(ours, ref)
    T = O.BICUBIC_TABLE
    assert T[0] == 1.0 and T[1] == 0.0                                                # t = 0: the sample itself
    assert np.allclose([T[2 * 512 + 1], T[2 * 512]], [-0.09375, 0.59375], atol=1e-7)  # t = 1/2: Keys' kernel, A = -0.75
    # torch's bicubic (align_corners=False, A = -0.75) samples a 2x upscale at t = 1/4, 3/4: the same table rows
    imp = np.zeros((1, 1, 1, 9), np.float32)
    imp[0, 0, 0, 4] = 1.0
    up = torch.nn.functional.interpolate(torch.from_numpy(imp), size=(1, 18), mode="bicubic", align_corners=False).numpy()[0, 0, 0]
    assert np.allclose(up[[5, 7, 9, 11]], [T[2 * 768 + 1], T[2 * 768], T[2 * 256], T[2 * 256 + 1]], atol=1e-6)  # distances 1.75, .75, .25, 1.25 from the impulse
    w = rng.standard_normal(50).astype(np.float64)
    tw = torch.nn.Parameter(torch.from_numpy(w.copy()))
    opt = torch.optim.Adam([tw], lr=1e-3, betas=(0.9, 0.999), eps=1e-300)
    m, v, ours_w = np.zeros_like(w), np.zeros_like(w), w.copy()
    for t in range(1, 6):
        g = rng.standard_normal(50)
        tw.grad = torch.from_numpy(g.copy())
        opt.step()
        ours_w, m, v = O.adam_tf(ours_w, g, m, v, t, 1e-3, eps=1e-300, dtype=np.float64)
        assert np.abs(ours_w - tw.detach().numpy()).max() < 1e-12
    a, b = rng.standard_normal((2, 4, 4, 3)), rng.standard_normal((2, 4, 4, 3))
    assert abs(O.mse_mean(a, b) - float(torch.nn.functional.mse_loss(torch.from_numpy(a), torch.from_numpy(b)))) < 1e-14

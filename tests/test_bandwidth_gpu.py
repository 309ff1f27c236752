"""Bandwidth kernels vs the oracle: index work bit-exact, fp32 arithmetic to fp32 round-off."""
import numpy as np
import pytest
import torch

from oracle import models as OM
from oracle import ops as O

pytestmark = pytest.mark.gpu


def _dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("n,h,w,c,r", [(1, 4, 5, 3, 3), (2, 7, 3, 1, 2), (1, 1, 1, 3, 4), (2, 36, 64, 3, 3)])
def test_pixel_shuffle_bit_exact(srk_ops, n, h, w, c, r):
    x = np.random.default_rng(0).standard_normal((n, h, w, c * r * r)).astype(np.float32)
    y = srk_ops.pixel_shuffle(_dev(x), r).cpu().numpy()
    assert np.array_equal(y, O.pixel_shuffle(x, r))
    if c == 3:
        assert np.array_equal(y[0], O.pixel_shuffle_reference_literal(x[0], r))
    back = srk_ops.pixel_unshuffle(_dev(y), r).cpu().numpy()
    assert np.array_equal(back, x)


@pytest.mark.parametrize("n,h,w,c,oh,ow", [(2, 243, 243, 3, 81, 81), (2, 81, 81, 3, 243, 243), (1, 11, 11, 1, 33, 33),
                                            (1, 33, 33, 1, 11, 11), (1, 7, 13, 3, 20, 31)])
def test_resize_bicubic_tf1_bit_exact(srk_ops, n, h, w, c, oh, ow):
    x = np.random.default_rng(1).uniform(-1, 1, (n, h, w, c)).astype(np.float32)
    y = srk_ops.resize_bicubic_tf1(_dev(x), oh, ow).cpu().numpy()
    ref = O.resize_bicubic_tf1(x, oh, ow)
    assert np.array_equal(y, ref), f"max diff {np.abs(y - ref).max()}"


@pytest.mark.parametrize("h,w", [(41, 41), (128, 128), (37, 53)])
def test_degrade_gauss_bilinear(srk_ops, h, w):
    rng = np.random.default_rng(2)
    hd = rng.uniform(0, 1, (6, h, w, 3)).astype(np.float32)
    scales = np.array([2, 3, 4, 2, 3, 4], np.float32)
    sd = srk_ops.degrade_gauss_bilinear(_dev(hd), _dev(scales)).cpu().numpy()
    ref = np.stack([O.hd_image_to_sd_image(hd[i], float(scales[i])) for i in range(6)])
    np.testing.assert_allclose(sd, ref, rtol=0, atol=2e-6)


def test_mse_l2norm_and_optimisers(srk_ops):
    rng = np.random.default_rng(3)
    a = rng.uniform(-1, 1, (8, 41, 41, 3)).astype(np.float32)
    b = rng.uniform(-1, 1, (8, 41, 41, 3)).astype(np.float32)
    loss = torch.zeros(1, device="cuda")
    d = torch.empty(a.shape, device="cuda")
    srk_ops.mse_fwd_bwd(_dev(a), _dev(b), loss, d)
    assert abs(float(loss) - O.mse_mean(a, b)) <= 1e-5
    np.testing.assert_allclose(d.cpu().numpy(), 2 * (a - b) / a.size, rtol=1e-5, atol=1e-9)
    # SRCNN row-L2-norm loss
    sr = rng.uniform(-1, 1, (4, 21, 21, 3)).astype(np.float32)
    hi = rng.uniform(-1, 1, (4, 21, 21, 3)).astype(np.float32)
    loss.zero_()
    d2 = torch.empty(sr.shape, device="cuda")
    srk_ops.l2norm_rows_mean_fwd_bwd(_dev(sr), _dev(hi), 21 * 21, loss, d2)
    ref_loss, ref_grad = O.l2norm_rows_mean(sr, hi, 21 * 21)
    assert abs(float(loss) - ref_loss) <= 1e-4 * ref_loss
    np.testing.assert_allclose(d2.cpu().numpy(), ref_grad, rtol=1e-4, atol=1e-8)
    # same with the gradient taken through the reconstruction layer's tanh (sr = tanh(pre))
    loss.zero_()
    srk_ops.l2norm_rows_mean_fwd_bwd(_dev(sr), _dev(hi), 21 * 21, loss, d2, sr_act="tanh")
    np.testing.assert_allclose(d2.cpu().numpy(), ref_grad * (1.0 - sr.astype(np.float64) ** 2), rtol=1e-4, atol=1e-8)
    # TF-Adam, 3 steps, and Momentum+clip
    n = 10007
    w, g = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32) * 0.1
    wd, gd = _dev(w), _dev(g)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    wr, mr, vr = w.copy(), np.zeros(n, np.float32), np.zeros(n, np.float32)
    for t in (1, 2, 3):
        srk_ops.adam_step(wd, gd, m, v, 1e-3, t)
        wr, mr, vr = O.adam_tf(wr, g, mr, vr, t, 1e-3)
    np.testing.assert_allclose(wd.cpu().numpy(), wr, rtol=1e-5, atol=1e-6)
    acc = torch.zeros(n, device="cuda")
    w2 = _dev(w)
    srk_ops.momentum_clip_step(w2, gd, acc, 0.1)
    wr2, _ = O.momentum_clip_tf(w, g, np.zeros(n, np.float32), 0.1)
    np.testing.assert_allclose(w2.cpu().numpy(), wr2, rtol=1e-6, atol=1e-7)


def test_fpa_upsample2_and_backward(srk_ops):
    rng = np.random.default_rng(4)
    x = O.bf16_round(rng.standard_normal((2, 5, 7, 64)).astype(np.float32))
    xf = srk_ops.fpa_from_nhwc(_dev(x))
    up = srk_ops.fpa_to_nhwc(srk_ops.fpa_upsample2(xf)).cpu().numpy()
    assert np.array_equal(up, O.resize_nearest_tf1(x, 10, 14))
    dy = O.bf16_round(rng.standard_normal((2, 10, 14, 64)).astype(np.float32))
    dx = srk_ops.fpa_to_nhwc(srk_ops.fpa_upsample2_bwd(srk_ops.fpa_from_nhwc(_dev(dy)))).cpu().numpy()
    ref = dy.reshape(2, 5, 2, 7, 2, 64).sum(axis=(2, 4))
    np.testing.assert_allclose(dx, ref, rtol=2.0 ** -7, atol=1e-3)


def test_vdsr_device_input_pipeline_matches_reference_generator(srk_ops):
    """batches cut from the device-resident image pool equal the reference's python generator for the same
    seed -- hd bit for bit (crop, flip, x/255, x*2-1 in fp32), sd within the degrade kernel's fp32 tolerance -- with and
    without the prefetch stream (vdsr/vdsr/dataset.py:41-128)."""
    from ml_super_resolution_b200.vdsr import dataset as D
    rng = np.random.default_rng(9)
    images = [rng.integers(0, 256, (int(rng.integers(45, 90)), int(rng.integers(45, 120)), 3), dtype=np.uint8) for _ in range(7)]
    images.append(rng.integers(0, 256, (30, 200, 3), dtype=np.uint8))  # too small: the generator must skip it
    ref = O.vdsr_image_batches(images, [2.0, 3.0, 4.0], 41, 6, np.random.RandomState(123))
    for prefetch in (False, True):
        ref = O.vdsr_image_batches(images, [2.0, 3.0, 4.0], 41, 6, np.random.RandomState(123))
        got = D.image_batches(images, [2.0, 3.0, 4.0], 41, 6, seed=123, prefetch=prefetch)
        for _ in range(4):
            sd_r, hd_r = next(ref)
            sd_g, hd_g = next(got)
            torch.cuda.current_stream().synchronize()
            assert np.array_equal(hd_g.cpu().numpy(), hd_r.astype(np.float32))
            assert np.abs(sd_g.cpu().numpy() - sd_r).max() <= 1e-5


def test_evaluation_post_pass(srk_ops):
    """psnr / ssim per image (max_val 2.0 on [-1,1] images and 1.0 in ESPCN's clipped packed space, RGB and
    Y), and the saturate-cast uint8 hand-off, against the restated tf.image semantics."""
    from ml_super_resolution_b200 import metrics as M
    hd = OM.synthetic_images(91, 3, 57, 83, 3)
    sr = np.clip(hd + 0.05 * np.random.default_rng(1).standard_normal(hd.shape).astype(np.float32), -1.2, 1.2).astype(np.float32)
    a, b = _dev(hd), _dev(sr)
    np.testing.assert_allclose(M.psnr(a, b, 2.0).cpu().numpy(), O.psnr(hd, sr, 2.0), rtol=0, atol=2e-4)
    np.testing.assert_allclose(M.ssim(a, b, 2.0).cpu().numpy(), O.ssim_tf(hd, sr, 2.0), rtol=0, atol=2e-5)
    # ESPCN: packed space, both score spaces
    hrp = O.pixel_unshuffle(OM.synthetic_images(92, 2, 51, 66, 3), 3)
    srp = (hrp + 0.03 * np.random.default_rng(2).standard_normal(hrp.shape)).astype(np.float32)
    for space in ("rgb", "y"):
        p, s = M.espcn_scores(_dev(srp), _dev(np.ascontiguousarray(hrp)), 3, space)
        rp, rs = O.espcn_scores(srp, hrp, 3, space)
        np.testing.assert_allclose(p.cpu().numpy(), rp, rtol=0, atol=3e-4)
        np.testing.assert_allclose(s.cpu().numpy(), rs, rtol=0, atol=3e-5)
    x = np.linspace(-1.3, 1.3, 4001, dtype=np.float32)
    assert np.array_equal(M.saturate_cast_u8(_dev(x)).cpu().numpy(), O.saturate_cast_u8(x))
    fm = (np.random.default_rng(4).standard_normal((1, 13, 17, 64)) * 0.7).astype(np.float32)
    assert np.array_equal(M.feature_mosaic_u8(_dev(fm)).cpu().numpy(), O.feature_mosaic_u8(fm))


def test_enet_device_input_pipeline_is_bit_identical_to_pillow_path(srk_ops):
    """EnhanceNet's batches cut and resized on the device (uint8 crop, Pillow-arithmetic 25 % bilinear and 400 % bicubic,
    /127.5 - 1) equal the restated reference generator bit for bit (the restatement itself is pinned against Pillow in the
    CPU suite); enet/enet/datasets.py:78-127."""
    from ml_super_resolution_b200.enet import datasets as ED
    rng = np.random.default_rng(21)
    images = [rng.integers(0, 256, (int(rng.integers(256, 300)), int(rng.integers(256, 330)), 3), dtype=np.uint8) for _ in range(3)]
    ref = O.enet_image_batches(images, 5, np.random.RandomState(9))
    got = ED.image_batches(images, 4, 5, seed=9)
    for _ in range(3):
        sd_r, bq_r, hd_r = next(ref)
        sd_g, bq_g, hd_g = next(got)
        assert np.array_equal(hd_g.cpu().numpy(), hd_r) and np.array_equal(sd_g.cpu().numpy(), sd_r) and np.array_equal(bq_g.cpu().numpy(), bq_r)

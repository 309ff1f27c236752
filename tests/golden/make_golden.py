"""Generates the committed golden fixtures under tests/golden/ (run once, in the build container):

    python tests/golden/make_golden.py

The reference (/root/reference) cannot be imported here -- every module imports tensorflow / skimage at the
top, neither of which exists in this image -- so the fixtures come from two sources:
  * `pixel_shuffle_*`: the reference's OWN numpy code path (np.split / reshape / concatenate), restated call
    for call in oracle.ops.pixel_*_reference_literal (espcn/espcn/experiment_test.py:91-96,173-177).  This
    part of the path is pure numpy in the reference, so these vectors are true reference outputs.
  * everything else: the fp64 oracle (parity unpinned, see oracle/__init__.py), cross-checked against
    independent implementations (explicit im2col GEMM, cv2.INTER_LINEAR, scipy.ndimage) at generation time.
Fixtures are small (a few hundred KB) and seeded; tests compare both the oracle (CPU) and the CUDA kernels
(GPU) against them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import models as OM  # noqa: E402
from oracle import ops as O  # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    # ---- pixel shuffle: reference-literal numpy
    packed = rng.standard_normal((6, 5, 27)).astype(np.float32)
    out["ps_packed"] = packed
    out["ps_shuffled_ref_literal"] = O.pixel_shuffle_reference_literal(packed, 3)
    hr = rng.standard_normal((8, 12, 3)).astype(np.float32)
    out["pu_hr"] = hr
    out["pu_packed_ref_literal_r2"] = O.pixel_unshuffle_reference_literal(hr, 2)
    # ---- conv: SAME 3x3 relu, VALID 9x9 relu, SAME 5x5 tanh (fp64 oracle, cross-checked vs im2col)
    for name, (k, cin, cout, pad, act, h, w) in {"conv_a": (3, 3, 8, "SAME", "relu", 7, 9), "conv_b": (9, 1, 4, "VALID", "relu", 13, 12),
                                                 "conv_c": (5, 3, 6, "SAME", "tanh", 8, 8)}.items():
        x = rng.uniform(-1, 1, (2, h, w, cin)).astype(np.float32)
        wt = (rng.standard_normal((k, k, cin, cout)) * 0.2).astype(np.float32)
        b = (rng.standard_normal(cout) * 0.1).astype(np.float32)
        y = O.conv2d_nhwc(x, wt, b, pad, act)
        assert np.abs(y - O.conv2d_nhwc_im2col(x, wt, b, pad, act)).max() < 1e-12
        out[f"{name}_x"], out[f"{name}_w"], out[f"{name}_b"], out[f"{name}_y"] = x, wt, b, y
    # ---- TF1 bicubic: 3x up and 3x down
    img = rng.uniform(-1, 1, (1, 12, 15, 3)).astype(np.float32)
    out["bicubic_x"] = img
    out["bicubic_up"] = O.resize_bicubic_tf1(img, 36, 45)
    out["bicubic_down"] = O.resize_bicubic_tf1(img, 4, 5)
    assert np.array_equal(out["bicubic_down"], img[:, ::3, ::3])
    out["bicubic_w341"] = O.bicubic_taps_tf1(243, 81)[1][1]
    # ---- VDSR degrade (cross-checked vs scipy / cv2 when available)
    hd = rng.uniform(0, 1, (41, 41, 3))
    for s in (2, 3, 4):
        out[f"degrade_s{s}"] = O.hd_image_to_sd_image(hd, s)
    out["degrade_hd"] = hd
    try:
        import cv2
        import scipy.ndimage as ndi
        bl = ndi.gaussian_filter(hd, [1.0, 1.0, 0], mode="nearest", truncate=4.0)
        assert np.abs(bl - O.gaussian_blur_nearest(hd, 1.0)).max() < 1e-12
        assert np.abs(cv2.resize(bl, (13, 13), interpolation=cv2.INTER_LINEAR) - O.resize_bilinear_edge(bl, 13, 13)).max() < 1e-9
    except ImportError:
        pass
    # ---- Adam / momentum one step
    w = rng.standard_normal(64).astype(np.float32)
    g = rng.standard_normal(64).astype(np.float32)
    w1, m1, v1 = O.adam_tf(w, g, np.zeros(64, np.float32), np.zeros(64, np.float32), 1, 0.01)
    out["adam_w"], out["adam_g"], out["adam_w1"], out["adam_m1"], out["adam_v1"] = w, g, w1, m1, v1
    # ---- tiny VDSR (4 layers) forward + loss + grads, tiny ESPCN forward
    pv = OM.vdsr_init(seed=7, num_layers=4)
    sd = OM.synthetic_images(1, 2, 9, 10, 3)
    hdv = OM.synthetic_images(2, 2, 9, 10, 3)
    loss, mse, grads, sr = OM.vdsr_loss_and_grads(pv, sd, hdv, num_layers=4)
    out["vdsr_sd"], out["vdsr_hd"], out["vdsr_sr"], out["vdsr_loss"] = sd, hdv, sr, np.float64(loss)
    for k, v in pv.items():
        out["vdsr_p/" + k] = v
        out["vdsr_g/" + k] = grads[k].astype(np.float32)
    pe = OM.espcn_init(seed=8, scaling_factor=3, channels=3)
    pe = {k: (v * 5).astype(np.float32) for k, v in pe.items()}
    lr = OM.synthetic_images(3, 1, 6, 7, 3)
    out["espcn_lr"], out["espcn_packed"] = lr, OM.espcn_forward(pe, lr)
    for k, v in pe.items():
        out["espcn_p/" + k] = v
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()

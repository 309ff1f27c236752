"""Reference initialisers, numpy host-side: they run once at model creation."""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


def xavier_uniform(rng, kh, kw, cin, cout):
    """tf.contrib.layers.xavier_initializer() -- vdsr/vdsr/model_vdsr.py:27."""
    lim = np.sqrt(6.0 / (kh * kw * cin + kh * kw * cout))
    return rng.uniform(-lim, lim, size=(kh, kw, cin, cout)).astype(np.float32)


def truncated_normal(rng, shape, stddev):
    """tf.truncated_normal_initializer(stddev) -- espcn model_espcn.py:21, enet model_enet.py:11,47, srcnn.py:84."""
    out = rng.standard_normal(size=shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * stddev).astype(np.float32)


def tf_conv_name(i: int) -> str:
    """tf.layers.conv2d auto-naming: conv2d, conv2d_1, ... in creation order."""
    return "conv2d" if i == 0 else f"conv2d_{i}"


def vdsr_params(seed=0, num_layers=20, channels=3) -> "OrderedDict[str, np.ndarray]":
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for i in range(num_layers):
        cin = channels if i == 0 else 64
        cout = channels if i == num_layers - 1 else 64
        p[f"{tf_conv_name(i)}/kernel:0"] = xavier_uniform(rng, 3, 3, cin, cout)
        p[f"{tf_conv_name(i)}/bias:0"] = np.zeros(cout, np.float32)
    return p


def espcn_params(seed=0, scaling_factor=3, channels=3) -> "OrderedDict[str, np.ndarray]":
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for name, s in (("f1", (5, 5, channels, 64)), ("f2", (3, 3, 64, 32)), ("f3", (3, 3, 32, channels * scaling_factor ** 2))):
        p[f"{name}/kernel:0"] = truncated_normal(rng, s, 0.02)
        p[f"{name}/bias:0"] = np.zeros(s[3], np.float32)
    return p


def srcnn_params(seed=0, channels=3, f=(9, 1, 5), n=(64, 32)) -> "OrderedDict[str, np.ndarray]":
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for name, s in (("patch_extraction", (f[0], f[0], channels, n[0])), ("non_linear_mapping", (f[1], f[1], n[0], n[1])),
                    ("reconstruction", (f[2], f[2], n[1], channels))):
        p[f"{name}/weights:0"] = truncated_normal(rng, s, 0.001)
        p[f"{name}/biases:0"] = np.zeros(s[3], np.float32)
    return p


ENET_G_LAYERS = ([(3, 3, 64)] + [(3, 64, 64), (1, 64, 64)] * 10 + [(3, 64, 64)] * 3 + [(3, 64, 3)])


def enet_g_params(seed=0, scope="g_") -> "OrderedDict[str, np.ndarray]":
    rng = np.random.default_rng(seed)
    p = OrderedDict()
    for i, (k, cin, cout) in enumerate(ENET_G_LAYERS):
        p[f"{scope}/{tf_conv_name(i)}/kernel:0"] = truncated_normal(rng, (k, k, cin, cout), 0.02)
        p[f"{scope}/{tf_conv_name(i)}/bias:0"] = np.zeros(cout, np.float32)
    return p

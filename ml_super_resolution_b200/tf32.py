"""The tf32 form of the models' convolutions (north_star: "NHWC bf16/tf32 tiles ... conv outputs within max-abs 2e-2 on [0,1]
images in bf16 (1e-3 in tf32)").  The reference computes in fp32 everywhere (vdsr/vdsr/experiment_train.py:17-26 declares fp32
placeholders; TF-1.8 cuDNN convolutions); this path keeps activations and weights in fp32 storage and multiplies them on the
tensor cores as tf32 (tcgen05 kind::tf32: 10-bit mantissa operands, fp32 accumulation in tensor memory), with exact `tanhf` in
the epilogue -- im2col + the generic GEMM of csrc/gemm_tc.cu, any kernel size / channel count / padding.  It is the accuracy
form (about 8x the HBM traffic of the fused bf16 kernels); the bf16 kernels remain the throughput path.

  vdsr_forward   vdsr/vdsr/model_vdsr.py:47-104        20 x 3x3 conv (+ReLU), global residual
  espcn_forward  espcn/espcn/model_espcn.py:117-134    5x5 tanh, 3x3 tanh, 3x3 linear (packed output)
  srcnn_forward  srcnn/srcnn.py:100-130                9-1-5 VALID, ReLU, ReLU, tanh
"""
from __future__ import annotations

import numpy as np
import torch

from . import nn
from .initializers import tf_conv_name


def _dev(a, device="cuda"):
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(device)


def _conv(x, w, b, pad, act, stride=1):
    layer = nn.Conv(_dev(w), _dev(b), stride, pad, act, dtype=torch.float32)
    return layer.forward(x)


def vdsr_forward(params: dict, sd: torch.Tensor, num_layers: int = 20) -> torch.Tensor:
    """sr = sd + conv_L(relu(conv_{L-1}(... relu(conv_1(sd)))))  (model_vdsr.py:47-104), tf32 tensor cores, fp32 storage."""
    t = sd
    for i in range(num_layers):
        name = tf_conv_name(i)
        t = _conv(t, params[f"{name}/kernel:0"], params[f"{name}/bias:0"], "SAME", "relu" if i < num_layers - 1 else None)
    out = sd.clone()
    nn.axpby(t, out, 1.0, 1.0)
    return out


def espcn_forward(params: dict, lr: torch.Tensor) -> torch.Tensor:
    """Packed `sr_result` [n, h, w, C r^2] (model_espcn.py:117-134)."""
    t = _conv(lr, params["f1/kernel:0"], params["f1/bias:0"], "SAME", "tanh")
    t = _conv(t, params["f2/kernel:0"], params["f2/bias:0"], "SAME", "tanh")
    return _conv(t, params["f3/kernel:0"], params["f3/bias:0"], "SAME", None)


def srcnn_forward(params: dict, lo: torch.Tensor) -> torch.Tensor:
    """srcnn/srcnn.py:100-130 on the already degraded image `lo`."""
    t = _conv(lo, params["patch_extraction/weights:0"], params["patch_extraction/biases:0"], "VALID", "relu")
    t = _conv(t, params["non_linear_mapping/weights:0"], params["non_linear_mapping/biases:0"], "VALID", "relu")
    return _conv(t, params["reconstruction/weights:0"], params["reconstruction/biases:0"], "VALID", "tanh")

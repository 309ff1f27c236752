"""General-shape layers on the generic tcgen05 GEMM (csrc/gemm_tc.cu): convolution (any channel count, kernel size, stride,
TF 'SAME' / 'VALID' padding), dense, 2x2 max-pool -- forward, data gradient and weight gradient -- for the parts of the reference
that are not 3x3 stride-1 stacks of 64 channels:

  * EnhanceNet's training losses: discriminator and VGG-19 (`enet/losses.py`)
        enet/enet/model_enet.py:118-161, enet/enet/model_vgg.py:11-99
  * the tf32 form of the four models' convolutions (`tf32.py`): fp32 storage, tcgen05 kind::tf32, exact tanh.

A convolution is im2col (srk_im2col) + one GEMM with bias and activation fused in its epilogue; its data gradient is one GEMM
against the HWIO kernel as stored + col2im (srk_col2im, gather form: no atomics); its weight gradient is one GEMM over the
transposed im2col and the transposed output gradient.  Activations are NHWC `torch.Tensor`s (bf16, or fp32 for the tf32 form);
torch only allocates them.
"""
from __future__ import annotations

import torch

from . import _ffi, ops
from ._ffi import check

DT_BF16, DT_F32 = 0, 1
ACT = {None: 0, "none": 0, "linear": 0, "relu": 1, "tanh": 2, "leaky_relu": 3, "sigmoid": 4}
PAD = {"SAME": 0, "same": 0, "VALID": 1, "valid": 1}


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return DT_BF16
    assert t.dtype == torch.float32, t.dtype
    return DT_F32


def _align(n: int, dtype: torch.dtype) -> int:
    q = 8 if dtype == torch.bfloat16 else 4  # rows of a K-major operand start on 16-byte boundaries
    return (n + q - 1) // q * q


def out_size(n: int, k: int, stride: int, pad: str) -> int:
    return int(_ffi.lib().srk_conv_out_size(n, k, stride, PAD[pad]))


PRECISE_TF32 = True  # fp32 operands: 3xTF32 (hi.hi + lo.hi + hi.lo, fp32-level accuracy) instead of one truncating tf32 product


def tf32_split(x: torch.Tensor, K: int, side: int) -> torch.Tensor:
    """fp32 [.., R, >=K] -> [.., R, 3*Kp] of round-to-nearest tf32 parts (srk_tf32_split)."""
    batch = x.shape[0] if x.dim() == 3 else 1
    R, Kp = x.shape[-2], _align(K, torch.float32)
    y = torch.empty(((batch, R, 3 * Kp) if x.dim() == 3 else (R, 3 * Kp)), dtype=torch.float32, device=x.device)
    check(_ffi.lib().srk_tf32_split(ops.handle(), ops._ptr_any(x), batch, R, K, x.stride(-2), x.stride(0) if x.dim() == 3 else 0, ops._ptr(y), Kp, side,
                                    ops._stream()), "srk_tf32_split")
    return y


def gemm(a: torch.Tensor, b: torch.Tensor, bias: torch.Tensor | None = None, act=None, leaky: float = 0.2, out_dtype=None,
         k: int | None = None, out: torch.Tensor | None = None, precise: bool | None = None) -> torch.Tensor:
    """D[.., M, N] = act(A[.., M, K] x B[.., N, K]^T + bias).  `a`, `b`: 2-D or 3-D (batched) views whose last dimension is
    contiguous; `k` = the true contraction length when the rows are padded (defaults to a.shape[-1]).  fp32 operands run on
    the tf32 tensor cores: `precise` (default PRECISE_TF32) selects the 3xTF32 split, else one tf32 product (operands truncated
    to 10 mantissa bits by the hardware)."""
    assert a.dtype == b.dtype and a.dim() == b.dim() and a.stride(-1) == 1 and b.stride(-1) == 1
    batch = a.shape[0] if a.dim() == 3 else 1
    M, N = a.shape[-2], b.shape[-2]
    K = k if k is not None else a.shape[-1]
    assert K <= a.shape[-1] and K <= b.shape[-1]
    if a.dtype == torch.float32 and (PRECISE_TF32 if precise is None else precise):
        a, b = tf32_split(a, K, 0), tf32_split(b, K, 1)
        K = a.shape[-1]
    out_dtype = out_dtype or a.dtype
    if out is None:
        out = torch.empty((batch, M, N) if a.dim() == 3 else (M, N), dtype=out_dtype, device=a.device)
    sa, sb, sd = (a.stride(0), b.stride(0), out.stride(0)) if a.dim() == 3 else (0, 0, 0)
    check(_ffi.lib().srk_gemm_tc(ops.handle(), ops._ptr_any(a), ops._ptr_any(b), ops._ptr_any(out), M, N, K, batch, a.stride(-2), b.stride(-2),
                                 out.stride(-2), sa, sb, sd, ops._ptr(bias), ACT[act], float(leaky), _dt(a), _dt(out), ops._stream()), "srk_gemm_tc")
    return out


def im2col(x: torch.Tensor, k: int, stride: int, pad: str, transposed: bool = False) -> torch.Tensor:
    """x NHWC -> col [M, Kp] (or colT [K, Mp]); M = n * Ho * Wo, K = k * k * C (padded row lengths, see srk_im2col)."""
    n, H, W, C = x.shape
    Ho, Wo = out_size(H, k, stride, pad), out_size(W, k, stride, pad)
    M, K = n * Ho * Wo, k * k * C
    if transposed:
        Mp = _align(M, x.dtype)
        col = torch.empty((K, Mp), dtype=x.dtype, device=x.device)
        Kp = K
    else:
        Kp, Mp = _align(K, x.dtype), M
        col = torch.empty((M, Kp), dtype=x.dtype, device=x.device)
    check(_ffi.lib().srk_im2col(ops.handle(), ops._ptr_any(x), _dt(x), n, H, W, C, k, stride, PAD[pad], ops._ptr_any(col), Kp, Mp, int(transposed),
                                ops._stream()), "srk_im2col")
    return col


def col2im(dcol: torch.Tensor, x_shape, k: int, stride: int, pad: str) -> torch.Tensor:
    n, H, W, C = x_shape
    dx = torch.empty(tuple(x_shape), dtype=dcol.dtype, device=dcol.device)
    check(_ffi.lib().srk_col2im(ops.handle(), ops._ptr_any(dcol), _dt(dcol), n, H, W, C, k, stride, PAD[pad], dcol.shape[1], ops._ptr_any(dx), ops._stream()),
          "srk_col2im")
    return dx


def transpose(x: torch.Tensor, pad_rows_to: int | None = None) -> torch.Tensor:
    """[.., R, C] -> [.., C, Rp] (Rp = R rounded up so that rows start on 16-byte boundaries; the padding is never read)."""
    batch = x.shape[0] if x.dim() == 3 else 1
    R, Cc = x.shape[-2], x.shape[-1]
    Rp = pad_rows_to or _align(R, x.dtype)
    y = torch.empty((batch, Cc, Rp) if x.dim() == 3 else (Cc, Rp), dtype=x.dtype, device=x.device)
    check(_ffi.lib().srk_transpose(ops.handle(), ops._ptr_any(x), _dt(x), batch, R, Cc, x.stride(-2), x.stride(0) if x.dim() == 3 else 0, ops._ptr_any(y), Rp,
                                   y.stride(0) if x.dim() == 3 else 0, ops._stream()), "srk_transpose")
    return y


def convert(x: torch.Tensor, dtype: torch.dtype, scale: float = 1.0) -> torch.Tensor:
    y = torch.empty(x.shape, dtype=dtype, device=x.device)
    check(_ffi.lib().srk_convert(ops.handle(), ops._ptr_any(x), _dt(x), x.numel(), float(scale), ops._ptr_any(y), _dt(y), ops._stream()), "srk_convert")
    return y


def axpby(x: torch.Tensor, y: torch.Tensor, alpha: float = 1.0, beta: float = 1.0) -> torch.Tensor:
    """y = alpha * x + beta * y (fp32, in place)."""
    assert x.dtype == torch.float32 and y.dtype == torch.float32 and x.numel() == y.numel()
    check(_ffi.lib().srk_axpby(ops.handle(), ops._ptr(x), x.numel(), float(alpha), float(beta), ops._ptr(y), ops._stream()), "srk_axpby")
    return y


def colsum(x2d: torch.Tensor, out: torch.Tensor | None = None, accumulate: bool = False) -> torch.Tensor:
    M, Cc = x2d.shape
    if out is None:
        out = torch.empty(Cc, dtype=torch.float32, device=x2d.device)
    check(_ffi.lib().srk_colsum(ops.handle(), ops._ptr_any(x2d), _dt(x2d), M, Cc, ops._ptr(out), int(accumulate), ops._stream()), "srk_colsum")
    return out


def act_bwd(dy: torch.Tensor, y: torch.Tensor, act, leaky: float = 0.2) -> torch.Tensor:
    if ACT[act] == 0:
        return dy
    dx = torch.empty_like(dy)
    check(_ffi.lib().srk_act_bwd(ops.handle(), ops._ptr_any(dy), ops._ptr_any(y), _dt(dy), dy.numel(), ACT[act], float(leaky), ops._ptr_any(dx), ops._stream()),
          "srk_act_bwd")
    return dx


def maxpool2x2(x: torch.Tensor) -> torch.Tensor:
    n, H, W, C = x.shape
    y = torch.empty((n, (H + 1) // 2, (W + 1) // 2, C), dtype=x.dtype, device=x.device)
    check(_ffi.lib().srk_maxpool2x2(ops.handle(), ops._ptr_any(x), _dt(x), n, H, W, C, ops._ptr_any(y), ops._stream()), "srk_maxpool2x2")
    return y


def maxpool2x2_bwd(x: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    n, H, W, C = x.shape
    dx = torch.empty_like(x)
    check(_ffi.lib().srk_maxpool2x2_bwd(ops.handle(), ops._ptr_any(x), ops._ptr_any(dy), _dt(x), n, H, W, C, ops._ptr_any(dx), ops._stream()), "srk_maxpool2x2_bwd")
    return dx


# ------------------------------------------------------------------------------------------------ layers
class Conv:
    """tf.layers.conv2d / tf.nn.conv2d + bias_add + activation.  `w` fp32 HWIO [k,k,Cin,Cout] and `b` fp32 [Cout] are views of
    the owner's parameter arena; `pack()` refreshes the GEMM-ready copies after an optimiser step."""

    def __init__(self, w: torch.Tensor, b: torch.Tensor, stride=1, pad="SAME", act=None, leaky=0.2, dtype=torch.bfloat16, in_dtype=None, precise=None):
        self.w, self.b, self.stride, self.pad, self.act, self.leaky = w, b, stride, pad, act, leaky
        self.precise = precise              # fp32 operands: 3xTF32 (None = nn.PRECISE_TF32) or one tf32 product
        self.k, self.cin, self.cout = w.shape[0], w.shape[2], w.shape[3]
        self.dtype = dtype                  # storage type of the OUTPUT activation
        self.in_dtype = in_dtype or dtype   # operand type of this layer's GEMMs (fp32 = tf32 tensor cores)
        self.pack()

    def pack(self):
        K = self.k * self.k * self.cin
        flat = self.w.reshape(K, self.cout)
        self.w_kn = convert(flat, self.in_dtype)                                                           # [K, Cout]: dgrad operand, as stored
        self.w_nk = transpose(self.w_kn, pad_rows_to=_align(K, self.in_dtype))                             # [Cout, Kp]: forward operand

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n, H, W, _ = x.shape
        if x.dtype != self.in_dtype:
            x = convert(x, self.in_dtype)
        col = x.reshape(-1, self.cin) if (self.k == 1 and self.stride == 1) else im2col(x, self.k, self.stride, self.pad)
        y = gemm(col, self.w_nk, self.b, self.act, self.leaky, out_dtype=self.dtype, k=self.k * self.k * self.cin, precise=self.precise)
        return y.view(n, out_size(H, self.k, self.stride, self.pad), out_size(W, self.k, self.stride, self.pad), self.cout)

    def backward(self, x: torch.Tensor, y: torch.Tensor, dy: torch.Tensor, need_dx=True, grads: dict | None = None, names=None):
        """dy = d loss / d y (post-activation).  Returns dx; with `grads` also stores dW (HWIO fp32) and db under `names`."""
        dz = act_bwd(dy, y, self.act, self.leaky)
        if dz.dtype != self.in_dtype:
            dz = convert(dz, self.in_dtype)
        dz2 = dz.reshape(-1, self.cout)
        if grads is not None:
            xin = x if x.dtype == self.in_dtype else convert(x, self.in_dtype)
            colT = im2col(xin, self.k, self.stride, self.pad, transposed=True)        # [K, Mp]
            dzT = transpose(dz2, pad_rows_to=colT.shape[1])                            # [Cout, Mp]
            dw = gemm(colT, dzT, out_dtype=torch.float32, k=dz2.shape[0], precise=self.precise)  # [K, Cout] = HWIO flat
            grads[names[0]] = dw.view(self.k, self.k, self.cin, self.cout)
            grads[names[1]] = colsum(dz2)
        if not need_dx:
            return None
        dcol = gemm(dz2, self.w_kn, out_dtype=self.in_dtype, precise=self.precise)    # [M, K]
        return col2im(dcol, x.shape, self.k, self.stride, self.pad)


class Dense:
    """tf.layers.dense: y = act(x W + b), W fp32 [in, out].  Output widths that would leave rows off 16-byte boundaries (the
    discriminator's last layer has ONE unit) are zero-padded to 8 columns inside; callers see [M, out]."""

    def __init__(self, w: torch.Tensor, b: torch.Tensor, act=None, leaky=0.2, dtype=torch.bfloat16, precise=None):
        self.w, self.b, self.act, self.leaky, self.dtype, self.precise = w, b, act, leaky, dtype, precise
        self.nin, self.nout = w.shape
        self.nout_p = _align(self.nout, dtype)
        self.pack()

    def pack(self):
        if self.nout_p == self.nout:
            self.w_kn = convert(self.w, self.dtype)             # [in, out]
            self.b_p = self.b
        else:
            wp = torch.zeros((self.nin, self.nout_p), dtype=torch.float32, device=self.w.device)
            wp[:, : self.nout].copy_(self.w)
            self.w_kn = convert(wp, self.dtype)
            self.b_p = torch.zeros(self.nout_p, dtype=torch.float32, device=self.w.device)
            self.b_p[: self.nout].copy_(self.b)
        self.w_nk = transpose(self.w_kn)                        # [out_p, in_p]

    def forward(self, x2d: torch.Tensor, out_dtype=None) -> torch.Tensor:
        y = gemm(x2d, self.w_nk, self.b_p, self.act, self.leaky, out_dtype=out_dtype or self.dtype, k=self.nin, precise=self.precise)
        return y if self.nout_p == self.nout else y[:, : self.nout].contiguous()

    def backward(self, x2d, y, dy, grads: dict | None = None, names=None):
        dz = act_bwd(dy, y, self.act, self.leaky)
        if dz.dtype != self.dtype:
            dz = convert(dz, self.dtype)
        if self.nout_p != self.nout:
            dzp = torch.zeros((dz.shape[0], self.nout_p), dtype=dz.dtype, device=dz.device)
            dzp[:, : self.nout].copy_(dz)
            dz = dzp
        if grads is not None:
            xT = transpose(x2d)                                   # [in, Mp]
            dzT = transpose(dz, pad_rows_to=xT.shape[1])          # [out_p, Mp]
            dw = gemm(xT, dzT, out_dtype=torch.float32, k=x2d.shape[0], precise=self.precise)
            grads[names[0]] = dw if self.nout_p == self.nout else dw[:, : self.nout].contiguous()
            grads[names[1]] = colsum(dz)[: self.nout]
        return gemm(dz, self.w_kn, out_dtype=self.dtype, precise=self.precise)  # dx = dz W^T: B = W as stored [in][out_p]

"""Host-side operator layer: one Python function per libsrk entry point (include/srk.h).

torch is used only as the device allocator and stream provider (`.data_ptr()`,
`torch.cuda.current_stream().cuda_stream`); all arithmetic happens in the hand-written sm_100a
kernels.  Every function is asynchronous on torch's current stream and CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _ffi
from ._ffi import SrkPanel, check

ACT = {None: 0, "none": 0, "linear": 0, "relu": 1, "tanh": 2}
PAD = {"SAME": 0, "same": 0, "VALID": 1, "valid": 1}
PACK_FWD, PACK_DGRAD, PACK_ROT180T_F32, PACK_FIRST, PACK_FIRST_ROT180T = 0, 1, 2, 3, 4

_handles: dict[int, C.c_void_p] = {}


def handle(device: int | None = None) -> C.c_void_p:
    """Per-device libsrk handle (created on first use; fails loudly without an sm_100 GPU)."""
    if device is None:
        device = torch.cuda.current_device()
    h = _handles.get(device)
    if h is None:
        h = C.c_void_p()
        check(_ffi.lib().srk_create(device, C.byref(h)), "srk_create")
        _handles[device] = h
    return h


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "libsrk takes contiguous device tensors"
    return C.c_void_p(t.data_ptr())


def _ptr_any(t: torch.Tensor | None) -> C.c_void_p:
    """Device pointer of a tensor or a strided VIEW of one (the callee is given the strides explicitly)."""
    if t is None:
        return C.c_void_p(0)
    assert t.is_cuda
    return C.c_void_p(t.data_ptr())


def _f32(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.float32, t.dtype
    return t


# ------------------------------------------------------------------------------------------
# FPA buffers
# ------------------------------------------------------------------------------------------


@dataclass
class Fpa:
    """Flat padded activation (include/srk.h): bf16 [rows, C], pixel (n,y,x) at row n*S+(y+1)*Wp+x."""
    data: torch.Tensor  # bf16 [rows_alloc, C]
    n_img: int
    H: int
    W: int

    @property
    def C(self) -> int:
        return self.data.shape[1]


def fpa_rows(n_img: int, H: int, W: int) -> int:
    return int(_ffi.lib().srk_fpa_rows(n_img, H, W))


def fpa_empty(n_img: int, H: int, W: int, C_: int = 64, device=None) -> Fpa:
    rows = fpa_rows(n_img, H, W)
    return Fpa(torch.empty((rows, C_), dtype=torch.bfloat16, device=device or "cuda"), n_img, H, W)


def fpa_from_nhwc(x: torch.Tensor, C_: int | None = None) -> Fpa:
    n, h, w, c = x.shape
    out = fpa_empty(n, h, w, C_ or c, x.device)
    assert out.C == c, "channel padding not supported here"
    check(_ffi.lib().srk_nhwc_to_fpa(handle(), _ptr(_f32(x)), c, n, h, w, _ptr(out.data), _stream()), "srk_nhwc_to_fpa")
    return out


def fpa_to_nhwc(a: Fpa) -> torch.Tensor:
    y = torch.empty((a.n_img, a.H, a.W, a.C), dtype=torch.float32, device=a.data.device)
    check(_ffi.lib().srk_fpa_to_nhwc(handle(), _ptr(a.data), a.C, a.n_img, a.H, a.W, _ptr(y), _stream()), "srk_fpa_to_nhwc")
    return y


def fpa_halo_exchange(a: Fpa, panels: torch.Tensor, max_cols: int) -> Fpa:
    """Refresh in place every column a panel does not own from the neighbouring panel that owns it (srk_fpa_halo_exchange)."""
    check(_ffi.lib().srk_fpa_halo_exchange(handle(), _ptr(a.data), a.C, _ptr(panels), a.n_img, a.H, a.W, max_cols, _stream()),
          "srk_fpa_halo_exchange")
    return a


def make_panels(entries, device="cuda") -> torch.Tensor:
    """entries: iterable of (frame, y0, x0, own_y0, own_y1, own_x0, own_x1) -> int32 [n,8] device tensor."""
    arr = np.zeros((len(entries), 8), np.int32)
    for i, e in enumerate(entries):
        arr[i, :7] = e
    return torch.from_numpy(arr).to(device)


# ------------------------------------------------------------------------------------------
# conv family
# ------------------------------------------------------------------------------------------


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def pad_cout(cout: int) -> int:
    return 16 if cout <= 16 else (32 if cout <= 32 else 64)


def pack_conv_weights(w_hwio: torch.Tensor, mode: int = PACK_FWD, np_: int | None = None, cinp: int | None = None,
                      out: torch.Tensor | None = None) -> torch.Tensor:
    """fp32 HWIO [k,k,cin,cout] -> bf16 [k*k, np, cinp] GEMM-B blocks (see srk_pack_conv_weights)."""
    k, k2, cin, cout = w_hwio.shape
    assert k == k2
    n_true, k_true = (cout, cin) if mode == PACK_FWD else (cin, cout)
    np_ = np_ or pad_cout(n_true)
    cinp = cinp or (32 if k_true <= 32 else 64)
    if out is None:
        out = torch.empty((k * k, np_, cinp), dtype=torch.bfloat16, device=w_hwio.device)
    check(_ffi.lib().srk_pack_conv_weights(handle(), _ptr(_f32(w_hwio)), k, cin, cout, mode, np_, cinp, _ptr(out), _stream()),
          "srk_pack_conv_weights")
    return out


def pad_bias(b: torch.Tensor | None, np_: int) -> torch.Tensor | None:
    if b is None:
        return None
    if b.numel() == np_:
        return b
    out = torch.zeros(np_, dtype=torch.float32, device=b.device)
    out[: b.numel()] = b
    return out


def conv_first(x: torch.Tensor, w_hwio: torch.Tensor, bias: torch.Tensor | None, padding="SAME", act=None,
               panels: torch.Tensor | None = None, panel_hw: tuple[int, int] | None = None, out: Fpa | None = None,
               relu_mask: Fpa | None = None) -> Fpa:
    """Small-Cin first layer: fp32 NHWC frames -> FPA (srk_conv_first)."""
    nf, fh, fw, cin = x.shape
    k = w_hwio.shape[0]
    assert w_hwio.shape[3] == 64
    halo = k - 1 if PAD[padding] == 1 else 0
    if panels is None:
        n_img, H, W = nf, fh - halo, fw - halo
    else:
        n_img = panels.shape[0]
        H, W = panel_hw[0] - halo, panel_hw[1] - halo
    if out is None:
        out = fpa_empty(n_img, H, W, 64, x.device)
    check(_ffi.lib().srk_conv_first(handle(), _ptr(_f32(x)), nf, fh, fw, cin, _ptr(_f32(w_hwio)), _ptr(bias), k, PAD[padding],
                                    ACT[act], _ptr(panels), n_img, H, W, _ptr(out.data),
                                    _ptr(relu_mask.data if relu_mask is not None else None), _stream()), "srk_conv_first")
    return out


def first_blocks(k: int, c: int) -> int:
    """64-element K blocks of a first-layer GEMM with K = k*k*c (padded to 16)."""
    return ((k * k * c + 15) // 16 * 16 + 63) // 64


def pack_first_weights(w_hwio: torch.Tensor, mode: int = PACK_FIRST) -> torch.Tensor:
    """fp32 HWIO -> bf16 [blocks, 64, 64] first-layer form (srk_conv_first_tc)."""
    k, _, cin, cout = w_hwio.shape
    c = cin if mode == PACK_FIRST else cout
    out = torch.empty((first_blocks(k, c), 64, 64), dtype=torch.bfloat16, device=w_hwio.device)
    check(_ffi.lib().srk_pack_conv_weights(handle(), _ptr(_f32(w_hwio)), k, cin, cout, mode, 64, 64, _ptr(out), _stream()),
          "srk_pack_conv_weights")
    return out


def conv_first_tc(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor | None, k: int, padding="SAME", act=None,
                  panels: torch.Tensor | None = None, panel_hw: tuple[int, int] | None = None, out: Fpa | None = None,
                  mask_src: Fpa | None = None, mask_kind=None) -> Fpa:
    """Small-Cin first layer on tensor cores: fp32 NHWC frames -> FPA (srk_conv_first_tc)."""
    nf, fh, fw, cin = x.shape
    halo = k - 1 if PAD[padding] == 1 else 0
    if panels is None:
        n_img, H, W = nf, fh - halo, fw - halo
    else:
        n_img = panels.shape[0]
        H, W = panel_hw[0] - halo, panel_hw[1] - halo
    if out is None:
        out = fpa_empty(n_img, H, W, 64, x.device)
    check(_ffi.lib().srk_conv_first_tc(handle(), _ptr(_f32(x)), nf, fh, fw, cin, _ptr(w_packed), _ptr(bias), k, PAD[padding], ACT[act],
                                       _ptr(panels), n_img, H, W, _ptr(out.data), _ptr(mask_src.data if mask_src is not None else None),
                                       ACT[mask_kind], _stream()), "srk_conv_first_tc")
    return out


CONV_FORMS = {"auto": 0, "flat": 1, "strip": 2}
STRIP_W = 126  # pixels per column strip (csrc/conv_strip.cu)


def strip_lanes(width: int) -> float:
    """Share of a 126-lane tile of the column-strip form that carries pixels for images `width` wide (cs_geom in conv_strip.cu):
    narrow images sit side by side, each with its zero column; wide ones are cut into strips."""
    wp = width + 1
    if 2 * wp <= STRIP_W:
        return (STRIP_W // wp) * wp / STRIP_W
    return wp / (-(-wp // STRIP_W) * STRIP_W)


class conv_form:
    """`with ops.conv_form(frame_width): ...` -- kernel form of the plain 3x3 64->64 layers inside the block (srk_set_conv_form),
    chosen from the FRAME width so that a frame cut into panels runs the same arithmetic as the un-tiled frame (bit-identical
    tiling); "auto" outside such blocks (training patches, single calls)."""

    def __init__(self, frame_width_or_form):
        f = frame_width_or_form
        self.form = f if isinstance(f, str) else ("strip" if (f >= 100 and (strip_lanes(f) >= 0.8 or f > 254)) else "flat")

    def __enter__(self):
        check(_ffi.lib().srk_set_conv_form(handle(), CONV_FORMS[self.form]), "srk_set_conv_form")
        return self

    def __exit__(self, *exc):
        check(_ffi.lib().srk_set_conv_form(handle(), CONV_FORMS["auto"]), "srk_set_conv_form")
        return False


def conv_tc(x: Fpa, w_packed: torch.Tensor, bias: torch.Tensor | None, k: int, act=None, out: Fpa | None = None,
            mask_src: Fpa | None = None, mask_kind=None, addend: Fpa | None = None, relu_after_add=False) -> Fpa:
    """Tensor-core conv FPA -> FPA (srk_conv_tc).  w_packed: [k*k, cout_p, cin_p] bf16."""
    cout_p, cin_p = w_packed.shape[1], w_packed.shape[2]
    assert cin_p == x.C, (cin_p, x.C)
    if out is None:
        out = fpa_empty(x.n_img, x.H, x.W, cout_p, x.data.device)
    check(_ffi.lib().srk_conv_tc(handle(), _ptr(x.data), cin_p, _ptr(w_packed), _ptr(bias), k, cout_p, ACT[act], x.n_img, x.H, x.W,
                                 _ptr(out.data), _ptr(mask_src.data if mask_src is not None else None), ACT[mask_kind],
                                 _ptr(addend.data if addend is not None else None), int(relu_after_add), _stream()), "srk_conv_tc")
    return out


class ConvChain:
    """A prepared srk_conv_tc_chain call: n plain 3x3 64->64 layers of one geometry in one persistent launch.  The pointer arrays
    live in this object (host memory the C call reads at launch time)."""

    def __init__(self, xs: list[Fpa], w_packed: list[torch.Tensor], biases: list[torch.Tensor | None], acts: list, ys: list[Fpa],
                 masks: list[Fpa | None] | None = None):
        n = len(xs)
        assert n == len(w_packed) == len(biases) == len(acts) == len(ys) and 1 <= n <= 20
        g = xs[0]
        assert all((a.n_img, a.H, a.W, a.C) == (g.n_img, g.H, g.W, 64) for a in list(xs) + list(ys))
        assert all(w.shape == (9, 64, 64) for w in w_packed)
        masks = masks or [None] * n
        vp = C.c_void_p * n
        self._keep = (xs, w_packed, biases, ys, masks)  # the buffers must outlive the launches
        self.x = vp(*[a.data.data_ptr() for a in xs])
        self.w = vp(*[w.data_ptr() for w in w_packed])
        self.b = vp(*[(b.data_ptr() if b is not None else None) for b in biases])
        self.a = (C.c_int * n)(*[ACT[a] for a in acts])
        self.y = vp(*[a.data.data_ptr() for a in ys])
        self.m = vp(*[(m.data.data_ptr() if m is not None else None) for m in masks])
        self.n, self.geom = n, (g.n_img, g.H, g.W)
        self.sync = torch.zeros(4, dtype=torch.int32, device=g.data.device)

    def run(self) -> Fpa:
        check(_ffi.lib().srk_conv_tc_chain(handle(), self.n, self.x, self.w, self.b, self.a, self.y, self.m, *self.geom, _ptr(self.sync), _stream()),
              "srk_conv_tc_chain")
        return self._keep[3][-1]


def conv_tc_last(x: Fpa, w_packed: torch.Tensor, bias: torch.Tensor | None, k: int, cout: int, act=None,
                 addend: torch.Tensor | None = None, shuffle_r: int = 1, panels: torch.Tensor | None = None,
                 frame_shape: tuple[int, int, int] | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """Last layer FPA -> fp32 NHWC with fused residual add / pixel shuffle / panel crop (srk_conv_tc_last)."""
    cout_p, cin_p = w_packed.shape[1], w_packed.shape[2]
    assert cin_p == x.C
    if panels is None:
        nf, fh, fw = x.n_img, x.H, x.W
    else:
        nf, fh, fw = frame_shape
    r = shuffle_r
    if out is None:
        out = torch.empty((nf, fh * r, fw * r, cout // (r * r)), dtype=torch.float32, device=x.data.device)
    check(_ffi.lib().srk_conv_tc_last(handle(), _ptr(x.data), cin_p, _ptr(w_packed), _ptr(bias), k, cout, cout_p, ACT[act], x.n_img,
                                      x.H, x.W, _ptr(panels), nf, fh, fw, r, _ptr(addend), _ptr(out), _stream()), "srk_conv_tc_last")
    return out


def u8_to_pm1_f64(x_u8: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """uint8 -> float32 `x / 127.5 - 1.0` with numpy's float64 arithmetic (srk_u8_to_pm1_f64): the reference drivers' input
    normalisation, on the device."""
    assert x_u8.dtype == torch.uint8 and out.dtype == torch.float32 and x_u8.numel() == out.numel() and x_u8.is_contiguous() and out.is_contiguous()
    check(_ffi.lib().srk_u8_to_pm1_f64(handle(), _ptr(x_u8), x_u8.numel(), _ptr(out), _stream()), "srk_u8_to_pm1_f64")
    return out


def copy_region(dst: torch.Tensor, src: torch.Tensor, box: tuple[int, int, int, int]) -> None:
    """dst[:, y0:y1, x0:x1] = src[:, y0:y1, x0:x1] between a page-locked host tensor and a device tensor of one NHWC shape (either
    direction), one strided DMA transfer per frame on the current stream (srk_memcpy2d_async)."""
    y0, y1, x0, x1 = box
    n, H, W, C_ = src.shape
    assert dst.shape == src.shape and dst.dtype == src.dtype and dst.is_contiguous() and src.is_contiguous() and dst.is_cuda != src.is_cuda
    e = src.element_size()
    pitch, off = W * C_ * e, (y0 * W + x0) * C_ * e
    for f in range(n):
        base = f * H * pitch + off
        check(_ffi.lib().srk_memcpy2d_async(dst.data_ptr() + base, pitch, src.data_ptr() + base, pitch, (x1 - x0) * C_ * e, y1 - y0, int(dst.is_cuda),
                                            _stream()), "srk_memcpy2d_async")


OUT_F32, OUT_U8 = 0, 1


def espcn_forward(lr: torch.Tensor, w1p: torch.Tensor, b1: torch.Tensor, w2p: torch.Tensor, b2: torch.Tensor, w3p: torch.Tensor,
                  b3: torch.Tensor, scaling_factor: int, shuffle: bool = True, out: torch.Tensor | None = None,
                  rows: tuple[int, int] | None = None, uint8: bool = False) -> torch.Tensor:
    """The whole ESPCN test graph in one persistent kernel (srk_espcn_forward): lr fp32 [n,H,W,C] -> fp32 | uint8
    [n,H*r,W*r,C] (shuffle) or [n,H,W,C*r^2] (packed).  `rows` = (y_begin, y_end): only that LR row band of every frame is
    produced (a rank's share of a tiled frame); the rest of `out` is left untouched."""
    n, H, W, C_ = lr.shape
    r = scaling_factor
    cout = C_ * r * r
    assert w1p.shape[1:] == (64, 64) and w2p.shape == (9, 32, 64) and w3p.shape == (9, _round_up(cout, 16), 32), "packed ESPCN kernels"
    if out is None:
        shape = (n, H * r, W * r, C_) if shuffle else (n, H, W, cout)
        out = torch.empty(shape, dtype=torch.uint8 if uint8 else torch.float32, device=lr.device)
    assert out.dtype == (torch.uint8 if uint8 else torch.float32) and out.numel() == n * H * W * cout
    y0, y1 = rows if rows is not None else (0, H)
    net = _ffi.SrkEspcnNet(w1p.data_ptr(), w2p.data_ptr(), w3p.data_ptr(), _f32(b1).data_ptr(), _f32(b2).data_ptr(), _f32(b3).data_ptr(), C_, r)
    check(_ffi.lib().srk_espcn_forward(handle(), C.byref(net), _ptr(_f32(lr)), n, H, W, y0, y1, int(shuffle), OUT_U8 if uint8 else OUT_F32,
                                       _ptr(out), _stream()), "srk_espcn_forward")
    return out


def espcn_forward_host(lr_host: torch.Tensor, out_host: torch.Tensor, w1p, b1, w2p, b2, w3p, b3, scaling_factor: int, shuffle: bool, uint8: bool,
                       lr_dev: torch.Tensor, lr_u8_dev: torch.Tensor | None, out_dev: torch.Tensor, band_rows: int = 540) -> torch.Tensor:
    """Host tensors in / out around the fused ESPCN kernel with the copies pipelined by row bands on three streams
    (srk_espcn_forward_host: the band loop runs in C).  `lr_host`: fp32 frames or the raw uint8 image (normalised on the device)."""
    n, H, W, C_ = lr_host.shape
    raw = lr_host.dtype == torch.uint8
    assert lr_host.is_contiguous() and out_host.is_contiguous() and not lr_host.is_cuda and not out_host.is_cuda
    assert out_host.dtype == (torch.uint8 if uint8 else torch.float32) and out_host.shape == out_dev.shape and out_dev.dtype == out_host.dtype
    net = _ffi.SrkEspcnNet(w1p.data_ptr(), w2p.data_ptr(), w3p.data_ptr(), _f32(b1).data_ptr(), _f32(b2).data_ptr(), _f32(b3).data_ptr(), C_, scaling_factor)
    check(_ffi.lib().srk_espcn_forward_host(handle(), C.byref(net), lr_host.data_ptr(), int(raw), n, H, W, int(shuffle), OUT_U8 if uint8 else OUT_F32,
                                            out_host.data_ptr(), _ptr(lr_dev), _ptr(lr_u8_dev) if raw else None, _ptr(out_dev), band_rows, _stream()),
          "srk_espcn_forward_host")
    return out_host


_wgrad_ws: dict = {}


def wgrad_workspace(n_img: int, H: int, W: int, device="cuda") -> torch.Tensor:
    """Per-shape workspace for srk_conv_wgrad_tc's per-CTA partial blocks (allocated once, reused by every layer)."""
    nbytes = int(_ffi.lib().srk_conv_wgrad_tc_workspace_bytes(handle(), n_img, H, W))
    key = (str(device), nbytes)
    if key not in _wgrad_ws:
        _wgrad_ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return _wgrad_ws[key]


def conv_wgrad_tc(x: Fpa, dy: Fpa, dw: torch.Tensor | None, dbias: torch.Tensor | None, accumulate=False,
                  workspace: torch.Tensor | None = None) -> None:
    """dw [3,3,64,64] (=|+=) X^T dY per tap, dbias [64] (=|+=) sum dY (srk_conv_wgrad_tc, deterministic).
    dw=None leaves the per-CTA partials in `workspace` for a later `wgrad_reduce_many`."""
    ws = workspace if workspace is not None else wgrad_workspace(x.n_img, x.H, x.W, x.data.device)
    check(_ffi.lib().srk_conv_wgrad_tc(handle(), _ptr(x.data), _ptr(dy.data), x.n_img, x.H, x.W, _ptr(dw), _ptr(dbias),
                                       int(accumulate), _ptr(ws), ws.numel(), _stream()), "srk_conv_wgrad_tc")
    if dw is not None:
        _ffi.launch_count += 1  # the reduce kernel


def conv_wgrad_tc_batched(xs, dys, workspace: torch.Tensor, layer_stride_bytes: int, dsts: torch.Tensor, accumulate=False) -> None:
    """Weight gradients of len(xs) layers of one geometry with one wgrad launch + one reduce launch (srk_conv_wgrad_tc_batched)."""
    n = len(xs)
    assert n == len(dys) and n > 0
    g = xs[0]
    xp = (C.c_void_p * n)(*[x.data.data_ptr() for x in xs])
    dp = (C.c_void_p * n)(*[d.data.data_ptr() for d in dys])
    check(_ffi.lib().srk_conv_wgrad_tc_batched(handle(), xp, dp, n, g.n_img, g.H, g.W, _ptr(workspace), layer_stride_bytes, _ptr(dsts),
                                               int(accumulate), _stream()), "srk_conv_wgrad_tc_batched")
    _ffi.launch_count += 1  # the reduce kernel


def make_wgrad_dsts(entries, device="cuda") -> torch.Tensor:
    """entries: iterable of (dw tensor, db tensor, ci_n, co_n) -> device array of srk_wgrad_dst."""
    arr = (_ffi.SrkWgradDst * len(entries))()
    for i, (dw, db, ci_n, co_n) in enumerate(entries):
        arr[i] = _ffi.SrkWgradDst(dw.data_ptr(), db.data_ptr(), ci_n, co_n)
    return torch.from_numpy(np.frombuffer(bytes(arr), dtype=np.uint8).copy()).to(device)


def wgrad_reduce_many(workspace: torch.Tensor, layer_stride_bytes: int, n_layers: int, n_img: int, H: int, W: int,
                      dsts: torch.Tensor, accumulate=False) -> None:
    """Fold the deferred per-CTA partial blocks of n_layers layers into their dw/dbias with one launch."""
    check(_ffi.lib().srk_wgrad_reduce_many(handle(), _ptr(workspace), layer_stride_bytes, n_layers, n_img, H, W, _ptr(dsts),
                                           int(accumulate), _stream()), "srk_wgrad_reduce_many")


def nhwc_to_fpa_pad(x: torch.Tensor, Cp: int, out: Fpa | None = None) -> Fpa:
    """fp32 NHWC [n,h,w,c] -> FPA with the channels zero-padded to Cp (feeds 3-channel tensors to the 64-channel kernels)."""
    n, h, w, c = x.shape
    if out is None:
        out = fpa_empty(n, h, w, Cp, x.device)
    check(_ffi.lib().srk_nhwc_to_fpa_pad(handle(), _ptr(_f32(x)), c, Cp, n, h, w, _ptr(out.data), _stream()), "srk_nhwc_to_fpa_pad")
    return out


def wgrad_workspace_bytes(n_img: int, H: int, W: int) -> int:
    return int(_ffi.lib().srk_conv_wgrad_tc_workspace_bytes(handle(), n_img, H, W))


def conv_first_wgrad(x: torch.Tensor, dy: Fpa, k: int, dw: torch.Tensor, dbias: torch.Tensor) -> None:
    n, h, w, cin = x.shape
    check(_ffi.lib().srk_conv_first_wgrad(handle(), _ptr(_f32(x)), n, h, w, cin, k, _ptr(dy.data), _ptr(dw), _ptr(dbias), _stream()),
          "srk_conv_first_wgrad")


def conv_last_wgrad(x: Fpa, dy: torch.Tensor, dw: torch.Tensor, dbias: torch.Tensor) -> None:
    cout = dy.shape[-1]
    check(_ffi.lib().srk_conv_last_wgrad(handle(), _ptr(x.data), _ptr(_f32(dy)), x.n_img, x.H, x.W, cout, _ptr(dw), _ptr(dbias),
                                         _stream()), "srk_conv_last_wgrad")


# ------------------------------------------------------------------------------------------
# bandwidth kernels
# ------------------------------------------------------------------------------------------


def pixel_shuffle(x: torch.Tensor, r: int) -> torch.Tensor:
    n, h, w, k = x.shape
    c = k // (r * r)
    y = torch.empty((n, h * r, w * r, c), dtype=torch.float32, device=x.device)
    check(_ffi.lib().srk_pixel_shuffle(handle(), _ptr(_f32(x)), n, h, w, c, r, _ptr(y), _stream()), "srk_pixel_shuffle")
    return y


def pixel_unshuffle(x: torch.Tensor, r: int) -> torch.Tensor:
    n, hh, ww, c = x.shape
    h, w = hh // r, ww // r
    y = torch.empty((n, h, w, c * r * r), dtype=torch.float32, device=x.device)
    check(_ffi.lib().srk_pixel_unshuffle(handle(), _ptr(_f32(x)), n, h, w, c, r, _ptr(y), _stream()), "srk_pixel_unshuffle")
    return y


def resize_bicubic_tf1(x: torch.Tensor, oh: int, ow: int) -> torch.Tensor:
    n, h, w, c = x.shape
    y = torch.empty((n, oh, ow, c), dtype=torch.float32, device=x.device)
    check(_ffi.lib().srk_resize_bicubic_tf1(handle(), _ptr(_f32(x)), n, h, w, c, oh, ow, _ptr(y), _stream()), "srk_resize_bicubic_tf1")
    return y


def degrade_gauss_bilinear(hd: torch.Tensor, scales: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    n, h, w, c = hd.shape
    if out is None:
        out = torch.empty_like(hd)
    check(_ffi.lib().srk_degrade_gauss_bilinear(handle(), _ptr(_f32(hd)), n, h, w, c, _ptr(_f32(scales)), _ptr(out), _stream()),
          "srk_degrade_gauss_bilinear")
    return out


def fpa_upsample2(x: Fpa, out: Fpa | None = None) -> Fpa:
    assert x.C == 64
    if out is None:
        out = fpa_empty(x.n_img, 2 * x.H, 2 * x.W, 64, x.data.device)
    check(_ffi.lib().srk_fpa_upsample2(handle(), _ptr(x.data), x.n_img, x.H, x.W, _ptr(out.data), _stream()), "srk_fpa_upsample2")
    return out


def fpa_upsample2_bwd(dy: Fpa, out: Fpa | None = None) -> Fpa:
    assert dy.C == 64
    h, w = dy.H // 2, dy.W // 2
    if out is None:
        out = fpa_empty(dy.n_img, h, w, 64, dy.data.device)
    check(_ffi.lib().srk_fpa_upsample2_bwd(handle(), _ptr(dy.data), dy.n_img, h, w, _ptr(out.data), _stream()), "srk_fpa_upsample2_bwd")
    return out


def mse_fwd_bwd(sr: torch.Tensor, hd: torch.Tensor, loss_accum: torch.Tensor, dsr: torch.Tensor | None = None,
                numel_total: float | None = None) -> None:
    n = sr.numel()
    check(_ffi.lib().srk_mse_fwd_bwd(handle(), _ptr(_f32(sr)), _ptr(_f32(hd)), n, float(numel_total or n), _ptr(loss_accum), _ptr(dsr),
                                     _stream()), "srk_mse_fwd_bwd")


def l2norm_rows_mean_fwd_bwd(sr: torch.Tensor, hd: torch.Tensor, cols: int, loss_accum: torch.Tensor,
                             dsr: torch.Tensor | None = None, sr_act=None) -> None:
    """sr_act='tanh': sr is a tanh output and dsr receives the gradient w.r.t. its pre-activation."""
    rows = sr.numel() // cols
    check(_ffi.lib().srk_l2norm_rows_mean_fwd_bwd(handle(), _ptr(_f32(sr)), _ptr(_f32(hd)), rows, cols, _ptr(loss_accum), _ptr(dsr),
                                                  ACT[sr_act], _stream()), "srk_l2norm_rows_mean_fwd_bwd")


def adam_step(w, g, m, v, lr: float, t: int, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, decay_mask=None) -> None:
    check(_ffi.lib().srk_adam_step(handle(), _ptr(_f32(w)), _ptr(_f32(g)), _ptr(m), _ptr(v), w.numel(), lr, beta1, beta2, eps, t,
                                   weight_decay, _ptr(decay_mask), _stream()), "srk_adam_step")


def momentum_clip_step(w, g, accum, lr: float, momentum=0.9, gradient_cap=0.01, weight_decay=0.0, decay_mask=None) -> None:
    check(_ffi.lib().srk_momentum_clip_step(handle(), _ptr(_f32(w)), _ptr(_f32(g)), _ptr(accum), w.numel(), lr, momentum, gradient_cap,
                                            weight_decay, _ptr(decay_mask), _stream()), "srk_momentum_clip_step")


def adam_step_dev(w, g, m, v, lr_t_dev: torch.Tensor, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, decay_mask=None) -> None:
    """Adam with the bias-corrected rate read from device memory (graph-replayable)."""
    check(_ffi.lib().srk_adam_step_dev(handle(), _ptr(_f32(w)), _ptr(_f32(g)), _ptr(m), _ptr(v), w.numel(), _ptr(_f32(lr_t_dev)), beta1,
                                       beta2, eps, weight_decay, _ptr(decay_mask), _stream()), "srk_adam_step_dev")


_comm_ready: set = set()


def comm_init(group=None) -> None:
    """srk_comm_init on this rank's handle: rank 0 creates the ncclUniqueId, torch.distributed (any backend) hands it round."""
    import torch.distributed as dist
    dev = torch.cuda.current_device()
    if dev in _comm_ready:
        return
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = (C.c_char * 128)()
    if rank == 0:
        check(_ffi.lib().srk_comm_unique_id(buf), "srk_comm_unique_id")
    ids = [bytes(buf)]
    dist.broadcast_object_list(ids, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    check(_ffi.lib().srk_comm_init(handle(), C.c_char_p(ids[0]), rank, world), "srk_comm_init")
    _comm_ready.add(dev)


def allreduce_grads(flat: torch.Tensor) -> None:
    """Sum all-reduce of the flat gradient arena over the ranks of `comm_init` (srk_allreduce_grads; graph-capturable)."""
    check(_ffi.lib().srk_allreduce_grads(handle(), _ptr(_f32(flat)), flat.numel(), _stream()), "srk_allreduce_grads")


_peer_ready: dict = {}


def peer_init(grad_elems: int, group=None) -> bool:
    """Set up the NVLink peer-memory exchange of this rank's handle for a flat gradient arena of `grad_elems` floats
    (srk_peer_alloc / srk_peer_open): every rank allocates its region, the 64-byte IPC handles travel through torch.distributed
    (any backend).  Returns False (and leaves the NCCL path in charge) when the ranks are not one process per GPU of one node or
    number more than 8."""
    import torch.distributed as dist
    dev = torch.cuda.current_device()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    key = (dev, world)
    if _peer_ready.get(key, 0) >= grad_elems:
        return True
    if world > 8:
        return False
    if key in _peer_ready:  # a larger arena than the region holds: every rank drains its work on the old mappings first
        torch.cuda.synchronize()
        dist.barrier(group)
        check(_ffi.lib().srk_peer_close(handle()), "srk_peer_close")
        dist.barrier(group)
    grad_elems = max(int(grad_elems), 1 << 22)  # (16 MB per staging buffer: every model of this repository fits without a re-allocation)
    buf = (C.c_char * 64)()
    ok = _ffi.lib().srk_peer_alloc(handle(), grad_elems, buf) == 0
    handles = [None] * world
    dist.all_gather_object(handles, bytes(buf) if ok else None, group=group)
    if ok and all(h is not None for h in handles):
        blob = (C.c_char * (64 * world))(*b"".join(handles))
        ok = _ffi.lib().srk_peer_open(handle(), rank, world, blob) == 0
    else:
        ok = False
    # the ranks must agree: one that could not map a peer (no IPC between these processes) sends everybody to the NCCL form
    oks = [None] * world
    dist.all_gather_object(oks, bool(ok), group=group)
    if not all(oks):
        _ffi.lib().srk_peer_close(handle())
        _peer_ready.pop(key, None)
        return False
    _peer_ready[key] = grad_elems
    return True


def allreduce_adam_step_dev(w, g, m, v, lr_t_dev: torch.Tensor, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, decay_mask=None) -> None:
    """g <- sum over ranks of g (rank order, over NVLink peer memory), then Adam with the device-resident rate: ONE kernel
    (srk_allreduce_adam_step_dev; graph-capturable).  Needs peer_init()."""
    check(_ffi.lib().srk_allreduce_adam_step_dev(handle(), _ptr(_f32(w)), _ptr(_f32(g)), _ptr(m), _ptr(v), w.numel(), _ptr(_f32(lr_t_dev)), beta1,
                                                 beta2, eps, weight_decay, _ptr(decay_mask), _stream()), "srk_allreduce_adam_step_dev")


def make_exchange_and_adam(arena, group=None, peer_exchange: bool = True, weight_decay: float = 0.0, decay_mask=None):
    """The data-parallel exchange + Adam update of a graphed training step as one callable `f(lr_t_dev)`: fused over NVLink peer
    memory (allreduce_adam_step_dev) when the ranks are peers of one node, NCCL all-reduce + Adam otherwise, plain Adam for a
    single rank.  Returns (f, fused)."""
    world = 1
    if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
        world = torch.distributed.get_world_size(group)
    fused = world > 1 and peer_exchange and peer_init(arena.w.numel(), group)
    if world > 1 and not fused:
        comm_init(group)

    def f(lr_t):
        if fused:
            allreduce_adam_step_dev(arena.w, arena.g, arena.m, arena.v, lr_t, weight_decay=weight_decay, decay_mask=decay_mask)
        else:
            if world > 1:
                allreduce_grads(arena.g)
            adam_step_dev(arena.w, arena.g, arena.m, arena.v, lr_t, weight_decay=weight_decay, decay_mask=decay_mask)

    return f, bool(fused)


class PinnedScalarFeed:
    """Feeds one host-computed fp32 scalar per step (the bias-corrected Adam rate) into device memory without a host sync.
    The copy reads pinned memory when the GPU EXECUTES it, not when it is queued, and with graph replay the host runs several
    steps ahead: a single reused pinned word would be overwritten before earlier steps' copies had run.  So the values live in a
    ring of pinned slots with one event each; a slot is rewritten only after the copy that read it has executed."""

    def __init__(self, slots: int = 16):
        self.buf = torch.zeros(slots, dtype=torch.float32).pin_memory()
        self.events = [None] * slots
        self.i = 0

    def push(self, value: float, dst: torch.Tensor) -> None:
        k = self.i % len(self.events)
        self.i += 1
        if self.events[k] is not None:
            self.events[k].synchronize()  # (long since executed unless the host is > `slots` steps ahead)
        self.buf[k] = value
        dst.copy_(self.buf[k:k + 1], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev


def graph_training_step(body, arena, beta1: float = 0.9, beta2: float = 0.999):
    """Capture `body(lr_t)` -- a whole training step: forward, loss, backward, (all-reduce,) `adam_step_dev(..., lr_t)`, re-pack --
    into ONE CUDA graph and return `step(lr)`.  `lr_t` is a device scalar holding the bias-corrected Adam rate of the step being
    replayed (fed through a PinnedScalarFeed ring, so the host may run many replays ahead).  The warm-up run outside capture
    (allocations, kernel attributes, NCCL channels) runs with lr_t == 0 and the moments are cleared afterwards: no update."""
    import math
    lr_t = torch.zeros(1, dtype=torch.float32, device=arena.w.device)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        body(lr_t)
    torch.cuda.current_stream().wait_stream(side)
    arena.m.zero_()
    arena.v.zero_()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        body(lr_t)
    feed = PinnedScalarFeed()
    t = [0]

    def step(lr: float):
        t[0] += 1
        feed.push(lr * math.sqrt(1.0 - beta2 ** t[0]) / (1.0 - beta1 ** t[0]), lr_t)
        graph.replay()

    step.graph = graph
    return step


def sumsq_masked(w: torch.Tensor, mask: torch.Tensor | None, scale: float, out_accum: torch.Tensor) -> None:
    check(_ffi.lib().srk_sumsq_masked(handle(), _ptr(_f32(w)), _ptr(mask), w.numel(), scale, _ptr(out_accum), _stream()), "srk_sumsq_masked")


class PackPlan:
    """Device-resident job table for srk_pack_conv_weights_batched: re-packs every conv kernel of a model
    from the flat fp32 parameter arena into one byte arena with a single launch."""

    def __init__(self, device="cuda"):
        self.jobs = []  # (src_offset, k, cin, cout, mode, np, cinp)
        self.device = device
        self.out = None
        self.views = []

    def add(self, src_offset: int, k: int, cin: int, cout: int, mode: int, np_: int = 0, cinp: int = 0) -> int:
        self.jobs.append((src_offset, k, cin, cout, mode, np_, cinp))
        return len(self.jobs) - 1

    def finalize(self) -> None:
        n = len(self.jobs)
        arr = (_ffi.SrkPackJob * n)()
        dst, elem = 0, 0
        layout = []
        for i, (src, k, cin, cout, mode, np_, cinp) in enumerate(self.jobs):
            if mode in (PACK_FIRST, PACK_FIRST_ROT180T):
                nb = first_blocks(k, cin if mode == PACK_FIRST else cout)
                count, esz, shape, dt = nb * 4096, 2, (nb, 64, 64), torch.bfloat16
            elif mode == PACK_ROT180T_F32:
                count, esz, shape, dt = k * k * cout * cin, 4, (k, k, cout, cin), torch.float32
            else:
                count, esz, shape, dt = k * k * np_ * cinp, 2, (k * k, np_, cinp), torch.bfloat16
            dst = _round_up(dst, 1024)
            arr[i] = _ffi.SrkPackJob(src, dst, elem, k, cin, cout, mode, np_, cinp)
            layout.append((dst, count * esz, shape, dt))
            dst += count * esz
            elem += count
        self.total = elem
        self.out = torch.zeros(_round_up(dst, 1024), dtype=torch.uint8, device=self.device)
        host = np.frombuffer(bytes(arr), dtype=np.uint8).copy()
        self.jobs_dev = torch.from_numpy(host).to(self.device)
        self.views = [self.out[o:o + nb].view(dt).view(shape) for (o, nb, shape, dt) in layout]

    def run(self, arena: torch.Tensor) -> None:
        check(_ffi.lib().srk_pack_conv_weights_batched(handle(), _ptr(_f32(arena)), _ptr(self.jobs_dev), len(self.jobs), self.total,
                                                       _ptr(self.out), _stream()), "srk_pack_conv_weights_batched")

"""ctypes binding of libsrk.so (include/srk.h).  No torch types cross this boundary: only raw
device pointers, sizes and a cudaStream_t.  There is no CPU fallback -- a missing library or a
failing call raises."""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

_lib = None


class SrkError(RuntimeError):
    pass


class SrkPackJob(C.Structure):
    _fields_ = [("src_offset", C.c_int64), ("dst_offset", C.c_int64), ("elem_begin", C.c_int64), ("k", C.c_int32),
                ("cin", C.c_int32), ("cout", C.c_int32), ("mode", C.c_int32), ("np", C.c_int32), ("cinp", C.c_int32)]


class SrkWgradDst(C.Structure):
    _fields_ = [("dw", C.c_void_p), ("db", C.c_void_p), ("ci_n", C.c_int32), ("co_n", C.c_int32)]


class SrkPanel(C.Structure):
    _fields_ = [("frame", C.c_int32), ("y0", C.c_int32), ("x0", C.c_int32), ("own_y0", C.c_int32),
                ("own_y1", C.c_int32), ("own_x0", C.c_int32), ("own_x1", C.c_int32), ("reserved", C.c_int32)]


class SrkEspcnNet(C.Structure):
    _fields_ = [("w1_packed", C.c_void_p), ("w2_packed", C.c_void_p), ("w3_packed", C.c_void_p), ("b1", C.c_void_p),
                ("b2", C.c_void_p), ("b3", C.c_void_p), ("channels", C.c_int32), ("scaling_factor", C.c_int32)]


_P, _I, _F, _SZ, _I64, _D = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64, C.c_double

# name -> (restype, argtypes); mirrors include/srk.h declaration by declaration
SIGNATURES = {
    "srk_version": (_I, []),
    "srk_last_error": (C.c_char_p, []),
    "srk_create": (_I, [_I, C.POINTER(_P)]),
    "srk_destroy": (_I, [_P]),
    "srk_num_sms": (_I, [_P]),
    "srk_set_conv_form": (_I, [_P, _I]),
    "srk_memcpy2d_async": (_I, [_P, _SZ, _P, _SZ, _SZ, _SZ, _I, _P]),
    "srk_espcn_forward_host": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _P]),
    "srk_peer_alloc": (_I, [_P, _SZ, _P]),
    "srk_peer_open": (_I, [_P, _I, _I, _P]),
    "srk_allreduce_adam_step_dev": (_I, [_P, _P, _P, _P, _P, _SZ, _P, _F, _F, _F, _F, _P, _P]),
    "srk_peer_close": (_I, [_P]),
    "srk_conv_tc_chain": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "srk_fpa_rows": (_I64, [_I, _I, _I]),
    "srk_pack_conv_weights": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "srk_pack_conv_weights_batched": (_I, [_P, _P, _P, _I, _I64, _P, _P]),
    "srk_sumsq_masked": (_I, [_P, _P, _P, _SZ, _F, _P, _P]),
    "srk_adam_step_dev": (_I, [_P, _P, _P, _P, _P, _SZ, _P, _F, _F, _F, _F, _P, _P]),
    "srk_conv_first": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _P, _I, _I, _I, _P, _P, _P]),
    "srk_conv_first_tc": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _P, _I, _I, _I, _P, _P, _I, _P]),
    "srk_conv_tc": (_I, [_P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, _I, _P]),
    "srk_conv_tc_last": (_I, [_P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I, _I, _I, _P, _P, _P]),
    "srk_conv_wgrad_tc_workspace_bytes": (_SZ, [_P, _I, _I, _I]),
    "srk_conv_wgrad_tc": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _I, _P, _SZ, _P]),
    "srk_conv_wgrad_tc_batched": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _SZ, _P, _I, _P]),
    "srk_wgrad_reduce_many": (_I, [_P, _P, _SZ, _I, _I, _I, _I, _P, _I, _P]),
    "srk_nhwc_to_fpa_pad": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "srk_conv_first_wgrad": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P]),
    "srk_conv_last_wgrad": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "srk_pixel_shuffle": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "srk_pixel_unshuffle": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "srk_resize_bicubic_tf1": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P]),
    "srk_degrade_gauss_bilinear": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "srk_fpa_upsample2": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "srk_fpa_upsample2_bwd": (_I, [_P, _P, _I, _I, _I, _P, _P]),
    "srk_mse_fwd_bwd": (_I, [_P, _P, _P, _SZ, _D, _P, _P, _P]),
    "srk_l2norm_rows_mean_fwd_bwd": (_I, [_P, _P, _P, _I, _I, _P, _P, _I, _P]),
    "srk_adam_step": (_I, [_P, _P, _P, _P, _P, _SZ, _F, _F, _F, _F, _I64, _F, _P, _P]),
    "srk_momentum_clip_step": (_I, [_P, _P, _P, _P, _SZ, _F, _F, _F, _F, _P, _P]),
    "srk_psnr": (_I, [_P, _P, _P, _I, _I64, _F, _P, _P, _P]),
    "srk_ssim": (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _P, _P, _P]),
    "srk_rgb_to_y": (_I, [_P, _P, _I64, _F, _F, _F, _F, _P, _P]),
    "srk_saturate_cast_u8": (_I, [_P, _P, _SZ, _F, _F, _P, _P]),
    "srk_feature_mosaic_u8": (_I, [_P, _P, _I, _I, _P, _P]),
    "srk_crop_u8": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P]),
    "srk_resample_u8": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _P, _P]),
    "srk_u8_to_pm1": (_I, [_P, _P, _SZ, _P, _P]),
    "srk_u8_to_pm1_f64": (_I, [_P, _P, _SZ, _P, _P]),
    "srk_crop_flip_u8": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "srk_affine_f32": (_I, [_P, _P, _SZ, _F, _F, _P, _P]),
    "srk_gemm_tc": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I64, _I64, _I64, _I64, _I64, _I64, _P, _I, _F, _I, _I, _P]),
    "srk_conv_out_size": (_I, [_I, _I, _I, _I]),
    "srk_im2col": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _I, _I64, _I, _P]),
    "srk_col2im": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "srk_transpose": (_I, [_P, _P, _I, _I, _I, _I, _I64, _I64, _P, _I64, _I64, _P]),
    "srk_maxpool2x2": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "srk_maxpool2x2_bwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "srk_act_bwd": (_I, [_P, _P, _P, _I, _I64, _I, _F, _P, _P]),
    "srk_vgg_preprocess": (_I, [_P, _P, _I64, _F, _F, _I, _P, _P]),
    "srk_vgg_preprocess_bwd": (_I, [_P, _P, _I, _I64, _F, _P, _I, _P]),
    "srk_normalize_channels": (_I, [_P, _P, _I, _I64, _I, _P, _P]),
    "srk_normalize_channels_bwd": (_I, [_P, _P, _P, _I, _I64, _I, _P, _P]),
    "srk_extract_patches16": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "srk_extract_patches16_bwd": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "srk_log_loss": (_I, [_P, _P, _I64, _F, _F, _P, _P, _P]),
    "srk_axpby": (_I, [_P, _P, _I64, _F, _F, _P, _P]),
    "srk_convert": (_I, [_P, _P, _I, _I64, _F, _P, _I, _P]),
    "srk_colsum": (_I, [_P, _P, _I, _I64, _I, _P, _I, _P]),
    "srk_tf32_split": (_I, [_P, _P, _I, _I64, _I, _I64, _I64, _P, _I, _I, _P]),
    "srk_comm_unique_id": (_I, [_P]),
    "srk_comm_init": (_I, [_P, _P, _I, _I]),
    "srk_allreduce_grads": (_I, [_P, _P, _SZ, _P]),
    "srk_comm_destroy": (_I, [_P]),
    "srk_espcn_forward": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "srk_fpa_halo_exchange": (_I, [_P, _P, _I, _P, _I, _I, _I, _I, _P]),
    "srk_fpa_to_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    "srk_nhwc_to_fpa": (_I, [_P, _P, _I, _I, _I, _I, _P, _P]),
}


def lib() -> C.CDLL:
    """Load libsrk.so (built by `python -m ml_super_resolution_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SrkError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); there is no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = _CountingLib(l)
    return _lib


# kernels launched per C-ABI call (everything else: 0) -- feeds bench.py's `gpu_launches`
KERNELS_PER_CALL = {name: 1 for name in SIGNATURES if name not in
                    ("srk_version", "srk_last_error", "srk_create", "srk_destroy", "srk_num_sms", "srk_set_conv_form", "srk_fpa_rows",
                     "srk_conv_wgrad_tc_workspace_bytes", "srk_conv_out_size", "srk_comm_unique_id", "srk_comm_init", "srk_allreduce_grads", "srk_comm_destroy", "srk_peer_alloc", "srk_peer_open", "srk_peer_close", "srk_memcpy2d_async")}
KERNELS_PER_CALL["srk_conv_wgrad_tc"] = 1  # +1 when it also runs the reduce (counted as srk_wgrad_reduce_many otherwise)
launch_count = 0


class _CountingLib:
    """Thin proxy over the CDLL that counts kernel launches issued through the C ABI."""

    def __init__(self, cdll):
        self._cdll = cdll
        for name in SIGNATURES:
            fn = getattr(cdll, name)
            k = KERNELS_PER_CALL.get(name, 0)
            setattr(self, name, self._wrap(fn, k) if k else fn)

    @staticmethod
    def _wrap(fn, k):
        def call(*args):
            global launch_count
            launch_count += k
            return fn(*args)
        return call


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise SrkError(f"{what} failed ({rc}): {lib().srk_last_error().decode()}")

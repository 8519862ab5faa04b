"""ctypes binding of libmmcodec.so (C ABI declared in include/mmcodec.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the error is
raised to the caller.  Build with ``python __graft_entry__.py`` (or ``make`` in the package dir).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmcodec.so")

MMC_OK, MMC_EINVAL, MMC_ECUDA, MMC_EUNSUPPORTED, MMC_EDOMAIN = 0, -1, -2, -3, -4
MEANS_NONE, MEANS_FULL, MEANS_PER_CHANNEL = 0, 1, 2
F32, BF16 = 0, 1
NCHW, NHWC, NHWC_PAD8 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_LEAKY_RELU, ACT_ABS, ACT_QRELU8 = 0, 1, 2, 3, 4
GDN_NONE, GDN_FORWARD, GDN_INVERSE = 0, 1, 2

c_i64, c_int, c_f32, c_vp = ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_void_p


class EbParams(ctypes.Structure):
    _fields_ = [("matrix", c_vp * 5), ("bias", c_vp * 5), ("factor", c_vp * 4), ("medians", c_vp)]


class ConvDesc(ctypes.Structure):
    _fields_ = [(n, c_int) for n in ("transposed", "B", "H", "W", "Cin", "Cout", "k", "stride", "in_dtype",
                                     "in_layout", "out_dtype", "out_layout", "act", "gdn", "out2_bf16")]


_PROTOS = {
    "mmc_version": (c_int, []),
    "mmc_last_error": (ctypes.c_char_p, []),
    "mmc_launch_count": (c_i64, []),
    "mmc_reset_launch_count": (None, []),
    "mmc_quantize_symbols": (c_int, [c_vp, c_vp, c_int, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "mmc_quantize_dequantize": (c_int, [c_vp, c_vp, c_int, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "mmc_quantize_noise": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "mmc_dequantize": (c_int, [c_vp, c_vp, c_int, c_i64, c_i64, c_i64, c_vp, c_vp]),
    "mmc_lower_bound": (c_int, [c_vp, c_f32, c_i64, c_vp, c_vp]),
    "mmc_lower_bound_bwd": (c_int, [c_vp, c_vp, c_f32, c_i64, c_vp, c_vp]),
    "mmc_build_indexes": (c_int, [c_vp, c_vp, c_int, c_f32, c_i64, c_vp, c_vp]),
    "mmc_channel_indexes": (c_int, [c_i64, c_i64, c_i64, c_vp, c_vp]),
    "mmc_eb_forward": (c_int, [c_vp, c_vp, ctypes.POINTER(EbParams), c_f32, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mmc_eb_build_lut": (c_int, [ctypes.POINTER(EbParams), c_f32, c_i64, c_int, c_vp, c_vp]),
    "mmc_eb_forward_lut": (c_int, [c_vp, ctypes.POINTER(EbParams), c_vp, c_int, c_f32, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mmc_eb_logits_cumulative": (c_int, [c_vp, ctypes.POINTER(EbParams), c_i64, c_i64, c_i64, c_vp, c_vp]),
    "mmc_gc_forward": (c_int, [c_vp, c_vp, c_vp, c_vp, c_f32, c_f32, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mmc_bits": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "mmc_pmf_to_quantized_cdf_host": (c_int, [c_vp, c_int, c_int, c_vp]),
    "mmc_rans_encode_batch_host": (c_int, [c_vp, c_vp, c_int, c_i64, c_vp, c_int, c_int, c_vp, c_vp, c_vp, ctypes.c_size_t, c_vp]),
    "mmc_rans_selftest": (c_i64, []),
    "mmc_rans_decode_batch_host": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_vp, c_int, c_int, c_vp, c_vp, c_vp]),
    "mmc_gdn_reparam": (c_int, [c_vp, c_vp, c_int, c_f32, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp]),
    "mmc_gdn_forward": (c_int, [c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_i64, c_int, c_vp, c_vp]),
    "mmc_conv_out_size": (c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "mmc_conv_pack_weights": (c_int, [ctypes.POINTER(ConvDesc), c_vp, c_vp, ctypes.POINTER(ctypes.c_size_t), c_vp]),
    "mmc_conv_forward_direct": (c_int, [ctypes.POINTER(ConvDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mmc_conv_forward_tc": (c_int, [ctypes.POINTER(ConvDesc), c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mmc_conv_forward_tc2": (c_int, [ctypes.POINTER(ConvDesc), c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mmc_conv_pad8_size": (c_int, [ctypes.POINTER(ConvDesc), ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "mmc_pad_nchw_to_nhwc8": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "mmc_nchw_f32_to_nhwc_bf16": (c_int, [c_vp, c_i64, c_int, c_i64, c_vp, c_vp]),
    "mmc_nhwc_bf16_to_nchw_f32": (c_int, [c_vp, c_i64, c_int, c_i64, c_vp, c_vp]),
    "mmc_nhwc_f32_to_nchw_f32": (c_int, [c_vp, c_i64, c_int, c_i64, c_vp, c_vp]),
    "mmc_f32_to_bf16": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "mmc_split_f32_bf16x3": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp]),
    "mmc_gdn_apply_f32": (c_int, [c_vp, c_vp, c_int, c_i64, c_vp, c_vp]),
    "mmc_gaussian_volume_workspace": (c_int, [c_i64, c_int, c_int, ctypes.POINTER(ctypes.c_size_t)]),
    "mmc_gaussian_volume": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_int, c_int, c_vp, ctypes.c_size_t, c_vp, c_vp]),
    "mmc_scale_space_warp": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "mmc_add": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "mmc_layernorm_bf16": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_f32, c_vp, c_vp, c_vp]),
    "mmc_gelu_bf16": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "mmc_image_bits": (c_int, [c_vp, c_int, c_i64, c_f32, c_vp, c_vp]),
    "mmc_u8_to_f32": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "mmc_image_sse": (c_int, [c_vp, c_vp, c_int, c_i64, c_f32, c_vp, c_vp]),
    "mmc_conv3x3_mean_workspace": (c_int, [c_int, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_size_t)]),
    "mmc_conv3x3_mean": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_vp]),
    "mmc_maxpool_nhwc_bf16": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "mmc_upsample_bilinear_add_bf16": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_vp, c_vp]),
    "mmc_sigmoid_gate_bf16": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "mmc_channel_mean_workspace": (c_int, [c_int, c_i64, c_int, ctypes.POINTER(ctypes.c_size_t)]),
    "mmc_channel_mean": (c_int, [c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp]),
    "mmc_channel_affine_bf16": (c_int, [c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_vp]),
    "mmc_window_attention": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_f32, c_vp, c_vp]),
    "mmc_color_convert": (c_int, [c_vp, c_i64, c_i64, c_int, c_vp, c_vp]),
    "mmc_avg_pool2": (c_int, [c_vp, c_i64, c_i64, c_int, c_int, c_vp, c_vp]),
    "mmc_upsample2x_bilinear": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_i64, c_vp]),
    "mmc_wgrad_tc": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "mmc_act_bwd": (c_int, [c_vp, c_vp, c_int, c_i64, c_vp, c_vp]),
    "mmc_colsum_bf16": (c_int, [c_vp, c_i64, c_int, c_f32, c_vp, c_vp]),
    "mmc_square_bf16": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "mmc_gdn_bwd_t": (c_int, [c_vp, c_vp, c_vp, c_int, c_i64, c_vp, c_vp]),
    "mmc_gdn_bwd_dx": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_vp, c_vp]),
    "mmc_reparam_bwd": (c_int, [c_vp, c_vp, c_f32, c_i64, c_vp, c_vp]),
    "mmc_abs_to_bf16": (c_int, [c_vp, c_i64, c_vp, c_vp]),
    "mmc_abs_bwd": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "mmc_gc_backward": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, c_f32, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "mmc_pmf_to_quantized_cdf": (c_int, [c_vp, c_i64, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "mmc_eb_backward": (c_int, [c_vp, c_vp, c_vp, ctypes.POINTER(EbParams), c_f32, c_i64, c_i64, c_i64, c_vp, c_vp, c_vp]),
    "mmc_im2col8": (c_int, [c_vp, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "mmc_wgrad_finalize": (c_int, [c_vp, c_int, c_int, c_int, c_f32, c_vp, c_int, c_vp, c_vp]),
    "mmc_maxpool_nhwc_bf16_idx": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "mmc_maxpool_nhwc_bf16_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "mmc_upsample_bilinear_bwd_bf16": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "mmc_sigmoid_gate_bwd_bf16": (c_int, [c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "mmc_gelu_bwd_bf16": (c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
    "mmc_layernorm_bwd_bf16": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_f32, c_vp, c_vp, c_vp, c_vp]),
    "mmc_rans_lanes_default": (c_int, [c_i64]),
    "mmc_rans_device_workspace": (c_int, [c_int, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_size_t)]),
    "mmc_rans_encode_device": (c_int, [c_vp, c_vp, c_int, c_i64, c_vp, c_int, c_int, c_vp, c_vp, c_int, c_vp, ctypes.c_size_t, c_vp, c_vp, c_vp, c_vp]),
    "mmc_rans_decode_device": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_i64, c_int, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "mmc_window_attention_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_f32, c_vp, c_vp, c_vp, c_vp]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None


class MmcodecError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile libmmcodec.so for sm_100a with nvcc (make in the package directory)."""
    pkg = os.path.dirname(_HERE)
    subprocess.check_call(["make", "-C", pkg, "-j", str(min(8, os.cpu_count() or 1))],
                          stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MmcodecError(
                f"{LIB_PATH} not found: the CUDA library has not been built (run `python __graft_entry__.py` "
                "or `make` in the package directory). mmcodec has no CPU or eager fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)  # raises AttributeError if a declared symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc == MMC_OK:
        return
    msg = lib().mmc_last_error().decode("utf-8", "replace")
    if rc in (MMC_EINVAL, MMC_EDOMAIN):
        raise ValueError(msg)
    if rc == MMC_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise MmcodecError(msg)

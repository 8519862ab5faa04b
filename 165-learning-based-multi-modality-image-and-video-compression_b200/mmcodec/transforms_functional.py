"""Host-side mirror of ``compressai.transforms.functional`` (compressai/transforms/functional.py:26-137): the colour
transforms of the video evaluation pipeline, same names, arguments and errors; the arithmetic runs in libmmcodec kernels on
planar fp32 CUDA tensors."""
from __future__ import annotations

from typing import Tuple, Union

import torch
from torch import Tensor

from . import _lib as L
from . import ops

__all__ = ["rgb2ycbcr", "ycbcr2rgb", "yuv_444_to_420", "yuv_420_to_444", "YCBCR_WEIGHTS"]

YCBCR_WEIGHTS = {"ITU-R_BT.709": (0.2126, 0.7152, 0.0722)}


def _check_input_tensor(tensor: Tensor) -> None:
    if (not isinstance(tensor, Tensor) or not tensor.is_floating_point() or not len(tensor.size()) in (3, 4)
            or not tensor.size(-3) == 3):
        raise ValueError("Expected a 3D or 4D tensor with shape (Nx3xHxW) or (3xHxW) as input")


def _convert(x: Tensor, to_ycbcr: bool) -> Tensor:
    _check_input_tensor(x)
    ops._require_cuda(x)
    xc = ops._f32c(x)
    n = xc.shape[0] if xc.dim() == 4 else 1
    out = torch.empty_like(xc)
    L.check(L.lib().mmc_color_convert(xc.data_ptr(), n, xc.shape[-1] * xc.shape[-2], int(to_ycbcr), out.data_ptr(), ops._stream()))
    return out


def rgb2ycbcr(rgb: Tensor) -> Tensor:
    """functional.py:26-45"""
    return _convert(rgb, True)


def ycbcr2rgb(ycbcr: Tensor) -> Tensor:
    """functional.py:48-66"""
    return _convert(ycbcr, False)


def _avg_pool2(t: Tensor) -> Tensor:
    ops._require_cuda(t)
    t = ops._f32c(t)
    n, c, h, w = t.shape
    if w % 2:
        t = t[..., : w - 1].contiguous()      # F.avg_pool2d floors odd sizes
        w -= 1
    out = torch.empty((n, c, h // 2, w // 2), dtype=torch.float32, device=t.device)
    L.check(L.lib().mmc_avg_pool2(t.data_ptr(), h * w, n * c, h, w, out.data_ptr(), ops._stream()))
    return out


def yuv_444_to_420(yuv: Union[Tensor, Tuple[Tensor, Tensor, Tensor]], mode: str = "avg_pool") -> Tuple[Tensor, Tensor, Tensor]:
    """functional.py:69-99"""
    if mode not in ("avg_pool",):
        raise ValueError(f'Invalid downsampling mode "{mode}".')
    if isinstance(yuv, torch.Tensor):
        y, u, v = yuv.chunk(3, 1)
    else:
        y, u, v = yuv
    return (y, _avg_pool2(u), _avg_pool2(v))


def yuv_420_to_444(yuv: Tuple[Tensor, Tensor, Tensor], mode: str = "bilinear", return_tuple: bool = False):
    """functional.py:102-137 (bilinear chroma upsampling on the device; the other modes are not on the accelerated path)"""
    if len(yuv) != 3 or any(not isinstance(c, torch.Tensor) for c in yuv):
        raise ValueError("Expected a tuple of 3 torch tensors")
    if mode not in ("bilinear", "bicubic", "nearest"):
        raise ValueError(f'Invalid upsampling mode "{mode}".')
    if mode != "bilinear":
        raise NotImplementedError(f'upsampling mode "{mode}" is not on the accelerated path (bilinear is)')
    y, u, v = yuv
    ops._require_cuda(y, u, v)
    y, u, v = ops._f32c(y), ops._f32c(u), ops._f32c(v)
    n, _, h, w = u.shape
    if tuple(y.shape) != (n, 1, 2 * h, 2 * w):
        raise ValueError("luma must be twice the chroma resolution")
    out = torch.empty((n, 3, 2 * h, 2 * w), dtype=torch.float32, device=y.device)
    out[:, 0:1].copy_(y)
    hw = 4 * h * w
    for k, c in ((1, u), (2, v)):       # each chroma plane is written straight into its slot of the (N, 3, H, W) result
        L.check(L.lib().mmc_upsample2x_bilinear(c.data_ptr(), n, h, w, out.data_ptr() + k * hw * 4, 3 * hw, ops._stream()))
    if return_tuple:
        return out[:, 0:1], out[:, 1:2], out[:, 2:3]
    return out

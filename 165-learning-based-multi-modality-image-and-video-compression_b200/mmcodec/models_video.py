"""ssf2020 video codec -- host-side mirror of ``compressai.models.video.google.ScaleSpaceFlow``
(compressai/models/video/google.py:55-508) with the reference's constructor arguments, sub-module names (hence
``state_dict`` keys) and return structures.

Every encoder / decoder / hyper stack runs on the fused tcgen05 conv kernels (ReLU and QReLU in the epilogue, NHWC bf16
between layers), the entropy stage on the HBM-bound likelihood kernels, and the scale-space prediction (Gaussian volume,
trilinear warp, residual) on the stencil / gather kernels of csrc/scale_space.cu.  Frames of one GOP are processed in order
(frame t needs the reconstruction of frame t-1, google.py:224-230); GOPs are independent and shard across GPUs.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import ops
from .entropy_models import GaussianConditional
from .layers import conv, deconv
from .models_mm import _to_nhwc_bf16
from .models import CompressionModel, _nhwc_to_logical, _resize_registered_buffers, get_scale_table
from .transforms import QReLU8, TransformStack, run_layers

__all__ = ["ScaleSpaceFlow", "gaussian_kernel1d"]


def gaussian_kernel1d(kernel_size: int, sigma: float, device=None, dtype=torch.float32) -> Tensor:
    """compressai/models/utils.py:155-162 (host-side constant, like the scale table)"""
    khalf = (kernel_size - 1) / 2.0
    x = torch.linspace(-khalf, khalf, steps=kernel_size, dtype=dtype, device=device)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    return pdf / pdf.sum()


def _to_nhwc_pair(y: Tensor):
    """Logical (B, C, H, W) fp32 latent -> (fp32 NHWC, bf16 NHWC)."""
    ops._require_cuda(y)
    y = y.float()
    y_nhwc = y.permute(0, 2, 3, 1) if ops._is_channels_last(y) else y.permute(0, 2, 3, 1).contiguous()
    return y_nhwc, ops.to_bf16(y_nhwc)


class Encoder(TransformStack):
    """models/video/google.py:80-93"""

    def __init__(self, in_planes: int, mid_planes: int = 128, out_planes: int = 192):
        super().__init__(conv(in_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         conv(mid_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         conv(mid_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         conv(mid_planes, out_planes, kernel_size=5, stride=2))


class Decoder(TransformStack):
    """models/video/google.py:95-108"""

    def __init__(self, out_planes: int, in_planes: int = 192, mid_planes: int = 128):
        super().__init__(deconv(in_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         deconv(mid_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         deconv(mid_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         deconv(mid_planes, out_planes, kernel_size=5, stride=2))


class HyperEncoder(TransformStack):
    """models/video/google.py:110-120"""

    def __init__(self, in_planes: int = 192, mid_planes: int = 192, out_planes: int = 192):
        super().__init__(conv(in_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         conv(mid_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         conv(mid_planes, mid_planes, kernel_size=5, stride=2))


class HyperDecoder(TransformStack):
    """models/video/google.py:122-132"""

    def __init__(self, in_planes: int = 192, mid_planes: int = 192, out_planes: int = 192):
        super().__init__(deconv(in_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         deconv(mid_planes, mid_planes, kernel_size=5, stride=2), nn.ReLU(inplace=True),
                         deconv(mid_planes, out_planes, kernel_size=5, stride=2))


class HyperDecoderWithQReLU(nn.Module):
    """models/video/google.py:134-158: three deconvs, each followed by QReLU(bit_depth=8) = clamp(x, 0, 255)."""

    def __init__(self, in_planes: int = 192, mid_planes: int = 192, out_planes: int = 192):
        super().__init__()
        self.deconv1 = deconv(in_planes, mid_planes, kernel_size=5, stride=2)
        self.deconv2 = deconv(mid_planes, mid_planes, kernel_size=5, stride=2)
        self.deconv3 = deconv(mid_planes, out_planes, kernel_size=5, stride=2)
        self._q = [QReLU8()]   # in a list: not a registered child, like the reference's plain-function attributes

    def layers(self):
        q = self._q[0]
        return [self.deconv1, q, self.deconv2, q, self.deconv3, q]

    def forward(self, x: Tensor) -> Tensor:
        return run_layers(self.layers(), x, "nchw_f32", "nchw_f32")


class Hyperprior(CompressionModel):
    """models/video/google.py:160-208"""

    def __init__(self, planes: int = 192, mid_planes: int = 192):
        super().__init__(entropy_bottleneck_channels=mid_planes)
        self.hyper_encoder = HyperEncoder(planes, mid_planes, planes)
        self.hyper_decoder_mean = HyperDecoder(planes, mid_planes, planes)
        self.hyper_decoder_scale = HyperDecoderWithQReLU(planes, mid_planes, planes)
        self.gaussian_conditional = GaussianConditional(None)

    def _scales_means(self, z_hat_bf16_nhwc: Tensor):
        scales = run_layers(self.hyper_decoder_scale.layers(), z_hat_bf16_nhwc, "nhwc_bf16", "nhwc_f32")
        means = run_layers(list(self.hyper_decoder_mean), z_hat_bf16_nhwc, "nhwc_bf16", "nhwc_f32")
        return _nhwc_to_logical(scales), _nhwc_to_logical(means)

    def forward_internal(self, y: Tensor, y_bf16: Tensor):
        """y: fp32 NHWC, y_bf16: its bf16 copy -> (y_hat bf16 NHWC, {"y": lik, "z": lik})   (google.py:171-180)"""
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        z = run_layers(list(self.hyper_encoder), y_bf16, "nhwc_bf16", "nhwc_f32")
        z_l = _nhwc_to_logical(z)
        z_noise = torch.empty_like(z_l).uniform_(-0.5, 0.5) if self.training else None
        _, z_lik, z_hat_bf16 = ops.eb_forward(z_l, eb._params(), z_noise, eb._lik_bound(), want_bf16=True,
                                              lut=None if self.training else eb._eval_lut())
        scales, means = self._scales_means(z_hat_bf16.permute(0, 2, 3, 1))
        y_l = _nhwc_to_logical(y)
        bound, lb = gc.lower_bound_scale._sync_bound(), gc._lik_bound()
        if self.training:
            _, y_lik = ops.gc_forward(y_l, scales, means, torch.empty_like(y_l).uniform_(-0.5, 0.5), bound, lb)
            y_hat_bf16 = ops.to_bf16(ops.quantize_dequantize(y_l, means))     # quantize_ste(y - means) + means
        else:
            _, y_lik, y_hat_bf16 = ops.gc_forward(y_l, scales, means, None, bound, lb, want_bf16=True)
        return y_hat_bf16.permute(0, 2, 3, 1), {"y": y_lik, "z": z_lik}

    def forward(self, y: Tensor):
        y_nhwc, y_bf16 = _to_nhwc_pair(y)
        y_hat_bf16, lik = self.forward_internal(y_nhwc, y_bf16)
        return _nhwc_to_logical(y_hat_bf16).float(), lik

    def compress_internal(self, y: Tensor, y_bf16: Tensor):
        """google.py:182-195: (y_hat bf16 NHWC, {"strings": [y_string, z_string], "shape": z.size()[-2:]})"""
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        z = _nhwc_to_logical(run_layers(list(self.hyper_encoder), y_bf16, "nhwc_bf16", "nhwc_f32"))
        z_string = eb.compress(z)
        z_hat = eb.decompress(z_string, z.size()[-2:])
        scales, means = self._scales_means(_to_nhwc_bf16(z_hat))
        y_l = _nhwc_to_logical(y)
        indexes = gc.build_indexes(scales)
        y_string = gc.compress(y_l, indexes, means)
        y_hat = gc.quantize(y_l, "dequantize", means)
        return _to_nhwc_bf16(y_hat), {"strings": [y_string, z_string], "shape": z.size()[-2:]}

    def compress(self, y: Tensor):
        y_nhwc, y_bf16 = _to_nhwc_pair(y)
        y_hat_bf16, out = self.compress_internal(y_nhwc, y_bf16)
        return _nhwc_to_logical(y_hat_bf16).float(), out

    def decompress_internal(self, strings, shape) -> Tensor:
        """google.py:197-208 -> y_hat bf16 NHWC"""
        assert isinstance(strings, list) and len(strings) == 2
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        z_hat = eb.decompress(strings[1], shape)
        scales, means = self._scales_means(_to_nhwc_bf16(z_hat))
        indexes = gc.build_indexes(scales)
        y_hat = gc.decompress(strings[0], indexes, z_hat.dtype, means)
        return _to_nhwc_bf16(y_hat)

    def decompress(self, strings, shape):
        return _nhwc_to_logical(self.decompress_internal(strings, shape)).float()


class ScaleSpaceFlow(nn.Module):
    """models/video/google.py:55-508"""

    def __init__(self, num_levels: int = 5, sigma0: float = 1.5, scale_field_shift: float = 1.0):
        super().__init__()
        self.img_encoder = Encoder(3)
        self.img_decoder = Decoder(3)
        self.img_hyperprior = Hyperprior()
        self.res_encoder = Encoder(3)
        self.res_decoder = Decoder(3, in_planes=384)
        self.res_hyperprior = Hyperprior()
        self.motion_encoder = Encoder(2 * 3)
        self.motion_decoder = Decoder(2 + 1)
        self.motion_hyperprior = Hyperprior()
        self.sigma0 = sigma0
        self.num_levels = num_levels
        self.scale_field_shift = scale_field_shift
        self._grid_cache = {}
        for name, m in self.named_modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                m._mmc_name = name

    # ---- frame-level pieces ------------------------------------------------------------------------------
    def forward(self, frames):
        if not isinstance(frames, List):
            raise RuntimeError(f"Invalid number of frames: {len(frames)}.")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("ScaleSpaceFlow has no backward path yet (QReLU, Gaussian-volume and warp adjoints): "
                                      "call it under torch.no_grad()")
        reconstructions, frames_likelihoods = [], []
        x_hat, likelihoods = self.forward_keyframe(frames[0])
        reconstructions.append(x_hat)
        frames_likelihoods.append(likelihoods)
        x_ref = x_hat.detach()
        for i in range(1, len(frames)):
            x_ref, likelihoods = self.forward_inter(frames[i], x_ref)
            reconstructions.append(x_ref)
            frames_likelihoods.append(likelihoods)
        return {"x_hat": reconstructions, "likelihoods": frames_likelihoods}

    def forward_keyframe(self, x):
        """google.py:232-236"""
        y, y_bf16 = run_layers(list(self.img_encoder), x, "nchw_f32", "nhwc_f32", out2=2)
        y_hat_bf16, likelihoods = self.img_hyperprior.forward_internal(y, y_bf16)
        x_hat = run_layers(list(self.img_decoder), y_hat_bf16, "nhwc_bf16", "nchw_f32")
        return x_hat, {"keyframe": likelihoods}

    def encode_keyframe(self, x):
        y, y_bf16 = run_layers(list(self.img_encoder), x, "nchw_f32", "nhwc_f32", out2=2)
        y_hat_bf16, out_keyframe = self.img_hyperprior.compress_internal(y, y_bf16)
        x_hat = run_layers(list(self.img_decoder), y_hat_bf16, "nhwc_bf16", "nchw_f32")
        return x_hat, out_keyframe

    def decode_keyframe(self, strings, shape):
        y_hat_bf16 = self.img_hyperprior.decompress_internal(strings, shape)
        return run_layers(list(self.img_decoder), y_hat_bf16, "nhwc_bf16", "nchw_f32")

    def _inter(self, x_cur, x_ref, motion_fn, res_fn):
        """Shared body of forward_inter / encode_inter (google.py:246-309)."""
        ops._require_cuda(x_cur, x_ref)
        x = torch.cat((x_cur.float(), x_ref.float()), dim=1)
        y_motion, y_motion_bf16 = run_layers(list(self.motion_encoder), x, "nchw_f32", "nhwc_f32", out2=2)
        y_motion_hat, out_motion = motion_fn(y_motion, y_motion_bf16)
        motion_info = run_layers(list(self.motion_decoder), y_motion_hat, "nhwc_bf16", "nchw_f32")
        x_pred, x_res = self._prediction(x_ref, motion_info, x_cur)
        y_res, y_res_bf16 = run_layers(list(self.res_encoder), x_res, "nchw_f32", "nhwc_f32", out2=2)
        y_res_hat, out_res = res_fn(y_res, y_res_bf16)
        y_combine = torch.cat((y_res_hat, y_motion_hat), dim=-1)
        x_res_hat = run_layers(list(self.res_decoder), y_combine, "nhwc_bf16", "nchw_f32")
        return ops.add(x_pred, x_res_hat), out_motion, out_res

    def forward_inter(self, x_cur, x_ref):
        x_rec, motion_likelihoods, res_likelihoods = self._inter(x_cur, x_ref, self.motion_hyperprior.forward_internal,
                                                                 self.res_hyperprior.forward_internal)
        return x_rec, {"motion": motion_likelihoods, "residual": res_likelihoods}

    def encode_inter(self, x_cur, x_ref):
        x_rec, out_motion, out_res = self._inter(x_cur, x_ref, self.motion_hyperprior.compress_internal,
                                                 self.res_hyperprior.compress_internal)
        return x_rec, {"strings": {"motion": out_motion["strings"], "residual": out_res["strings"]},
                       "shape": {"motion": out_motion["shape"], "residual": out_res["shape"]}}

    def decode_inter(self, x_ref, strings, shapes):
        y_motion_hat = self.motion_hyperprior.decompress_internal(strings["motion"], shapes["motion"])
        motion_info = run_layers(list(self.motion_decoder), y_motion_hat, "nhwc_bf16", "nchw_f32")
        x_pred = self._prediction(x_ref, motion_info, None)
        y_res_hat = self.res_hyperprior.decompress_internal(strings["residual"], shapes["residual"])
        y_combine = torch.cat((y_res_hat, y_motion_hat), dim=-1)
        x_res_hat = run_layers(list(self.res_decoder), y_combine, "nhwc_bf16", "nchw_f32")
        return ops.add(x_pred, x_res_hat)

    # ---- scale-space prediction -----------------------------------------------------------------------------
    def _base_grid(self, H: int, W: int, device):
        """Rows of meshgrid2d (utils.py:192-195): computed once per frame size by the same torch call the reference makes."""
        key = (H, W, str(device))
        if key not in self._grid_cache:
            g = F.affine_grid(torch.eye(2, 3).unsqueeze(0), (1, 1, H, W), align_corners=False)
            self._grid_cache[key] = (g[0, 0, :, 0].contiguous().to(device), g[0, :, 0, 1].contiguous().to(device))
        return self._grid_cache[key]

    def gaussian_volume(self, x, sigma: float, num_levels: int):
        """google.py:331-355"""
        k = 2 * int(math.ceil(3 * sigma)) + 1
        return ops.gaussian_volume(x, gaussian_kernel1d(k, sigma), num_levels)

    def warp_volume(self, volume, flow, scale_field, padding_mode: str = "border"):
        """google.py:357-375"""
        if volume.ndimension() != 5:
            raise ValueError(f"Invalid number of dimensions for volume {volume.ndimension()}")
        if padding_mode != "border":
            raise NotImplementedError("only border padding is on the accelerated path")
        N, C, _, H, W = volume.size()
        bx, by = self._base_grid(H, W, volume.device)
        return ops.scale_space_warp(volume, torch.cat((flow, scale_field), dim=1), bx, by)

    def _prediction(self, x_ref, motion_info, x_cur=None):
        volume = self.gaussian_volume(x_ref, self.sigma0, self.num_levels)
        bx, by = self._base_grid(x_ref.shape[2], x_ref.shape[3], x_ref.device)
        return ops.scale_space_warp(volume, motion_info, bx, by, x_cur)

    def forward_prediction(self, x_ref, motion_info):
        """google.py:377-382"""
        return self._prediction(x_ref, motion_info, None)

    # ---- sequence-level API ---------------------------------------------------------------------------------
    def aux_loss(self):
        return [m.aux_loss() for m in self.modules() if isinstance(m, CompressionModel)]

    def compress(self, frames):
        if not isinstance(frames, List):
            raise RuntimeError(f"Invalid number of frames: {len(frames)}.")
        frame_strings, shape_infos = [], []
        x_ref, out_keyframe = self.encode_keyframe(frames[0])
        frame_strings.append(out_keyframe["strings"])
        shape_infos.append(out_keyframe["shape"])
        for i in range(1, len(frames)):
            x_ref, out_interframe = self.encode_inter(frames[i], x_ref)
            frame_strings.append(out_interframe["strings"])
            shape_infos.append(out_interframe["shape"])
        return frame_strings, shape_infos

    def decompress(self, strings, shapes):
        if not isinstance(strings, List) or not isinstance(shapes, List):
            raise RuntimeError(f"Invalid number of frames: {len(strings)}.")
        assert len(strings) == len(shapes), f"Number of information should match {len(strings)} != {len(shapes)}."
        dec_frames = []
        x_ref = self.decode_keyframe(strings[0], shapes[0])
        dec_frames.append(x_ref)
        for i in range(1, len(strings)):
            x_ref = self.decode_inter(x_ref, strings[i], shapes[i])
            dec_frames.append(x_ref)
        return dec_frames

    def load_state_dict(self, state_dict, strict: bool = True):
        """google.py:458-500"""
        for hp in ("img_hyperprior", "res_hyperprior", "motion_hyperprior"):
            m = getattr(self, hp)
            _resize_registered_buffers(m.gaussian_conditional, f"{hp}.gaussian_conditional",
                                       ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
            _resize_registered_buffers(m.entropy_bottleneck, f"{hp}.entropy_bottleneck",
                                       ["_quantized_cdf", "_offset", "_cdf_length"], state_dict)
        return nn.Module.load_state_dict(self, state_dict, strict=strict)

    @classmethod
    def from_state_dict(cls, state_dict):
        net = cls()
        net.load_state_dict(state_dict)
        return net

    def update(self, scale_table=None, force=False):
        """google.py:502-508 (+ the per-hyperprior update of CompressionModel)"""
        if scale_table is None:
            scale_table = get_scale_table()
        updated = False
        for hp in (self.img_hyperprior, self.res_hyperprior, self.motion_hyperprior):
            updated |= hp.gaussian_conditional.update_scale_table(scale_table, force=force)
            updated |= hp.update(force=force)
        return updated

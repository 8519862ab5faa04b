"""Multi-modality (RGB + depth / thermal) two-branch codec -- host-side mirror of the fork's own models
``JointAutoregressiveHierarchicalPriors_R`` (guide / RGB branch, exposes six hidden feature maps) and
``JointAutoregressiveHierarchicalPriors_D`` (second-modality branch, fuses those maps through
``eg_ext* -> tran_conv* -> ESA``), compressai/models/google.py:696-1459, with the reference's constructor arguments,
attribute names and ``state_dict`` keys.

Every conv / deconv (+GDN / IGDN / ReLU / LeakyReLU), the masked context convolution, the 1x1 entropy-parameter
convs and the entropy stage run on the libmmcodec kernels with NHWC bf16 activations; channel concatenations are
views-plus-one-copy on the NHWC tensors.  The ESA gate (1x1 -> 3x3 s2 p0 -> maxpool 7/3 -> 3x3 x3 -> bilinear
upsample -> 1x1 -> sigmoid, google.py:1432-1459) runs entirely in libmmcodec in inference (convs on
the tensor-core kernel, pool / upsample+add / gate in csrc/esa.cu); under autograd it runs on torch ops (bf16, channels-last).
``forward`` is implemented for eval and for the training-mode forward pass
(uniform-noise quantisation); the autoregressive ``compress`` / ``decompress`` (google.py:836-1003, a serial
per-pixel Python loop in the reference) are out of scope and raise.
"""
from __future__ import annotations

import os
from typing import Any, Dict

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import autograd as AG
from . import ops
from .entropy_models import GaussianConditional
from .layers import GDN, Conv2d, conv, deconv
from .models import MeanScaleHyperprior, _nhwc_to_logical
from .transforms import TransformStack, run_layers

__all__ = ["MaskedConv2d", "ESA", "Encoder1", "Decoder1", "JointAutoregressiveHierarchicalPriors", "JointAutoregressiveHierarchicalPriors_R", "Guided_compresser",
           "JointAutoregressiveHierarchicalPriors_D"]


class MaskedConv2d(Conv2d):
    """Masked 2-D convolution (compressai/layers/layers.py:52-78): the mask is folded into the weights, which is
    exactly what the reference does in place before every forward (``self.weight.data *= self.mask``)."""

    def __init__(self, *args: Any, mask_type: str = "A", **kwargs: Any):
        super().__init__(*args, **kwargs)
        if mask_type not in ("A", "B"):
            raise ValueError(f'Invalid "mask_type" value "{mask_type}"')
        self.register_buffer("mask", torch.ones_like(self.weight.data))
        _, _, h, w = self.mask.size()
        self.mask[:, :, h // 2, w // 2 + (mask_type == "B"):] = 0
        self.mask[:, :, h // 2 + 1:] = 0

    def _apply_mask(self):
        with torch.no_grad():
            self.weight.mul_(self.mask)   # idempotent; bumps the version so the packed copy is refreshed once

    def packed_weight(self, d):
        if getattr(self, "_masked_version", None) != (self.weight._version, self.weight.data_ptr()):
            self._apply_mask()
            self._masked_version = (self.weight._version, self.weight.data_ptr())
        return super().packed_weight(d)

    def f32_weight(self):
        self._apply_mask()
        return super().f32_weight()


class ESA(nn.Module):
    """Enhanced spatial attention gate (google.py:1432-1459).

    Every convolution (1x1 conv1 / conv_f / conv4, 3x3 conv_max / conv3 / conv3_, and the stride-2 3x3 conv2) runs on the
    tensor-core conv kernel with NHWC bf16 activations, forward and backward; the max-pool (7 / 3), the bilinear upsampling fused
    with the add, and the sigmoid gate are single libmmcodec passes (csrc/esa.cu), with their adjoints in csrc/fusion_bwd.cu when
    the gate trains on the kernels (``train_on_kernels``; otherwise the whole gate runs on torch's bf16 ops).  conv2 has padding 0,
    which the kernel's padding-k/2 addressing expresses exactly as the padding-1 convolution of the map shifted by one pixel:
    conv_p0(x)[o] = conv_p1(pad_top_left(x))[o + 1]."""

    def __init__(self, n_feats: int):
        super().__init__()
        f = n_feats // 4
        self.conv1 = Conv2d(n_feats, f, 1)
        self.conv_f = Conv2d(f, f, 1)
        self.conv_max = Conv2d(f, f, 3, padding=1)
        self.conv2 = Conv2d(f, f, 3, stride=2, padding=0)
        self.conv3 = Conv2d(f, f, 3, padding=1)
        self.conv3_ = Conv2d(f, f, 3, padding=1)
        self.conv4 = Conv2d(f, n_feats, 1)
        self.sigmoid = nn.Sigmoid()
        self.relu = nn.ReLU(inplace=True)

    def _conv2_p1(self) -> nn.Module:
        """conv2 as a padding-1 layer sharing conv2's parameters (not a registered child: the state_dict is the reference's)."""
        m = getattr(self, "_conv2_alias", None)
        if m is None:
            c = self.conv2
            m = Conv2d(c.in_channels, c.out_channels, 3, stride=2, padding=1)
            m._mmc_name = getattr(c, "_mmc_name", "conv2")
            object.__setattr__(self, "_conv2_alias", m)
        m._parameters["weight"], m._parameters["bias"] = self.conv2.weight, self.conv2.bias
        return m

    # Training keeps the gate on torch's bf16 ops: its seven small convolutions per block (42 layers in the depth branch) each cost
    # ~5 launches in the backward pass and made the step host-bound (measured 52.3 vs 42.9 ms per 4-pair step); the kernel path
    # below is what inference uses (19.0 -> 17.7 ms per 8 pairs).  Both are covered by the parity tests.
    train_on_kernels = os.environ.get("MMC_ESA_TRAIN_KERNELS", "0") == "1"

    def _forward_torch(self, x: Tensor) -> Tensor:
        """(B, C, H, W) channels-last view -> same, stock torch ops (google.py:1445-1459)."""
        conv = lambda m, t: nn.Conv2d.forward(m, t)      # the reference's op, not the fused executor
        c1_ = conv(self.conv1, x)
        c1 = conv(self.conv2, c1_)
        v_max = F.max_pool2d(c1, kernel_size=7, stride=3)
        v_range = self.relu(conv(self.conv_max, v_max))
        c3 = self.relu(conv(self.conv3, v_range))
        c3 = conv(self.conv3_, c3)
        c3 = F.interpolate(c3, (x.size(2), x.size(3)), mode="bilinear", align_corners=False)
        cf = conv(self.conv_f, c1_)
        return x * self.sigmoid(conv(self.conv4, c3 + cf))

    def forward_nhwc_bf16(self, x: Tensor) -> Tensor:
        """x: (B, H, W, C) bf16 -> same (records an autograd graph when x or the parameters require grad)."""
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if needs_grad and not self.train_on_kernels:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = self._forward_torch(x.permute(0, 3, 1, 2))
            return y.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous()
        B, H, W, _ = x.shape
        c1_ = run_layers([self.conv1], x, "nhwc_bf16", "nhwc_bf16")
        ho, wo = (H - 3) // 2 + 1, (W - 3) // 2 + 1
        c1 = run_layers([self._conv2_p1()], F.pad(c1_, (0, 0, 1, 0, 1, 0)), "nhwc_bf16", "nhwc_bf16")[:, 1:1 + ho, 1:1 + wo]
        if needs_grad:
            # training on the kernels: the same pool / upsample + add / gate passes as inference, recorded with their adjoints
            # (csrc/fusion_bwd.cu: arg-max gather, bilinear adjoint, gate backward)
            v_max = AG.maxpool_nhwc(c1.contiguous(), 7, 3)
            c3 = run_layers([self.conv_max, self.relu, self.conv3, self.relu, self.conv3_], v_max, "nhwc_bf16", "nhwc_bf16")
            cf = run_layers([self.conv_f], c1_, "nhwc_bf16", "nhwc_bf16")
            c4 = run_layers([self.conv4], AG.upsample_add(c3, cf), "nhwc_bf16", "nhwc_bf16")
            return AG.sigmoid_gate(x, c4)
        # inference: pool, upsample + add and the gate are single libmmcodec passes (csrc/esa.cu)
        v_max = ops.maxpool_nhwc_bf16(c1, 7, 3)
        c3 = run_layers([self.conv_max, self.relu, self.conv3, self.relu, self.conv3_], v_max, "nhwc_bf16", "nhwc_bf16")
        cf = run_layers([self.conv_f], c1_, "nhwc_bf16", "nhwc_bf16")
        c4 = run_layers([self.conv4], ops.upsample_bilinear_add_bf16(c3, cf), "nhwc_bf16", "nhwc_bf16")
        return ops.sigmoid_gate_bf16(x, c4)

    def forward(self, x: Tensor) -> Tensor:
        """(B, C, H, W) fp32 in / out, as the reference's module."""
        return _nhwc_to_logical(self.forward_nhwc_bf16(_to_nhwc_bf16(x))).float()


class Encoder1(nn.Module):
    """google.py:696-718: g_a with its three post-GDN feature maps exposed."""

    def __init__(self, N, M, channel=3, first_stride=2, **kwargs):
        super().__init__()
        self.g_a_conv1 = conv(channel, N, kernel_size=5, stride=first_stride)
        self.g_a_gdn1 = GDN(N)
        self.g_a_conv2 = conv(N, N, kernel_size=5, stride=2)
        self.g_a_gdn2 = GDN(N)
        self.g_a_conv3 = conv(N, N, kernel_size=5, stride=2)
        self.g_a_gdn3 = GDN(N)
        self.g_a_conv4 = conv(N, M, kernel_size=5, stride=2)

    def forward_internal(self, x: Tensor):
        """fp32 NCHW image -> (y fp32 NHWC, y bf16 NHWC, g1, g2, g3 bf16 NHWC)."""
        g1 = run_layers([self.g_a_conv1, self.g_a_gdn1], x, "nchw_f32", "nhwc_bf16")
        g2 = run_layers([self.g_a_conv2, self.g_a_gdn2], g1, "nhwc_bf16", "nhwc_bf16")
        g3 = run_layers([self.g_a_conv3, self.g_a_gdn3], g2, "nhwc_bf16", "nhwc_bf16")
        y, y_bf16 = run_layers([self.g_a_conv4], g3, "nhwc_bf16", "nhwc_f32", out2=2)
        return y, y_bf16, g1, g2, g3

    def forward(self, x):
        y, _, g1, g2, g3 = self.forward_internal(x)
        return tuple(_nhwc_to_logical(t) for t in (y, g1, g2, g3))


class Decoder1(nn.Module):
    """google.py:720-742"""

    def __init__(self, N, M, channel=3, first_stride=2, **kwargs):
        super().__init__()
        self.g_s_conv1 = deconv(M, N, kernel_size=5, stride=2)
        self.g_s_gdn1 = GDN(N, inverse=True)
        self.g_s_conv2 = deconv(N, N, kernel_size=5, stride=2)
        self.g_s_gdn2 = GDN(N, inverse=True)
        self.g_s_conv3 = deconv(N, N, kernel_size=5, stride=2)
        self.g_s_gdn3 = GDN(N, inverse=True)
        self.g_s_conv4 = deconv(N, channel, kernel_size=5, stride=first_stride)

    def forward_internal(self, y_hat_bf16: Tensor):
        g1 = run_layers([self.g_s_conv1, self.g_s_gdn1], y_hat_bf16, "nhwc_bf16", "nhwc_bf16")
        g2 = run_layers([self.g_s_conv2, self.g_s_gdn2], g1, "nhwc_bf16", "nhwc_bf16")
        g3 = run_layers([self.g_s_conv3, self.g_s_gdn3], g2, "nhwc_bf16", "nhwc_bf16")
        x_hat = run_layers([self.g_s_conv4], g3, "nhwc_bf16", "nchw_f32")
        return x_hat, g1, g2, g3

    def forward(self, y_hat):
        ops._require_cuda(y_hat)
        x_hat, g1, g2, g3 = self.forward_internal(_to_nhwc_bf16(y_hat))
        return (x_hat,) + tuple(_nhwc_to_logical(t) for t in (g1, g2, g3))


def _to_nhwc_bf16(t: Tensor) -> Tensor:
    """Logical (B, C, H, W) tensor of any float dtype / layout -> contiguous (B, H, W, C) bf16."""
    ops._require_cuda(t)
    if t.dtype == torch.bfloat16 and ops._is_channels_last(t):
        return t.permute(0, 2, 3, 1)
    if t.dtype == torch.float32 and t.is_contiguous():
        return ops.nchw_to_nhwc_bf16(t)
    if t.dtype == torch.float32 and ops._is_channels_last(t):
        return ops.to_bf16(t).permute(0, 2, 3, 1)
    return ops.to_bf16(t.float().contiguous(memory_format=torch.channels_last)).permute(0, 2, 3, 1)


class _ContextModelMixin:
    """Entropy stage shared by the two branches (google.py:800-822 / 1196-1211): hyperprior + masked context conv
    + 1x1 entropy-parameter convs -> (scales, means) -> Gaussian likelihood of y."""

    def _entropy_stage(self, y: Tensor, y_bf16: Tensor):
        eb, gc = self.entropy_bottleneck, self.gaussian_conditional
        M = self.M
        z = run_layers(list(self.h_a), y_bf16, "nhwc_bf16", "nhwc_f32")
        z_l = _nhwc_to_logical(z)
        noise = getattr(self, "_noise_override", None) or {}      # tests pass the oracle's noise tensors
        draw = lambda key, like: noise[key].to(like.device) if key in noise else torch.empty_like(like).uniform_(-0.5, 0.5)
        if self.training:
            z_hat, z_lik = AG.eb_forward(z_l, eb, draw("z", z_l))
            z_hat_bf16 = AG.cast_bf16(z_hat)
        else:
            _, z_lik, z_hat_bf16 = ops.eb_forward(z_l, eb._params(), None, eb._lik_bound(), want_bf16=True, lut=eb._eval_lut())
        params = run_layers(list(self.h_s), z_hat_bf16.permute(0, 2, 3, 1), "nhwc_bf16", "nhwc_bf16")
        # y_hat = quantize(y, noise | dequantize) WITHOUT means (google.py:805-807): this is what the context model
        # and the synthesis transform see, while the likelihood below is evaluated at round(y - mu) + mu
        y_l = _nhwc_to_logical(y)
        if self.training:
            y_hat = AG.add_noise(y_l, draw("y_hat", y_l))
        else:
            y_hat = ops.quantize_dequantize(y_l)
        y_hat_bf16 = AG.cast_bf16(y_hat).permute(0, 2, 3, 1)
        ctx = run_layers([self.context_prediction], y_hat_bf16, "nhwc_bf16", "nhwc_bf16")
        gp = run_layers(list(self.entropy_parameters), (params, ctx), "nhwc_bf16", "nhwc_f32")
        scales_hat, means_hat = _nhwc_to_logical(gp[..., :M]), _nhwc_to_logical(gp[..., M:])
        y_noise = draw("y", y_l) if self.training else None
        _, y_lik = AG.gc_forward(y_l, scales_hat, means_hat, y_noise, gc.lower_bound_scale._sync_bound(), gc._lik_bound())
        return y_hat_bf16, y_lik, z_lik

    def _init_entropy_stage(self, N, M):
        self.h_a = TransformStack(conv(M, N, stride=1, kernel_size=3), nn.LeakyReLU(inplace=True),
                                  conv(N, N, stride=2, kernel_size=5), nn.LeakyReLU(inplace=True),
                                  conv(N, N, stride=2, kernel_size=5))
        self.h_s = TransformStack(deconv(N, M, stride=2, kernel_size=5), nn.LeakyReLU(inplace=True),
                                  deconv(M, M * 3 // 2, stride=2, kernel_size=5), nn.LeakyReLU(inplace=True),
                                  conv(M * 3 // 2, M * 2, stride=1, kernel_size=3))
        self.entropy_parameters = TransformStack(Conv2d(M * 12 // 3, M * 10 // 3, 1), nn.LeakyReLU(inplace=True),
                                                 Conv2d(M * 10 // 3, M * 8 // 3, 1), nn.LeakyReLU(inplace=True),
                                                 Conv2d(M * 8 // 3, M * 6 // 3, 1))
        self.context_prediction = MaskedConv2d(M, 2 * M, kernel_size=5, padding=2, stride=1)
        self.gaussian_conditional = GaussianConditional(None)

    def compress(self, *a, **k):
        raise NotImplementedError("autoregressive compress() (serial per-pixel context loop, google.py:836-876) is out of scope")

    def decompress(self, *a, **k):
        raise NotImplementedError("autoregressive decompress() (google.py:920-1003) is out of scope")


class JointAutoregressiveHierarchicalPriors(_ContextModelMixin, MeanScaleHyperprior):
    """The zoo's mbt2018 model (google.py:421-520): stock g_a / g_s around the hyperprior + masked-context entropy stage.
    ``forward`` (eval and training) runs on the kernels; the serial autoregressive coder (google.py:565-692) is out of scope."""

    def __init__(self, N=192, M=192, channel=3, **kwargs):
        super().__init__(N=N, M=M, **kwargs)
        self.g_a = TransformStack(conv(channel, N, kernel_size=5, stride=2), GDN(N), conv(N, N, kernel_size=5, stride=2), GDN(N),
                                  conv(N, N, kernel_size=5, stride=2), GDN(N), conv(N, M, kernel_size=5, stride=2))
        self.g_s = TransformStack(deconv(M, N, kernel_size=5, stride=2), GDN(N, inverse=True), deconv(N, N, kernel_size=5, stride=2),
                                  GDN(N, inverse=True), deconv(N, N, kernel_size=5, stride=2), GDN(N, inverse=True),
                                  deconv(N, channel, kernel_size=5, stride=2))
        self._init_entropy_stage(N, M)
        self.N, self.M = int(N), int(M)
        self._tag_layer_names()

    @property
    def downsampling_factor(self) -> int:
        return 2 ** (4 + 2)

    def forward(self, x):
        y, y_bf16 = run_layers(list(self.g_a), x, "nchw_f32", "nhwc_f32", out2=2)
        y_hat_bf16, y_lik, z_lik = self._entropy_stage(y, y_bf16)
        x_hat = run_layers(list(self.g_s), y_hat_bf16, "nhwc_bf16", "nchw_f32")
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}}

    @classmethod
    def from_state_dict(cls, state_dict, channel=3):
        """google.py:522-529"""
        N = state_dict["g_a.0.weight"].size(0)
        M = state_dict["g_a.6.weight"].size(0)
        net = cls(N, M, channel)
        net.load_state_dict(state_dict)
        return net


class JointAutoregressiveHierarchicalPriors_R(_ContextModelMixin, MeanScaleHyperprior):
    """Guide (RGB) branch, google.py:746-825.  ``forward`` also returns the six hidden maps the second branch fuses;
    they are logical (B, N, H, W) bf16 tensors in channels-last memory (the kernels' native activation format)."""

    def __init__(self, N=192, M=192, channel=3, first_stride=2, **kwargs):
        super().__init__(N=N, M=M, **kwargs)
        self.first_stride = first_stride
        self.enc1 = Encoder1(N, M, channel, first_stride)
        self.dec1 = Decoder1(N, M, channel, first_stride)
        self._init_entropy_stage(N, M)
        self.N = int(N)
        self.M = int(M)
        self._tag_layer_names()

    def forward(self, x):
        y, y_bf16, ga1, ga2, ga3 = self.enc1.forward_internal(x)
        y_hat_bf16, y_lik, z_lik = self._entropy_stage(y, y_bf16)
        x_hat, gs1, gs2, gs3 = self.dec1.forward_internal(y_hat_bf16)
        hidden = {k: _nhwc_to_logical(v) for k, v in (("ga1", ga1), ("ga2", ga2), ("ga3", ga3),
                                                    ("gs1", gs1), ("gs2", gs2), ("gs3", gs3))}
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}, "hidden": hidden}


class Guided_compresser(JointAutoregressiveHierarchicalPriors_R):
    """The guide-modality codec of the RGB-T reproduction (compressai/models/master.py:1167-1300): the same network as
    ``JointAutoregressiveHierarchicalPriors_R`` with a configurable input channel count (default 1: thermal / depth) and
    first-layer stride; same sub-module names, ``state_dict`` keys and ``forward`` result (incl. the six hidden maps)."""

    def __init__(self, N=192, M=192, channel=1, first_stride=2, **kwargs):
        super().__init__(N=N, M=M, channel=channel, first_stride=first_stride, **kwargs)

    @property
    def downsampling_factor(self) -> int:
        return 2 ** (4 + 2)


class JointAutoregressiveHierarchicalPriors_D(_ContextModelMixin, MeanScaleHyperprior):
    """Second-modality (depth / thermal, 1 channel) branch with cross-modality fusion, google.py:1006-1248."""

    def __init__(self, N=192, M=192, **kwargs):
        super().__init__(N=N, M=M, **kwargs)
        self.pic2_g_a_conv1 = conv(1, N)
        self.pic2_g_a_gdn1 = GDN(N)
        self.pic2_g_a_conv2 = conv(2 * N, N)
        self.pic2_g_a_gdn2 = GDN(N)
        self.pic2_g_a_conv3 = conv(2 * N, N)
        self.pic2_g_a_gdn3 = GDN(N)
        self.pic2_g_a_conv4 = conv(2 * N, M)
        self.pic2_g_s_conv1 = deconv(M, N)
        self.pic2_g_s_gdn1 = GDN(N, inverse=True)
        self.pic2_g_s_conv2 = deconv(2 * N, N)
        self.pic2_g_s_gdn2 = GDN(N, inverse=True)
        self.pic2_g_s_conv3 = deconv(2 * N, N)
        self.pic2_g_s_gdn3 = GDN(N, inverse=True)
        self.pic2_g_s_conv4 = deconv(2 * N, 1)
        for i in range(1, 7):
            setattr(self, f"tran_conv{i}", conv(2 * N, N, stride=1))
        self._init_entropy_stage(N, M)
        self.N = int(N)
        self.M = int(M)
        for i in range(1, 7):
            setattr(self, f"r{i}", nn.ReLU())
        for i in range(1, 7):
            setattr(self, f"attention{i}", ESA(N))
        for i in range(1, 13):
            setattr(self, f"eg_ext{i}", TransformStack(Conv2d(N, N, stride=1, kernel_size=3, padding=1), nn.ReLU(inplace=True)))
        self._tag_layer_names()

    def _fuse(self, i: int, x: Tensor, guide: Tensor) -> Tensor:
        """eg_ext(2i-1)(x), eg_ext(2i)(guide) -> cat -> tran_conv_i -> ESA_i   (google.py:1151-1156 and five repeats)"""
        e_own = run_layers(list(getattr(self, f"eg_ext{2 * i - 1}")), x, "nhwc_bf16", "nhwc_bf16")
        e_guide = run_layers(list(getattr(self, f"eg_ext{2 * i}")), guide, "nhwc_bf16", "nhwc_bf16")
        f = run_layers([getattr(self, f"tran_conv{i}")], (e_own, e_guide), "nhwc_bf16", "nhwc_bf16")
        return getattr(self, f"attention{i}").forward_nhwc_bf16(f)

    def _analysis(self, x, g):
        """pic2_g_a with the three encoder-side fusions (google.py:1148-1194): fp32 NCHW depth map -> (y fp32, y bf16), NHWC."""
        a = run_layers([self.pic2_g_a_conv1, self.pic2_g_a_gdn1], x, "nchw_f32", "nhwc_bf16")
        f1 = self._fuse(1, a, g["ga1"])
        a = run_layers([self.pic2_g_a_conv2, self.pic2_g_a_gdn2], (a, f1), "nhwc_bf16", "nhwc_bf16")
        f2 = self._fuse(2, a, g["ga2"])
        a = run_layers([self.pic2_g_a_conv3, self.pic2_g_a_gdn3], (a, f2), "nhwc_bf16", "nhwc_bf16")
        f3 = self._fuse(3, a, g["ga3"])
        return run_layers([self.pic2_g_a_conv4], (a, f3), "nhwc_bf16", "nhwc_f32", out2=2)

    def _synthesis(self, y_hat_bf16, g):
        """pic2_g_s with the three decoder-side fusions (google.py:1213-1246): y_hat bf16 NHWC -> x_hat fp32 NCHW."""
        s = run_layers([self.pic2_g_s_conv1, self.pic2_g_s_gdn1], y_hat_bf16, "nhwc_bf16", "nhwc_bf16")
        f4 = self._fuse(4, s, g["gs1"])
        s = run_layers([self.pic2_g_s_conv2, self.pic2_g_s_gdn2], (s, f4), "nhwc_bf16", "nhwc_bf16")
        f5 = self._fuse(5, s, g["gs2"])
        s = run_layers([self.pic2_g_s_conv3, self.pic2_g_s_gdn3], (s, f5), "nhwc_bf16", "nhwc_bf16")
        f6 = self._fuse(6, s, g["gs3"])
        return run_layers([self.pic2_g_s_conv4], (s, f6), "nhwc_bf16", "nchw_f32")

    def forward(self, x, hidden: Dict[str, Tensor]):
        ops._require_cuda(x)
        g = {k: _to_nhwc_bf16(hidden[k]) for k in ("ga1", "ga2", "ga3", "gs1", "gs2", "gs3")}
        y, y_bf16 = self._analysis(x, g)
        y_hat_bf16, y_lik, z_lik = self._entropy_stage(y, y_bf16)
        x_hat = self._synthesis(y_hat_bf16, g)
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}}


# zoo registration (compressai/zoo/image.py:56,220-229): build_model("mbt2018", q)
from . import models as _models  # noqa: E402

_models.MODELS["mbt2018"] = JointAutoregressiveHierarchicalPriors
_models.CFGS["mbt2018"] = {q: ((192, 192) if q <= 4 else (192, 320)) for q in range(1, 9)}

"""RGB-T paper reproduction, master side -- host-side mirror of ``Master_compresser`` and its building blocks
(compressai/models/master.py:29-216, 386-951) with the reference's constructor arguments, attribute names and ``state_dict``
keys.  Its guide codec ``Guided_compresser`` lives in models_mm.py (it is the ``_R`` network).

Data flow (master.py:904-951): the master image and the DECODED guide image go through ``Feature_encoder`` (3x3 conv + three
residual blocks, 64 channels); ``Channel_aligner`` predicts a per-sample, per-channel affine (gamma from the guide features,
beta from the master features: shared 4-conv trunk, global average) and applies it to the guide features; the 128-channel
concatenation is coded by g_a / hyperprior / masked context model exactly like the two-branch codec; ``Master_decoder``
upsamples with deconv+IGDN and, after every stage, ``Spatial_aligner`` cross-attends (4x4 windows, 3 heads, dim 96, second
block shifted by 2) from the master map to the guide codec's hidden map of the same scale; ``Feature_decoder`` reconstructs.

Every conv / deconv (+GDN / IGDN / LeakyReLU), the entropy stage and -- in inference -- the token-wise Linear layers of the
attention blocks (1x1 tensor-core convs on the token grid: LayerNorm, Linear and GELU are per-token, so they commute with the
cyclic shift and the window partition), the LayerNorms and the window attention itself (``mmc_window_attention``: shift,
partition, QK^T + relative-position bias + shift mask, softmax, PV and the inverse permutation in one kernel) run in
libmmcodec.  Under autograd the same kernels run, each step recorded with its backward kernel (mmcodec.autograd, csrc/fusion_bwd.cu);
``attention_train_on_kernels = False`` restores the torch-op training path (bf16 autocast) the tests compare against.  ``compress`` / ``decompress`` (serial per-pixel context loop) are out of scope.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import _lib as L
from . import autograd as AG
from . import ops
from . import transforms as T
from .layers import GDN, Conv2d, conv, deconv
from .models import MeanScaleHyperprior, _nhwc_to_logical
from .models_mm import _ContextModelMixin, _to_nhwc_bf16
from .transforms import TransformStack, run_layers

__all__ = ["ResidualBlock", "Feature_encoder", "Feature_decoder", "Channel_aligner", "PatchEmbed", "Mlp", "WindowAttention",
           "SwinTransformerBlock", "Spatial_aligner", "Master_decoder", "Master_compresser"]

# The attention blocks run on the libmmcodec kernels; tests flip this to cross-check against the torch-op path.
attention_on_kernels = True
# ... under autograd too (backward kernels in csrc/fusion_bwd.cu); False = the round-1 training path (torch ops in bf16 autocast)
attention_train_on_kernels = True


def _attn_kernels() -> bool:
    return attention_on_kernels and (attention_train_on_kernels or not torch.is_grad_enabled())


def _conv3x3(cin, cout, stride=1):
    return conv(cin, cout, kernel_size=3, stride=stride)


def _conv1x1(cin, cout):
    return conv(cin, cout, kernel_size=1, stride=1)


def _cat_if_training(x):
    """A (x1, x2) channel pair stays a pair for the two-source kernels (forward and, with ``transforms.two_source_train``,
    backward); otherwise it is concatenated once under autograd."""
    if isinstance(x, (tuple, list)) and torch.is_grad_enabled() and not T.two_source_train:
        return torch.cat(tuple(x), dim=-1)
    return x


class ResidualBlock(nn.Module):
    """master.py:29-62: conv3x3 -> LeakyReLU -> conv3x3 -> LeakyReLU, plus identity (1x1 ``skip`` conv if widths differ)."""

    def __init__(self, in_ch: int, out_ch: int):
        super().__init__()
        self.conv1 = _conv3x3(in_ch, out_ch)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = _conv3x3(out_ch, out_ch)
        self.skip = _conv1x1(in_ch, out_ch) if in_ch != out_ch else None

    def forward_nhwc(self, x):
        """NHWC bf16 (or a channel pair of them) -> NHWC bf16."""
        body = run_layers([self.conv1, self.leaky_relu, self.conv2, self.leaky_relu], x, "nhwc_bf16", "nhwc_bf16")
        if self.skip is not None:
            return body + run_layers([self.skip], x, "nhwc_bf16", "nhwc_bf16")
        return body + x

    def forward(self, x: Tensor) -> Tensor:
        return _nhwc_to_logical(self.forward_nhwc(_to_nhwc_bf16(x)))


class Feature_encoder(nn.Module):
    """master.py:68-89"""

    def __init__(self, in_channel=3, out_channel=64, stride=1) -> None:
        super().__init__()
        self.conv1 = _conv3x3(in_channel, out_channel, stride)
        self.resblock1 = ResidualBlock(64, 64)
        self.resblock2 = ResidualBlock(64, 64)
        self.resblock3 = ResidualBlock(64, 64)

    def forward_nhwc(self, x: Tensor) -> Tensor:
        """fp32 NCHW image -> NHWC bf16 features."""
        first = run_layers([self.conv1], x, "nchw_f32", "nhwc_bf16")
        out = first
        for blk in (self.resblock1, self.resblock2, self.resblock3):
            out = blk.forward_nhwc(out)
        return out + first

    def forward(self, x: Tensor) -> Tensor:
        return _nhwc_to_logical(self.forward_nhwc(x))


class Feature_decoder(nn.Module):
    """master.py:101-118"""

    def __init__(self, in_channel=64 * 3, out_channel=3, stride=1) -> None:
        super().__init__()
        self.resblock1 = ResidualBlock(in_channel, 64)
        self.resblock2 = ResidualBlock(64, 64)
        self.resblock3 = ResidualBlock(64, 64)
        self.deconv1 = deconv(64, out_channel, kernel_size=3, stride=stride)
        self.conv = _conv1x1(in_channel, 64)

    def forward_nhwc(self, x) -> Tensor:
        """NHWC bf16 features (or a channel pair) -> fp32 logical NCHW reconstruction."""
        x = _cat_if_training(x)
        out = self.resblock1.forward_nhwc(x)
        out = self.resblock2.forward_nhwc(out)
        out = self.resblock3.forward_nhwc(out)
        out = out + run_layers([self.conv], x, "nhwc_bf16", "nhwc_bf16")
        return run_layers([self.deconv1], out, "nhwc_bf16", "nchw_f32")

    def forward(self, x: Tensor) -> Tensor:
        return self.forward_nhwc(_to_nhwc_bf16(x))


class Channel_aligner(nn.Module):
    """master.py:158-210.  The 4-conv trunk is shared by the two feature maps, so both go through it as ONE batch of 2B.  In
    inference the two heads (conv5 / conv6 followed by a global average) are evaluated by linearity from channel sums of the trunk
    output (``heads_by_linearity``; set False to run the convolutions and average their outputs, which the tests compare)."""

    heads_by_linearity = True

    def __init__(self) -> None:
        super().__init__()
        self.conv1 = _conv3x3(64, 256)
        self.leaky_relu1 = nn.LeakyReLU(inplace=True)
        self.conv2 = _conv3x3(256, 256)
        self.leaky_relu2 = nn.LeakyReLU(inplace=True)
        self.conv3 = _conv3x3(256, 256)
        self.leaky_relu3 = nn.LeakyReLU(inplace=True)
        self.conv4 = _conv3x3(256, 256)
        self.leaky_relu4 = nn.LeakyReLU(inplace=True)
        self.conv5 = _conv3x3(256, 64)
        self.conv6 = _conv3x3(256, 64)
        self.avgpool1 = nn.AdaptiveAvgPool2d(1)
        self.avgpool2 = nn.AdaptiveAvgPool2d(1)

    def forward_nhwc(self, feature1: Tensor, feature2: Tensor):
        """(master features, guide features) NHWC bf16 -> (aligned guide features NHWC bf16, beta, gamma (B, 64, 1, 1) fp32)."""
        B = feature1.shape[0]
        trunk = [self.conv1, self.leaky_relu1, self.conv2, self.leaky_relu2, self.conv3, self.leaky_relu3, self.conv4, self.leaky_relu4]
        t = run_layers(trunk, torch.cat((feature1, feature2), dim=0), "nhwc_bf16", "nhwc_bf16")
        if torch.is_grad_enabled() and (t.requires_grad or feature2.requires_grad or any(p.requires_grad for p in self.conv5.parameters())):
            beta = run_layers([self.conv5], t[:B], "nhwc_bf16", "nhwc_f32").mean(dim=(1, 2))
            gamma = run_layers([self.conv6], t[B:], "nhwc_bf16", "nhwc_f32").mean(dim=(1, 2))
            out = torch.addcmul(beta[:, None, None, :], gamma[:, None, None, :], feature2.float()).to(torch.bfloat16)
        elif self.heads_by_linearity:
            # the heads are only ever averaged over all positions: mean(conv(t)) from border-corrected sums of t, no convolution
            beta = ops.conv3x3_mean(t[:B], self.conv5.weight, self.conv5.bias)
            gamma = ops.conv3x3_mean(t[B:], self.conv6.weight, self.conv6.bias)
            out = ops.channel_affine_bf16(feature2, gamma, beta)
        else:
            beta = ops.channel_mean(run_layers([self.conv5], t[:B], "nhwc_bf16", "nhwc_f32"))
            gamma = ops.channel_mean(run_layers([self.conv6], t[B:], "nhwc_bf16", "nhwc_f32"))
            out = ops.channel_affine_bf16(feature2, gamma, beta)
        return out, beta[:, :, None, None], gamma[:, :, None, None]

    def forward(self, feature1: Tensor, feature2: Tensor):
        out, beta, gamma = self.forward_nhwc(_to_nhwc_bf16(feature1), _to_nhwc_bf16(feature2))
        return _nhwc_to_logical(out), beta, gamma


# ---- window cross-attention ------------------------------------------------------------------------------------------------
class _TokenLinear(nn.Linear):
    """nn.Linear (same parameters / keys) that can also run as a 1x1 tensor-core conv over a (B, H, W, C) token grid."""

    def _alias(self) -> nn.Module:
        """The same layer as a 1x1 Conv2d over the token grid (not a registered child; its parameters are views of this module's)."""
        m = getattr(self, "_alias1", None)
        if m is None:
            m = Conv2d(self.in_features, self.out_features, 1, bias=self.bias is not None)
            m._mmc_name = getattr(self, "_mmc_name", "linear")
            object.__setattr__(self, "_alias1", m)
        m._parameters["weight"] = self.weight.detach().view(self.out_features, self.in_features, 1, 1)
        m._parameters["bias"] = self.bias.detach() if self.bias is not None else None
        return m

    def on_grid(self, x: Tensor, act: int = L.ACT_NONE, out_f32: bool = False) -> Tensor:
        B, H, W, C = x.shape
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
            # recorded for backward: dgrad / wgrad of the 1x1 layer on the tensor-core kernels, gradients land on weight / bias
            if act != L.ACT_NONE or out_f32:
                raise NotImplementedError("_TokenLinear.on_grid: the training path has no fused activation / fp32 output")
            return AG.conv_alias(x, self.weight.view(self.out_features, C, 1, 1), self.bias, self._alias())
        d = ops.conv_desc(False, B, H, W, C, self.out_features, 1, 1, L.BF16, L.NHWC, L.F32 if out_f32 else L.BF16, L.NHWC, act=act)
        key = (self.weight._version, self.weight.data_ptr())
        if getattr(self, "_pack_key", None) != key:
            self._pack = ops.conv_pack_weights(d, self.weight.detach().view(self.out_features, C, 1, 1))
            self._pack_key = key
        return ops.conv_forward_tc(d, x, self._pack, self.bias.detach() if self.bias is not None else None,
                                   name=getattr(self, "_mmc_name", "linear"))


class PatchEmbed(nn.Module):
    """master.py:386-428: non-overlapping ``patch_size`` x ``patch_size`` patches -> ``embed_dim`` tokens (a stride-p conv)."""

    def __init__(self, img_size=(224, 224), patch_size=4, in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        p = patch_size if isinstance(patch_size, tuple) else (patch_size, patch_size)
        self.img_size, self.patch_size = tuple(img_size), p
        self.patches_resolution = [img_size[0] // p[0], img_size[1] // p[1]]
        self.num_patches = self.patches_resolution[0] * self.patches_resolution[1]
        self.in_chans, self.embed_dim = in_chans, embed_dim
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=p, stride=p)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def _check(self, H, W):
        if (H, W) != self.img_size:
            raise ValueError(f"Input image size ({H}*{W}) doesn't match model ({self.img_size[0]}*{self.img_size[1]}).")

    def _patch_matrix(self) -> Tensor:
        """proj.weight (E, C, p, p) as the (E, p*p*C) matrix that multiplies space-to-depth patches laid out (dy, dx, c)."""
        w = self.proj.weight
        return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)

    def _as_3x3(self) -> nn.Module:
        """The 2x2 stride-2 projection as a 3x3 stride-2 padding-1 layer whose first kernel row and column are zero
        (out[i] = w[1] x[2i] + w[2] x[2i+1]): the tensor-core conv reads the NHWC map in place, no space-to-depth copy.
        Not a registered child (the state_dict is the reference's); refreshed when ``proj`` changes."""
        m = getattr(self, "_alias3", None)
        if m is None or m.weight.device != self.proj.weight.device:
            m = conv(self.in_chans, self.embed_dim, kernel_size=3, stride=2).to(self.proj.weight.device)
            m.requires_grad_(False)
            m._mmc_name = getattr(self.proj, "_mmc_name", "patch_embed")
            object.__setattr__(self, "_alias3", m)
        key = (self.proj.weight._version, self.proj.weight.data_ptr(), self.proj.bias._version)
        if getattr(self, "_alias3_key", None) != key:
            with torch.no_grad():
                m.weight.copy_(F.pad(self.proj.weight.detach(), (1, 0, 1, 0)))
                m.bias.copy_(self.proj.bias.detach())
            self._alias3_key = key
        return m

    def forward_grid(self, x: Tensor) -> Tensor:
        """NHWC bf16 map -> (B, H/p, W/p, E) bf16 token grid."""
        B, H, W, C = x.shape
        self._check(H, W)
        p, q = self.patch_size
        on_kernels = _attn_kernels()
        if on_kernels and (p, q) == (2, 2) and H % 2 == 0 and W % 2 == 0 and self.norm is None:
            m = self._as_3x3()
            if torch.is_grad_enabled() and (x.requires_grad or self.proj.weight.requires_grad):
                # the zero-padded 3x3 kernel as a differentiable function of proj.weight: the gradient of the padding is dropped by F.pad
                return AG.conv_alias(x, F.pad(self.proj.weight, (1, 0, 1, 0)), self.proj.bias, m)
            return run_layers([m], x, "nhwc_bf16", "nhwc_bf16")
        patches = x.reshape(B, H // p, p, W // q, q, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H // p, W // q, p * q * C)
        if on_kernels and not torch.is_grad_enabled():
            d = ops.conv_desc(False, B, H // p, W // q, p * q * C, self.embed_dim, 1, 1, L.BF16, L.NHWC, L.BF16, L.NHWC)
            key = (self.proj.weight._version, self.proj.weight.data_ptr())
            if getattr(self, "_pack_key", None) != key:
                self._pack = ops.conv_pack_weights(d, self._patch_matrix().detach().reshape(self.embed_dim, -1, 1, 1).contiguous())
                self._pack_key = key
            tok = ops.conv_forward_tc(d, patches.contiguous(), self._pack, self.proj.bias.detach(), name=getattr(self.proj, "_mmc_name", "patch"))
        else:
            tok = F.linear(patches, self._patch_matrix().to(patches.dtype), self.proj.bias.to(patches.dtype))
        if self.norm is not None:
            tok = self.norm(tok.float()).to(tok.dtype)
        return tok

    def forward(self, x: Tensor) -> Tensor:
        tok = self.forward_grid(_to_nhwc_bf16(x))
        return tok.reshape(tok.shape[0], -1, tok.shape[-1])


def window_partition(x: Tensor, window_size: int = 4) -> Tensor:
    """master.py:431-443: (B, H, W, C) -> (B * nW, ws, ws, C)"""
    B, H, W, C = x.shape
    ws = window_size
    return x.reshape(B, H // ws, ws, W // ws, ws, C).transpose(2, 3).reshape(-1, ws, ws, C)


def window_reverse(windows: Tensor, window_size: int, H: int, W: int) -> Tensor:
    """master.py:446-460"""
    ws = window_size
    B = windows.shape[0] // ((H // ws) * (W // ws))
    return windows.reshape(B, H // ws, W // ws, ws, ws, -1).transpose(2, 3).reshape(B, H, W, -1)


class Mlp(nn.Module):
    """master.py:463-481"""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        self.fc1 = _TokenLinear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = _TokenLinear(hidden_features or in_features, out_features or in_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x: Tensor) -> Tensor:
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class WindowAttention(nn.Module):
    """master.py:484-568: multi-head attention inside one window with queries from ``x`` (qkv1) and keys / values from
    ``guided`` (qkv2), a learned relative-position bias and an optional additive (0 / -100) shift mask."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=True, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, tuple(window_size), num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        wh, ww = self.window_size
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * wh - 1) * (2 * ww - 1), num_heads))
        iy, ix = torch.meshgrid(torch.arange(wh), torch.arange(ww), indexing="ij")
        iy, ix = iy.flatten(), ix.flatten()
        index = (iy[:, None] - iy[None, :] + wh - 1) * (2 * ww - 1) + (ix[:, None] - ix[None, :] + ww - 1)
        self.register_buffer("relative_position_index", index)
        self.qkv1 = _TokenLinear(dim, dim, bias=qkv_bias)
        self.qkv2 = _TokenLinear(dim, dim * 2, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = _TokenLinear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)

    def position_bias(self) -> Tensor:
        """(heads, N, N) fp32"""
        n = self.window_size[0] * self.window_size[1]
        return self.relative_position_bias_table[self.relative_position_index.reshape(-1)].reshape(n, n, -1).permute(2, 0, 1)

    def forward(self, x: Tensor, guided: Tensor, mask=None) -> Tensor:
        """x, guided: (nW * B, N, C)"""
        Bw, N, C = x.shape
        h = self.num_heads
        q = self.qkv1(x).reshape(Bw, N, h, C // h).transpose(1, 2) * self.scale
        k, v = self.qkv2(guided).reshape(Bw, N, 2, h, C // h).permute(2, 0, 3, 1, 4)
        att = (q @ k.transpose(-2, -1)).float() + self.position_bias()
        if mask is not None:
            att = (att.reshape(-1, mask.shape[0], h, N, N) + mask[None, :, None]).reshape(Bw, h, N, N)
        att = self.attn_drop(self.softmax(att)).to(v.dtype)
        return self.proj_drop(self.proj((att @ v).transpose(1, 2).reshape(Bw, N, C)))


class SwinTransformerBlock(nn.Module):
    """master.py:572-706: pre-norm cross-attention block; ``norm1`` normalises BOTH the master and the guide tokens."""

    def __init__(self, dim, input_resolution, num_heads, window_size=7, shift_size=0, mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, fused_window_process=False):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, tuple(input_resolution), num_heads
        self.window_size, self.shift_size, self.mlp_ratio = window_size, shift_size, mlp_ratio
        if min(self.input_resolution) <= self.window_size:
            self.shift_size, self.window_size = 0, min(self.input_resolution)
        if not 0 <= self.shift_size < self.window_size:
            raise ValueError("shift_size must in 0-window_size")
        if drop_path > 0.:
            raise NotImplementedError("stochastic depth is not used by Spatial_aligner (drop_path=0)")
        self.norm1 = norm_layer(dim)
        self.attn = WindowAttention(dim, window_size=(self.window_size, self.window_size), num_heads=num_heads, qkv_bias=qkv_bias,
                                    qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)
        attn_mask = None
        if self.shift_size > 0:
            H, W = self.input_resolution
            region = torch.zeros(1, H, W, 1)
            bands = (slice(0, -self.window_size), slice(-self.window_size, -self.shift_size), slice(-self.shift_size, None))
            for n, (hs, wsl) in enumerate((a, b) for a in bands for b in bands):
                region[:, hs, wsl, :] = n
            ids = window_partition(region, self.window_size).reshape(-1, self.window_size * self.window_size)
            diff = ids[:, None, :] - ids[:, :, None]
            attn_mask = torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))
        self.register_buffer("attn_mask", attn_mask)
        self.fused_window_process = fused_window_process

    def forward_grid(self, x: Tensor, guided: Tensor) -> Tensor:
        """(B, H, W, C) bf16 token grids -> (B, H, W, C) bf16, all on the libmmcodec kernels; with autograd on, every step is
        recorded with its backward kernel (mmcodec.autograd: LayerNorm, the 1x1 Linear layers, window attention, GELU)."""
        a = self.attn
        q = a.qkv1.on_grid(AG.layernorm(x, self.norm1))
        kv = a.qkv2.on_grid(AG.layernorm(guided, self.norm1))
        ctx = AG.window_attention(q, kv, a.relative_position_bias_table, self.window_size, self.shift_size, a.num_heads, a.scale)
        x, normed = AG.layernorm(x, self.norm2, delta=a.proj.on_grid(ctx))
        return x + self.mlp.fc2.on_grid(AG.gelu(self.mlp.fc1.on_grid(normed)))

    def forward(self, x: Tensor, guided: Tensor) -> Tensor:
        """(B, L, C) tokens, torch ops (differentiable)."""
        H, W = self.input_resolution
        B, Ln, C = x.shape
        if Ln != H * W:
            raise ValueError("input feature has wrong size")
        ws, sh = self.window_size, self.shift_size
        grid = lambda t: self.norm1(t).reshape(B, H, W, C)
        xs, gs = grid(x), grid(guided)
        if sh > 0:
            xs, gs = (torch.roll(t, shifts=(-sh, -sh), dims=(1, 2)) for t in (xs, gs))
        win = lambda t: window_partition(t, ws).reshape(-1, ws * ws, C)
        o = self.attn(win(xs), win(gs), mask=self.attn_mask)
        o = window_reverse(o.reshape(-1, ws, ws, C), ws, H, W)
        if sh > 0:
            o = torch.roll(o, shifts=(sh, sh), dims=(1, 2))
        x = x + self.drop_path(o.reshape(B, H * W, C))
        return x + self.drop_path(self.mlp(self.norm2(x)))


class Spatial_aligner(nn.Module):
    """master.py:708-742.  NOTE the reference reinterprets the (B, L, 96) token tensor as (B, 96, H/2, W/2) with ``.view``
    (no transpose) before the 2x2 stride-2 ``recovery`` transposed conv; reproduced as is."""

    def __init__(self, in_channel=192, out_channel=192, input_resolution=(224, 224)) -> None:
        super().__init__()
        self.window_size, self.patch_size, self.embed_dim = 4, 2, 96
        self.input_resolution = tuple(input_resolution)
        self.patch_embeding1 = PatchEmbed(img_size=self.input_resolution, patch_size=2, in_chans=in_channel, embed_dim=96)
        self.patch_embeding2 = PatchEmbed(img_size=self.input_resolution, patch_size=2, in_chans=in_channel, embed_dim=96)
        res = (self.input_resolution[0] // 2, self.input_resolution[1] // 2)
        self.blocks = nn.ModuleList([SwinTransformerBlock(dim=96, num_heads=3, window_size=4, input_resolution=res,
                                                          shift_size=0 if i % 2 == 0 else 2) for i in range(2)])
        self.recovery = nn.ConvTranspose2d(96, out_channel, kernel_size=2, stride=2)

    def _recover(self, tok: Tensor, B: int, H: int, W: int) -> Tensor:
        """(B, h, w, 96) token grid -> NHWC bf16 (B, H, W, out): the reference's reinterpretation, then the 2x2 s2 deconv (on the
        transposed-conv kernel in inference; as a per-pixel matrix product followed by depth-to-space under autograd)."""
        E, h, w = self.embed_dim, H // 2, W // 2
        grid = tok.reshape(B, E, h * w).transpose(1, 2).reshape(B, h, w, E)    # memory reinterpreted as NCHW, then made NHWC
        wt = self.recovery.weight                                                 # (E, O, 2, 2)
        O = wt.shape[1]
        if _attn_kernels():
            # the 2x2 stride-2 transposed conv as a 3x3 stride-2 one with kernel row / column 0 zero (out[2i + d] = x[i] w[1 + d]):
            # the transposed-conv kernel writes the full-resolution NHWC map directly, no depth-to-space copy
            m = getattr(self, "_recovery3", None)
            if m is None or m.weight.device != wt.device:
                m = deconv(E, O, kernel_size=3, stride=2).to(wt.device)
                m.requires_grad_(False)
                m._mmc_name = getattr(self.recovery, "_mmc_name", "recovery")
                object.__setattr__(self, "_recovery3", m)
            key = (wt._version, wt.data_ptr(), self.recovery.bias._version)
            if getattr(self, "_recovery3_key", None) != key:
                with torch.no_grad():
                    m.weight.copy_(F.pad(wt.detach(), (1, 0, 1, 0)))
                    m.bias.copy_(self.recovery.bias.detach())
                self._recovery3_key = key
            if torch.is_grad_enabled() and (grid.requires_grad or wt.requires_grad):
                return AG.conv_alias(grid.contiguous(), F.pad(wt, (1, 0, 1, 0)), self.recovery.bias, m)
            return run_layers([m], grid.contiguous(), "nhwc_bf16", "nhwc_bf16")
        mat = wt.permute(2, 3, 1, 0).reshape(4 * O, E)
        out = F.linear(grid, mat.to(grid.dtype), self.recovery.bias.repeat(4).to(grid.dtype))
        return out.reshape(B, h, w, 2, 2, O).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, O)

    def forward_nhwc(self, x: Tensor, guided: Tensor) -> Tensor:
        B, H, W, _ = x.shape
        tok, gtok = self.patch_embeding1.forward_grid(x), self.patch_embeding2.forward_grid(guided)
        if _attn_kernels():
            for blk in self.blocks:
                tok = blk.forward_grid(tok, gtok)
        else:
            t, g = tok.reshape(B, -1, self.embed_dim), gtok.reshape(B, -1, self.embed_dim)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                for blk in self.blocks:
                    t = blk(t, g)
            tok = t.to(torch.bfloat16)
        return self._recover(tok.contiguous(), B, H, W)

    def forward(self, x: Tensor, guided: Tensor) -> Tensor:
        return _nhwc_to_logical(self.forward_nhwc(_to_nhwc_bf16(x), _to_nhwc_bf16(guided)))


class Master_decoder(nn.Module):
    """master.py:745-811"""

    def __init__(self, N=192, M=192, channel=64 * 2, width=224, height=224, first_stride=2, master_chl=3) -> None:
        super().__init__()
        self.encoder_first_stride = first_stride
        width //= first_stride
        height //= first_stride
        self.g_s_conv1 = deconv(M, N, kernel_size=5, stride=2)
        self.g_s_gdn1 = GDN(N, inverse=True)
        self.sp_aligner1 = Spatial_aligner(input_resolution=(width // 4, height // 4))
        self.g_s_conv2 = deconv(2 * N, N, kernel_size=5, stride=2)
        self.g_s_gdn2 = GDN(N, inverse=True)
        self.sp_aligner2 = Spatial_aligner(input_resolution=(width // 2, height // 2))
        self.g_s_conv3 = deconv(2 * N, N, kernel_size=5, stride=2)
        self.g_s_gdn3 = GDN(N, inverse=True)
        self.sp_aligner3 = Spatial_aligner(input_resolution=(width, height))
        self.g_s_conv4 = deconv(2 * N, channel, kernel_size=5, stride=first_stride)
        self.master_chl = master_chl
        if master_chl == 1:
            self.downsample1 = conv(N, N, kernel_size=5, stride=2)
            self.downsample2 = conv(N, N, kernel_size=5, stride=2)
            self.downsample3 = conv(N, N, kernel_size=5, stride=2)

    def forward_nhwc(self, y_hat: Tensor, guide_hidden: Dict[str, Tensor]) -> Tensor:
        """NHWC bf16 latent + the guide codec's gs1..gs3 maps (logical NCHW) -> NHWC bf16 feature reconstruction."""
        g = [_to_nhwc_bf16(guide_hidden[k]) for k in ("gs1", "gs2", "gs3")]
        if self.master_chl == 1:
            g = [run_layers([m], t.contiguous(), "nhwc_bf16", "nhwc_bf16") for m, t in zip((self.downsample1, self.downsample2, self.downsample3), g)]
        s = y_hat
        stages = ((self.g_s_conv1, self.g_s_gdn1, self.sp_aligner1), (self.g_s_conv2, self.g_s_gdn2, self.sp_aligner2),
                  (self.g_s_conv3, self.g_s_gdn3, self.sp_aligner3))
        for (dc, gd, al), gm in zip(stages, g):
            own = run_layers([dc, gd], s, "nhwc_bf16", "nhwc_bf16")
            s = (al.forward_nhwc(own, gm.contiguous()).contiguous(), own)      # cat([aligned, identity]) feeds the next layer
        return run_layers([self.g_s_conv4], s, "nhwc_bf16", "nhwc_bf16")

    def forward(self, x: Tensor, guide_hidden: Dict[str, Tensor]):
        return {"x_feature_hat": _nhwc_to_logical(self.forward_nhwc(_to_nhwc_bf16(x), guide_hidden))}


class Master_compresser(_ContextModelMixin, MeanScaleHyperprior):
    """master.py:837-951.  ``width`` / ``height`` are the first / second spatial size of the GUIDE image (= half the 3-channel
    master image); ``channel`` is the master image's channel count (3: RGB master + 1-channel guide, 1: the reverse)."""

    forward_takes_guide_image = True      # TrainStep: net(x, guided, hidden) as in examples/train.py:224

    def __init__(self, width=256, height=256, channel=3, N=192, M=192) -> None:
        super().__init__(M, M)
        master_chl, guided_chl, master_stride, guided_stride = (3, 1, 2, 1) if channel != 1 else (1, 3, 1, 2)
        self.fencoder1 = Feature_encoder(in_channel=master_chl, out_channel=64, stride=master_stride)
        self.fencoder2 = Feature_encoder(in_channel=guided_chl, out_channel=64, stride=guided_stride)
        self.ch_aligner = Channel_aligner()
        self.g_a = TransformStack(conv(64 * 2, N, kernel_size=5, stride=2), GDN(N), conv(N, N, kernel_size=5, stride=2), GDN(N),
                                  conv(N, N, kernel_size=5, stride=2), GDN(N), conv(N, M, kernel_size=5, stride=2))
        self._init_entropy_stage(N, M)
        self.N, self.M = int(N), int(M)
        self.decoder = Master_decoder(N=192, M=192, channel=64 * 2, width=width, height=height, first_stride=2, master_chl=master_chl)
        self.fdecoder = Feature_decoder(in_channel=64 * 3, out_channel=master_chl, stride=master_stride)
        self._tag_layer_names()
        for name, m in self.named_modules():
            if isinstance(m, nn.Linear):
                m._mmc_name = name

    @property
    def downsampling_factor(self) -> int:
        return 2 ** (4 + 2 + 1)

    def forward(self, x: Tensor, guided_hat: Tensor, guided_hidden: Dict[str, Tensor]):
        x_feature = self.fencoder1.forward_nhwc(x)
        guided_feature = self.fencoder2.forward_nhwc(guided_hat)
        guided_align, beta, gamma = self.ch_aligner.forward_nhwc(x_feature, guided_feature)
        y, y_bf16 = run_layers(list(self.g_a), (x_feature, guided_align), "nhwc_bf16", "nhwc_f32", out2=2)
        y_hat_bf16, y_lik, z_lik = self._entropy_stage(y, y_bf16)
        feature_hat = self.decoder.forward_nhwc(y_hat_bf16, guided_hidden)
        x_hat = self.fdecoder.forward_nhwc((feature_hat, guided_align))
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}}

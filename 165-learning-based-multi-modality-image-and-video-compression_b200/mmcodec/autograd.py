"""Training path: ``torch.autograd.Function``s over the libmmcodec kernels.

Forward passes are the inference kernels (same numerics); backward passes are
  * input gradient of conv() = deconv() with the same weight tensor and vice versa (``mmc_conv_forward_tc`` with the
    adjoint descriptor: compressai/models/utils.py:128-146 are each other's adjoints for even sizes),
  * weight gradient on tensor cores (``mmc_wgrad_tc``), bias gradient as a column sum,
  * GDN / IGDN backward as two 1x1 tensor-core contractions around two elementwise kernels, dgamma through the weight-gradient
    kernel with k = 1,
  * likelihood backward kernels of the entropy models.
Activations are saved as NHWC bf16 (what the forward kernels produce anyway); parameter gradients are fp32.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib as L
from . import ops

__all__ = ["run_layers_train", "eb_forward", "gc_forward", "gdn_forward", "cast_bf16", "add_noise", "wants_grad", "maxpool_nhwc",
           "upsample_add", "sigmoid_gate", "layernorm", "gelu", "window_attention", "conv_alias"]


def wants_grad(layers, x: Tensor) -> bool:
    if not torch.is_grad_enabled():
        return False
    if x.requires_grad:
        return True
    return any(p.requires_grad for m in layers for p in m.parameters(recurse=False))


# ---------------------------------------------------------------------------------------------------------
# conv / deconv (+ ReLU / LeakyReLU) and conv + GDN
# ---------------------------------------------------------------------------------------------------------
def _adjoint(conv: nn.Module) -> nn.Module:
    """deconv() sharing ``conv``'s weight tensor (and the other way round): the input-gradient operator."""
    from .layers import Conv2d, ConvTranspose2d
    adj = getattr(conv, "_mmc_adjoint", None)
    if adj is None:
        k, s = conv.kernel_size[0], conv.stride[0]
        if isinstance(conv, nn.ConvTranspose2d):
            adj = Conv2d(conv.out_channels, conv.in_channels, kernel_size=k, stride=s, padding=k // 2, bias=False)
        else:
            adj = ConvTranspose2d(conv.out_channels, conv.in_channels, kernel_size=k, stride=s, padding=k // 2,
                                  output_padding=s - 1, bias=False)
        adj._mmc_name = getattr(conv, "_mmc_name", "conv") + ".dgrad"
        object.__setattr__(conv, "_mmc_adjoint", adj)          # not a registered child: no state_dict entry
    adj._parameters["weight"] = conv.weight                    # same tensor (masked convs have already applied the mask)
    return adj


def _run(layers, x, in_fmt, out_fmt, out2=0):
    from .transforms import run_layers
    with torch.no_grad():
        return run_layers(layers, x, in_fmt, out_fmt, out2=out2, _train_dispatch=False)


def _grad_nhwc_bf16(g: Tensor, out_fmt: str, cout: int):
    """Upstream gradient in the forward output's format -> (NHWC bf16 gradient with >= 8 channels, planar fp32 or None)."""
    if out_fmt == "nhwc_bf16":
        return g.contiguous(), None
    if out_fmt == "nhwc_f32":
        return ops.to_bf16(g.contiguous()), None
    # planar NCHW fp32 (reconstruction layers, Cout <= 4)
    g = g.float().contiguous()
    return ops.nchw_to_nhwc8(g), g


def _conv_backward(conv, x_saved, in_fmt, g, g_planar, need_dx):
    """(dx, dweight, dbias) of one conv / deconv layer given its NHWC bf16 output gradient g."""
    transposed = isinstance(conv, nn.ConvTranspose2d)
    k, s = conv.kernel_size[0], conv.stride[0]
    cin, cout = conv.in_channels, conv.out_channels
    name = getattr(conv, "_mmc_name", "conv")
    dbias = None
    if conv.bias is not None:
        dbias = ops.colsum(g)[:cout].to(conv.bias.dtype)
    x_nhwc = ops.nchw_to_nhwc8(x_saved) if in_fmt == "nchw_f32" else x_saved
    if transposed:      # S = input (low resolution), L = grad_output
        dw = (ops.wgrad_edge(x_nhwc, g, k, s, name=name) if g.shape[-1] == 8 and cout <= 8 else ops.wgrad(x_nhwc, g, k, s, name=name))[:cin, :cout]
    else:               # S = grad_output, L = input
        dw = (ops.wgrad_edge(g, x_nhwc, k, s, name=name) if x_nhwc.shape[-1] == 8 and cin <= 8 else ops.wgrad(g, x_nhwc, k, s, name=name))[:cout, :cin]
    mask = getattr(conv, "mask", None)
    if mask is not None:
        dw = dw * mask          # MaskedConv2d: masked taps never receive gradient (their weights are re-zeroed every forward)
    dx = None
    if need_dx:
        adj = _adjoint(conv)
        if g_planar is not None:
            dx = _run([adj], g_planar, "nchw_f32", "nhwc_bf16")
        else:
            dx = _run([adj], g[..., :cout] if g.shape[-1] != cout else g, "nhwc_bf16", "nhwc_bf16")
        if not transposed and in_fmt != "nchw_f32" and dx.shape[1:3] != x_saved.shape[1:3]:
            # odd input size: the adjoint deconv also produces the gradient of the (non-existent) padding row / column
            dx = dx[:, : x_saved.shape[1], : x_saved.shape[2]].contiguous()
    return dx, dw.contiguous().to(conv.weight.dtype), dbias


def _gdn_helpers(gdn):
    """1x1 conv modules computing norm = beta' + gamma' x^2 and u = gamma'^T t on the tensor-core kernel."""
    from .layers import Conv2d
    beta_eff, gamma_eff, _ = gdn.effective_params()
    key = (beta_eff.data_ptr(), gamma_eff.data_ptr(), gdn._cache_key)
    if getattr(gdn, "_bwd_key", None) != key:
        C = beta_eff.numel()
        fwd, bwd = Conv2d(C, C, 1, bias=True), Conv2d(C, C, 1, bias=False)
        fwd._parameters["weight"] = nn.Parameter(gamma_eff.reshape(C, C, 1, 1), requires_grad=False)
        fwd._parameters["bias"] = nn.Parameter(beta_eff, requires_grad=False)
        bwd._parameters["weight"] = nn.Parameter(gamma_eff.t().contiguous().reshape(C, C, 1, 1), requires_grad=False)
        fwd._mmc_name, bwd._mmc_name = "gdn.norm", "gdn.normT"
        object.__setattr__(gdn, "_bwd_convs", (fwd, bwd))
        gdn._bwd_key = key
    return gdn._bwd_convs


def _gdn_backward(gdn, x_pre, g):
    """GDN / IGDN backward (layers/gdn.py:77-92): returns (dx bf16 NHWC, dbeta, dgamma) w.r.t. the RAW parameters."""
    fwd, bwd = _gdn_helpers(gdn)
    inv = gdn.inverse
    x2 = ops.square(x_pre)
    norm = _run([fwd], x2, "nhwc_bf16", "nhwc_f32")
    t = ops.gdn_bwd_t(g, x_pre, norm, inv)
    u = _run([bwd], t, "nhwc_bf16", "nhwc_f32")
    dx = ops.gdn_bwd_dx(g, x_pre, norm, u, inv)
    sign = 0.5 if inv else -0.5
    dbeta_eff = ops.colsum(t, scale=sign)
    C = dbeta_eff.numel()
    dgamma_eff = ops.wgrad(t, x2, 1, 1, scale=sign, name="gdn.dgamma").reshape(C, C)
    br, gr = gdn.beta_reparam, gdn.gamma_reparam
    dbeta = ops.reparam_bwd(gdn.beta, dbeta_eff, (br.minimum + br.reparam_offset ** 2) ** 0.5)
    dgamma = ops.reparam_bwd(gdn.gamma, dgamma_eff, (gr.minimum + gr.reparam_offset ** 2) ** 0.5)
    return dx, dbeta, dgamma


class _ConvFn(torch.autograd.Function):
    """conv()/deconv() + bias (+ ReLU / LeakyReLU)."""

    @staticmethod
    def forward(ctx, x, weight, bias, conv, act_module, in_fmt, out_fmt):
        layers = [conv] + ([act_module] if act_module is not None else [])
        y = _run(layers, x, in_fmt, out_fmt)
        ctx.conv, ctx.act_module, ctx.in_fmt, ctx.out_fmt = conv, act_module, in_fmt, out_fmt
        ctx.save_for_backward(x, y if act_module is not None else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, y = ctx.saved_tensors
        conv = ctx.conv
        g, g_planar = _grad_nhwc_bf16(gy, ctx.out_fmt, conv.out_channels)
        if ctx.act_module is not None:
            if ctx.out_fmt == "nchw_f32":
                raise NotImplementedError("activation backward is implemented for NHWC layer outputs")
            act = L.ACT_LEAKY_RELU if isinstance(ctx.act_module, nn.LeakyReLU) else L.ACT_RELU
            g = ops.act_bwd(g, y if y.dtype == torch.bfloat16 else ops.to_bf16(y), act)   # only the sign of y matters
        dx, dw, db = _conv_backward(conv, x, ctx.in_fmt, g, g_planar, ctx.needs_input_grad[0])
        return dx, dw, db, None, None, None, None


class _ConvGdnFn(torch.autograd.Function):
    """conv()/deconv() + bias + GDN / IGDN in one forward launch; the pre-GDN activations are its secondary output."""

    @staticmethod
    def forward(ctx, x, weight, bias, beta, gamma, conv, gdn, in_fmt):
        y, x_pre = _run([conv, gdn], x, in_fmt, "nhwc_bf16", out2=3)
        ctx.conv, ctx.gdn, ctx.in_fmt = conv, gdn, in_fmt
        ctx.save_for_backward(x, x_pre)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, x_pre = ctx.saved_tensors
        g_pre, dbeta, dgamma = _gdn_backward(ctx.gdn, x_pre, gy.contiguous())
        dx, dw, db = _conv_backward(ctx.conv, x, ctx.in_fmt, g_pre, None, ctx.needs_input_grad[0])
        return dx, dw, db, dbeta.to(ctx.gdn.beta.dtype), dgamma.to(ctx.gdn.gamma.dtype), None, None, None


def _adjoint_part(conv: nn.Module, lo: int, hi: int) -> nn.Module:
    """Adjoint of ``conv`` restricted to input channels [lo, hi): the input-gradient operator of ONE source of a two-source layer
    (its weight is the matching slice of ``conv.weight``, so no gradient for the other source's channels is computed or split off)."""
    from .layers import Conv2d, ConvTranspose2d
    parts = conv.__dict__.setdefault("_mmc_adjoint_parts", {})
    adj = parts.get((lo, hi))
    k, s = conv.kernel_size[0], conv.stride[0]
    transposed = isinstance(conv, nn.ConvTranspose2d)
    if adj is None:
        if transposed:
            adj = Conv2d(conv.out_channels, hi - lo, kernel_size=k, stride=s, padding=k // 2, bias=False)
        else:
            adj = ConvTranspose2d(conv.out_channels, hi - lo, kernel_size=k, stride=s, padding=k // 2, output_padding=s - 1, bias=False)
        adj._mmc_name = getattr(conv, "_mmc_name", "conv") + f".dgrad[{lo}:{hi}]"
        parts[(lo, hi)] = adj
    w = conv.weight.detach()
    adj._parameters["weight"] = w[lo:hi] if transposed else w[:, lo:hi]     # a view: shares the version counter, own data_ptr
    return adj


def _conv_backward2(conv, x1, x2, g, g_planar, need1, need2):
    """(dx1, dx2, dweight, dbias) of a conv / deconv layer whose input is the channel concatenation of the NHWC bf16 maps x1 and x2
    (google.py:1153 and its repeats), which is never materialised: the weight gradient is computed per source and joined along
    the input-channel axis (a parameter-sized copy), each input gradient by the adjoint layer of that source's weight slice."""
    transposed = isinstance(conv, nn.ConvTranspose2d)
    k, s = conv.kernel_size[0], conv.stride[0]
    c1, cout = x1.shape[-1], conv.out_channels
    name = getattr(conv, "_mmc_name", "conv")
    dbias = ops.colsum(g)[:cout].to(conv.bias.dtype) if conv.bias is not None else None
    if transposed:      # S = input (low resolution), L = grad_output: (cin, cout, k, k)
        edge = g.shape[-1] == 8 and cout <= 8
        dw = torch.cat([(ops.wgrad_edge(xi, g, k, s, name=name) if edge else ops.wgrad(xi, g, k, s, name=name))[:, :cout] for xi in (x1, x2)], dim=0)
    else:               # S = grad_output, L = input: (cout, cin, k, k)
        dw = torch.cat([ops.wgrad(g, x1, k, s, name=name), ops.wgrad(g, x2, k, s, name=name)], dim=1)
    mask = getattr(conv, "mask", None)
    if mask is not None:
        dw = dw * mask
    dxs = []
    for need, x_saved, (lo, hi) in ((need1, x1, (0, c1)), (need2, x2, (c1, conv.in_channels))):
        dx = None
        if need:
            adj = _adjoint_part(conv, lo, hi)
            if g_planar is not None:
                dx = _run([adj], g_planar, "nchw_f32", "nhwc_bf16")
            else:
                dx = _run([adj], g[..., :cout] if g.shape[-1] != cout else g, "nhwc_bf16", "nhwc_bf16")
            if not transposed and dx.shape[1:3] != x_saved.shape[1:3]:
                dx = dx[:, : x_saved.shape[1], : x_saved.shape[2]].contiguous()
        dxs.append(dx)
    return dxs[0], dxs[1], dw.contiguous().to(conv.weight.dtype), dbias


class _Conv2Fn(torch.autograd.Function):
    """Two-source conv()/deconv() + bias (+ ReLU / LeakyReLU): the K loop of the forward kernel reads both maps
    (mmc_conv_forward_tc2), the backward pass never builds the concatenation either."""

    @staticmethod
    def forward(ctx, x1, x2, weight, bias, conv, act_module, out_fmt):
        layers = [conv] + ([act_module] if act_module is not None else [])
        y = _run(layers, (x1, x2), "nhwc_bf16", out_fmt)
        ctx.conv, ctx.act_module, ctx.out_fmt = conv, act_module, out_fmt
        ctx.save_for_backward(x1, x2, y if act_module is not None else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x1, x2, y = ctx.saved_tensors
        conv = ctx.conv
        g, g_planar = _grad_nhwc_bf16(gy, ctx.out_fmt, conv.out_channels)
        if ctx.act_module is not None:
            if ctx.out_fmt == "nchw_f32":
                raise NotImplementedError("activation backward is implemented for NHWC layer outputs")
            act = L.ACT_LEAKY_RELU if isinstance(ctx.act_module, nn.LeakyReLU) else L.ACT_RELU
            g = ops.act_bwd(g, y if y.dtype == torch.bfloat16 else ops.to_bf16(y), act)
        dx1, dx2, dw, db = _conv_backward2(conv, x1, x2, g, g_planar, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return dx1, dx2, dw, db, None, None, None


class _ConvGdn2Fn(torch.autograd.Function):
    """Two-source conv()/deconv() + bias + GDN / IGDN in one forward launch (pic2_g_a_conv2..4 / pic2_g_s_conv2..3 with their
    GDN layers, google.py:1160-1246)."""

    @staticmethod
    def forward(ctx, x1, x2, weight, bias, beta, gamma, conv, gdn):
        y, x_pre = _run([conv, gdn], (x1, x2), "nhwc_bf16", "nhwc_bf16", out2=3)
        ctx.conv, ctx.gdn = conv, gdn
        ctx.save_for_backward(x1, x2, x_pre)
        return y

    @staticmethod
    def backward(ctx, gy):
        x1, x2, x_pre = ctx.saved_tensors
        g_pre, dbeta, dgamma = _gdn_backward(ctx.gdn, x_pre, gy.contiguous())
        dx1, dx2, dw, db = _conv_backward2(ctx.conv, x1, x2, g_pre, None, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return dx1, dx2, dw, db, dbeta.to(ctx.gdn.beta.dtype), dgamma.to(ctx.gdn.gamma.dtype), None, None


class _GdnFn(torch.autograd.Function):
    """Stand-alone GDN / IGDN module call (compressai/layers/gdn.py:77-92) on a logical (B, C, H, W) fp32 tensor: forward on the
    fp32 kernel, backward through the same tensor-core contractions as the fused conv + GDN layers (bf16 operands)."""

    @staticmethod
    def forward(ctx, x, beta, gamma, gdn):
        beta_eff, gamma_eff, _ = gdn.effective_params()
        y = ops.gdn_forward(x, beta_eff, gamma_eff, gdn.inverse)
        ctx.gdn = gdn
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        gdn = ctx.gdn
        x_pre = ops.nchw_to_nhwc_bf16(x.float().contiguous())
        g = ops.nchw_to_nhwc_bf16(gy.float().contiguous())
        dx, dbeta, dgamma = _gdn_backward(gdn, x_pre, g)
        return dx.permute(0, 3, 1, 2).float(), dbeta.to(gdn.beta.dtype), dgamma.to(gdn.gamma.dtype), None


def gdn_forward(gdn, x: Tensor) -> Tensor:
    """GDN.forward with autograd when the input or the module's parameters need it."""
    if torch.is_grad_enabled() and (x.requires_grad or gdn.beta.requires_grad or gdn.gamma.requires_grad):
        return _GdnFn.apply(x, gdn.beta, gdn.gamma, gdn)
    beta_eff, gamma_eff, _ = gdn.effective_params()
    with torch.no_grad():
        return ops.gdn_forward(x, beta_eff, gamma_eff, gdn.inverse)


def run_layers_train(layers, x: Tensor, in_fmt: str, out_fmt: str, out2: int = 0):
    """``transforms.run_layers`` with autograd: one Function per fused launch."""
    from .transforms import parse_layers
    steps = parse_layers(layers)
    if not steps:
        raise ValueError("empty transform stack")
    if in_fmt == "nchw_f32" and steps[0].conv.in_channels > 8:
        raise NotImplementedError("training path takes NHWC bf16 activations (or an image with <= 8 channels)")
    pair = None
    if isinstance(x, (tuple, list)):
        # two NHWC bf16 sources feeding the first layer: neither pass materialises their concatenation
        if in_fmt != "nhwc_bf16":
            raise NotImplementedError("two-source layers take NHWC bf16 maps")
        pair = (x[0].contiguous(), x[1].contiguous())
        x = pair[0]
    cur, fmt = x, in_fmt
    for i, s in enumerate(steps):
        last = i == len(steps) - 1
        c = s.conv
        if hasattr(c, "_apply_mask"):
            c._apply_mask()
        ofmt = "nhwc_bf16"
        if last:
            ofmt = out_fmt
            if out_fmt == "nchw_f32" and c.out_channels > 4:
                ofmt = "nhwc_f32"
        if s.gdn is not None:
            if ofmt != "nhwc_bf16":
                raise NotImplementedError("conv + GDN layers write NHWC bf16 on the training path")
            if i == 0 and pair is not None:
                cur = _ConvGdn2Fn.apply(pair[0], pair[1], c.weight, c.bias, s.gdn.beta, s.gdn.gamma, c, s.gdn)
            else:
                cur = _ConvGdnFn.apply(cur, c.weight, c.bias, s.gdn.beta, s.gdn.gamma, c, s.gdn, fmt)
        else:
            act_module = None
            if s.act == L.ACT_RELU:
                act_module = nn.ReLU()
            elif s.act == L.ACT_LEAKY_RELU:
                act_module = nn.LeakyReLU()
            elif s.act != L.ACT_NONE:
                raise NotImplementedError("this activation has no backward kernel yet")
            if i == 0 and pair is not None:
                cur = _Conv2Fn.apply(pair[0], pair[1], c.weight, c.bias, c, act_module, ofmt)
            else:
                cur = _ConvFn.apply(cur, c.weight, c.bias, c, act_module, fmt, ofmt)
        fmt = ofmt
    if out_fmt == "nchw_f32" and fmt == "nhwc_f32":
        cur = cur.permute(0, 3, 1, 2)
    if out2:
        src = cur.permute(0, 2, 3, 1) if (out_fmt == "nchw_f32" and fmt == "nhwc_f32") else cur
        if src.dtype != torch.float32:
            raise NotImplementedError("the secondary output is derived from an fp32 primary output on the training path")
        return cur, (_AbsBf16Fn.apply(src) if out2 == 1 else cast_bf16(src))
    return cur


# ---------------------------------------------------------------------------------------------------------
# small differentiable plumbing
# ---------------------------------------------------------------------------------------------------------
class _CastBf16Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return ops.to_bf16(x)

    @staticmethod
    def backward(ctx, g):
        return g.float()


def cast_bf16(x: Tensor) -> Tensor:
    return _CastBf16Fn.apply(x) if (torch.is_grad_enabled() and x.requires_grad) else ops.to_bf16(x)


class _AbsBf16Fn(torch.autograd.Function):
    """torch.abs(y) as the bf16 input of h_a (models/google.py:283)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        out = torch.empty_like(x, dtype=torch.bfloat16)
        L.check(L.lib().mmc_abs_to_bf16(x.data_ptr(), x.numel(), out.data_ptr(), ops._stream()))
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        L.check(L.lib().mmc_abs_bwd(g.contiguous().data_ptr(), x.data_ptr(), x.numel(), dx.data_ptr(), ops._stream()))
        return dx


class _AddNoiseFn(torch.autograd.Function):
    """EntropyModel.quantize(x, "noise") (entropy_models.py:163-167): identity gradient."""

    @staticmethod
    def forward(ctx, x, noise):
        xv = x.float()
        if not (xv.is_contiguous() or ops._is_channels_last(xv)):
            xv = xv.contiguous()
        nz = ops._like_layout(noise, xv)
        out = torch.empty_like(xv)
        L.check(L.lib().mmc_quantize_noise(xv.data_ptr(), nz.data_ptr(), xv.numel(), out.data_ptr(), ops._stream()))
        return out

    @staticmethod
    def backward(ctx, g):
        return g, None


def add_noise(x: Tensor, noise: Tensor) -> Tensor:
    """x + noise in x's memory layout (contiguous or channels-last)."""
    return _AddNoiseFn.apply(x, noise)


# ---------------------------------------------------------------------------------------------------------
# entropy models
# ---------------------------------------------------------------------------------------------------------
class _GcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scales, means, noise, scale_bound, lik_bound):
        xv = x.float()
        if not (xv.is_contiguous() or ops._is_channels_last(xv)):
            xv = xv.contiguous()
        s = ops._like_layout(scales, xv)
        m = ops._like_layout(means, xv) if means is not None else None
        nz = ops._like_layout(noise, xv) if noise is not None else None
        x_hat, lik = ops.gc_forward(xv, s, m, nz, scale_bound, lik_bound)
        ctx.save_for_backward(xv, s, m, nz)
        ctx.bounds = (scale_bound, lik_bound)
        if noise is None:
            ctx.mark_non_differentiable(x_hat)
        return x_hat, lik

    @staticmethod
    def backward(ctx, g_xhat, g_lik):
        xv, s, m, nz = ctx.saved_tensors
        g = ops._like_layout(g_lik, xv)
        dx, ds, dm = ops.gc_backward(xv, s, m, nz, g, *ctx.bounds)
        if nz is not None and g_xhat is not None:
            dx = dx + g_xhat        # x_hat = x + noise
        return dx, ds, dm, None, None, None


def gc_forward(x, scales, means, noise, scale_bound, lik_bound):
    """GaussianConditional.forward with autograd when any input needs it (entropy_models.py:715-731)."""
    needs = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, scales, means))
    if needs:
        return _GcFn.apply(x, scales, means, noise, float(scale_bound), float(lik_bound))
    return ops.gc_forward(x, scales, means, noise, scale_bound, lik_bound)


_EB_SIZES = (3, 9, 9, 9, 3, 3, 3, 3, 3, 1, 3, 3, 3, 3)     # _matrix0..4, _bias0..4, _factor0..3 per channel


class _EbFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, noise, eb, *params):
        xv, outer, C, inner = ops._view_oci(x.float())
        nz = ops._like_layout(noise, xv)
        pk = eb._params()
        x_hat, lik = ops.eb_forward(xv, pk, nz, eb._lik_bound())
        ctx.save_for_backward(xv, nz)
        ctx.eb, ctx.dims = eb, (outer, C, inner)
        ctx.shapes = [tuple(p.shape) for p in params]
        return x_hat, lik

    @staticmethod
    def backward(ctx, g_xhat, g_lik):
        xv, nz = ctx.saved_tensors
        eb = ctx.eb
        outer, C, inner = ctx.dims
        g = ops._like_layout(g_lik, xv)
        dx, dparams = ops.eb_backward(xv, nz, g, eb._params(), eb._lik_bound(), outer, C, inner)
        if g_xhat is not None:
            dx = dx + g_xhat
        grads, off = [], 0
        for n, shp in zip(_EB_SIZES, ctx.shapes):
            grads.append(dparams[:, off:off + n].reshape(shp))
            off += n
        return (dx, None, None) + tuple(grads)


def eb_forward(x, eb, noise):
    """EntropyBottleneck.forward in noise mode with autograd (entropy_models.py:495-540): returns (x_hat, likelihood)."""
    params = [getattr(eb, f"_matrix{i}") for i in range(5)] + [getattr(eb, f"_bias{i}") for i in range(5)] + \
             [getattr(eb, f"_factor{i}") for i in range(4)]
    needs = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
    if needs:
        return _EbFn.apply(x, noise, eb, *params)
    return ops.eb_forward(x, eb._params(), noise, eb._lik_bound())


# ---------------------------------------------------------------------------------------------------------
# non-convolution steps of the fusion layers (csrc/esa.cu, csrc/attention.cu forward; csrc/fusion_bwd.cu backward)
# ---------------------------------------------------------------------------------------------------------
class _MaxPoolFn(torch.autograd.Function):
    """F.max_pool2d(kernel_size=k, stride=s) on an NHWC bf16 map (ESA, google.py:1448); the arg-max map is the saved state."""

    @staticmethod
    def forward(ctx, x, k, stride):
        y, idx = ops.maxpool_nhwc_bf16_idx(x, k, stride)
        ctx.save_for_backward(idx)
        ctx.geom = (tuple(x.shape), k, stride)
        return y

    @staticmethod
    def backward(ctx, gy):
        (idx,) = ctx.saved_tensors
        shape, k, stride = ctx.geom
        return ops.maxpool_nhwc_bf16_bwd(gy.contiguous(), idx, shape, k, stride), None, None


class _UpsampleAddFn(torch.autograd.Function):
    """F.interpolate(small, size of add, "bilinear") + add (google.py:1453-1455)."""

    @staticmethod
    def forward(ctx, small, add):
        ctx.small_hw = (small.shape[1], small.shape[2])
        return ops.upsample_bilinear_add_bf16(small, add)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        d_small = ops.upsample_bilinear_bwd_bf16(g, *ctx.small_hw) if ctx.needs_input_grad[0] else None
        return d_small, (g if ctx.needs_input_grad[1] else None)


class _SigmoidGateFn(torch.autograd.Function):
    """x * sigmoid(c) (google.py:1456-1459)."""

    @staticmethod
    def forward(ctx, x, c):
        ctx.save_for_backward(x, c)
        return ops.sigmoid_gate_bf16(x, c)

    @staticmethod
    def backward(ctx, g):
        x, c = ctx.saved_tensors
        return ops.sigmoid_gate_bwd_bf16(g.contiguous(), x, c)


def maxpool_nhwc(x: Tensor, k: int, stride: int) -> Tensor:
    return _MaxPoolFn.apply(x, k, stride) if (torch.is_grad_enabled() and x.requires_grad) else ops.maxpool_nhwc_bf16(x, k, stride)


def upsample_add(small: Tensor, add: Tensor) -> Tensor:
    if torch.is_grad_enabled() and (small.requires_grad or add.requires_grad):
        return _UpsampleAddFn.apply(small, add)
    return ops.upsample_bilinear_add_bf16(small, add)


def sigmoid_gate(x: Tensor, c: Tensor) -> Tensor:
    if torch.is_grad_enabled() and (x.requires_grad or c.requires_grad):
        return _SigmoidGateFn.apply(x, c)
    return ops.sigmoid_gate_bf16(x, c)


class _LayerNormFn(torch.autograd.Function):
    """nn.LayerNorm over the channels of a bf16 token grid, optionally of x + delta (the residual add of master.py:699 fused in
    front of norm2); with ``delta`` the outputs are (sum, normed), the sum being the updated residual stream."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, delta):
        ctx.eps, ctx.fused = float(eps), delta is not None
        if delta is not None:
            s, y = ops.layernorm_bf16(x, weight, bias, eps, delta=delta, want_sum=True)
            ctx.save_for_backward(s, weight)
            return s, y
        y = ops.layernorm_bf16(x, weight, bias, eps)
        ctx.save_for_backward(x if x.is_contiguous() else x.contiguous(), weight)
        return y

    @staticmethod
    def backward(ctx, *grads):
        v, weight = ctx.saved_tensors
        if ctx.fused:
            g_sum, g = grads
        else:
            g_sum, g = None, grads[0]
        if g is None:
            g = torch.zeros_like(v)
        dv, dw, db = ops.layernorm_bwd_bf16(g.contiguous(), v, weight, ctx.eps, g_sum=g_sum.contiguous() if g_sum is not None else None)
        return dv, dw.to(weight.dtype), db.to(weight.dtype), None, (dv if ctx.fused else None)


def layernorm(x: Tensor, norm: nn.LayerNorm, delta: Optional[Tensor] = None):
    """LayerNorm(x [+ delta]) on the kernels with autograd; returns normed, or (sum, normed) when ``delta`` is given."""
    needs = torch.is_grad_enabled() and (x.requires_grad or norm.weight.requires_grad or (delta is not None and delta.requires_grad))
    if needs:
        return _LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps, delta)
    if delta is not None:
        return ops.layernorm_bf16(x, norm.weight, norm.bias, norm.eps, delta=delta, want_sum=True)
    return ops.layernorm_bf16(x, norm.weight, norm.bias, norm.eps)


class _GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.gelu_bf16(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return ops.gelu_bwd_bf16(g.contiguous(), x)


def gelu(x: Tensor) -> Tensor:
    return _GeluFn.apply(x) if (torch.is_grad_enabled() and x.requires_grad) else ops.gelu_bf16(x)


class _WindowAttnFn(torch.autograd.Function):
    """WindowAttention core (master.py:535-568) with the roll / window partition of SwinTransformerBlock.forward (master.py:652-697)
    as index arithmetic; backward recomputes the probabilities (nothing but q and kv is saved)."""

    @staticmethod
    def forward(ctx, q, kv, table, window, shift, heads, scale):
        ctx.save_for_backward(q, kv, table)
        ctx.cfg = (window, shift, heads, scale)
        return ops.window_attention(q, kv, table, window, shift, heads, scale)

    @staticmethod
    def backward(ctx, g):
        q, kv, table = ctx.saved_tensors
        dq, dkv, dtable = ops.window_attention_bwd(q, kv, table, g.contiguous(), *ctx.cfg)
        return dq, dkv, dtable.to(table.dtype), None, None, None, None


def window_attention(q: Tensor, kv: Tensor, table: Tensor, window: int, shift: int, heads: int, scale: float) -> Tensor:
    if torch.is_grad_enabled() and (q.requires_grad or kv.requires_grad or table.requires_grad):
        return _WindowAttnFn.apply(q, kv, table, window, shift, heads, scale)
    return ops.window_attention(q, kv, table, window, shift, heads, scale)


def conv_alias(x: Tensor, weight4d: Tensor, bias: Optional[Tensor], alias: nn.Module, act_module=None, out_fmt: str = "nhwc_bf16") -> Tensor:
    """One conv layer on an NHWC bf16 map whose weight is a differentiable function of some other parameter (a Linear weight viewed as
    1x1 kernels, a 2x2 patch projection zero-padded to 3x3): ``alias`` is the layer the kernels see (its ``weight`` holds the same
    values as ``weight4d``), gradients flow to ``weight4d`` / ``bias``."""
    if torch.is_grad_enabled() and (x.requires_grad or weight4d.requires_grad or (bias is not None and bias.requires_grad)):
        return _ConvFn.apply(x, weight4d, bias, alias, act_module, "nhwc_bf16", out_fmt)
    layers = [alias] + ([act_module] if act_module is not None else [])
    return _run(layers, x, "nhwc_bf16", out_fmt)

"""Executor for the analysis / synthesis / hyper transform stacks (g_a, g_s, h_a, h_s).

A stack is the reference's ``nn.Sequential`` of conv/deconv layers, each optionally followed by GDN
/ IGDN / ReLU / LeakyReLU (compressai/models/google.py:143-161,254-269,363-377).  The executor
groups every conv with the op that follows it into ONE fused kernel launch, keeps activations
between layers as NHWC bf16 and touches fp32 only at the API edges.

Kernel choice per layer is by shape, not by backend: the tcgen05 implicit-GEMM kernel
(``mmc_conv_forward_tc``) takes every layer with Cin % 8 == 0, Cin >= 32 and Cout % 16 == 0; the CUDA-core
kernel (``mmc_conv_forward_direct``) takes the rest (3-channel image edges, odd test shapes).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib as L
from . import ops

FORMATS = ("nchw_f32", "nhwc_bf16", "nhwc_f32")

# Set by mmcodec.config; tests flip it to cross-check the two kernels against each other.
use_tensor_cores = True

# Two-source layers (cat([a, b]) -> conv, google.py:1153 ...) under autograd: True = the two-source kernel forward and per-source
# gradients backward (no concatenated tensor in either pass); False = torch.cat first (the round-1 training path; tests compare both).
two_source_train = True

# Arithmetic of the transform stacks (BASELINE.json north_star: "bf16/tf32 with fp32 accumulate ... 1e-2 relative in bf16, 1e-4 in
# fp32"): "bf16" = bf16 operands, fp32 accumulate, fused GDN (the fast path); "fp32" = every layer evaluated to ~1e-5 relative on
# the SAME bf16 tensor-core kernel through a three-term operand split (see _run_layers_fp32), fp32 activations between layers,
# fp32 GDN.  Inference only.
_precision = "bf16"


class precision:
    """``with mmcodec.precision("fp32"): out = net(x)`` -- context manager (also usable as ``mmcodec.precision.set("fp32")``)."""

    MODES = ("bf16", "fp32")

    def __init__(self, mode: str):
        if mode not in self.MODES:
            raise ValueError(f'precision must be one of {self.MODES}, got "{mode}"')
        self.mode = mode

    def __enter__(self):
        global _precision
        self.prev, _precision = _precision, self.mode
        return self

    def __exit__(self, *a):
        global _precision
        _precision = self.prev

    @classmethod
    def set(cls, mode: str):
        global _precision
        if mode not in cls.MODES:
            raise ValueError(f'precision must be one of {cls.MODES}, got "{mode}"')
        _precision = mode


def current_precision() -> str:
    return _precision


class QReLU8(nn.Module):
    """Marker for QReLU.apply(x, bit_depth=8, beta=100) (layers/layers.py:247-277; forward = clamp(x, 0, 255)) in a fused
    stack; used by the ssf2020 scale hyper-decoder (models/video/google.py:128-148)."""

    def forward(self, x: Tensor) -> Tensor:   # stand-alone use: same kernel family, no conv in front
        raise NotImplementedError("QReLU8 is only available fused behind a conv / deconv layer")


@dataclass
class Step:
    conv: nn.Module
    transposed: bool
    act: int = L.ACT_NONE
    gdn: Optional[nn.Module] = None


def parse_layers(layers) -> List[Step]:
    from .layers import GDN
    steps: List[Step] = []
    for m in layers:
        if isinstance(m, nn.ConvTranspose2d):
            steps.append(Step(m, True))
        elif isinstance(m, nn.Conv2d):
            steps.append(Step(m, False))
        elif isinstance(m, GDN):
            if not steps or steps[-1].gdn is not None or steps[-1].act != L.ACT_NONE:
                raise NotImplementedError("GDN must directly follow a conv/deconv layer in a fused stack")
            steps[-1].gdn = m
        elif isinstance(m, nn.LeakyReLU):
            if not steps or steps[-1].gdn is not None or abs(m.negative_slope - 0.01) > 1e-12:
                raise NotImplementedError("only LeakyReLU(0.01) directly after a conv/deconv is fused")
            steps[-1].act = L.ACT_LEAKY_RELU
        elif isinstance(m, QReLU8):
            if not steps or steps[-1].gdn is not None:
                raise NotImplementedError("QReLU must directly follow a conv/deconv layer in a fused stack")
            steps[-1].act = L.ACT_QRELU8
        elif isinstance(m, nn.ReLU):
            if not steps or steps[-1].gdn is not None:
                raise NotImplementedError("ReLU must directly follow a conv/deconv layer in a fused stack")
            steps[-1].act = L.ACT_RELU
        else:
            raise NotImplementedError(f"layer {type(m).__name__} is not part of the accelerated transform path")
    for s in steps:
        c = s.conv
        k = c.kernel_size[0]
        if (c.kernel_size[0] != c.kernel_size[1] or k not in (1, 3, 5) or c.stride[0] != c.stride[1]
                or c.stride[0] not in (1, 2) or c.padding != (k // 2, k // 2) or c.dilation != (1, 1) or c.groups != 1
                or (s.transposed and c.output_padding != (c.stride[0] - 1, c.stride[0] - 1))):
            raise NotImplementedError("only conv()/deconv()-style layers (k in {1,3,5}, stride in {1,2}, padding=k//2, "
                                      "output_padding=stride-1) are on the accelerated path")
    return steps


def _tc_eligible(cin: int, cout: int, gdn: bool) -> bool:
    """Shapes the tcgen05 kernel takes (mmc_conv_forward_tc): 64-channel K boxes (ragged tail zero-filled),
    N tiles that are multiples of 16; fused GDN needs the whole channel vector in one N tile."""
    if not use_tensor_cores or cin % 8 != 0 or cin < 32 or cout % 16 != 0 or cout > 1024:
        return False
    return (not gdn) or cout in (64, 128, 192)


def run_layers(layers, x: Tensor, in_fmt: str, out_fmt: str, out2: int = 0, _train_dispatch: bool = True):
    """Run a conv stack.  ``x``: (B,C,H,W) fp32 for "nchw_f32", (B,H,W,C) for the NHWC formats.
    Returns the output in ``out_fmt``; "nchw_f32" results may be channels-last strided views (same
    logical shape and values as the reference's NCHW tensor).  With ``out2`` (1: |output|, 2: output)
    also returns that tensor as NHWC bf16 (the h_a input, models/google.py:283,381)."""
    assert in_fmt in FORMATS and out_fmt in FORMATS
    pair = None
    if isinstance(x, (tuple, list)):
        # two NHWC bf16 feature maps whose channel concatenation feeds the first layer (google.py:1153 ...): the tensor-core kernel
        # reads both sources in its K loop; anything else (training, odd channel counts) concatenates first
        x1, x2 = x
        ops._require_cuda(x1, x2)
        first = next((m for m in layers if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d))), None)
        narrow_out = first is not None and isinstance(first, nn.ConvTranspose2d) and first.out_channels <= 4 and out_fmt == "nchw_f32"
        fusable = (in_fmt == "nhwc_bf16" and use_tensor_cores and first is not None and x1.shape[-1] % 64 == 0 and x2.shape[-1] % 8 == 0
                   and x1.shape[:3] == x2.shape[:3] and (first.out_channels % 16 == 0 or narrow_out))
        wants = torch.is_grad_enabled() and (x1.requires_grad or x2.requires_grad or (first is not None and any(p.requires_grad for p in first.parameters())))
        if fusable and wants and _train_dispatch and _precision != "fp32" and two_source_train:
            # training: the same two-source kernel forward, per-source weight / input gradients backward (no concatenation either way)
            from . import autograd as AG
            return AG.run_layers_train(layers, (x1, x2), in_fmt, out_fmt, out2)
        if fusable and not wants:
            pair = (x1.contiguous(), x2.contiguous())
            x = pair[0]
        else:
            x = torch.cat((x1, x2), dim=-1)
    ops._require_cuda(x)
    if _train_dispatch and torch.is_grad_enabled() and pair is None:
        from . import autograd as AG
        if AG.wants_grad(layers, x) and _precision == "fp32":
            raise NotImplementedError('mmcodec.precision("fp32") is an inference mode; the training path computes in bf16 / fp32-accumulate')
        if AG.wants_grad(layers, x):
            return AG.run_layers_train(layers, x, in_fmt, out_fmt, out2)   # same kernels, recorded for backward
    steps = parse_layers(layers)
    if not steps:
        raise ValueError("empty transform stack")
    if _precision == "fp32":
        if pair is not None:
            x = torch.cat((pair[0].float(), pair[1].float()), dim=-1)
        return _run_layers_fp32(steps, x, in_fmt, out_fmt, out2)
    with torch.no_grad():
        cur, fmt = x, in_fmt
        if fmt == "nchw_f32":
            if x.dim() != 4:
                raise ValueError("expected a 4-D (B, C, H, W) input")
            cur = x.float()
            if ops._is_channels_last(cur):
                cur, fmt = cur.permute(0, 2, 3, 1), "nhwc_f32"   # already NHWC in memory
            else:
                cur = cur.contiguous()
        y2 = None
        for i, s in enumerate(steps):
            last = i == len(steps) - 1
            c = s.conv
            cin, cout = (c.in_channels, c.out_channels)
            if fmt == "nchw_f32":
                B, C, H, W = cur.shape
            else:
                B, H, W, C = cur.shape
            if i == 0 and pair is not None:
                C += pair[1].shape[-1]
            if C != cin:
                raise ValueError(f"expected {cin} input channels, got {C}")
            stride, k = c.stride[0], c.kernel_size[0]
            gdn_ok = s.gdn is None or cout in (64, 128, 192)
            # shape class of this layer (see mmc_conv_forward_tc)
            edge_in = (use_tensor_cores and not s.transposed and cin <= 8 and cout % 16 == 0 and cout <= 1024 and gdn_ok
                       and k in (3, 5) and fmt == "nchw_f32")
            edge_out = (use_tensor_cores and s.transposed and cout <= 4 and stride == 2 and last and out_fmt == "nchw_f32"
                        and s.gdn is None and not out2 and cin % 8 == 0 and cin >= 32)
            tc = edge_in or edge_out or _tc_eligible(cin, cout, s.gdn is not None)
            in_layout = None
            if edge_in:
                dpad = ops.conv_desc(False, B, H, W, cin, cout, k, stride, L.BF16, L.NHWC_PAD8, L.BF16, L.NHWC)
                cur, fmt, in_layout = ops.pad_to_nhwc8(cur, dpad), "nhwc_bf16", L.NHWC_PAD8
            elif tc and fmt != "nhwc_bf16":
                # API-edge input of a tensor-core layer: one conversion pass to NHWC bf16
                cur = ops.nchw_to_nhwc_bf16(cur) if fmt == "nchw_f32" else ops.to_bf16(cur)
                fmt = "nhwc_bf16"
            if not last:
                ofmt = "nhwc_bf16"
            elif out_fmt == "nchw_f32":
                # narrow outputs (the 3-channel image) are written planar; wide ones stay NHWC in memory
                ofmt = "nchw_f32" if (edge_out or (not tc and cout <= 4)) else "nhwc_f32"
            else:
                ofmt = out_fmt
            if in_layout is None:
                in_layout = L.NCHW if fmt.startswith("nchw") else L.NHWC
            d = ops.conv_desc(s.transposed, B, H, W, cin, cout, k, stride,
                              L.F32 if fmt.endswith("f32") else L.BF16, in_layout,
                              L.F32 if ofmt.endswith("f32") else L.BF16, L.NCHW if ofmt.startswith("nchw") else L.NHWC,
                              act=s.act, gdn=(L.GDN_NONE if s.gdn is None else (L.GDN_INVERSE if s.gdn.inverse else L.GDN_FORWARD)),
                              out2=(out2 if last else 0))
            beta_eff = gamma_eff = gamma_bf16 = None
            if s.gdn is not None:
                beta_eff, gamma_eff, gamma_bf16 = s.gdn.effective_params()
            bias = c.bias.detach() if c.bias is not None else None
            if bias is not None and bias.dtype != torch.float32:
                bias = bias.float()
            name = getattr(c, "_mmc_name", "conv")
            if tc:
                src = pair if (i == 0 and pair is not None) else cur
                out = ops.conv_forward_tc(d, src, c.packed_weight(d), bias, beta_eff, gamma_bf16, name=name)
            else:
                out = ops.conv_forward_direct(d, cur, c.f32_weight(), bias, beta_eff, gamma_eff, name=name)
            if d.out2_bf16:
                out, y2 = out
            cur, fmt = out, ofmt
        if out_fmt == "nchw_f32" and fmt == "nhwc_f32":
            cur = cur.permute(0, 3, 1, 2)   # logical NCHW, channels-last memory
        return (cur, y2) if out2 else cur


def _fp32_shadow(conv: nn.Module) -> nn.Module:
    """The layer with its input channels tripled and the weights laid out [w_hi | w_hi | w_lo] (w_hi = bf16(w), w_lo = bf16(w - w_hi)):
    against activations split [x_hi | x_lo | x_hi] (ops.split_bf16x3) the bf16 tensor-core kernel accumulates
    x_hi w_hi + x_lo w_hi + x_hi w_lo in fp32.  Derived data, rebuilt when the weight's version or storage changes."""
    from .layers import Conv2d, ConvTranspose2d
    key = (conv.weight._version, conv.weight.data_ptr())
    cached = getattr(conv, "_mmc_fp32_shadow", None)
    if cached is None or cached[0] != key:
        w = conv.weight.detach().float()
        hi = w.bfloat16().float()
        lo = (w - hi).bfloat16().float()
        k, s = conv.kernel_size[0], conv.stride[0]
        with torch.device("meta"):
            if isinstance(conv, nn.ConvTranspose2d):
                m = ConvTranspose2d(3 * conv.in_channels, conv.out_channels, kernel_size=k, stride=s, padding=k // 2, output_padding=s - 1, bias=False)
            else:
                m = Conv2d(3 * conv.in_channels, conv.out_channels, kernel_size=k, stride=s, padding=k // 2, bias=False)
        cat_dim = 0 if isinstance(conv, nn.ConvTranspose2d) else 1          # the input-channel dimension of the weight tensor
        m._parameters["weight"] = nn.Parameter(torch.cat([hi, hi, lo], dim=cat_dim).contiguous(), requires_grad=False)
        m._mmc_name = getattr(conv, "_mmc_name", "conv") + ".x3"
        cached = (key, m)
        object.__setattr__(conv, "_mmc_fp32_shadow", cached)    # not a registered child: no state_dict entry
    return cached[1]


def _fp32_gdn(gdn: nn.Module, x_nhwc: Tensor) -> Tensor:
    """GDN / IGDN on an fp32 NHWC map (layers/gdn.py:77-92): norm = beta + gamma x^2 as a 1x1 tensor-core convolution over the
    three-term split of x^2 (weights [g_hi | g_hi | g_lo], bias beta), then y = x * rsqrt(norm) (IGDN: * sqrt) elementwise."""
    from .layers import Conv2d
    beta_eff, gamma_eff, _ = gdn.effective_params()
    key = (beta_eff.data_ptr(), gamma_eff.data_ptr(), gdn._cache_key)
    cached = getattr(gdn, "_mmc_fp32_norm", None)
    if cached is None or cached[0] != key:
        C = beta_eff.numel()
        g = gamma_eff.detach().float().reshape(C, C, 1, 1)
        hi = g.bfloat16().float()
        lo = (g - hi).bfloat16().float()
        with torch.device("meta"):
            m = Conv2d(3 * C, C, kernel_size=1, stride=1, padding=0, bias=True)
        m._parameters["weight"] = nn.Parameter(torch.cat([hi, hi, lo], dim=1).contiguous(), requires_grad=False)
        m._parameters["bias"] = nn.Parameter(beta_eff.detach().float().clone(), requires_grad=False)
        m._mmc_name = "gdn.norm.x3"
        cached = (key, m)
        object.__setattr__(gdn, "_mmc_fp32_norm", cached)
    m = cached[1]
    B, H, W, C = x_nhwc.shape
    if not _tc_eligible(3 * C, C, False):
        return ops.gdn_forward(x_nhwc.permute(0, 3, 1, 2), beta_eff, gamma_eff, gdn.inverse).permute(0, 2, 3, 1)   # fp32 CUDA-core kernel
    x2 = ops.split_bf16x3(x_nhwc, square=True)
    d = ops.conv_desc(False, B, H, W, 3 * C, C, 1, 1, L.BF16, L.NHWC, L.F32, L.NHWC)
    norm = ops.conv_forward_tc(d, x2, m.packed_weight(d), m.bias.detach(), None, None, name="gdn.norm.x3")
    return ops.gdn_apply_f32(x_nhwc, norm, gdn.inverse)


def _fp32_edge_conv(conv: nn.Module, x_nchw: Tensor, act: int, name: str) -> Tensor:
    """Image-edge convolution (Cin <= 4) in fp32 mode on the tensor-core image-edge kernel: the three product terms need 3 Cin > 8
    staged channels, so they are two launches -- [x_hi | x_lo] against [w_hi | w_hi] (+ bias) and x_hi against w_lo -- summed in
    fp32.  (The fp32 CUDA-core kernel took 124 ms for the 64 x 768x512 batch; this takes ~1.5 ms.)  Returns NHWC fp32."""
    from .layers import Conv2d
    cin, cout, k, stride = conv.in_channels, conv.out_channels, conv.kernel_size[0], conv.stride[0]
    key = (conv.weight._version, conv.weight.data_ptr())
    cached = getattr(conv, "_mmc_fp32_edge", None)
    if cached is None or cached[0] != key:
        w = conv.weight.detach().float()
        hi = w.bfloat16().float()
        lo = (w - hi).bfloat16().float()
        with torch.device("meta"):
            ma = Conv2d(2 * cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=False)
            mb = Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=False)
        ma._parameters["weight"] = nn.Parameter(torch.cat([hi, hi], dim=1).contiguous(), requires_grad=False)
        mb._parameters["weight"] = nn.Parameter(lo.contiguous(), requires_grad=False)
        ma._mmc_name, mb._mmc_name = name + ".x2", name + ".lo"
        cached = (key, ma, mb)
        object.__setattr__(conv, "_mmc_fp32_edge", cached)
    _, ma, mb = cached
    B, C, H, W = x_nchw.shape
    x_hi = x_nchw.bfloat16().float()
    x6 = torch.cat([x_hi, (x_nchw - x_hi).bfloat16().float()], dim=1).contiguous()
    bias = conv.bias.detach().float() if conv.bias is not None else None
    outs = []
    for m, xin, b in ((ma, x6, bias), (mb, x_hi.contiguous(), None)):
        ci = xin.shape[1]
        d = ops.conv_desc(False, B, H, W, ci, cout, k, stride, L.BF16, L.NHWC_PAD8, L.F32, L.NHWC)
        outs.append(ops.conv_forward_tc(d, ops.pad_to_nhwc8(xin, d), m.packed_weight(d), b, None, None, name=m._mmc_name))
    out = outs[0].add_(outs[1])
    if act == L.ACT_RELU:
        out = torch.relu_(out)
    elif act == L.ACT_LEAKY_RELU:
        out = torch.nn.functional.leaky_relu_(out, 0.01)
    elif act == L.ACT_QRELU8:
        out = out.clamp_(0, 255)
    return out


def _run_layers_fp32(steps: List[Step], x: Tensor, in_fmt: str, out_fmt: str, out2: int):
    """precision("fp32"): the stack with fp32 activations between layers (NHWC), each tensor-core-eligible conv / deconv as ONE bf16
    tensor-core launch over three-term split operands (fp32 accumulate, bias and ReLU / LeakyReLU fused, fp32 out), GDN / IGDN
    on the fp32 kernel (csrc/gdn.cu), the image-edge convolution (3 input channels) on the fp32 CUDA-core kernel with its GDN
    fused.  Inputs may be fp32 whatever ``in_fmt`` says (the models hand fp32 latents on in this mode).  Outputs are fp32:
    "nchw_f32" as in the fast path, "nhwc_f32" AND "nhwc_bf16" as (B, H, W, C) fp32; ``out2`` likewise fp32."""
    with torch.no_grad():
        cur = x.float()
        nchw = in_fmt == "nchw_f32"
        if nchw and cur.dim() != 4:
            raise ValueError("expected a 4-D (B, C, H, W) input")
        if nchw and ops._is_channels_last(cur):
            cur, nchw = cur.permute(0, 2, 3, 1), False
        cur = cur.contiguous()
        for i, s in enumerate(steps):
            last = i == len(steps) - 1
            c = s.conv
            cin, cout, stride, k = c.in_channels, c.out_channels, c.stride[0], c.kernel_size[0]
            if nchw:
                B, C, H, W = cur.shape
            else:
                B, H, W, C = cur.shape
            if C != cin:
                raise ValueError(f"expected {cin} input channels, got {C}")
            bias = c.bias.detach().float() if c.bias is not None else None
            name = getattr(c, "_mmc_name", "conv")
            beta_eff = gamma_eff = None
            if s.gdn is not None:
                beta_eff, gamma_eff, _ = s.gdn.effective_params()
            narrow = s.transposed and cout <= 4 and stride == 2 and last and out_fmt == "nchw_f32" and not out2
            tc = use_tensor_cores and cin % 4 == 0 and _tc_eligible(3 * cin, cout, False) or \
                (use_tensor_cores and narrow and cin % 8 == 0 and cin >= 32)
            if tc:
                if nchw:
                    cur, nchw = cur.permute(0, 2, 3, 1).contiguous(), False
                x3 = ops.split_bf16x3(cur)
                sh = _fp32_shadow(c)
                d3 = ops.conv_desc(s.transposed, B, H, W, 3 * cin, cout, k, stride, L.BF16, L.NHWC, L.F32, L.NCHW if narrow else L.NHWC,
                                   act=s.act, gdn=L.GDN_NONE, out2=0)
                cur = ops.conv_forward_tc(d3, x3, sh.packed_weight(d3), bias, None, None, name=name + ".x3")
                nchw = narrow
                if s.gdn is not None:
                    cur = _fp32_gdn(s.gdn, cur)
            elif (use_tensor_cores and nchw and not s.transposed and cin <= 4 and cout % 16 == 0 and cout <= 1024 and k in (3, 5)):
                cur, nchw = _fp32_edge_conv(c, cur, s.act, name), False
                if s.gdn is not None:
                    cur = _fp32_gdn(s.gdn, cur)
            else:
                # image-edge / odd shapes: the fp32 CUDA-core kernel (GDN fused in fp32 when it fits)
                planar_out = last and out_fmt == "nchw_f32" and cout <= 4
                d = ops.conv_desc(s.transposed, B, H, W, cin, cout, k, stride, L.F32, L.NCHW if nchw else L.NHWC, L.F32,
                                  L.NCHW if planar_out else L.NHWC, act=s.act,
                                  gdn=(L.GDN_NONE if s.gdn is None else (L.GDN_INVERSE if s.gdn.inverse else L.GDN_FORWARD)), out2=0)
                cur = ops.conv_forward_direct(d, cur, c.f32_weight(), bias, beta_eff, gamma_eff, name=name)
                nchw = planar_out
        y2 = None
        if out2:
            src = cur.permute(0, 2, 3, 1) if nchw else cur
            y2 = torch.abs(src) if out2 == 1 else src
        if out_fmt == "nchw_f32" and not nchw:
            cur = cur.permute(0, 3, 1, 2)        # logical NCHW, channels-last memory (as on the fast path)
        elif out_fmt != "nchw_f32" and nchw:
            cur = cur.permute(0, 2, 3, 1).contiguous()
        return (cur, y2) if out2 else cur


class TransformStack(nn.Sequential):
    """``nn.Sequential`` with the same children / state_dict keys as the reference's g_a, g_s, h_a,
    h_s; ``forward`` runs the fused executor instead of calling the children one by one."""

    def forward(self, x: Tensor) -> Tensor:
        return run_layers(list(self), x, "nchw_f32", "nchw_f32")

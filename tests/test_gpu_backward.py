"""GPU parity tests of the backward (training) kernels against torch CPU autograd of the oracle's ops in float64.
Operands are rounded to bf16 first, so the comparison isolates accumulation order: rel-RMS <= 2e-3."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import mmcodec  # noqa: E402
from mmcodec import ops  # noqa: E402


def dev():
    return torch.device("cuda", 0)


def rel_rms(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(torch.sqrt(((a - b) ** 2).mean() / (b ** 2).mean().clamp_min(1e-30)))


def bf16r(t):
    return t.to(torch.bfloat16).to(torch.float64)


@pytest.mark.parametrize("transposed,B,Cin,Cout,H,W,k,s", [
    (False, 2, 64, 128, 16, 24, 5, 2),
    (False, 1, 192, 192, 12, 20, 3, 1),
    (False, 2, 96, 48, 9, 70, 1, 1),       # ragged M / N tiles, W > 64 (two row segments), 1x1
    (False, 2, 8, 128, 32, 48, 5, 2),      # image-edge layer: 3 channels zero-padded to 8 -> one 16-wide N tile
    (True, 2, 128, 64, 8, 12, 5, 2),
    (True, 1, 384, 16, 10, 14, 5, 2),      # three M tiles
    (True, 2, 192, 192, 6, 9, 3, 1),
    (False, 1, 384, 192, 24, 136, 5, 1),   # tran_conv shape class
])
def test_wgrad_tc(transposed, B, Cin, Cout, H, W, k, s):
    g = torch.Generator().manual_seed(H * W + Cin)
    x = bf16r(torch.randn(B, Cin, H, W, generator=g))
    pad = k // 2
    if transposed:
        w = torch.zeros(Cin, Cout, k, k, dtype=torch.float64, requires_grad=True)
        y = F.conv_transpose2d(x, w, stride=s, padding=pad, output_padding=s - 1)
    else:
        w = torch.zeros(Cout, Cin, k, k, dtype=torch.float64, requires_grad=True)
        y = F.conv2d(x, w, stride=s, padding=pad)
    dy = bf16r(torch.randn(y.shape, generator=g))
    y.backward(dy)
    ref = w.grad
    x_nhwc = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev())
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).to(dev())
    # conv: S = grad_output (low resolution), L = input;  deconv: S = input, L = grad_output
    dw = ops.wgrad(x_nhwc, dy_nhwc, k, s) if transposed else ops.wgrad(dy_nhwc, x_nhwc, k, s)
    assert tuple(dw.shape) == tuple(ref.shape)
    assert rel_rms(dw, ref) < 2e-3

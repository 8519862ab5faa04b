"""Reference CPU execution path, restated -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference has no kernels of its own: its CPU path is a sequence of stock torch ops
(``nn.Conv2d`` / ``nn.ConvTranspose2d`` / ``F.conv2d`` / elementwise ATen).  This module restates
that sequence functionally on a plain ``state_dict`` so that it can (a) run on the GPU box, where
/root/reference does not exist, as the timed CPU baseline of ``bench.py`` and (b) serve as the
stage-wise parity reference at sizes where the plain-C oracle would take minutes.  It is pinned
to the real reference by tests/golden (tests/test_oracle_golden.py).

All citations are relative to /root/reference/CompressAI.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ---- layers ---------------------------------------------------------------------------------
def conv(sd: Dict[str, Tensor], name: str, x: Tensor, stride: int = 2) -> Tensor:
    """compressai/models/utils.py:128-135"""
    w = sd[name + ".weight"]
    return F.conv2d(x, w, sd[name + ".bias"], stride=stride, padding=w.shape[-1] // 2)


def deconv(sd: Dict[str, Tensor], name: str, x: Tensor, stride: int = 2) -> Tensor:
    """compressai/models/utils.py:138-146"""
    w = sd[name + ".weight"]
    return F.conv_transpose2d(x, w, sd[name + ".bias"], stride=stride, padding=w.shape[-1] // 2,
                              output_padding=stride - 1)


class _LowerBoundFn(torch.autograd.Function):
    """compressai/ops/bound_ops.py:36-56: max(x, bound) whose gradient also passes where it pushes x up towards the bound"""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        pass_through_if = (x >= bound) | (grad_output < 0)
        return pass_through_if.type(grad_output.dtype) * grad_output, None


def lower_bound(x: Tensor, bound: float) -> Tensor:
    """compressai/ops/bound_ops.py:36-80 (forward: torch.max(x, bound))"""
    return _LowerBoundFn.apply(x, torch.tensor([bound], dtype=x.dtype))


def gdn(sd: Dict[str, Tensor], name: str, x: Tensor, inverse: bool = False, beta_min: float = 1e-6) -> Tensor:
    """compressai/layers/gdn.py:77-92 + compressai/ops/parametrizers.py:47-64"""
    pedestal = (2.0 ** -18) ** 2
    beta = lower_bound(sd[name + ".beta"], (beta_min + pedestal) ** 0.5) ** 2 - pedestal
    gamma = lower_bound(sd[name + ".gamma"], pedestal ** 0.5) ** 2 - pedestal
    C = x.shape[1]
    norm = F.conv2d(x ** 2, gamma.reshape(C, C, 1, 1), beta)
    norm = torch.sqrt(norm) if inverse else torch.rsqrt(norm)
    return x * norm


# ---- entropy models -------------------------------------------------------------------------
def quantize(x: Tensor, mode: str, means: Optional[Tensor] = None, noise: Optional[Tensor] = None) -> Tensor:
    """compressai/entropy_models/entropy_models.py:157-182 (noise passed in)"""
    if mode == "noise":
        return x + noise
    out = x.clone()
    if means is not None:
        out -= means
    out = torch.round(out)
    if mode == "dequantize":
        if means is not None:
            out += means
        return out
    assert mode == "symbols"
    return out.int()


def eb_logits_cumulative(sd: Dict[str, Tensor], name: str, v: Tensor) -> Tensor:
    """entropy_models.py:457-477; v is (C, 1, L)"""
    logits = v
    for i in range(5):
        logits = torch.matmul(F.softplus(sd[f"{name}._matrix{i}"]), logits)
        logits = logits + sd[f"{name}._bias{i}"]
        if i < 4:
            logits = logits + torch.tanh(sd[f"{name}._factor{i}"]) * torch.tanh(logits)
    return logits


def eb_likelihood(sd, name, v):
    """entropy_models.py:480-492"""
    lower = eb_logits_cumulative(sd, name, v - 0.5)
    upper = eb_logits_cumulative(sd, name, v + 0.5)
    sign = -torch.sign(lower + upper)
    return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))


def eb_forward(sd, name, x, noise=None, lik_bound: float = 1e-9):
    """entropy_models.py:495-540; returns (x_hat, likelihood) shaped like x (N, C, ...)."""
    C = x.shape[1]
    perm = [1, 0] + list(range(2, x.dim()))
    xp = x.permute(*perm).contiguous()
    shape = xp.shape
    values = xp.reshape(C, 1, -1)
    medians = sd[f"{name}.quantiles"][:, :, 1:2]
    if noise is not None:
        outputs = values + noise.permute(*perm).reshape(C, 1, -1)
    else:
        outputs = quantize(values, "dequantize", medians)
    lik = eb_likelihood(sd, name, outputs)
    if lik_bound > 0:
        lik = lower_bound(lik, lik_bound)
    outputs = outputs.reshape(shape).permute(*perm).contiguous()
    lik = lik.reshape(shape).permute(*perm).contiguous()
    return outputs, lik


def gc_likelihood(x_hat, scales, means=None, scale_bound: float = 0.11):
    """entropy_models.py:692-709"""
    values = x_hat - means if means is not None else x_hat
    scales = lower_bound(scales, scale_bound)
    values = torch.abs(values)
    const = float(-(2 ** -0.5))
    upper = 0.5 * torch.erfc(const * ((0.5 - values) / scales))
    lower = 0.5 * torch.erfc(const * ((-0.5 - values) / scales))
    return upper - lower


def gc_forward(x, scales, means=None, noise=None, scale_bound=0.11, lik_bound=1e-9):
    """entropy_models.py:715-731"""
    out = quantize(x, "noise" if noise is not None else "dequantize", means, noise)
    lik = gc_likelihood(out, scales, means, scale_bound)
    if lik_bound > 0:
        lik = lower_bound(lik, lik_bound)
    return out, lik


def get_scale_table(lo: float = 0.11, hi: float = 256.0, levels: int = 64) -> Tensor:
    """compressai/models/google.py:208-214"""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels))


def build_indexes(scales: Tensor, scale_table: Tensor, scale_bound: float = 0.11) -> Tensor:
    """entropy_models.py:735-740"""
    scales = lower_bound(scales, scale_bound)
    indexes = scales.new_full(scales.size(), len(scale_table) - 1).int()
    for s in scale_table[:-1]:
        indexes -= (scales <= s).int()
    return indexes


def eb_build_indexes(size) -> Tensor:
    """entropy_models.py:542-553"""
    C = size[1]
    view = [1] * len(size)
    view[1] = -1
    return torch.arange(C).view(*view).int().repeat(size[0], 1, *size[2:])


# ---- models ---------------------------------------------------------------------------------
def _seq_g_a(sd, x):
    """models/google.py:143-151 (same layer stack in all three model families)"""
    for i in (0, 2, 4):
        x = gdn(sd, f"g_a.{i + 1}", conv(sd, f"g_a.{i}", x))
    return conv(sd, "g_a.6", x)


def _seq_g_s(sd, y_hat):
    """models/google.py:153-161"""
    x = y_hat
    for i in (0, 2, 4):
        x = gdn(sd, f"g_s.{i + 1}", deconv(sd, f"g_s.{i}", x), inverse=True)
    return deconv(sd, "g_s.6", x)


def factorized_forward(sd, x, noise=None):
    """FactorizedPrior.forward: models/google.py:172-182"""
    y = _seq_g_a(sd, x)
    y_hat, y_lik = eb_forward(sd, "entropy_bottleneck", y, noise)
    x_hat = _seq_g_s(sd, y_hat)
    return {"x_hat": x_hat, "likelihoods": {"y": y_lik}, "y": y, "y_hat": y_hat}


def _h_a(sd, y, act):
    z = act(conv(sd, "h_a.0", y, stride=1))
    z = act(conv(sd, "h_a.2", z))
    return conv(sd, "h_a.4", z)


def hyperprior_forward(sd, x, noise=None):
    """ScaleHyperprior.forward: models/google.py:281-295; training mode when ``noise`` = {"z", "y"} holds the uniform draws"""
    y = _seq_g_a(sd, x)
    z = _h_a(sd, torch.abs(y), F.relu)
    z_hat, z_lik = eb_forward(sd, "entropy_bottleneck", z, None if noise is None else noise["z"])
    s = F.relu(deconv(sd, "h_s.0", z_hat))
    s = F.relu(deconv(sd, "h_s.2", s))
    scales_hat = F.relu(conv(sd, "h_s.4", s, stride=1))
    y_hat, y_lik = gc_forward(y, scales_hat, noise=None if noise is None else noise["y"])
    x_hat = _seq_g_s(sd, y_hat)
    return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}, "y": y, "z": z, "z_hat": z_hat,
            "scales_hat": scales_hat, "y_hat": y_hat}


def _mean_scale_params(sd, z_hat):
    """MeanScaleHyperprior h_s: models/google.py:371-377"""
    s = F.leaky_relu(deconv(sd, "h_s.0", z_hat))
    s = F.leaky_relu(deconv(sd, "h_s.2", s))
    return conv(sd, "h_s.4", s, stride=1)


def mean_scale_forward(sd, x, noise=None):
    """MeanScaleHyperprior.forward: models/google.py:379-391; ``noise``: see hyperprior_forward"""
    y = _seq_g_a(sd, x)
    z = _h_a(sd, y, F.leaky_relu)
    z_hat, z_lik = eb_forward(sd, "entropy_bottleneck", z, None if noise is None else noise["z"])
    gaussian_params = _mean_scale_params(sd, z_hat)
    scales_hat, means_hat = gaussian_params.chunk(2, 1)
    y_hat, y_lik = gc_forward(y, scales_hat, means_hat, noise=None if noise is None else noise["y"])
    x_hat = _seq_g_s(sd, y_hat)
    return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}, "y": y, "z": z, "z_hat": z_hat,
            "scales_hat": scales_hat, "means_hat": means_hat, "y_hat": y_hat}


def mean_scale_compress_symbols(sd, x, scale_table):
    """MeanScaleHyperprior.compress up to the int32 symbols/indexes handed to the rANS coder:
    models/google.py:393-404 + entropy_models.py:237-246,559-566.  z_hat is what
    entropy_bottleneck.decompress would return: dequantize(symbols, medians)."""
    y = _seq_g_a(sd, x)
    z = _h_a(sd, y, F.leaky_relu)
    medians = sd["entropy_bottleneck.quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
    z_symbols = quantize(z, "symbols", medians)
    z_indexes = eb_build_indexes(z.size())
    z_hat = z_symbols.float() + medians
    gaussian_params = _mean_scale_params(sd, z_hat)
    scales_hat, means_hat = gaussian_params.chunk(2, 1)
    y_indexes = build_indexes(scales_hat, scale_table)
    y_symbols = quantize(y, "symbols", means_hat)
    return {"y_symbols": y_symbols, "y_indexes": y_indexes, "z_symbols": z_symbols, "z_indexes": z_indexes,
            "y": y, "z": z, "scales_hat": scales_hat, "means_hat": means_hat}


def hyperprior_compress_symbols(sd, x, scale_table):
    """ScaleHyperprior.compress up to symbols/indexes: models/google.py:324-332"""
    y = _seq_g_a(sd, x)
    z = _h_a(sd, torch.abs(y), F.relu)
    medians = sd["entropy_bottleneck.quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
    z_symbols = quantize(z, "symbols", medians)
    z_indexes = eb_build_indexes(z.size())
    z_hat = z_symbols.float() + medians
    s = F.relu(deconv(sd, "h_s.0", z_hat))
    s = F.relu(deconv(sd, "h_s.2", s))
    scales_hat = F.relu(conv(sd, "h_s.4", s, stride=1))
    y_indexes = build_indexes(scales_hat, scale_table)
    y_symbols = quantize(y, "symbols")
    return {"y_symbols": y_symbols, "y_indexes": y_indexes, "z_symbols": z_symbols, "z_indexes": z_indexes,
            "y": y, "z": z, "scales_hat": scales_hat}


FORWARD = {"factorized": factorized_forward, "hyperprior": hyperprior_forward, "mean-scale": mean_scale_forward}


def bpp(out, num_pixels: int) -> float:
    """utils/eval_model/__main__t.py:197-200"""
    return float(sum(torch.log(l).sum() / (-math.log(2) * num_pixels) for l in out["likelihoods"].values()))


# ---- multi-modality two-branch codec (compressai/models/google.py:696-1459) ---------------------------------
def _context_entropy_stage(sd, y, noise=None):
    """Hyperprior + masked context model + entropy parameters (google.py:800-822, 1196-1211); eval mode, or training mode
    when ``noise`` = {"z", "y_hat", "y"} holds the three uniform draws the reference makes."""
    z = _h_a(sd, y, F.leaky_relu)
    z_hat, z_lik = eb_forward(sd, "entropy_bottleneck", z, noise=None if noise is None else noise["z"])
    params = _mean_scale_params(sd, z_hat)
    if noise is None:
        y_hat = quantize(y, "dequantize")                              # no means: what the context model / decoder see
    else:
        y_hat = quantize(y, "noise", noise=noise["y_hat"])
    w = sd["context_prediction.weight"] * sd["context_prediction.mask"]  # layers/layers.py:75-78
    ctx = F.conv2d(y_hat, w, sd["context_prediction.bias"], padding=2)
    g = torch.cat((params, ctx), dim=1)
    for i in (0, 2, 4):
        g = F.conv2d(g, sd[f"entropy_parameters.{i}.weight"], sd[f"entropy_parameters.{i}.bias"])
        if i < 4:
            g = F.leaky_relu(g)
    scales_hat, means_hat = g.chunk(2, 1)
    _, y_lik = gc_forward(y, scales_hat, means_hat, noise=None if noise is None else noise["y"])
    return y_hat, y_lik, z_lik, {"z": z, "scales_hat": scales_hat, "means_hat": means_hat}


def mbt2018_forward(sd, x, noise=None):
    """JointAutoregressiveHierarchicalPriors.forward (the zoo's mbt2018, google.py:499-520): stock g_a / g_s around the
    context-model entropy stage."""
    y = _seq_g_a(sd, x)
    y_hat, y_lik, z_lik, extra = _context_entropy_stage(sd, y, noise)
    return {"x_hat": _seq_g_s(sd, y_hat), "likelihoods": {"y": y_lik, "z": z_lik}, "y": y, "y_hat": y_hat, **extra}


def mm_r_forward(sd, x):
    """JointAutoregressiveHierarchicalPriors_R.forward (google.py:800-825)."""
    h = {}
    a = x
    for i in (1, 2, 3):
        a = gdn(sd, f"enc1.g_a_gdn{i}", conv(sd, f"enc1.g_a_conv{i}", a))
        h[f"ga{i}"] = a
    y = conv(sd, "enc1.g_a_conv4", a)
    y_hat, y_lik, z_lik, extra = _context_entropy_stage(sd, y)
    s = y_hat
    for i in (1, 2, 3):
        s = gdn(sd, f"dec1.g_s_gdn{i}", deconv(sd, f"dec1.g_s_conv{i}", s), inverse=True)
        h[f"gs{i}"] = s
    x_hat = deconv(sd, "dec1.g_s_conv4", s)
    return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}, "hidden": h, "y": y, "y_hat": y_hat, **extra}


def esa(sd, name, x):
    """ESA.forward (google.py:1445-1459)."""
    c = lambda n, t, **k: F.conv2d(t, sd[f"{name}.{n}.weight"], sd[f"{name}.{n}.bias"], **k)
    c1_ = c("conv1", x)
    c1 = c("conv2", c1_, stride=2)
    v_max = F.max_pool2d(c1, kernel_size=7, stride=3)
    v_range = F.relu(c("conv_max", v_max, padding=1))
    c3 = F.relu(c("conv3", v_range, padding=1))
    c3 = c("conv3_", c3, padding=1)
    c3 = F.interpolate(c3, (x.size(2), x.size(3)), mode="bilinear", align_corners=False)
    cf = c("conv_f", c1_)
    return x * torch.sigmoid(c("conv4", c3 + cf))


def mm_fuse(sd, i, own, guide):
    """One cross-modality fusion block: eg_ext(2i-1)(own), eg_ext(2i)(guide) -> cat -> tran_conv_i -> ESA_i
    (google.py:1151-1156 and its five repeats)."""
    e1 = F.relu(F.conv2d(own, sd[f"eg_ext{2 * i - 1}.0.weight"], sd[f"eg_ext{2 * i - 1}.0.bias"], padding=1))
    e2 = F.relu(F.conv2d(guide, sd[f"eg_ext{2 * i}.0.weight"], sd[f"eg_ext{2 * i}.0.bias"], padding=1))
    f = conv(sd, f"tran_conv{i}", torch.cat((e1, e2), dim=1), stride=1)
    return esa(sd, f"attention{i}", f)


def mm_d_forward(sd, x, hidden, noise=None):
    """JointAutoregressiveHierarchicalPriors_D.forward (google.py:1140-1248); ``noise``: see _context_entropy_stage."""
    fuse = lambda i, own, guide: mm_fuse(sd, i, own, guide)

    a = gdn(sd, "pic2_g_a_gdn1", conv(sd, "pic2_g_a_conv1", x))
    f = fuse(1, a, hidden["ga1"])
    a = gdn(sd, "pic2_g_a_gdn2", conv(sd, "pic2_g_a_conv2", torch.cat((a, f), 1)))
    f = fuse(2, a, hidden["ga2"])
    a = gdn(sd, "pic2_g_a_gdn3", conv(sd, "pic2_g_a_conv3", torch.cat((a, f), 1)))
    f = fuse(3, a, hidden["ga3"])
    y = conv(sd, "pic2_g_a_conv4", torch.cat((a, f), 1))
    y_hat, y_lik, z_lik, extra = _context_entropy_stage(sd, y, noise)
    s = gdn(sd, "pic2_g_s_gdn1", deconv(sd, "pic2_g_s_conv1", y_hat), inverse=True)
    f = fuse(4, s, hidden["gs1"])
    s = gdn(sd, "pic2_g_s_gdn2", deconv(sd, "pic2_g_s_conv2", torch.cat((s, f), 1)), inverse=True)
    f = fuse(5, s, hidden["gs2"])
    s = gdn(sd, "pic2_g_s_gdn3", deconv(sd, "pic2_g_s_conv3", torch.cat((s, f), 1)), inverse=True)
    f = fuse(6, s, hidden["gs3"])
    x_hat = deconv(sd, "pic2_g_s_conv4", torch.cat((s, f), 1))
    return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}, "y": y, "y_hat": y_hat, **extra}



# ---- Master_compresser (RGB-T paper reproduction, compressai/models/master.py) -------------------------------------------
def _c3(sd, name, x, stride=1):
    """conv3x3 / conv1x1 (master.py:18-25): padding = k // 2."""
    w = sd[name + ".weight"]
    return F.conv2d(x, w, sd[name + ".bias"], stride=stride, padding=w.shape[-1] // 2)


def residual_block(sd, name, x):
    """ResidualBlock.forward (master.py:48-62): LeakyReLU after BOTH convs, optional 1x1 skip."""
    out = F.leaky_relu(_c3(sd, name + ".conv2", F.leaky_relu(_c3(sd, name + ".conv1", x))))
    return out + (_c3(sd, name + ".skip", x) if (name + ".skip.weight") in sd else x)


def feature_encoder(sd, name, x, stride):
    """Feature_encoder.forward (master.py:78-89)."""
    first = _c3(sd, name + ".conv1", x, stride)
    out = first
    for i in (1, 2, 3):
        out = residual_block(sd, f"{name}.resblock{i}", out)
    return out + first


def feature_decoder(sd, name, x, stride):
    """Feature_decoder.forward (master.py:110-118): 3 residual blocks + 1x1 shortcut, then a k=3 transposed conv."""
    out = x
    for i in (1, 2, 3):
        out = residual_block(sd, f"{name}.resblock{i}", out)
    out = out + _c3(sd, name + ".conv", x)
    return deconv(sd, name + ".deconv1", out, stride=stride)


def channel_aligner(sd, name, master_feat, guide_feat):
    """Channel_aligner.forward (master.py:177-210): one shared 4-conv trunk applied to each feature map, conv5 / conv6 heads,
    global average -> beta (from the master features) and gamma (from the guide features); guide * gamma + beta."""
    def trunk(t):
        for i in (1, 2, 3, 4):
            t = F.leaky_relu(_c3(sd, f"{name}.conv{i}", t))
        return t
    beta = _c3(sd, name + ".conv5", trunk(master_feat)).mean(dim=(2, 3), keepdim=True)
    gamma = _c3(sd, name + ".conv6", trunk(guide_feat)).mean(dim=(2, 3), keepdim=True)
    return gamma * guide_feat + beta, beta, gamma


def _to_windows(t, ws):
    """window_partition (master.py:431-443) on (B, H, W, C) -> (B * nW, ws * ws, C)."""
    B, H, W, C = t.shape
    return t.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def _from_windows(t, ws, B, H, W):
    """window_reverse (master.py:446-460)."""
    return t.view(B, H // ws, W // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def relative_position_index(ws: int) -> Tensor:
    """master.py:512-522: index into the (2 ws - 1)^2 bias table for every (query, key) pair of a ws x ws window."""
    r = torch.arange(ws)
    ci, cj = torch.meshgrid(r, r, indexing="ij")
    ci, cj = ci.flatten(), cj.flatten()
    return (ci[:, None] - ci[None, :] + ws - 1) * (2 * ws - 1) + (cj[:, None] - cj[None, :] + ws - 1)


def shift_attention_mask(H: int, W: int, ws: int, shift: int) -> Tensor:
    """master.py:625-643: 0 / -100 mask between tokens that the cyclic shift brought together from different image regions."""
    region = torch.zeros(H, W)
    bands = (slice(0, -ws), slice(-ws, -shift), slice(-shift, None))
    n = 0
    for hs in bands:
        for wsl in bands:
            region[hs, wsl] = n
            n += 1
    m = _to_windows(region.view(1, H, W, 1), ws).squeeze(-1)          # (nW, ws*ws)
    diff = m[:, None, :] - m[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def swin_cross_block(sd, name, x, guide, H, W, ws, shift, heads=3):
    """SwinTransformerBlock.forward (master.py:652-706) with WindowAttention.forward (master.py:535-568): queries from the
    master tokens, keys / values from the guide tokens, both normalised by the SAME norm1."""
    B, L, C = x.shape
    ln = lambda t, n: F.layer_norm(t, (C,), sd[f"{name}.{n}.weight"], sd[f"{name}.{n}.bias"])
    lin = lambda t, n: F.linear(t, sd[f"{name}.{n}.weight"], sd[f"{name}.{n}.bias"])
    if min(H, W) <= ws:                                                # master.py:601-603
        ws, shift = min(H, W), 0
    q_src, kv_src = ln(x, "norm1").view(B, H, W, C), ln(guide, "norm1").view(B, H, W, C)
    if shift:
        q_src, kv_src = (torch.roll(t, (-shift, -shift), (1, 2)) for t in (q_src, kv_src))
    qw, kw = _to_windows(q_src, ws), _to_windows(kv_src, ws)
    nWB, N, hd = qw.shape[0], ws * ws, C // heads
    q = lin(qw, "attn.qkv1").view(nWB, N, heads, hd).transpose(1, 2) * hd ** -0.5
    kv = lin(kw, "attn.qkv2").view(nWB, N, 2, heads, hd)
    k, v = kv[:, :, 0].transpose(1, 2), kv[:, :, 1].transpose(1, 2)
    att = q @ k.transpose(-2, -1)
    table = sd[f"{name}.attn.relative_position_bias_table"]
    att = att + table[relative_position_index(ws).view(-1)].view(N, N, heads).permute(2, 0, 1)
    if shift:
        att = att.view(B, -1, heads, N, N) + shift_attention_mask(H, W, ws, shift).to(att)[None, :, None]
        att = att.view(nWB, heads, N, N)
    o = lin((att.softmax(-1) @ v).transpose(1, 2).reshape(nWB, N, C), "attn.proj")
    o = _from_windows(o, ws, B, H, W)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    x = x + o.reshape(B, L, C)
    return x + lin(F.gelu(lin(ln(x, "norm2"), "mlp.fc1")), "mlp.fc2")


def spatial_aligner(sd, name, x, guide):
    """Spatial_aligner.forward (master.py:730-742): 2x2 patch embedding of both maps, two cross-attention blocks (window 4,
    second one shifted by 2), then -- as the reference does -- the (B, L, 96) token tensor is REINTERPRETED (``.view``, no
    transpose) as (B, 96, H/2, W/2) and a 2x2 stride-2 transposed conv restores the resolution."""
    B, _, H, W = x.shape
    emb = lambda n, t: F.conv2d(t, sd[f"{name}.{n}.proj.weight"], sd[f"{name}.{n}.proj.bias"], stride=2).flatten(2).transpose(1, 2)
    tok, gtok = emb("patch_embeding1", x), emb("patch_embeding2", guide)
    for i in (0, 1):
        tok = swin_cross_block(sd, f"{name}.blocks.{i}", tok, gtok, H // 2, W // 2, 4, 0 if i == 0 else 2)
    grid = tok.contiguous().view(B, tok.shape[-1], H // 2, W // 2)
    return F.conv_transpose2d(grid, sd[name + ".recovery.weight"], sd[name + ".recovery.bias"], stride=2)


def master_decoder(sd, name, y_hat, hidden):
    """Master_decoder.forward (master.py:776-811)."""
    g = [hidden["gs1"], hidden["gs2"], hidden["gs3"]]
    if (name + ".downsample1.weight") in sd:                           # 1-channel master: guide maps are at twice the size
        g = [conv(sd, f"{name}.downsample{i + 1}", t) for i, t in enumerate(g)]
    s = y_hat
    for i in (1, 2, 3):
        s = gdn(sd, f"{name}.g_s_gdn{i}", deconv(sd, f"{name}.g_s_conv{i}", s), inverse=True)
        s = torch.cat((spatial_aligner(sd, f"{name}.sp_aligner{i}", s, g[i - 1]), s), dim=1)
    first_stride = 2                                                   # master.py:900 always builds the decoder with 2
    return deconv(sd, name + ".g_s_conv4", s, stride=first_stride)


def master_forward(sd, x, guided_hat, guided_hidden, noise=None):
    """Master_compresser.forward (master.py:904-951).  Strides follow the constructor (master.py:840-850): 3-channel master
    -> (master_stride, guided_stride) = (2, 1); 1-channel master -> (1, 2)."""
    ms, gs = (2, 1) if x.shape[1] == 3 else (1, 2)
    xf = feature_encoder(sd, "fencoder1", x, ms)
    gf = feature_encoder(sd, "fencoder2", guided_hat, gs)
    aligned, beta, gamma = channel_aligner(sd, "ch_aligner", xf, gf)
    a = torch.cat((xf, aligned), dim=1)
    for i in (0, 2, 4):
        a = gdn(sd, f"g_a.{i + 1}", conv(sd, f"g_a.{i}", a))
    y = conv(sd, "g_a.6", a)
    y_hat, y_lik, z_lik, extra = _context_entropy_stage(sd, y, noise)
    feat_hat = master_decoder(sd, "decoder", y_hat, guided_hidden)
    x_hat = feature_decoder(sd, "fdecoder", torch.cat((feat_hat, aligned), dim=1), ms)
    return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}, "y": y, "y_hat": y_hat, "beta": beta, "gamma": gamma,
            "x_feature": xf, "guided_align": aligned, "x_feature_hat": feat_hat, **extra}


def masked_conv_mask(weight_shape, mask_type: str = "A") -> Tensor:
    """MaskedConv2d mask buffer (compressai/layers/layers.py:64-72)."""
    mask = torch.ones(tuple(weight_shape))
    h, w = mask.shape[-2:]
    mask[:, :, h // 2, w // 2 + (mask_type == "B"):] = 0
    mask[:, :, h // 2 + 1:] = 0
    return mask


# ---- ssf2020 video codec (compressai/models/video/google.py:55-508) --------------------------------------------
def _ssf_stack(sd, prefix, x, transposed, n, act=F.relu):
    """Encoder / Decoder / HyperEncoder / HyperDecoder nn.Sequential: conv at indexes 0, 2, 4[, 6] with ReLU between."""
    op = deconv if transposed else conv
    for i in range(n):
        x = op(sd, f"{prefix}.{2 * i}", x)
        if i < n - 1:
            x = act(x)
    return x


def ssf_hyperprior(sd, p, y):
    """ScaleSpaceFlow.Hyperprior.forward, eval mode (google.py:171-180)."""
    z = _ssf_stack(sd, f"{p}.hyper_encoder", y, False, 3)
    z_hat, z_lik = eb_forward(sd, f"{p}.entropy_bottleneck", z)
    s = z_hat
    for i in (1, 2, 3):
        s = deconv(sd, f"{p}.hyper_decoder_scale.deconv{i}", s).clamp(min=0, max=255)   # QReLU fwd, layers/layers.py:277
    means = _ssf_stack(sd, f"{p}.hyper_decoder_mean", z_hat, True, 3)
    y_hat, y_lik = gc_forward(y, s, means)          # == quantize_ste(y - means) + means in the forward pass
    return y_hat, {"y": y_lik, "z": z_lik}, {"z": z, "z_hat": z_hat, "scales": s, "means": means}


def gaussian_volume(x, sigma: float, num_levels: int):
    """google.py:331-355 with models/utils.py:155-189"""
    k = 2 * int(math.ceil(3 * sigma)) + 1
    khalf = (k - 1) / 2.0
    t = torch.linspace(-khalf, khalf, steps=k, dtype=x.dtype)
    pdf = torch.exp(-0.5 * (t / sigma).pow(2))
    k1 = pdf / pdf.sum()
    kernel = torch.mm(k1[:, None], k1[None, :])

    def blur(v):
        pad = k // 2
        v = F.pad(v, (pad, pad, pad, pad), mode="replicate")
        return F.conv2d(v, kernel.expand(v.size(1), 1, k, k), groups=v.size(1))

    volume = [x.unsqueeze(2)]
    x = blur(x)
    volume += [x.unsqueeze(2)]
    for i in range(1, num_levels):
        x = F.avg_pool2d(x, kernel_size=(2, 2), stride=(2, 2))
        x = blur(x)
        interp = x
        for _ in range(0, i):
            interp = F.interpolate(interp, scale_factor=2, mode="bilinear", align_corners=False)
        volume.append(interp.unsqueeze(2))
    return torch.cat(volume, dim=2)


def warp_volume(volume, flow, scale_field):
    """google.py:357-375"""
    N, C, _, H, W = volume.size()
    theta = torch.eye(2, 3).unsqueeze(0).expand(N, 2, 3)
    grid = F.affine_grid(theta, (N, C, H, W), align_corners=False)
    update_grid = grid + flow.permute(0, 2, 3, 1).float()
    update_scale = scale_field.permute(0, 2, 3, 1).float()
    volume_grid = torch.cat((update_grid, update_scale), dim=-1).unsqueeze(1)
    out = F.grid_sample(volume.float(), volume_grid, padding_mode="border", align_corners=False)
    return out.squeeze(2)


def ssf_forward(sd, frames, num_levels: int = 5, sigma0: float = 1.5):
    """ScaleSpaceFlow.forward, eval mode (google.py:212-273)."""
    trace = []
    y = _ssf_stack(sd, "img_encoder", frames[0], False, 4)
    y_hat, lik, ex = ssf_hyperprior(sd, "img_hyperprior", y)
    x_ref = _ssf_stack(sd, "img_decoder", y_hat, True, 4)
    recs, liks = [x_ref], [{"keyframe": lik}]
    trace.append({"y": y, "y_hat": y_hat, **ex})
    for x_cur in frames[1:]:
        y_m = _ssf_stack(sd, "motion_encoder", torch.cat((x_cur, x_ref), dim=1), False, 4)
        y_m_hat, lik_m, ex_m = ssf_hyperprior(sd, "motion_hyperprior", y_m)
        motion_info = _ssf_stack(sd, "motion_decoder", y_m_hat, True, 4)
        flow, scale_field = motion_info.chunk(2, dim=1)
        volume = gaussian_volume(x_ref, sigma0, num_levels)
        x_pred = warp_volume(volume, flow, scale_field)
        x_res = x_cur - x_pred
        y_r = _ssf_stack(sd, "res_encoder", x_res, False, 4)
        y_r_hat, lik_r, ex_r = ssf_hyperprior(sd, "res_hyperprior", y_r)
        x_res_hat = _ssf_stack(sd, "res_decoder", torch.cat((y_r_hat, y_m_hat), dim=1), True, 4)
        trace.append({"x_ref": x_ref, "y_motion": y_m, "y_motion_hat": y_m_hat, "motion_info": motion_info, "volume": volume,
                      "x_pred": x_pred, "x_res": x_res, "y_res": y_r, "y_res_hat": y_r_hat, "x_res_hat": x_res_hat,
                      "motion": ex_m, "residual": ex_r})
        x_ref = x_pred + x_res_hat
        recs.append(x_ref)
        liks.append({"motion": lik_m, "residual": lik_r})
    return {"x_hat": recs, "likelihoods": liks, "trace": trace}


# ---- colour transforms (compressai/transforms/functional.py:26-137) ---------------------------------------------
def rgb2ycbcr(rgb):
    r, g, b = rgb.chunk(3, -3)
    Kr, Kg, Kb = 0.2126, 0.7152, 0.0722
    y = Kr * r + Kg * g + Kb * b
    cb = 0.5 * (b - y) / (1 - Kb) + 0.5
    cr = 0.5 * (r - y) / (1 - Kr) + 0.5
    return torch.cat((y, cb, cr), dim=-3)


def ycbcr2rgb(ycbcr):
    y, cb, cr = ycbcr.chunk(3, -3)
    Kr, Kg, Kb = 0.2126, 0.7152, 0.0722
    r = y + (2 - 2 * Kr) * (cr - 0.5)
    b = y + (2 - 2 * Kb) * (cb - 0.5)
    g = (y - Kr * r - Kb * b) / Kg
    return torch.cat((r, g, b), dim=-3)


def yuv_444_to_420(yuv):
    y, u, v = yuv.chunk(3, 1)
    return y, F.avg_pool2d(u, kernel_size=2, stride=2), F.avg_pool2d(v, kernel_size=2, stride=2)


def yuv_420_to_444(yuv):
    y, u, v = yuv
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=False)
    return torch.cat((y, up(u), up(v)), dim=1)

"""GPU parity tests of the backward (training) kernels against torch CPU autograd of the oracle's ops in float64.
Operands are rounded to bf16 first, so the comparison isolates accumulation order: rel-RMS <= 2e-3."""
import math

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


# ---------------------------------------------------------------------------------------------------------
# layer-level backward through the autograd Functions vs float64 autograd of the oracle's ops
# ---------------------------------------------------------------------------------------------------------
from oracle import torch_port as tp  # noqa: E402
from mmcodec import autograd as AG  # noqa: E402
from mmcodec.layers import GDN, conv, deconv  # noqa: E402
from mmcodec.transforms import run_layers  # noqa: E402


def cos(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return float((a * b).sum() / (a.norm() * b.norm()).clamp_min(1e-30))


@pytest.mark.parametrize("kind,cin,cout,k,s,act", [
    ("conv", 64, 128, 5, 2, "leaky"), ("deconv", 128, 64, 5, 2, "relu"), ("conv", 192, 96, 3, 1, None),
    ("conv", 384, 192, 5, 1, None), ("conv", 96, 64, 1, 1, "leaky"), ("deconv", 192, 288, 5, 2, "leaky")])
def test_conv_layer_backward(kind, cin, cout, k, s, act):
    torch.manual_seed(cin + cout)
    m = (conv if kind == "conv" else deconv)(cin, cout, kernel_size=k, stride=s).to(dev())
    with torch.no_grad():
        m.weight.copy_(m.weight.to(torch.bfloat16).float())       # bf16-exact weights: the comparison isolates the kernels
        m.bias.copy_(m.bias.to(torch.bfloat16).float())
    B, H, W = 2, 12, 20
    x = torch.randn(B, H, W, cin, device=dev()).to(torch.bfloat16).requires_grad_(True)
    layers = [m] + ([torch.nn.LeakyReLU()] if act == "leaky" else [torch.nn.ReLU()] if act == "relu" else [])
    y = run_layers(layers, x, "nhwc_bf16", "nhwc_bf16")
    assert y.requires_grad
    gy = torch.randn_like(y.float()).to(torch.bfloat16)
    y.backward(gy)
    # reference
    xr = x.detach().double().cpu().permute(0, 3, 1, 2).requires_grad_(True)
    wr, br = m.weight.detach().double().cpu().requires_grad_(True), m.bias.detach().double().cpu().requires_grad_(True)
    if kind == "conv":
        yr = F.conv2d(xr, wr, br, stride=s, padding=k // 2)
    else:
        yr = F.conv_transpose2d(xr, wr, br, stride=s, padding=k // 2, output_padding=s - 1)
    if act:
        yr = F.leaky_relu(yr) if act == "leaky" else F.relu(yr)
    assert rel_rms(y.detach().float().permute(0, 3, 1, 2), yr.detach()) < 1e-2
    yr.backward(gy.double().cpu().permute(0, 3, 1, 2))
    assert rel_rms(x.grad.float().permute(0, 3, 1, 2), xr.grad) < 2e-2
    assert rel_rms(m.weight.grad, wr.grad) < 2e-2 and rel_rms(m.bias.grad, br.grad) < 2e-2


def test_edge_layers_backward():
    """Image-edge conv (3 -> N, fp32 NCHW input, no input gradient) and reconstruction deconv (N -> 3, planar fp32 output)."""
    torch.manual_seed(3)
    m = conv(3, 128).to(dev())
    d = deconv(128, 3).to(dev())
    with torch.no_grad():
        for p in list(m.parameters()) + list(d.parameters()):
            p.copy_(p.to(torch.bfloat16).float())          # bf16-exact parameters: same ReLU mask on both sides
    x = torch.rand(2, 3, 32, 48, device=dev())
    y = run_layers([m, torch.nn.ReLU()], x, "nchw_f32", "nhwc_bf16")
    gy = torch.randn_like(y.float()).to(torch.bfloat16)
    y.backward(gy)
    xr = x.double().cpu().to(torch.bfloat16).double()
    wr, br = m.weight.detach().double().cpu().requires_grad_(True), m.bias.detach().double().cpu().requires_grad_(True)
    yr = F.relu(F.conv2d(xr, wr, br, stride=2, padding=2))
    yr.backward(gy.double().cpu().permute(0, 3, 1, 2))
    assert rel_rms(m.weight.grad, wr.grad) < 2e-2 and rel_rms(m.bias.grad, br.grad) < 2e-2
    h = torch.randn(2, 8, 12, 128, device=dev()).to(torch.bfloat16).requires_grad_(True)
    xh = run_layers([d], h, "nhwc_bf16", "nchw_f32")
    assert tuple(xh.shape) == (2, 3, 16, 24) and xh.dtype == torch.float32
    gx = torch.randn_like(xh)
    xh.backward(gx)
    hr = h.detach().double().cpu().permute(0, 3, 1, 2).requires_grad_(True)
    wr, br = d.weight.detach().double().cpu().requires_grad_(True), d.bias.detach().double().cpu().requires_grad_(True)
    yr = F.conv_transpose2d(hr, wr, br, stride=2, padding=2, output_padding=1)
    yr.backward(gx.double().cpu())
    assert rel_rms(h.grad.float().permute(0, 3, 1, 2), hr.grad) < 2e-2
    assert rel_rms(d.weight.grad, wr.grad) < 2e-2 and rel_rms(d.bias.grad, br.grad) < 2e-2


@pytest.mark.parametrize("inverse,C", [(False, 128), (True, 192)])
def test_conv_gdn_backward(inverse, C):
    """Fused conv + GDN / IGDN forward (pre-GDN activations as secondary output) and its backward."""
    torch.manual_seed(C)
    m = (deconv if inverse else conv)(64, C).to(dev())
    gdn = GDN(C, inverse=inverse).to(dev())
    with torch.no_grad():
        m.weight.mul_(3.0)
        gdn.gamma.add_(torch.rand_like(gdn.gamma) * 0.02)
        gdn.beta.add_(torch.rand_like(gdn.beta) * 0.5)
    x = torch.randn(2, 10, 14, 64, device=dev()).to(torch.bfloat16).requires_grad_(True)
    y = run_layers([m, gdn], x, "nhwc_bf16", "nhwc_bf16")
    gy = torch.randn_like(y.float()).to(torch.bfloat16)
    y.backward(gy)
    sd = {"c.weight": m.weight.detach().double().cpu().requires_grad_(True), "c.bias": m.bias.detach().double().cpu().requires_grad_(True),
          "g.beta": gdn.beta.detach().double().cpu().requires_grad_(True), "g.gamma": gdn.gamma.detach().double().cpu().requires_grad_(True)}
    xr = x.detach().double().cpu().permute(0, 3, 1, 2).requires_grad_(True)
    pre = (tp.deconv if inverse else tp.conv)(sd, "c", xr)
    yr = tp.gdn(sd, "g", pre, inverse=inverse)
    assert rel_rms(y.detach().float().permute(0, 3, 1, 2), yr.detach()) < 1e-2
    yr.backward(gy.double().cpu().permute(0, 3, 1, 2))
    assert rel_rms(x.grad.float().permute(0, 3, 1, 2), xr.grad) < 3e-2
    assert rel_rms(m.weight.grad, sd["c.weight"].grad) < 3e-2 and rel_rms(m.bias.grad, sd["c.bias"].grad) < 3e-2
    assert rel_rms(gdn.beta.grad, sd["g.beta"].grad) < 3e-2 and rel_rms(gdn.gamma.grad, sd["g.gamma"].grad) < 3e-2


def test_gaussian_conditional_backward():
    torch.manual_seed(5)
    shape = (2, 16, 6, 10)
    x = (torch.randn(shape) * 4).to(dev()).requires_grad_(True)
    scales = torch.exp(torch.empty(shape).uniform_(np.log(0.05), np.log(30.0))).to(dev()).requires_grad_(True)
    means = (torch.randn(shape) * 2).to(dev()).requires_grad_(True)
    noise = torch.empty(shape).uniform_(-0.5, 0.5).to(dev())
    _, lik = AG.gc_forward(x, scales, means, noise, 0.11, 1e-9)
    g = torch.randn(shape, device=dev())
    lik.backward(g)
    xr, sr, mr = (t.detach().double().cpu().requires_grad_(True) for t in (x, scales, means))
    _, lr = tp.gc_forward(xr, sr, mr, noise=noise.double().cpu())
    lr.backward(g.double().cpu())
    for mine, ref in ((x.grad, xr.grad), (scales.grad, sr.grad), (means.grad, mr.grad)):
        assert rel_rms(mine, ref) < 1e-3


def test_entropy_bottleneck_backward():
    torch.manual_seed(6)
    C = 24
    eb = mmcodec.EntropyBottleneck(C).to(dev())
    with torch.no_grad():
        for i in range(5):
            getattr(eb, f"_matrix{i}").add_(torch.randn_like(getattr(eb, f"_matrix{i}")) * 0.3)
            if i < 4:
                getattr(eb, f"_factor{i}").add_(torch.randn_like(getattr(eb, f"_factor{i}")) * 0.3)
    sd = {f"eb.{k}": v.detach().double().cpu().requires_grad_(True) for k, v in eb.named_parameters()}
    for layout in ("nchw", "channels_last"):
        eb.zero_grad()
        for v in sd.values():
            v.grad = None
        x = (torch.randn(2, C, 5, 7) * 3).to(dev())
        if layout == "channels_last":
            x = x.contiguous(memory_format=torch.channels_last)
        x.requires_grad_(True)
        noise = torch.empty(2, C, 5, 7).uniform_(-0.5, 0.5).to(dev())
        x_hat, lik = AG.eb_forward(x, eb, noise)
        g = torch.randn(2, C, 5, 7, device=dev())
        (lik * g).sum().backward()
        xr = x.detach().double().cpu().contiguous().requires_grad_(True)
        _, lr = tp.eb_forward(sd, "eb", xr, noise=noise.double().cpu())
        (lr * g.double().cpu()).sum().backward()
        assert rel_rms(lik.detach(), lr.detach()) < 1e-4
        assert rel_rms(x.grad, xr.grad) < 1e-3
        for k, p in eb.named_parameters():
            if k == "quantiles":
                continue
            assert rel_rms(p.grad, sd[f"eb.{k}"].grad) < 2e-3, (layout, k)
    # aux loss: gradient reaches the quantiles only
    eb.zero_grad()
    eb.loss().backward()
    assert eb.quantiles.grad is not None and eb._matrix0.grad is None


def test_module_forwards_are_differentiable_like_the_reference():
    """EntropyBottleneck / GaussianConditional / GDN module calls made the way the reference's training code makes
    them (``self.entropy_bottleneck(z)`` inside a custom model, entropy_models.py:495-540,715-731, layers/gdn.py:77-92) carry
    a graph: rate gradients reach the caller's tensors and the modules' parameters, and match the oracle's autograd."""
    torch.manual_seed(11)
    C = 128
    d = dev()
    # --- EntropyBottleneck(z) in training mode -------------------------------------------------------------------
    eb = mmcodec.EntropyBottleneck(C).to(d).train()
    z = (torch.randn(2, C, 4, 6) * 3).to(d).requires_grad_(True)
    torch.manual_seed(12)
    z_hat, lik = eb(z)
    assert lik.grad_fn is not None and z_hat.grad_fn is not None
    noise = (z_hat - z).detach()                                   # the draw the module made
    assert float(noise.abs().max()) <= 0.5
    torch.log(lik).sum().backward()
    sd = {f"eb.{k}": v.detach().double().cpu().requires_grad_(True) for k, v in eb.named_parameters()}
    zr = z.detach().double().cpu().requires_grad_(True)
    _, lr = tp.eb_forward(sd, "eb", zr, noise=noise.double().cpu())
    torch.log(lr).sum().backward()
    assert rel_rms(z.grad, zr.grad) < 2e-3
    assert eb._matrix0.grad is not None and rel_rms(eb._matrix0.grad, sd["eb._matrix0"].grad) < 5e-3
    # eval mode: values only (round() has no gradient in the reference either)
    eb.eval()
    z_hat_e, lik_e = eb(z)
    assert z_hat_e.grad_fn is None and lik_e.grad_fn is None
    # --- GaussianConditional(y, scales, means) -------------------------------------------------------------------
    gc = mmcodec.GaussianConditional(None).to(d).train()
    shape = (2, 16, 6, 10)
    y = (torch.randn(shape) * 4).to(d).requires_grad_(True)
    scales = torch.exp(torch.empty(shape).uniform_(np.log(0.05), np.log(30.0))).to(d).requires_grad_(True)
    means = (torch.randn(shape) * 2).to(d).requires_grad_(True)
    y_hat, lk = gc(y, scales, means)
    nz = (y_hat - y).detach()
    torch.log(lk).sum().backward()
    yr, sr, mr = (t.detach().double().cpu().requires_grad_(True) for t in (y, scales, means))
    _, lr = tp.gc_forward(yr, sr, mr, noise=nz.double().cpu())
    torch.log(lr).sum().backward()
    for mine, ref in ((y.grad, yr.grad), (scales.grad, sr.grad), (means.grad, mr.grad)):
        assert mine is not None and rel_rms(mine, ref) < 2e-3
    gc.eval()                                                       # eval mode: the gradient reaches the scales only
    s2 = scales.detach().clone().requires_grad_(True)
    _, lk = gc(y.detach(), s2, means.detach())
    torch.log(lk).sum().backward()
    sr2 = s2.detach().double().cpu().requires_grad_(True)
    _, lr = tp.gc_forward(y.detach().double().cpu(), sr2, means.detach().double().cpu())
    torch.log(lr).sum().backward()
    assert rel_rms(s2.grad, sr2.grad) < 2e-3
    # --- stand-alone GDN module ------------------------------------------------------------------------------------
    for inverse in (False, True):
        gdn = GDN(C, inverse=inverse).to(d)
        with torch.no_grad():
            gdn.gamma.add_(torch.rand_like(gdn.gamma) * 0.02)
            gdn.beta.add_(torch.rand_like(gdn.beta) * 0.5)
        x = torch.randn(2, C, 10, 12, device=d).to(torch.bfloat16).float().requires_grad_(True)
        out = gdn(x)
        assert out.grad_fn is not None
        gy = torch.randn_like(out).to(torch.bfloat16).float()
        out.backward(gy)
        sdg = {"g.beta": gdn.beta.detach().double().cpu().requires_grad_(True), "g.gamma": gdn.gamma.detach().double().cpu().requires_grad_(True)}
        xr = x.detach().double().cpu().requires_grad_(True)
        yr = tp.gdn(sdg, "g", xr, inverse=inverse)
        assert rel_rms(out.detach(), yr.detach()) < 1e-4
        yr.backward(gy.double().cpu())
        assert rel_rms(x.grad, xr.grad) < 3e-2
        assert rel_rms(gdn.beta.grad, sdg["g.beta"].grad) < 3e-2 and rel_rms(gdn.gamma.grad, sdg["g.gamma"].grad) < 3e-2
        with torch.no_grad():
            assert gdn(x).grad_fn is None


# ---------------------------------------------------------------------------------------------------------
# end to end: training-mode forward + backward of the second-modality branch vs the oracle's autograd (fp32 CPU)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("esa_on_kernels", [False, True])
def test_mm_branch_training_step_vs_oracle(esa_on_kernels):
    import json
    import os
    from weights import make_mm_state_dict
    from mmcodec import models_mm as mm
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_mm.npz"))

    def load(tag, cls, seed):
        shapes = {k: tuple(v[0]) for k, v in json.loads(str(g[f"{tag}_state_dict"])).items()}
        sd = {k: torch.from_numpy(v) for k, v in make_mm_state_dict(shapes, seed).items()}
        net = cls(192, 192)
        net.update()
        net.load_state_dict({**net.state_dict(), **sd})
        sd["context_prediction.mask"] = tp.masked_conv_mask(shapes["context_prediction.weight"], "A")
        return net.to(dev()), sd

    net_r, sd_r = load("r", mm.JointAutoregressiveHierarchicalPriors_R, 0)
    net_d, sd_d = load("d", mm.JointAutoregressiveHierarchicalPriors_D, 1)
    x, depth = torch.from_numpy(g["x"]), torch.from_numpy(g["depth"])
    torch.set_num_threads(8)
    with torch.no_grad():
        ref_r = tp.mm_r_forward(sd_r, x)
    gen = torch.Generator().manual_seed(9)
    noise = {"z": torch.empty(1, 192, 2, 3).uniform_(-0.5, 0.5, generator=gen), "y_hat": torch.empty(1, 192, 8, 12).uniform_(-0.5, 0.5, generator=gen),
             "y": torch.empty(1, 192, 8, 12).uniform_(-0.5, 0.5, generator=gen)}
    crit = mmcodec.RateDistortionLoss(3)
    # oracle: fp32 autograd over the restated reference ops, same noise
    sd_ref = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("mask",)) else v) for k, v in sd_d.items()}
    out_ref = tp.mm_d_forward(sd_ref, depth, ref_r["hidden"], noise=noise)
    loss_ref = crit(out_ref, depth)
    loss_ref["loss"].backward()
    # ours: guide maps from the reference (so both sides see identical inputs), training mode, same noise
    net_d.train()
    net_d._noise_override = noise
    mm.ESA.train_on_kernels = esa_on_kernels
    hidden = {k: v.to(dev()) for k, v in ref_r["hidden"].items()}
    out = net_d(depth.to(dev()), hidden)
    loss = crit(out, depth.to(dev()))
    loss["loss"].backward()
    loss = {k: v.detach() for k, v in loss.items()}
    loss_ref = {k: v.detach() for k, v in loss_ref.items()}
    assert abs(float(loss["bpp_loss"]) - float(loss_ref["bpp_loss"])) / float(loss_ref["bpp_loss"]) < 0.02
    assert abs(float(loss["loss"]) - float(loss_ref["loss"])) / float(loss_ref["loss"]) < 0.03
    checked, stats = 0, []
    for name, p in net_d.named_parameters():
        ref_g = sd_ref[name].grad if name in sd_ref else None
        if ref_g is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0 or name.endswith("quantiles"), name   # unused inherited transforms
            continue
        assert p.grad is not None, name
        if float(ref_g.norm()) < 1e-12:
            continue
        c = cos(p.grad, ref_g)
        ratio = float(p.grad.double().norm().cpu() / ref_g.double().norm())
        stats.append((c, ratio, name))
        checked += 1
    print("worst gradient directions:", sorted(stats)[:8])
    print("gradient norm ratios: min", min(s_[1] for s_ in stats), "max", max(s_[1] for s_ in stats))
    # bf16 activations / gradients through up to ~25 layers (and the bf16 forward feeding every activation mask):
    # direction and scale of EVERY parameter gradient; the median agreement is far tighter
    assert all(c > 0.93 and 0.85 < r < 1.15 for c, r, _ in stats), sorted(stats)[:5]
    assert sorted(s_[0] for s_ in stats)[len(stats) // 2] > 0.99
    assert checked > 100
    # one full optimisation step through the public training API changes the parameters and keeps everything finite
    step = mmcodec.TrainStep(net_d, net_r.eval(), quality=3)
    before = net_d.tran_conv1.weight.detach().clone()
    res = step(depth.to(dev()), x.to(dev()))
    mm.ESA.train_on_kernels = False
    assert all(bool(torch.isfinite(v).all()) for v in res.values())
    assert float((net_d.tran_conv1.weight.detach() - before).abs().max()) > 0


@pytest.mark.parametrize("arch,cls,N,M", [("factorized", mmcodec.FactorizedPrior, 128, 192), ("hyperprior", mmcodec.ScaleHyperprior, 128, 192),
                                          ("mean-scale", mmcodec.MeanScaleHyperprior, 128, 192)])
def test_zoo_models_training_step_vs_oracle(arch, cls, N, M):
    """CompressionModel.forward in training mode + RateDistortionLoss backward for the three zoo families vs fp32 autograd over
    the restated reference ops with the same noise; then one TrainStep (Adam + aux Adam) through the public training API."""
    from weights import make_image, make_state_dict
    sd = {k: torch.from_numpy(v) for k, v in make_state_dict(arch, N, M, seed=0).items()}
    net = cls(N, M)
    net.update()
    net.load_state_dict({**net.state_dict(), **sd})
    net = net.to(dev()).train()
    x = torch.from_numpy(make_image(2, 64, 128, seed=5))
    gen = torch.Generator().manual_seed(3)
    yshape, zshape = (2, M, 4, 8), (2, N, 1, 2)
    noise = {"z": torch.empty(yshape if arch == "factorized" else zshape).uniform_(-0.5, 0.5, generator=gen),
             "y": torch.empty(yshape).uniform_(-0.5, 0.5, generator=gen)}
    crit = mmcodec.RateDistortionLoss(3)
    sd_ref = {k: v.clone().requires_grad_(True) if v.is_floating_point() else v for k, v in sd.items()}
    torch.set_num_threads(8)
    out_ref = tp.FORWARD[arch](sd_ref, x, noise["z"]) if arch == "factorized" else tp.FORWARD[arch](sd_ref, x, noise)
    loss_ref = crit(out_ref, x)
    loss_ref["loss"].backward()
    net._noise_override = noise
    out = net(x.to(dev()))
    loss = crit(out, x.to(dev()))
    loss["loss"].backward()
    assert abs(float(loss["loss"].detach()) - float(loss_ref["loss"].detach())) / float(loss_ref["loss"].detach()) < 0.03
    stats = []
    for name, p in net.named_parameters():
        ref_g = sd_ref[name].grad
        if ref_g is None or float(ref_g.norm()) < 1e-12:
            continue
        assert p.grad is not None, name
        stats.append((cos(p.grad, ref_g), float(p.grad.double().norm().cpu() / ref_g.double().norm()), name))
    print(arch, "worst:", sorted(stats)[:4])
    assert all(c > 0.93 and 0.85 < r < 1.15 for c, r, _ in stats), sorted(stats)[:5]
    assert len(stats) > 30
    net._noise_override = None
    res = mmcodec.TrainStep(net, None, quality=3)(x.to(dev()))
    assert all(bool(torch.isfinite(v).all()) for v in res.values())


@pytest.mark.parametrize("family", ["mean-scale", "master"])
def test_graphed_train_step_matches_eager(family):
    """mmcodec.GraphedTrainStep (whole optimisation step as one CUDA graph, parameter-dependent caches rebuilt inside the graph)
    follows the eager TrainStep: same initial weights, same batches, same noise draws -> same loss trajectory and parameters
    (up to the fp32 summation order of the split-K weight-gradient atomics), and the loss goes down."""
    import copy
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(21)
    if family == "mean-scale":
        net_a = mmcodec.MeanScaleHyperprior(128, 192)
        guide = None
        xs = [torch.rand(2, 3, 128, 192, generator=gen).to(dev()) for _ in range(3)]
        gs = [None] * 3
        noise = {"y": torch.rand(2, 192, 8, 12, generator=gen) - 0.5, "z": torch.rand(2, 128, 2, 3, generator=gen) - 0.5}
        watch = ["g_a.0.weight", "g_s.6.weight", "h_s.2.bias", "g_a.1.gamma"]
    else:
        net_a = mmcodec.Master_compresser(width=64, height=128, channel=3)
        guide = mmcodec.Guided_compresser(channel=1).eval()
        guide.update()
        guide.to(dev())
        xs = [torch.rand(1, 3, 128, 256, generator=gen).to(dev()) for _ in range(3)]
        gs = [torch.rand(1, 1, 64, 128, generator=gen).to(dev()) for _ in range(3)]
        noise = {"z": torch.rand(1, 192, 1, 2, generator=gen) - 0.5, "y_hat": torch.rand(1, 192, 4, 8, generator=gen) - 0.5,
                 "y": torch.rand(1, 192, 4, 8, generator=gen) - 0.5}
        watch = ["fencoder1.conv1.weight", "decoder.sp_aligner2.blocks.1.attn.qkv2.weight", "ch_aligner.conv3.weight", "g_a.1.gamma"]
    net_a.update()
    net_a.to(dev())
    net_b = copy.deepcopy(net_a)
    noise = {k: v.to(dev()) for k, v in noise.items()}     # device-resident: a host-to-device copy cannot be captured
    for n in (net_a, net_b):
        n._noise_override = noise
    eager = mmcodec.TrainStep(net_a, guide, quality=3)
    graphed = mmcodec.GraphedTrainStep(net_b, guide, warmup=2, quality=3)
    la, lb = [], []
    for i in range(9):                                  # calls 1-2 eager warm-up, call 3 captures, calls 4-9 replay
        x, g_ = xs[i % 3], gs[i % 3]
        la.append(float(eager(x, g_)["loss"]))
        lb.append(float(graphed(x, g_)["loss"]))
    assert graphed.graph is not None and graphed.calls == 9
    assert all(math.isfinite(v) for v in la + lb)
    assert all(abs(a - b) / abs(a) < 2e-2 for a, b in zip(la, lb)), (la, lb)
    assert lb[-1] < lb[0] and lb[-2] < lb[1] and lb[-3] < lb[2]     # same batch three rounds later: the steps did optimise
    pa, pb = dict(net_a.named_parameters()), dict(net_b.named_parameters())
    for name in watch:
        d = float((pa[name].detach() - pb[name].detach()).abs().max())
        step_size = 9 * 1e-4                               # Adam moves a weight by at most ~lr per step
        assert d < 0.25 * step_size, (name, d)
    # eager inference after graphed training sees the CURRENT weights (version counters were bumped after every replay)
    net_b.eval()
    net_a.eval()
    with torch.no_grad():
        oa = net_a(xs[0], gs[0], guide(gs[0])["hidden"]) if guide is not None else net_a(xs[0])
        ob = net_b(xs[0], gs[0], guide(gs[0])["hidden"]) if guide is not None else net_b(xs[0])
    assert float((oa["x_hat"].float() - ob["x_hat"].float()).abs().max()) < 0.05 * float(oa["x_hat"].float().abs().max()) + 1e-3

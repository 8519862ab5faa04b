"""GPU parity tests of the RGB-T reproduction's master codec (Master_compresser, compressai/models/master.py:837-951) against
the reference's own run (tests/golden/models_master.npz) and, stage by stage, against the CPU oracle
(oracle/torch_port.py: master_forward and its parts) fed with the reference's inputs to that stage.  Tolerances follow
test_gpu_models_mm.py: bf16 activations -> rel-RMS <= 1e-2 per conv stage, 2e-2 for the deep chains (feature codecs, the
decoder with its three attention stages); bpp within 5 % end to end (quantiser flips included)."""
import json
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from oracle import torch_port as tp
from weights import make_master_state_dict

pytestmark = pytest.mark.gpu

import mmcodec  # noqa: E402
from mmcodec import models_master as mst  # noqa: E402
from mmcodec import ops  # noqa: E402


def dev():
    return torch.device("cuda", 0)


def rel_rms(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(torch.sqrt(((a - b) ** 2).mean() / (b ** 2).mean().clamp_min(1e-30)))


def bf(a):
    return torch.from_numpy(a).view(torch.bfloat16).float()


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_master.npz"))


@pytest.fixture(scope="module")
def setup(g):
    shapes = {k: tuple(v[0]) for k, v in json.loads(str(g["state_dict"])).items()}
    sd = {k: torch.from_numpy(v) for k, v in make_master_state_dict(shapes, 4).items()}
    net = mmcodec.Master_compresser(width=64, height=128, channel=3).eval()
    net.update()
    net.load_state_dict({**net.state_dict(), **sd})
    net.update(force=True)
    net = net.to(dev())
    sd["context_prediction.mask"] = tp.masked_conv_mask(shapes["context_prediction.weight"], "A")
    hidden = {k: bf(g[f"hidden_{k}_bf16"]) for k in ("gs1", "gs2", "gs3")}
    g_hat = bf(g["g_hat_bf16"])
    torch.set_num_threads(8)
    with torch.no_grad():
        ref = tp.master_forward(sd, torch.from_numpy(g["x"]), g_hat, hidden)
    return net, sd, ref, g_hat, hidden


def bpp_of(liks, npix):
    return sum(float(torch.log(l.double()).sum()) for l in liks.values()) / (-math.log(2) * npix)


# ---- token-side kernels ----------------------------------------------------------------------------------------------------
def test_layernorm_and_gelu_kernels():
    gen = torch.Generator().manual_seed(0)
    x = (torch.randn(3, 10, 12, 96, generator=gen) * 2 + 0.5).to(dev()).bfloat16()
    dl = torch.randn(3, 10, 12, 96, generator=gen).to(dev()).bfloat16()
    w = (1 + 0.1 * torch.randn(96, generator=gen)).to(dev())
    b = (0.1 * torch.randn(96, generator=gen)).to(dev())
    y = ops.layernorm_bf16(x, w, b, 1e-5)
    ref = F.layer_norm(x.float(), (96,), w, b, 1e-5)
    assert float((y.float() - ref).abs().max()) <= 2 ** -7 * float(ref.abs().max())      # one bf16 rounding of the result
    s, y2 = ops.layernorm_bf16(x, w, b, 1e-5, delta=dl, want_sum=True)
    assert torch.equal(s, x + dl)                                                           # the fused residual add is the bf16 add
    ref2 = F.layer_norm(s.float(), (96,), w, b, 1e-5)
    assert float((y2.float() - ref2).abs().max()) <= 2 ** -7 * float(ref2.abs().max())
    for C in (32, 100, 256):                                                                # ragged channel counts
        xc = torch.randn(7, C, generator=gen).to(dev()).bfloat16()
        wc, bc = torch.rand(C, generator=gen).to(dev()), torch.rand(C, generator=gen).to(dev())
        rc = F.layer_norm(xc.float(), (C,), wc, bc, 1e-5)
        assert float((ops.layernorm_bf16(xc, wc, bc).float() - rc).abs().max()) <= 2 ** -7 * float(rc.abs().max())
    ge = ops.gelu_bf16(x)
    assert float((ge.float() - F.gelu(x.float())).abs().max()) <= 2 ** -7 * float(x.float().abs().max())
    with pytest.raises((ValueError, NotImplementedError, mmcodec.MmcodecError)):
        ops.layernorm_bf16(torch.zeros(2, 512, device=dev(), dtype=torch.bfloat16), torch.ones(512, device=dev()), torch.zeros(512, device=dev()))


def test_channel_mean_and_affine_kernels():
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(3, 20, 28, 64, generator=gen).to(dev())
    m = ops.channel_mean(x)
    assert float((m - x.double().mean(dim=(1, 2)).float()).abs().max()) < 1e-6
    assert torch.equal(ops.channel_mean(x[1:2]), m[1:2])                      # a sample's mean does not depend on its batch
    f = torch.randn(3, 20, 28, 64, generator=gen).to(dev()).bfloat16()
    gm, bt = torch.randn(3, 64, generator=gen).to(dev()), torch.randn(3, 64, generator=gen).to(dev())
    ref = (gm[:, None, None, :] * f.float() + bt[:, None, None, :])
    out = ops.channel_affine_bf16(f, gm, bt)
    assert float((out.float() - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max())
    fr = torch.randn(2, 5, 7, 12, generator=gen).to(dev()).bfloat16()         # C % 8 != 0: scalar path
    gr, br = torch.randn(2, 12, generator=gen).to(dev()), torch.randn(2, 12, generator=gen).to(dev())
    rr = gr[:, None, None, :] * fr.float() + br[:, None, None, :]
    assert float((ops.channel_affine_bf16(fr, gr, br).float() - rr).abs().max()) <= 2 ** -8 * float(rr.abs().max())
    xr = torch.randn(2, 5, 7, 40, generator=gen).to(dev())                    # ragged channel count
    assert float((ops.channel_mean(xr) - xr.mean(dim=(1, 2))).abs().max()) < 1e-6


@pytest.mark.parametrize("H,W,ws,shift", [(8, 16, 4, 0), (8, 16, 4, 2), (16, 8, 4, 2), (4, 8, 4, 0), (2, 6, 2, 0), (2, 6, 2, 1), (12, 12, 4, 3)])
def test_window_attention_kernel_vs_fp32_reference(H, W, ws, shift):
    """mmc_window_attention against an fp32 restatement of roll -> partition -> attention -> reverse -> roll back
    (the oracle's own helper functions) on identical bf16 inputs."""
    gen = torch.Generator().manual_seed(H * 100 + W + shift)
    B, heads, C = 2, 3, 96
    q = torch.randn(B, H, W, C, generator=gen).bfloat16()
    kv = torch.randn(B, H, W, 2 * C, generator=gen).bfloat16()
    table = 0.5 * torch.randn((2 * ws - 1) ** 2, heads, generator=gen)
    scale = (C // heads) ** -0.5
    out = ops.window_attention(q.to(dev()), kv.to(dev()), table.to(dev()), ws, shift, heads, scale)
    # fp32 reference
    qf, kf, vf = q.float(), kv.float()[..., :C], kv.float()[..., C:]
    if shift:
        qf, kf, vf = (torch.roll(t, (-shift, -shift), (1, 2)) for t in (qf, kf, vf))
    N = ws * ws
    split = lambda t: tp._to_windows(t, ws).view(-1, N, heads, C // heads).transpose(1, 2)
    att = (split(qf) * scale) @ split(kf).transpose(-2, -1)
    att = att + table[tp.relative_position_index(ws).view(-1)].view(N, N, heads).permute(2, 0, 1)
    if shift:
        att = (att.view(B, -1, heads, N, N) + tp.shift_attention_mask(H, W, ws, shift)[None, :, None]).view(-1, heads, N, N)
    o = (att.softmax(-1) @ split(vf)).transpose(1, 2).reshape(-1, N, C)
    o = tp._from_windows(o, ws, B, H, W)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    assert float((out.float().cpu() - o).abs().max()) <= 2 ** -7 * float(o.abs().max()) + 1e-3
    with pytest.raises((ValueError, NotImplementedError, mmcodec.MmcodecError)):
        ops.window_attention(q.to(dev()), kv.to(dev()), torch.zeros(81, 3, device=dev()), 5, 0, heads, scale)


# ---- stages vs the oracle on the reference's tensors -----------------------------------------------------------------------
def test_conv3x3_mean_by_linearity_vs_convolution(setup):
    """mmc_conv3x3_mean (mean over positions of a padding-1 3x3 convolution from border-corrected channel sums) against the
    convolution itself in fp64, incl. degenerate maps (one row / one column / one pixel) and batch independence; then the two
    Channel_aligner head evaluations (by linearity / by running conv5, conv6 and averaging) against each other."""
    gen = torch.Generator().manual_seed(31)
    for (B, H, W, C, O) in ((3, 20, 28, 64, 24), (2, 1, 5, 16, 8), (2, 4, 1, 8, 3), (1, 1, 1, 8, 5), (2, 33, 17, 256, 64)):
        t = torch.randn(B, H, W, C, generator=gen).bfloat16()
        w = torch.randn(O, C, 3, 3, generator=gen) / (3 * C ** 0.5)
        b = torch.randn(O, generator=gen)
        ref = F.conv2d(t.double().permute(0, 3, 1, 2), w.double(), b.double(), padding=1).mean(dim=(2, 3))
        out = ops.conv3x3_mean(t.to(dev()), w.to(dev()), b.to(dev()))
        assert float((out.double().cpu() - ref).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max())), (B, H, W, C, O)
        assert torch.equal(ops.conv3x3_mean(t[B - 1:].to(dev()), w.to(dev()), b.to(dev())), out[B - 1:])
        nob = ops.conv3x3_mean(t.to(dev()), w.to(dev()), None)
        assert float((nob.double().cpu() - (ref - b.double())).abs().max()) < 1e-5 * max(1.0, float(ref.abs().max()))
    net, _, ref_out, _, _ = setup
    al = net.ch_aligner
    gf = torch.randn(2, 64, 24, 40, generator=gen).to(dev())
    xf = torch.randn(2, 64, 24, 40, generator=gen).to(dev())
    with torch.no_grad():
        a1, b1, g1 = al(xf, gf)
        al.heads_by_linearity = False
        try:
            a2, b2, g2 = al(xf, gf)
        finally:
            al.heads_by_linearity = True
    assert rel_rms(b1, b2) < 1e-2 and rel_rms(g1, g2) < 1e-2 and rel_rms(a1.float(), a2.float()) < 1e-2


def test_state_dict_keys_and_shapes(g):
    ref = {k: tuple(v[0]) for k, v in json.loads(str(g["state_dict"])).items()}
    net = mmcodec.Master_compresser(width=64, height=128, channel=3)
    mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert set(mine) == set(ref)
    data_dependent = ("_offset", "_quantized_cdf", "_cdf_length", "scale_table")
    assert all(mine[k] == ref[k] for k in ref if not k.endswith(data_dependent))


def test_feature_encoders_and_channel_aligner(g, setup):
    net, sd, ref, g_hat, _ = setup
    x = torch.from_numpy(g["x"]).to(dev())
    with torch.no_grad():
        xf = net.fencoder1(x)
        gf = net.fencoder2(g_hat.to(dev()))
        assert rel_rms(xf.float(), ref["x_feature"]) < 2e-2
        gf_ref = tp.feature_encoder(sd, "fencoder2", g_hat, 1)
        assert rel_rms(gf.float(), gf_ref) < 2e-2
        aligned, beta, gamma = net.ch_aligner(ref["x_feature"].to(dev()), gf_ref.to(dev()))
    assert tuple(beta.shape) == tuple(ref["beta"].shape) == (1, 64, 1, 1)
    assert rel_rms(beta, ref["beta"]) < 1e-2 and rel_rms(gamma, ref["gamma"]) < 1e-2
    assert rel_rms(aligned.float(), ref["guided_align"]) < 1e-2


@pytest.mark.parametrize("on_kernels", [True, False])
def test_spatial_aligner_vs_oracle(setup, on_kernels, monkeypatch):
    """One Spatial_aligner (patch embedding, plain + shifted cross-attention block, token reinterpretation, recovery deconv) on
    random maps of the stage-2 size, on the libmmcodec kernels and on the torch-op path."""
    net, sd, _, _, hidden = setup
    monkeypatch.setattr(mst, "attention_on_kernels", on_kernels)
    gen = torch.Generator().manual_seed(11)
    own = torch.randn(2, 192, 16, 32, generator=gen).bfloat16().float()
    guide = torch.cat((hidden["gs2"], hidden["gs2"].flip(3)), 0)
    with torch.no_grad():
        ref = tp.spatial_aligner(sd, "decoder.sp_aligner2", own, guide)
        out = net.decoder.sp_aligner2(own.to(dev()), guide.to(dev()))
    assert tuple(out.shape) == tuple(ref.shape)
    assert rel_rms(out.float(), ref) < 1.5e-2


def test_attention_paths_agree(setup, monkeypatch):
    """Kernel path vs torch-op path of the same block on the same bf16 tokens (both bf16 pipelines)."""
    net = setup[0]
    blk = net.decoder.sp_aligner3.blocks[1]
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(1, 16, 32, 96, generator=gen).to(dev()).bfloat16()
    gd = torch.randn(1, 16, 32, 96, generator=gen).to(dev()).bfloat16()
    with torch.no_grad():
        a = blk.forward_grid(x, gd)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            b = blk(x.reshape(1, -1, 96), gd.reshape(1, -1, 96)).reshape(1, 16, 32, 96)
    assert rel_rms(a.float(), b.float()) < 1e-2


def test_master_decoder_and_feature_decoder(setup):
    net, sd, ref, _, hidden = setup
    with torch.no_grad():
        fh = net.decoder(ref["y_hat"].to(dev()), {k: v.to(dev()) for k, v in hidden.items()})["x_feature_hat"]
        assert tuple(fh.shape) == tuple(ref["x_feature_hat"].shape)
        assert rel_rms(fh.float(), ref["x_feature_hat"]) < 2e-2
        x_hat = net.fdecoder(torch.cat((ref["x_feature_hat"], ref["guided_align"]), 1).to(dev()))
    assert tuple(x_hat.shape) == tuple(ref["x_hat"].shape)
    assert float((x_hat.float().cpu() - ref["x_hat"]).abs().max()) < 2e-2 * float(ref["x_hat"].abs().max())


def test_forward_vs_reference_golden(g, setup):
    net, _, ref, g_hat, hidden = setup
    x = torch.from_numpy(g["x"]).to(dev())
    with torch.no_grad():
        o = net(x, g_hat.to(dev()), {k: v.to(dev()) for k, v in hidden.items()})
    assert set(o) == {"x_hat", "likelihoods"} and tuple(o["x_hat"].shape) == g["x_hat"].shape
    assert tuple(o["likelihoods"]["y"].shape) == g["lik_y"].shape and tuple(o["likelihoods"]["z"].shape) == g["lik_z"].shape
    npix = x.shape[0] * x.shape[2] * x.shape[3]
    ref_bpp = sum(oracle.bits(g[f"lik_{k}"]) for k in o["likelihoods"]) / npix
    assert abs(bpp_of(o["likelihoods"], npix) - ref_bpp) / ref_bpp < 0.05
    assert rel_rms(o["x_hat"].float(), torch.from_numpy(g["x_hat"])) < 0.15


def test_graphed_forward_matches_eager(g, setup):
    net, _, _, g_hat, hidden = setup
    x = torch.from_numpy(g["x"]).to(dev())
    args = (x, g_hat.to(dev()), {k: v.to(dev()) for k, v in hidden.items()})
    with torch.no_grad():
        eager = net(*args)
        graphed = mmcodec.GraphedForward(net, *args)
        o = graphed(*args)
        o = graphed(*args)
    assert torch.equal(o["x_hat"], eager["x_hat"]) and torch.equal(o["likelihoods"]["y"], eager["likelihoods"]["y"])


def test_one_channel_master_variant():
    """channel=1: 1-channel master at stride 1, 3-channel guide at stride 2, guide maps downsampled by the extra convs
    (master.py:840-850, 765-768, 783-786)."""
    torch.manual_seed(0)
    net = mmcodec.Master_compresser(width=64, height=64, channel=1).eval()
    net.update()
    net = net.to(dev())
    gen = torch.Generator().manual_seed(1)
    x = torch.rand(1, 1, 64, 64, generator=gen).to(dev())
    g_hat = torch.rand(1, 3, 128, 128, generator=gen).to(dev())
    hidden = {"gs1": torch.randn(1, 192, 16, 16, generator=gen).to(dev()), "gs2": torch.randn(1, 192, 32, 32, generator=gen).to(dev()),
              "gs3": torch.randn(1, 192, 64, 64, generator=gen).to(dev())}
    sd = {k: v.detach().cpu().float() for k, v in net.state_dict().items()}
    with torch.no_grad():
        o = net(x, g_hat, hidden)
        ref = tp.master_forward(sd, x.cpu(), g_hat.cpu(), {k: v.cpu() for k, v in hidden.items()})
    assert tuple(o["x_hat"].shape) == (1, 1, 64, 64) == tuple(ref["x_hat"].shape)
    npix = 64 * 64
    assert abs(bpp_of(o["likelihoods"], npix) - bpp_of(ref["likelihoods"], npix)) / bpp_of(ref["likelihoods"], npix) < 0.05
    assert rel_rms(o["x_hat"].float(), ref["x_hat"]) < 0.1


def test_one_channel_master_vs_reference_golden():
    """Master_compresser(channel=1) against the reference's own run (tests/golden/models_master1.npz, seeded inputs)."""
    from weights import make_master1_inputs
    g1 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_master1.npz"))
    shapes = {k: tuple(v[0]) for k, v in json.loads(str(g1["state_dict"])).items()}
    sd = {k: torch.from_numpy(v) for k, v in make_master_state_dict(shapes, 5).items()}
    net = mmcodec.Master_compresser(width=64, height=64, channel=1).eval()
    assert set(net.state_dict()) == set(shapes)
    net.update()
    net.load_state_dict({**net.state_dict(), **sd})
    net.update(force=True)
    net = net.to(dev())
    x, g_hat, hidden = make_master1_inputs(10)
    with torch.no_grad():
        o = net(torch.from_numpy(x).to(dev()), torch.from_numpy(g_hat).to(dev()), {k: torch.from_numpy(v).to(dev()) for k, v in hidden.items()})
    assert tuple(o["x_hat"].shape) == g1["x_hat"].shape
    npix = 64 * 64
    ref_bpp = sum(oracle.bits(g1[f"lik_{k}"]) for k in o["likelihoods"]) / npix
    assert abs(bpp_of(o["likelihoods"], npix) - ref_bpp) / ref_bpp < 0.05
    assert rel_rms(o["x_hat"].float(), torch.from_numpy(g1["x_hat"])) < 0.15


def test_training_step_gradients(g, setup):
    """Training-mode forward + backward (examples/train.py:208-274: master loss = lambda * MSE + bpp): every parameter
    on the forward path receives a finite gradient; the unused inherited ``g_s`` does not; the
    rate-distortion loss and the main gradients match the oracle's autograd on the same noise draws."""
    net, sd, _, g_hat, hidden = setup
    net.train()
    try:
        x = torch.from_numpy(g["x"]).to(dev())
        gen = torch.Generator().manual_seed(9)
        noise = {"z": torch.rand(1, 192, 1, 2, generator=gen) - 0.5, "y_hat": torch.rand(1, 192, 4, 8, generator=gen) - 0.5,
                 "y": torch.rand(1, 192, 4, 8, generator=gen) - 0.5}
        net._noise_override = noise
        net.zero_grad(set_to_none=True)
        o = net(x, g_hat.to(dev()), {k: v.to(dev()) for k, v in hidden.items()})
        crit = mmcodec.RateDistortionLoss(0)                       # lambda = 256
        loss = crit(o, x)
        loss["loss"].backward()
        # oracle autograd
        sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("mask", "bound", "pedestal")) else v) for k, v in sd.items()}
        ro = tp.master_forward(sdg, x.cpu(), g_hat, hidden, noise=noise)
        npix = x.shape[2] * x.shape[3]
        r_bpp = sum(torch.log(l).sum() for l in ro["likelihoods"].values()) / (-math.log(2) * npix)
        r_loss = 256 * F.mse_loss(ro["x_hat"], x.cpu()) + r_bpp
        r_loss.backward()
        assert abs(float(loss["loss"].detach()) - float(r_loss.detach())) / float(r_loss.detach()) < 2e-2
        missing, checked = [], 0
        for name, p in net.named_parameters():
            if name.startswith("g_s.") or name.endswith("quantiles"):
                assert p.grad is None, name
                continue
            if p.grad is None:
                missing.append(name)
                continue
            assert torch.isfinite(p.grad).all(), name
            rg = sdg[name].grad
            if p.dim() >= 2 and rg is not None and float(rg.abs().max()) > 0:
                cos = float(F.cosine_similarity(p.grad.flatten().double().cpu(), rg.flatten().double(), dim=0))
                assert cos > 0.9, (name, cos)
                checked += 1
        assert not missing, missing
        assert checked > 60
    finally:
        net._noise_override = None
        net.eval()
        net.zero_grad(set_to_none=True)


def test_train_step_api_with_guide_codec():
    """mmcodec.TrainStep on the (master, guide) pair: net(x, guided, hidden) with hidden from the frozen guide codec
    (examples/train.py:208-233); one step changes the attention and conv parameters and keeps everything finite."""
    torch.manual_seed(0)
    guide = mmcodec.Guided_compresser(channel=1).eval()
    master = mmcodec.Master_compresser(width=64, height=128, channel=3)
    for n in (guide, master):
        n.update()
        n.to(dev())
    gen = torch.Generator().manual_seed(4)
    x = torch.rand(2, 3, 128, 256, generator=gen).to(dev())
    t = torch.rand(2, 1, 64, 128, generator=gen).to(dev())
    step = mmcodec.TrainStep(master, guide, quality=3)
    watch = {n: p.detach().clone() for n, p in master.named_parameters()
             if n in ("decoder.sp_aligner2.blocks.1.attn.qkv2.weight", "ch_aligner.conv3.weight", "fdecoder.deconv1.weight",
                      "decoder.sp_aligner1.blocks.0.attn.relative_position_bias_table", "entropy_bottleneck.quantiles")}
    assert len(watch) == 5
    res = step(x, t)
    assert all(bool(torch.isfinite(v).all()) for v in res.values())
    params = dict(master.named_parameters())
    for n, before in watch.items():
        assert float((params[n].detach() - before).abs().max()) > 0, n
    assert all(bool(torch.isfinite(p).all()) for p in master.parameters())


def test_full_size_properties_768x512():
    """BASELINE-size pair (3 x 512 x 768 master, 1 x 256 x 384 guide through Guided_compresser): shapes, finite outputs,
    likelihood ranges, batch independence."""
    torch.manual_seed(0)
    guide = mmcodec.Guided_compresser(channel=1).eval()
    master = mmcodec.Master_compresser(width=256, height=384, channel=3).eval()
    for n in (guide, master):
        n.update()
        n.to(dev())
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 512, 768, generator=gen).to(dev())
    t = torch.rand(2, 1, 256, 384, generator=gen).to(dev())
    with torch.no_grad():
        og = guide(t)
        o = master(x, og["x_hat"], og["hidden"])
        og0 = guide(t[:1])
        o0 = master(x[:1], og0["x_hat"], og0["hidden"])
    assert tuple(o["x_hat"].shape) == (2, 3, 512, 768)
    assert tuple(o["likelihoods"]["y"].shape) == (2, 192, 16, 24) and tuple(o["likelihoods"]["z"].shape) == (2, 192, 4, 6)
    assert torch.isfinite(o["x_hat"]).all()
    for l in o["likelihoods"].values():
        assert float(l.min()) >= 1e-9 and float(l.max()) <= 1.0
    assert torch.equal(o["x_hat"][:1], o0["x_hat"]) and torch.equal(o["likelihoods"]["y"][:1], o0["likelihoods"]["y"])

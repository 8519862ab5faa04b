"""GPU parity tests of the two-branch RGB + depth codec (JointAutoregressiveHierarchicalPriors_R / _D,
compressai/models/google.py:746-1248) against the reference's own run (tests/golden/models_mm.npz) and, stage by stage,
against the CPU oracle (oracle/torch_port.py: mm_r_forward / mm_d_forward) fed with the reference's inputs to that stage.
Tolerances as in test_gpu_models.py: bf16 transforms -> rel-RMS <= 1e-2 per stage (2e-2 for the fused analysis chain of the
depth branch, which is ten bf16 layers and three attention gates deep), bpp within 0.5 % on identical latents."""
import json
import math

import numpy as np
import pytest
import torch

import oracle
from oracle import torch_port as tp
from weights import make_mm_state_dict

pytestmark = pytest.mark.gpu

import mmcodec  # noqa: E402
from mmcodec import models_mm as mm  # noqa: E402


def dev():
    return torch.device("cuda", 0)


def rel_rms(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float(torch.sqrt(((a - b) ** 2).mean() / (b ** 2).mean().clamp_min(1e-30)))


@pytest.fixture(scope="module")
def g():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_mm.npz"))


def load(g, tag, cls, seed):
    shapes = {k: tuple(v[0]) for k, v in json.loads(str(g[f"{tag}_state_dict"])).items()}
    sd = {k: torch.from_numpy(v) for k, v in make_mm_state_dict(shapes, seed).items()}
    net = cls(192, 192).eval()
    net.update()
    net.load_state_dict({**net.state_dict(), **sd})
    net.update(force=True)   # CDF tables of the entropy bottleneck depend on the loaded parameters
    sd["context_prediction.mask"] = tp.masked_conv_mask(shapes["context_prediction.weight"], "A")
    return net.to(dev()), sd


@pytest.fixture(scope="module")
def setup(g):
    net_r, sd_r = load(g, "r", mm.JointAutoregressiveHierarchicalPriors_R, 0)
    net_d, sd_d = load(g, "d", mm.JointAutoregressiveHierarchicalPriors_D, 1)
    torch.set_num_threads(8)
    with torch.no_grad():
        ref_r = tp.mm_r_forward(sd_r, torch.from_numpy(g["x"]))
        ref_d = tp.mm_d_forward(sd_d, torch.from_numpy(g["depth"]), ref_r["hidden"])
    return net_r, net_d, ref_r, ref_d


def bpp_of(liks, npix):
    return sum(float(torch.log(l.double()).sum()) for l in liks.values()) / (-math.log(2) * npix)


def test_forward_structure_and_bpp_vs_reference(g, setup):
    net_r, net_d, ref_r, ref_d = setup
    x, d = torch.from_numpy(g["x"]).to(dev()), torch.from_numpy(g["depth"]).to(dev())
    with torch.no_grad():
        o_r = net_r(x)
        o_d = net_d(d, o_r["hidden"])
    assert set(o_r) == {"x_hat", "likelihoods", "hidden"} and set(o_d) == {"x_hat", "likelihoods"}
    assert set(o_r["hidden"]) == {"ga1", "ga2", "ga3", "gs1", "gs2", "gs3"}
    npix = x.shape[0] * x.shape[2] * x.shape[3]
    for tag, o in (("r", o_r), ("d", o_d)):
        assert tuple(o["x_hat"].shape) == g[f"{tag}_x_hat"].shape
        for k, lk in o["likelihoods"].items():
            assert tuple(lk.shape) == g[f"{tag}_lik_{k}"].shape
            assert float(lk.min()) >= 1e-9 * 0.999 and float(lk.max()) <= 1 + 1e-6
        ref_bpp = sum(oracle.bits(g[f"{tag}_lik_{k}"]) for k in o["likelihoods"]) / npix
        mine = bpp_of(o["likelihoods"], npix)
        # through the quantiser AND the context model (every flipped symbol moves the predicted mean / scale of its
        # neighbours): bpp agrees to a few per cent end to end; the 0.5 % gate is test_entropy_stage_on_reference_latents
        assert abs(mine - ref_bpp) / ref_bpp < 0.05, (tag, mine, ref_bpp)
        assert rel_rms(o["x_hat"].float(), torch.from_numpy(g[f"{tag}_x_hat"])) < 0.15
    for k, v in o_r["hidden"].items():
        assert v.shape == ref_r["hidden"][k].shape


def test_stagewise_transforms_vs_oracle(g, setup):
    net_r, net_d, ref_r, ref_d = setup
    x, d = torch.from_numpy(g["x"]).to(dev()), torch.from_numpy(g["depth"]).to(dev())
    with torch.no_grad():
        y, ga1, ga2, ga3 = net_r.enc1(x)
        assert rel_rms(y.float(), ref_r["y"]) < 1e-2
        for t, k in ((ga1, "ga1"), (ga2, "ga2"), (ga3, "ga3")):
            assert rel_rms(t.float(), ref_r["hidden"][k]) < 1e-2, k
        x_hat, gs1, gs2, gs3 = net_r.dec1(ref_r["y_hat"].to(dev()))
        assert rel_rms(x_hat.float(), ref_r["x_hat"]) < 1e-2
        for t, k in ((gs1, "gs1"), (gs2, "gs2"), (gs3, "gs3")):
            assert rel_rms(t.float(), ref_r["hidden"][k]) < 1e-2, k
        # depth branch fed with the REFERENCE's hidden maps
        hid = {k: mm._to_nhwc_bf16(v.to(dev())) for k, v in ref_r["hidden"].items()}
        y_d, _ = net_d._analysis(d, hid)
        assert rel_rms(y_d.permute(0, 3, 1, 2).float(), ref_d["y"]) < 2e-2
        xh_d = net_d._synthesis(mm._to_nhwc_bf16(ref_d["y_hat"].to(dev())), hid)
        assert rel_rms(xh_d.float(), ref_d["x_hat"]) < 2e-2
        # one fusion block on its own: eg_ext -> cat -> tran_conv -> ESA (google.py:1151-1156)
        a_ref = tp.gdn(tp_sd(net_d), "pic2_g_a_gdn1", tp.conv(tp_sd(net_d), "pic2_g_a_conv1", torch.from_numpy(g["depth"])))
        f_ref = tp.mm_fuse(tp_sd(net_d), 1, a_ref, ref_r["hidden"]["ga1"])
        f = net_d._fuse(1, mm._to_nhwc_bf16(a_ref.to(dev())), hid["ga1"])
        assert rel_rms(f.permute(0, 3, 1, 2).float(), f_ref) < 1e-2


def tp_sd(net):
    return {k: v.detach().float().cpu() for k, v in net.state_dict().items()}


def test_entropy_stage_on_reference_latents(g, setup):
    """Hyperprior + masked context conv + entropy-parameter convs + Gaussian likelihood on the reference's fp32 y:
    y_hat (what the context model sees) is bit-exact, bpp within 0.5 %."""
    net_r, net_d, ref_r, ref_d = setup
    npix = g["x"].shape[0] * g["x"].shape[2] * g["x"].shape[3]
    for net, ref in ((net_r, ref_r), (net_d, ref_d)):
        y = ref["y"].to(dev()).permute(0, 2, 3, 1).contiguous()
        with torch.no_grad():
            y_hat_bf16, y_lik, z_lik = net._entropy_stage(y, mmcodec.ops.to_bf16(y))
        assert torch.equal(y_hat_bf16.permute(0, 3, 1, 2).float().cpu(), ref["y_hat"])   # integers: exact in bf16
        mine = bpp_of({"y": y_lik, "z": z_lik}, npix)
        want = bpp_of(ref["likelihoods"], npix)
        assert abs(mine - want) / want < 5e-3, (mine, want)


def test_masked_conv_and_errors(setup):
    net_r, net_d, _, _ = setup
    w = net_r.context_prediction.weight.detach()
    net_r.context_prediction.packed_weight  # noqa: B018
    with torch.no_grad():
        net_r.context_prediction(torch.zeros(1, 192, 8, 8, device=dev()))
    assert float(w[:, :, 2, 2:].abs().max()) == 0.0 and float(w[:, :, 3:].abs().max()) == 0.0   # mask type A applied in place
    with pytest.raises(NotImplementedError):
        net_r.compress(torch.zeros(1, 3, 64, 64, device=dev()))
    with pytest.raises(RuntimeError):
        net_d(torch.zeros(1, 1, 64, 64), {})


def test_guided_compresser_vs_reference_golden():
    """Guided_compresser (compressai/models/master.py:1215-1300): 1-channel guide codec of the RGB-T reproduction."""
    import os
    gg = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_guided.npz"))
    shapes = {k: tuple(v[0]) for k, v in json.loads(str(gg["state_dict"])).items()}
    sd = {k: torch.from_numpy(v) for k, v in make_mm_state_dict(shapes, 2).items()}
    net = mmcodec.Guided_compresser(channel=1).eval()
    net.update()
    net.load_state_dict({**net.state_dict(), **sd})
    net.update(force=True)
    net = net.to(dev())
    x = torch.from_numpy(gg["x"]).to(dev())
    with torch.no_grad():
        o = net(x)
    assert set(o) == {"x_hat", "likelihoods", "hidden"} and tuple(o["x_hat"].shape) == gg["x_hat"].shape
    npix = x.shape[0] * x.shape[2] * x.shape[3]
    ref_bpp = sum(oracle.bits(gg[f"lik_{k}"]) for k in o["likelihoods"]) / npix
    assert abs(bpp_of(o["likelihoods"], npix) - ref_bpp) / ref_bpp < 0.05
    for k in ("ga1", "ga2", "ga3"):      # encoder-side hidden maps do not pass through a quantiser: stage-wise tolerance
        assert rel_rms(o["hidden"][k].float()[:, ::8, ::2, ::2], torch.from_numpy(gg[f"hidden_{k}"])) < 1e-2, k
    assert rel_rms(o["x_hat"].float(), torch.from_numpy(gg["x_hat"])) < 0.15


def test_full_size_properties_768x512():
    """BASELINE size (768x512 RGB + depth pair, random-init weights): shapes, likelihood ranges, finite outputs, batch
    independence (images are independent units: a batch of two equals two batches of one)."""
    torch.manual_seed(0)
    net_r = mm.JointAutoregressiveHierarchicalPriors_R(192, 192).eval()
    net_d = mm.JointAutoregressiveHierarchicalPriors_D(192, 192).eval()
    for n in (net_r, net_d):
        n.update()
        n.to(dev())
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(2, 3, 512, 768, generator=gen).to(dev())
    d = torch.rand(2, 1, 512, 768, generator=gen).to(dev())
    with torch.no_grad():
        o_r = net_r(x)
        o_d = net_d(d, o_r["hidden"])
        o_r0 = net_r(x[:1])
        o_d0 = net_d(d[:1], o_r0["hidden"])
    assert tuple(o_d["x_hat"].shape) == (2, 1, 512, 768) and tuple(o_r["x_hat"].shape) == (2, 3, 512, 768)
    assert tuple(o_r["hidden"]["ga1"].shape) == (2, 192, 256, 384) and tuple(o_r["hidden"]["gs3"].shape) == (2, 192, 256, 384)
    assert tuple(o_d["likelihoods"]["y"].shape) == (2, 192, 32, 48) and tuple(o_d["likelihoods"]["z"].shape) == (2, 192, 8, 12)
    for o in (o_r, o_d):
        assert bool(torch.isfinite(o["x_hat"]).all())
        for lk in o["likelihoods"].values():
            assert float(lk.min()) >= 1e-9 * 0.999 and float(lk.max()) <= 1 + 1e-6
    assert torch.equal(o_d0["x_hat"], o_d["x_hat"][:1]) and torch.equal(o_d0["likelihoods"]["y"], o_d["likelihoods"]["y"][:1])


def test_mbt2018_vs_reference_golden_and_training():
    """The zoo's JointAutoregressiveHierarchicalPriors (mbt2018, google.py:421-520) through build_model: state_dict keys, eval
    forward vs the reference's run (stage-wise against the oracle on the reference's latents), training-mode gradients vs the
    oracle's autograd on the same noise, q >= 5 configuration (M = 320: entropy_parameters widths that are not multiples of 16)."""
    import os
    from weights import make_mbt2018_state_dict
    gg = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_mbt2018.npz"))
    shapes = {k: tuple(v[0]) for k, v in json.loads(str(gg["state_dict"])).items()}
    sd = {k: torch.from_numpy(v) for k, v in make_mbt2018_state_dict(shapes, 3).items()}
    net = mmcodec.build_model("mbt2018", 3).eval()
    assert isinstance(net, mm.JointAutoregressiveHierarchicalPriors) and set(net.state_dict()) == set(shapes)
    net.update()
    net.load_state_dict({**net.state_dict(), **sd})
    net.update(force=True)
    net = net.to(dev())
    sd["context_prediction.mask"] = tp.masked_conv_mask(shapes["context_prediction.weight"], "A")
    x = torch.from_numpy(gg["x"])
    with torch.no_grad():
        o = net(x.to(dev()))
        ref = tp.mbt2018_forward(sd, x)
        x_hat_same_latent = net.g_s(ref["y_hat"].to(dev()))
    npix = x.shape[0] * x.shape[2] * x.shape[3]
    ref_bpp = sum(oracle.bits(gg[f"lik_{k}"]) for k in o["likelihoods"]) / npix
    assert abs(bpp_of(o["likelihoods"], npix) - ref_bpp) / ref_bpp < 0.05
    assert rel_rms(x_hat_same_latent.float(), ref["x_hat"]) < 1e-2
    assert rel_rms(o["x_hat"].float(), torch.from_numpy(gg["x_hat"])) < 0.15
    with pytest.raises(NotImplementedError):
        net.compress(x.to(dev()))
    # training-mode forward / backward on the same noise draws
    gen = torch.Generator().manual_seed(2)
    noise = {"z": torch.rand(1, 192, 2, 3, generator=gen) - 0.5, "y_hat": torch.rand(1, 192, 8, 12, generator=gen) - 0.5,
             "y": torch.rand(1, 192, 8, 12, generator=gen) - 0.5}
    net.train()
    net._noise_override = noise
    crit = mmcodec.RateDistortionLoss(3)
    loss = crit(net(x.to(dev())), x.to(dev()))
    loss["loss"].backward()
    sdg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("mask", "bound", "pedestal")) else v) for k, v in sd.items()}
    ro = tp.mbt2018_forward(sdg, x, noise=noise)
    r_loss = crit(ro, x)
    r_loss["loss"].backward()
    assert abs(float(loss["loss"]) - float(r_loss["loss"])) / float(r_loss["loss"]) < 3e-2
    checked = 0
    for name, p in net.named_parameters():
        rg = sdg[name].grad if name in sdg else None
        if rg is None or p.dim() < 2 or float(rg.norm()) < 1e-12:
            continue
        assert p.grad is not None, name
        c = float(torch.nn.functional.cosine_similarity(p.grad.flatten().double().cpu(), rg.flatten().double(), dim=0))
        assert c > 0.93, (name, c)
        checked += 1
    assert checked >= 20
    net._noise_override = None
    # the high-quality configuration
    big = mmcodec.build_model("mbt2018", 6).eval()
    big.update()
    big = big.to(dev())
    with torch.no_grad():
        ob = big(torch.rand(1, 3, 64, 128, generator=gen).to(dev()))
    assert tuple(ob["likelihoods"]["y"].shape) == (1, 320, 4, 8) and torch.isfinite(ob["x_hat"]).all()

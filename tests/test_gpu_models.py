"""GPU parity tests of the model-level path (CompressionModel.forward / symbols-and-indexes) against
the reference's own outputs (tests/golden/models.npz) and, stage by stage, against the CPU oracle.

Tolerances (BASELINE.json north_star):
  * transforms run in bf16 with fp32 accumulate -> stage-wise rel-RMS <= 1e-2 on the REFERENCE's
    input to that stage (elementwise relative error is meaningless near zero activations);
  * symbols / indexes bit-exact given identical fp32 latents;
  * likelihoods within 1e-4 relative given identical latents; bpp within 0.5 % end to end.
End to end through the quantiser only bpp is comparable (rounding is chaotic by construction)."""
import numpy as np
import pytest
import torch

import oracle
from oracle import torch_port as tp
from weights import make_image, make_state_dict

pytestmark = pytest.mark.gpu

import mmcodec  # noqa: E402
from mmcodec import ops  # noqa: E402
from mmcodec.transforms import run_layers  # noqa: E402

ARCHS = [("factorized", mmcodec.FactorizedPrior, 128, 192),
         ("hyperprior", mmcodec.ScaleHyperprior, 128, 192),
         ("mean-scale", mmcodec.MeanScaleHyperprior, 192, 320)]


def dev():
    return torch.device("cuda", 0)


# gates at 2x the measured values (round 2, B200, profiles/r02_parity_fullsize_measured.jsonl): x_hat rel-RMS through the quantiser
# 0.010 / 0.010 / 0.020; symbol / index disagreement y 2.2 % / 25 %, z 0.8 % / 0 (the round-1 gates 0.1 and 0.70 were far
# from the measured values)
X_HAT_REL_RMS = {"factorized": 0.021, "hyperprior": 0.021, "mean-scale": 0.04}
MIN_AGREE = {"y_symbols": 0.955, "y_indexes": 0.50, "z_symbols": 0.983, "z_indexes": 0.999}


def _record(test, **values):
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_measured.jsonl"), "a") as f:
            f.write(json.dumps({"test": test, **values}) + "\n")


def load(cls, arch, N, M):
    sd = {k: torch.from_numpy(v) for k, v in make_state_dict(arch, N, M, seed=0).items()}
    net = cls(N, M).eval()
    net.update()
    net.load_state_dict({**net.state_dict(), **sd})
    net.update(force=True)   # CDF tables of the entropy bottleneck depend on the loaded parameters
    return net.to(dev()), sd


def rel_rms(a, b):
    a, b = a.double(), b.double()
    return float(torch.sqrt(((a - b) ** 2).mean() / (b ** 2).mean().clamp_min(1e-30)))


@pytest.mark.parametrize("arch,cls,N,M", ARCHS)
def test_forward_vs_reference_golden(models_golden, arch, cls, N, M):
    """Output structure, shapes and bpp of CompressionModel.forward vs the reference run."""
    g = models_golden
    tag = arch.replace("-", "_")
    net, _ = load(cls, arch, N, M)
    x = torch.from_numpy(g["x"]).to(dev())
    with torch.no_grad():
        out = net(x)
    assert set(out) == {"x_hat", "likelihoods"}
    assert tuple(out["x_hat"].shape) == g[f"{tag}_x_hat"].shape
    npix = x.shape[0] * x.shape[2] * x.shape[3]
    ref_bits = 0.0
    for k, lk in out["likelihoods"].items():
        ref = g[f"{tag}_lik_{k}"]
        assert tuple(lk.shape) == ref.shape
        assert float(lk.min()) >= 1e-9 * 0.999 and float(lk.max()) <= 1.0 + 1e-6
        ref_bits += oracle.bits(ref)
    bpp, ref_bpp = net.bpp(out, npix), ref_bits / npix
    assert abs(bpp - ref_bpp) / ref_bpp < 5e-3, (bpp, ref_bpp)
    # reconstruction: bf16 transforms + chaotic rounding -> compare at the PSNR level
    xh, rxh = out["x_hat"].float().cpu(), torch.from_numpy(g[f"{tag}_x_hat"])
    _record("forward_vs_golden", arch=arch, x_hat_rel_rms=rel_rms(xh, rxh), bpp_rel=abs(bpp - ref_bpp) / ref_bpp)
    assert rel_rms(xh, rxh) < X_HAT_REL_RMS[arch]


@pytest.mark.parametrize("arch,cls,N,M", ARCHS)
def test_stagewise_transforms_vs_oracle(arch, cls, N, M):
    """Each transform stack fed with the REFERENCE's input to that stage."""
    net, sd = load(cls, arch, N, M)
    x = torch.from_numpy(make_image(2, 128, 192, seed=7))
    with torch.no_grad():
        ref = tp.FORWARD[arch](sd, x)
        y = net.g_a(x.to(dev()))
        assert tuple(y.shape) == tuple(ref["y"].shape)
        assert rel_rms(y.float().cpu(), ref["y"]) < 1e-2
        x_hat = net.g_s(ref["y_hat"].to(dev()))
        assert rel_rms(x_hat.float().cpu(), ref["x_hat"]) < 1e-2
        if arch == "factorized":
            return
        h_in = torch.abs(ref["y"]) if arch == "hyperprior" else ref["y"]
        z = net.h_a(h_in.to(dev()))
        assert rel_rms(z.float().cpu(), ref["z"]) < 1e-2
        p = net.h_s(ref["z_hat"].to(dev()))
        ref_p = ref["scales_hat"] if arch == "hyperprior" else torch.cat([ref["scales_hat"], ref["means_hat"]], 1)
        assert rel_rms(p.float().cpu(), ref_p) < 1e-2


@pytest.mark.parametrize("arch,cls,N,M", ARCHS[1:])
def test_entropy_stage_bit_exact_on_reference_latents(models_golden, arch, cls, N, M):
    """Symbols and CDF indexes handed to the rANS coder: bit-exact given the reference's fp32 latents
    (checked against what the reference itself passed to encode_with_indexes)."""
    g = models_golden
    tag = arch.replace("-", "_")
    net, sd = load(cls, arch, N, M)
    x = torch.from_numpy(g["x"])
    fn = tp.hyperprior_compress_symbols if arch == "hyperprior" else tp.mean_scale_compress_symbols
    with torch.no_grad():
        ref = fn(sd, x, tp.get_scale_table())
    B = x.shape[0]
    eb, gc = net.entropy_bottleneck, net.gaussian_conditional
    z = ref["z"].to(dev())
    z_sym, z_idx = eb.symbols_and_indexes(z)
    assert np.array_equal(z_sym.cpu().numpy().reshape(B, -1), g[f"{tag}_z_symbols"])
    assert np.array_equal(z_idx.cpu().numpy().reshape(B, -1), g[f"{tag}_z_indexes"])
    y_idx = gc.build_indexes(ref["scales_hat"].to(dev()))
    means = ref["means_hat"].to(dev()) if "means_hat" in ref else None
    y_sym, y_idx = gc.symbols_and_indexes(ref["y"].to(dev()), y_idx, means)
    assert np.array_equal(y_sym.cpu().numpy().reshape(B, -1), g[f"{tag}_y_symbols"])
    assert np.array_equal(y_idx.cpu().numpy().reshape(B, -1), g[f"{tag}_y_indexes"])
    assert np.abs(g[f"{tag}_y_symbols"]).max() > 5 and len(np.unique(g[f"{tag}_y_indexes"])) > 20
    # likelihoods on identical latents: 1e-4 relative
    full = tp.FORWARD[arch](sd, x)
    _, lik = gc(full["y"].to(dev()), full["scales_hat"].to(dev()), means)
    r = full["likelihoods"]["y"]
    assert float(((lik.cpu() - r).abs() / r.clamp_min(1e-9)).max()) < 1e-4
    _, zl = eb(full["z"].to(dev()))
    r = full["likelihoods"]["z"]
    assert float(((zl.cpu() - r).abs() / r.clamp_min(1e-9)).max()) < 1e-4


@pytest.mark.parametrize("arch,cls,N,M", ARCHS[1:])
def test_symbols_and_indexes_end_to_end(models_golden, arch, cls, N, M):
    """model.symbols_and_indexes(x) (compress() up to the coder): shapes/dtypes exact; the z path and
    most of the y path agree with the reference even through the bf16 transforms."""
    g = models_golden
    tag = arch.replace("-", "_")
    net, _ = load(cls, arch, N, M)
    x = torch.from_numpy(g["x"]).to(dev())
    with torch.no_grad():
        c = net.symbols_and_indexes(x)
    B = x.shape[0]
    for name in ("y_symbols", "y_indexes", "z_symbols", "z_indexes"):
        assert c[name].dtype == torch.int32
        mine, ref = c[name].cpu().numpy().reshape(B, -1), g[f"{tag}_{name}"]
        assert mine.shape == ref.shape
        agree = float((mine == ref).mean())
        _record("symbols_end_to_end", arch=arch, name=name, agree=agree)
        # SURVEY.md 8d probe: bf16 transforms flip a few % of y symbols and ~15-30 % of indexes (every error passes through
        # round()); the gates are set at twice the measured disagreement (profiles/r02_parity_fullsize_measured.jsonl)
        assert agree > MIN_AGREE[name], (name, agree)
    assert tuple(c["shape"]) == tuple(g[f"{tag}_shape"])


def test_public_stack_forward_accepts_channels_last_and_validates():
    net, sd = load(mmcodec.ScaleHyperprior, "hyperprior", 128, 192)
    x = torch.from_numpy(make_image(1, 64, 64, seed=3)).to(dev())
    with torch.no_grad():
        y0 = net.g_a(x)
        y1 = net.g_a(x.to(memory_format=torch.channels_last))
    assert rel_rms(y1.float().cpu(), y0.float().cpu()) < 1e-2
    with pytest.raises(ValueError):
        net.g_a(torch.rand(1, 5, 64, 64, device=dev()))
    with pytest.raises(ValueError):
        net.gaussian_conditional(torch.rand(1, 4, 2, 2, device=dev()), torch.rand(1, 4, 2, 3, device=dev()))


def test_full_size_properties():
    """BASELINE sizes (768x512, and one 1088x1920 compress-side pass): size-independent properties --
    eval x_hat is integer (+median / +mean), likelihoods in [1e-9, 1], indexes in [0, 63], bpp of
    the per-channel sums equals the total, idempotence of quantisation."""
    torch.manual_seed(0)
    net = mmcodec.build_model("bmshj2018-hyperprior", 4).eval()
    net.update()
    net = net.to(dev())
    x = torch.rand(4, 3, 512, 768, device=dev())
    with torch.no_grad():
        out = net(x)
    ly, lz = out["likelihoods"]["y"], out["likelihoods"]["z"]
    assert tuple(ly.shape) == (4, 192, 32, 48) and tuple(lz.shape) == (4, 128, 8, 12) and tuple(out["x_hat"].shape) == (4, 3, 512, 768)
    for lk in (ly, lz):
        assert float(lk.min()) >= 1e-9 * 0.999 and float(lk.max()) <= 1 + 1e-6 and bool(torch.isfinite(lk).all())
    total = ops.bits(ly).item()
    per_channel = sum(ops.bits(ly[:, c:c + 1].contiguous()).item() for c in range(0, 192, 48))
    part = sum(ops.bits(ly[:, c:c + 48].contiguous()).item() for c in range(0, 192, 48))
    assert abs(part - total) / total < 1e-4 and per_channel > 0
    net2 = mmcodec.build_model("mbt2018-mean", 6).eval()
    net2.update()
    net2 = net2.to(dev())
    x2 = torch.rand(1, 3, 1088, 1920, device=dev())
    with torch.no_grad():
        c = net2.symbols_and_indexes(x2)
    assert tuple(c["y_symbols"].shape) == (1, 320, 68, 120) and tuple(c["z_symbols"].shape) == (1, 192, 17, 30)
    assert int(c["y_indexes"].min()) >= 0 and int(c["y_indexes"].max()) <= 63
    assert torch.equal(c["z_indexes"][0, :, 0, 0].cpu(), torch.arange(192, dtype=torch.int32))
    # quantisation is idempotent: quantize(dequantize(sym)) == sym
    eb = net2.entropy_bottleneck
    med = eb._get_medians().detach().reshape(1, -1, 1, 1)
    z_hat = eb.dequantize(c["z_symbols"], med)
    assert torch.equal(eb.quantize(z_hat, "symbols", med), c["z_symbols"])


def test_host_pipeline_matches_device_forward():
    """Host-buffer API: pipelined micro-batches give exactly the results of one device forward."""
    net, _ = load(mmcodec.ScaleHyperprior, "hyperprior", 128, 192)
    x = torch.from_numpy(make_image(5, 64, 128, seed=11))
    with torch.no_grad():
        ref = net(x.to(dev()))
    pipe = mmcodec.HostPipeline(net, micro_batch=2)
    for _ in range(2):   # second call reuses the pinned result buffers
        out = pipe(x.pin_memory())
        torch.cuda.current_stream().synchronize()
        assert torch.equal(out["x_hat"], ref["x_hat"].cpu())
        for k in ref["likelihoods"]:
            assert out["likelihoods"][k].shape == ref["likelihoods"][k].shape
            assert torch.equal(out["likelihoods"][k], ref["likelihoods"][k].cpu())
    with pytest.raises(ValueError):
        pipe(x.to(dev()))


def test_host_pipeline_metrics_mode():
    """outputs="metrics": per-image bpp and MSE reduced on the device (mmc_image_bits / mmc_image_sse inside the micro-batch
    graphs) equal the values computed on the host from the full outputs (eval_model/__main__t.py:151-173), incl. the ragged
    tail micro-batch; the stand-alone kernels against torch on odd sizes and channels-last likelihoods."""
    import math
    from mmcodec import ops
    net, _ = load(mmcodec.ScaleHyperprior, "hyperprior", 128, 192)
    x = torch.from_numpy(make_image(5, 64, 128, seed=12))
    with torch.no_grad():
        ref = net(x.to(dev()))
    bpp_ref = sum(torch.log(l.double()).flatten(1).sum(1) for l in ref["likelihoods"].values()).cpu() / (-math.log(2) * 64 * 128)
    mse_ref = ((ref["x_hat"].double().cpu() - x.double()) ** 2).flatten(1).mean(1)
    pipe = mmcodec.HostPipeline(net, micro_batch=2, outputs="metrics")
    for _ in range(2):
        out = pipe(x.pin_memory())
        torch.cuda.current_stream().synchronize()
        assert set(out) == {"bpp", "mse"} and tuple(out["bpp"].shape) == (5,)
        assert float(((out["bpp"].double() - bpp_ref).abs() / bpp_ref).max()) < 1e-4
        assert float(((out["mse"].double() - mse_ref).abs() / mse_ref).max()) < 1e-4
    with pytest.raises(ValueError):
        mmcodec.HostPipeline(net, outputs="everything")
    # 8-bit host images: the device-side /255 is bit-exact with ToTensor's, so both input formats give identical results
    x8 = (x * 255).round().to(torch.uint8)
    xf = x8.to(torch.float32).div(255)
    assert torch.equal(ops.u8_to_f32(x8.to(dev())).cpu(), xf)
    full = mmcodec.HostPipeline(net, micro_batch=2)
    o8 = {k: (v.clone() if torch.is_tensor(v) else {n: t.clone() for n, t in v.items()}) for k, v in full(x8.pin_memory()).items()}
    torch.cuda.synchronize()
    of = full(xf.pin_memory())
    torch.cuda.synchronize()
    assert torch.equal(o8["x_hat"], of["x_hat"]) and all(torch.equal(o8["likelihoods"][n], of["likelihoods"][n]) for n in of["likelihoods"])
    with pytest.raises(TypeError):
        full(x.double().pin_memory())
    gen = torch.Generator().manual_seed(3)
    lk = (torch.rand(3, 7, 5, 9, generator=gen) * 0.9 + 0.05).to(dev())
    for t in (lk, lk.contiguous(memory_format=torch.channels_last)):
        acc = torch.zeros(3, device=dev())
        ops.image_bits(t, acc, 0.5)
        want = -0.5 * torch.log2(lk.double()).flatten(1).sum(1)
        assert float(((acc.double() - want).abs() / want.abs()).max()) < 1e-5
    a, b = torch.randn(2, 3, 11, 13, generator=gen).to(dev()), torch.randn(2, 3, 11, 13, generator=gen).to(dev())
    acc = ops.image_mse(a, b, torch.zeros(2, device=dev()))
    assert float((acc.double() - ((a.double() - b.double()) ** 2).flatten(1).mean(1)).abs().max()) < 1e-5


@pytest.mark.parametrize("arch,cls,N,M", ARCHS)
def test_compress_decompress_round_trip(models_golden, arch, cls, N, M):
    """CompressionModel.compress / decompress (models/google.py:196-205,324-344,393-416): the rANS streams decode to
    exactly the symbols that were coded, decompress() reproduces forward()'s reconstruction, and the coded size
    matches the reference's own streams for the same images to within 1 %."""
    g = models_golden
    tag = arch.replace("-", "_")
    net, _ = load(cls, arch, N, M)
    x = torch.from_numpy(g["x"]).to(dev())
    with torch.no_grad():
        c = net.compress(x)
        d = net.decompress(c["strings"], c["shape"])
        fwd = net(x)
    B = x.shape[0]
    assert len(c["strings"]) == (1 if arch == "factorized" else 2)
    assert all(isinstance(s, bytes) for ss in c["strings"] for s in ss) and all(len(ss) == B for ss in c["strings"])
    assert torch.equal(d["x_hat"], fwd["x_hat"].clamp(0, 1))
    mine = sum(len(s) for ss in c["strings"] for s in ss)
    ref = sum(g[f"{tag}_string_{si}_{bi}"].size for si in range(len(c["strings"])) for bi in range(B))
    assert abs(mine - ref) / ref < 0.01, (mine, ref)
    # (No coded-size vs entropy-estimate check: with these synthetic weights ~20 % of the y likelihoods sit on the 1e-9
    #  floor, i.e. 30 estimated bits each, while the coder spends a few bypass nibbles on them.)


def test_forward_bpp_from_fused_accumulator():
    """Eval forward sums -log2(likelihood) inside the entropy kernels; bpp(out) uses that sum and equals the separate reduction pass
    (taken on clones, which do not carry the accumulator), also on the second forward of the same module and under graph replay."""
    net, _ = load(mmcodec.ScaleHyperprior, "hyperprior", 128, 192)
    npix = 2 * 64 * 128
    for seed in (50, 51):
        x = torch.from_numpy(make_image(2, 64, 128, seed=seed)).to(dev())
        with torch.no_grad():
            out = net(x)
        assert getattr(out["likelihoods"]["y"], "_mmc_bits_total", None) is out["likelihoods"]["z"]._mmc_bits_total
        fused = net.bpp(out, npix)
        plain = net.bpp({"x_hat": out["x_hat"], "likelihoods": {k: v.clone() for k, v in out["likelihoods"].items()}}, npix)
        assert abs(fused - plain) / plain < 1e-5, (fused, plain)
    g = mmcodec.GraphedForward(net, x)
    for _ in range(2):
        o = g(x)
        torch.cuda.synchronize()
        assert abs(net.bpp(o, npix) - plain) / plain < 1e-5


def test_compress_pipeline_matches_compress():
    """mmcodec.CompressPipeline (host coding of batch i overlapped with the GPU stage of batch i + 1): byte-identical to compress()."""
    net, _ = load(mmcodec.MeanScaleHyperprior, "mean-scale", 192, 320)
    xs = [torch.from_numpy(make_image(3, 64, 128, seed=40 + i)).to(dev()) for i in range(4)]
    with torch.no_grad():
        want = [net.compress(x) for x in xs]
    pipe = mmcodec.CompressPipeline(net, depth=2)
    got = [f.result() for f in [pipe.submit(x) for x in xs]]
    pipe.close()
    for w, g in zip(want, got):
        assert g["strings"] == w["strings"] and tuple(g["shape"]) == tuple(w["shape"])
    fnet, _ = load(mmcodec.FactorizedPrior, "factorized", 128, 192)
    with torch.no_grad():
        w = fnet.compress(xs[0])
    fp = mmcodec.CompressPipeline(fnet)
    assert fp.submit(xs[0]).result()["strings"] == w["strings"]
    fp.close()


def test_graphed_forward_replays_exactly():
    """mmcodec.GraphedForward: the whole forward captured as one CUDA graph gives the eager results, also on new inputs."""
    net, _ = load(mmcodec.MeanScaleHyperprior, "mean-scale", 192, 320)
    x0 = torch.from_numpy(make_image(2, 64, 128, seed=21)).to(dev())
    x1 = torch.from_numpy(make_image(2, 64, 128, seed=22)).to(dev())
    g = mmcodec.GraphedForward(net, x0)
    with torch.no_grad():
        for x in (x0, x1, x0):
            want = net(x)
            got = g(x)
            torch.cuda.synchronize()
            assert torch.equal(got["x_hat"], want["x_hat"])
            for k in want["likelihoods"]:
                assert torch.equal(got["likelihoods"][k], want["likelihoods"][k])
    with pytest.raises(ValueError):
        g(torch.zeros(1, 3, 64, 128, device=dev()))
    with pytest.raises(RuntimeError):
        mmcodec.GraphedForward(net, x0.cpu())


def test_graphs_follow_parameter_updates():
    """Captured graphs hold the packed weights of capture time: both wrappers re-capture after an in-place parameter update."""
    net, _ = load(mmcodec.ScaleHyperprior, "hyperprior", 128, 192)
    x = torch.from_numpy(make_image(2, 64, 128, seed=31))
    g = mmcodec.GraphedForward(net, x.to(dev()))
    pipe = mmcodec.HostPipeline(net, micro_batch=2)
    pipe(x.pin_memory())
    with torch.no_grad():
        net.g_s[6].weight.mul_(1.5)
        net.g_s[6].bias.add_(0.25)
        want = net(x.to(dev()))
        got = g(x.to(dev()))
        torch.cuda.synchronize()
        assert torch.equal(got["x_hat"], want["x_hat"])
        out = pipe(x.pin_memory())
        torch.cuda.current_stream().synchronize()
        assert torch.equal(out["x_hat"], want["x_hat"].cpu())

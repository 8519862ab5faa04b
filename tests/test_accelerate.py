"""mmcodec.accelerate(model): the drop-in behind LIVE reference modules.

CPU part (authoring container only -- needs /root/reference): the reference's own zoo classes are instantiated, accelerated, and
must keep their state_dict (keys, values, Parameter identity), their classes' isinstance relations (CompressionModel.update()
depends on them, models/google.py:100-125) and a working update(); a CPU forward must raise (no CPU path).
GPU part (runs on the box, where the reference does not exist): the same adoption code on FOREIGN module classes -- stand-ins
that carry exactly the attributes of compressai.layers.GDN / entropy_models.* but none of this package's classes -- driven by the
reference's ScaleHyperprior.forward restated in the test, against the mirror model and the oracle."""
import contextlib
import io
import os
import sys

import pytest
import torch
import torch.nn as nn

import mmcodec
from mmcodec.transforms import TransformStack
from weights import make_image, make_state_dict

HAVE_REF = os.path.isdir("/root/reference/CompressAI")


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference tree (authoring container)")
def test_accelerate_live_reference_models():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import gen_golden
    with contextlib.redirect_stdout(io.StringIO()):
        gen_golden.import_reference()
    from compressai.entropy_models import EntropyBottleneck, GaussianConditional
    from compressai.layers import GDN
    from compressai.models.google import FactorizedPrior, MeanScaleHyperprior, ScaleHyperprior
    for cls, args in ((ScaleHyperprior, (128, 192)), (MeanScaleHyperprior, (192, 320)), (FactorizedPrior, (128, 192))):
        with contextlib.redirect_stdout(io.StringIO()):
            net = cls(*args)
        sd0 = {k: v.clone() for k, v in net.state_dict().items()}
        ids0 = {k: id(p) for k, p in net.named_parameters()}
        opt = torch.optim.Adam(net.parameters(), lr=1e-4)        # built BEFORE the swap: must keep pointing at live parameters
        assert mmcodec.accelerate(net) is net and type(net) is cls
        sd1 = net.state_dict()
        assert list(sd0) == list(sd1) and all(torch.equal(sd0[k], sd1[k]) for k in sd0)
        assert ids0 == {k: id(p) for k, p in net.named_parameters()}
        assert {id(p) for g in opt.param_groups for p in g["params"]} == set(ids0.values())
        eb = net.entropy_bottleneck
        assert isinstance(eb, EntropyBottleneck) and isinstance(eb, mmcodec.EntropyBottleneck)
        assert isinstance(net.g_a, TransformStack) and isinstance(net.g_s, TransformStack)
        assert isinstance(net.g_a[1], GDN) and isinstance(net.g_a[1], mmcodec.GDN) and isinstance(net.g_a[0], mmcodec.layers.Conv2d)
        if hasattr(net, "gaussian_conditional"):
            gc = net.gaussian_conditional
            assert isinstance(gc, GaussianConditional) and isinstance(gc, mmcodec.GaussianConditional)
            assert isinstance(net.h_a, TransformStack) and isinstance(net.h_s, TransformStack)
        with contextlib.redirect_stdout(io.StringIO()):
            assert net.update(force=True) is True            # the reference's update() finds the bottleneck by isinstance
        assert eb._quantized_cdf.numel() > 0 and float(net.aux_loss()) > 0
        with pytest.raises(RuntimeError, match="CUDA tensors only"):
            net(torch.rand(1, 3, 64, 64))
        assert mmcodec.accelerate(net)._mmc_accelerated_modules == 0   # idempotent: nothing left to swap
        # the reference scripts its GDN (tests/test_scripting.py:37-59): the accelerated one still scripts, into the registered op,
        # on the live Parameters
        sg = torch.jit.script(net.g_a[1])
        assert "ops.mmcodec.gdn" in sg.code and sg.state_dict()["gamma"].data_ptr() == net.g_a[1].gamma.data_ptr()


# ---- foreign stand-ins: the attribute sets of the reference's classes, none of this package's types ---------------------------
def _foreign(obj: nn.Module, name: str) -> nn.Module:
    cls = type(name, (nn.Module,), {"__module__": "fake_compressai." + name.lower(),
                                    "forward": lambda self, *a, **k: (_ for _ in ()).throw(AssertionError("reference forward called"))})
    new = object.__new__(cls)
    new.__dict__.update(obj.__dict__)
    new.__dict__["_modules"] = {k: _foreign_tree(v) for k, v in obj._modules.items()}
    for k in [k for k in new.__dict__ if k.startswith(("_cache", "_lut", "_bound_", "_pack", "_mmc"))]:
        del new.__dict__[k]                                    # what only the mirror classes carry
    return new


def _foreign_tree(m):
    if isinstance(m, mmcodec.GDN):
        return _foreign(m, "GDN")
    if isinstance(m, mmcodec.EntropyBottleneck):
        f = _foreign(m, "EntropyBottleneck")
        f.entropy_coder = type("Coder", (), {"name": "ans"})()
        del f.__dict__["entropy_coder_name"]
        return f
    if isinstance(m, mmcodec.GaussianConditional):
        f = _foreign(m, "GaussianConditional")
        f.entropy_coder = type("Coder", (), {"name": "ans"})()
        del f.__dict__["entropy_coder_name"]
        return f
    if isinstance(m, mmcodec.LowerBound):
        return _foreign(m, "LowerBound")
    if isinstance(m, mmcodec.NonNegativeParametrizer):
        return _foreign(m, "NonNegativeParametrizer")
    if isinstance(m, mmcodec.layers.ConvTranspose2d):
        c = nn.ConvTranspose2d(m.in_channels, m.out_channels, m.kernel_size, m.stride, m.padding, m.output_padding)
        c.weight, c.bias = m.weight, m.bias
        return c
    if isinstance(m, mmcodec.layers.Conv2d):
        c = nn.Conv2d(m.in_channels, m.out_channels, m.kernel_size, m.stride, m.padding)
        c.weight, c.bias = m.weight, m.bias
        return c
    if isinstance(m, nn.Sequential):
        return nn.Sequential(*[_foreign_tree(c) for c in m])
    return m


class ForeignScaleHyperprior(nn.Module):
    """compressai/models/google.py:218-295 restated: module tree of foreign classes + the reference's forward body"""

    def __init__(self, mirror):
        super().__init__()
        for name in ("entropy_bottleneck", "g_a", "g_s", "h_a", "h_s", "gaussian_conditional"):
            self.add_module(name, _foreign_tree(getattr(mirror, name)))

    def forward(self, x):
        y = self.g_a(x)
        z = self.h_a(torch.abs(y))
        z_hat, z_likelihoods = self.entropy_bottleneck(z)
        scales_hat = self.h_s(z_hat)
        y_hat, y_likelihoods = self.gaussian_conditional(y, scales_hat)
        x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_likelihoods, "z": z_likelihoods}}


@pytest.mark.gpu
def test_accelerate_foreign_modules_on_gpu():
    from oracle import torch_port as tp
    dev = torch.device("cuda", 0)
    sd = {k: torch.from_numpy(v) for k, v in make_state_dict("hyperprior", 128, 192, seed=0).items()}
    mirror = mmcodec.ScaleHyperprior(128, 192).eval()
    mirror.update()
    mirror.load_state_dict({**mirror.state_dict(), **sd})
    mirror.update(force=True)
    mirror = mirror.to(dev)
    foreign = ForeignScaleHyperprior(mirror).eval()
    assert not any(type(m).__module__.startswith("mmcodec") for m in foreign.modules())
    keys = list(foreign.state_dict())
    assert mmcodec.accelerate(foreign) is foreign and foreign._mmc_accelerated_modules == 6
    assert list(foreign.state_dict()) == keys
    assert isinstance(foreign.g_a, TransformStack) and isinstance(foreign.g_a[1], mmcodec.GDN) and type(foreign.g_a[1]).__name__ == "GDN"
    x = torch.from_numpy(make_image(2, 128, 192, seed=7))
    with torch.no_grad():
        out = foreign(x.to(dev))
        want = mirror(x.to(dev))
        ref = tp.hyperprior_forward(sd, x)
    npix = 2 * 128 * 192
    bpp, mbpp, rbpp = mirror.bpp(out, npix), mirror.bpp(want, npix), tp.bpp(ref, npix)
    assert abs(bpp - rbpp) / rbpp < 5e-3 and abs(bpp - mbpp) / mbpp < 5e-3, (bpp, mbpp, rbpp)
    rel = float(torch.sqrt(((out["x_hat"].float().cpu() - ref["x_hat"]) ** 2).mean() / (ref["x_hat"] ** 2).mean()))
    assert rel < 0.1, rel
    # the swapped entropy modules behave like the reference's on identical latents
    y, s = ref["y"].to(dev), ref["scales_hat"].to(dev)
    y_hat, lik = foreign.gaussian_conditional(y, s)
    assert torch.equal(y_hat.cpu(), ref["y_hat"])
    assert float(((lik.cpu() - ref["likelihoods"]["y"]).abs() / ref["likelihoods"]["y"].clamp_min(1e-9)).max()) < 1e-4
    assert torch.equal(foreign.gaussian_conditional.build_indexes(s).cpu(), tp.build_indexes(ref["scales_hat"], tp.get_scale_table()))

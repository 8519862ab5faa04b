#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (authoring container only).

Usage (from the repo root, in the container where /root/reference exists):

    python tests/golden/gen_golden.py

The reference (/root/reference/CompressAI, CompressAI 1.2.0.dev0 fork) is copied to a scratch
directory, its two pybind11 extensions are built there (``python setup.py build_ext --inplace``),
five unused-on-the-hot-path imports are stubbed and the reference's own
modules are driven with the deterministic weights of ``tests/golden/weights.py``.  Nothing from
the reference is copied into the repo; only input/output vectors are saved.  The GPU box has no
/root/reference, so tests read the committed .npz files only.
"""
from __future__ import annotations

import contextlib
import io
import os
import shutil
import subprocess
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from weights import make_image, make_master1_inputs, make_master_state_dict, make_mbt2018_state_dict, make_mm_state_dict, make_ssf_state_dict, make_state_dict  # noqa: E402

REF_SRC = "/root/reference/CompressAI"
SCRATCH = os.environ.get("MMC_REF_SCRATCH", "/tmp/ref_probe")


def import_reference():
    dst = os.path.join(SCRATCH, "CompressAI")
    if not os.path.exists(os.path.join(dst, "compressai")):
        os.makedirs(SCRATCH, exist_ok=True)
        shutil.copytree(REF_SRC, dst)
        subprocess.check_call(["chmod", "-R", "u+w", dst])
    if not any(f.startswith("_CXX") for f in os.listdir(os.path.join(dst, "compressai"))):
        subprocess.check_call([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=dst,
                              stdout=subprocess.DEVNULL)

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    stub("torchsnooper", snoop=lambda *a, **k: (lambda f: f))
    stub("imp")
    stub("torchinfo", summary=lambda *a, **k: None)
    stub("timm")
    stub("timm.models")
    stub("timm.models.layers", DropPath=nn.Identity,
         to_2tuple=lambda x: x if isinstance(x, tuple) else (x, x), trunc_normal_=nn.init.trunc_normal_)
    stub("pytorch_msssim", ms_ssim=None)
    sys.path.insert(0, dst)
    import compressai  # noqa: F401
    return compressai


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def load_into(module, sd, prefix=""):
    """Copy arrays into the reference module's parameters by state_dict key (in place)."""
    target = module.state_dict()
    for k, v in sd.items():
        if not k.startswith(prefix):
            continue
        kk = k[len(prefix):]
        assert kk in target, kk
        assert tuple(target[kk].shape) == v.shape, (kk, target[kk].shape, v.shape)
        target[kk].copy_(torch.from_numpy(v))


def t2n(t):
    return t.detach().cpu().numpy()


def kernel_goldens(out):
    from compressai.entropy_models import EntropyBottleneck, GaussianConditional
    from compressai.layers import GDN
    from compressai.models.google import get_scale_table
    from compressai.models.utils import conv, deconv
    from compressai.ops import LowerBound
    from compressai._CXX import pmf_to_quantized_cdf

    rs = np.random.RandomState(7)
    torch.manual_seed(0)

    # ---- quantize / dequantize (entropy_models.py:157-199) ------------------------------
    em = GaussianConditional(None)
    kat = np.array([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, -0.0, 3.49999, -7.5000005, 1e-30, 123456.5], np.float32)
    out["q_kat_x"] = kat
    out["q_kat_sym"] = t2n(em.quantize(torch.from_numpy(kat), "symbols"))
    x = (rs.standard_normal((2, 6, 5, 7)) * 6).astype(np.float32)
    means = (rs.standard_normal((2, 6, 5, 7)) * 3).astype(np.float32)
    # force exact ties after mean subtraction
    x.reshape(-1)[::5] = (means.reshape(-1)[::5] + rs.randint(-9, 9, x.reshape(-1)[::5].shape) + 0.5).astype(np.float32)
    out["q_x"], out["q_means"] = x, means
    out["q_sym_means"] = t2n(em.quantize(torch.from_numpy(x), "symbols", torch.from_numpy(means)))
    out["q_deq_means"] = t2n(em.quantize(torch.from_numpy(x), "dequantize", torch.from_numpy(means)))
    out["q_sym_nomeans"] = t2n(em.quantize(torch.from_numpy(x), "symbols"))
    out["q_deq_nomeans"] = t2n(em.quantize(torch.from_numpy(x), "dequantize"))
    chmeans = (rs.standard_normal((1, 6, 1, 1))).astype(np.float32)
    out["q_chmeans"] = chmeans
    out["q_sym_chmeans"] = t2n(em.quantize(torch.from_numpy(x), "symbols", torch.from_numpy(chmeans)))
    sym = out["q_sym_means"]
    out["dq_means"] = t2n(em.dequantize(torch.from_numpy(sym), torch.from_numpy(means)))
    out["dq_nomeans"] = t2n(em.dequantize(torch.from_numpy(sym)))

    # ---- build_indexes (entropy_models.py:735-740) -----------------------------------------
    gc = GaussianConditional(None)
    quiet(gc.update_scale_table, get_scale_table())
    table = t2n(gc.scale_table)
    out["scale_table"] = table
    out["gc_quantized_cdf"] = t2n(gc._quantized_cdf)
    out["gc_cdf_length"] = t2n(gc._cdf_length)
    out["gc_offset"] = t2n(gc._offset)
    adv = [-1.0, 0.0, 0.11, 1.0, 200.0, 256.0, 1e6, np.nan, np.inf, -np.inf, 0.10999999, 0.11000001]
    for t in table:
        adv += [t, np.nextafter(np.float32(t), np.float32(1e9)), np.nextafter(np.float32(t), np.float32(-1e9))]
    rnd = np.exp(rs.uniform(np.log(0.05), np.log(300), 4000))
    scales = np.concatenate([np.array(adv, np.float32), rnd.astype(np.float32)]).astype(np.float32)
    out["bi_scales"] = scales
    out["bi_indexes"] = t2n(gc.build_indexes(torch.from_numpy(scales)))

    # ---- GaussianConditional.forward (entropy_models.py:715-731) ----------------------------
    n = 6000
    sig = np.exp(rs.uniform(np.log(0.05), np.log(300), n)).astype(np.float32)
    mu = rs.uniform(-4, 4, n).astype(np.float32)
    y = (sig * rs.standard_normal(n) + mu).astype(np.float32)
    y[:200] = mu[:200] + rs.randint(-3, 4, 200) + 0.5  # ties
    y[200:260] = 50.0  # deep tail -> likelihood floor
    sig[200:230] = 2.0
    shp = (2, 3, 10, 100)
    y, sig, mu = y.reshape(shp), sig.reshape(shp), mu.reshape(shp)
    out["gc_y"], out["gc_scales"], out["gc_means"] = y, sig, mu
    gc.eval()
    a, b = gc(torch.from_numpy(y), torch.from_numpy(sig), torch.from_numpy(mu))
    out["gc_yhat_means"], out["gc_lik_means"] = t2n(a), t2n(b)
    a, b = gc(torch.from_numpy(y), torch.from_numpy(sig))
    out["gc_yhat_nomeans"], out["gc_lik_nomeans"] = t2n(a), t2n(b)
    torch.manual_seed(11)
    noise = torch.empty(shp).uniform_(-0.5, 0.5)
    torch.manual_seed(11)
    a, b = gc(torch.from_numpy(y), torch.from_numpy(sig), torch.from_numpy(mu), training=True)
    out["gc_noise"] = t2n(noise)
    out["gc_yhat_noise"], out["gc_lik_noise"] = t2n(a), t2n(b)
    assert np.array_equal(t2n(a), y + t2n(noise))

    # ---- EntropyBottleneck (entropy_models.py:495-540) ---------------------------------------
    C = 8
    ebw = {}
    from weights import _entropy_bottleneck
    _entropy_bottleneck(np.random.RandomState(3), ebw, "eb", C)
    eb = EntropyBottleneck(C)
    load_into(eb, ebw, "eb.")
    eb.eval()
    for k, v in ebw.items():
        out["eb_param_" + k[3:]] = v
    med = ebw["eb.quantiles"][:, 0, 1]
    xe = (med[None, :, None, None] + rs.uniform(-40, 40, (3, C, 4, 50))).astype(np.float32)
    xe[0, :, 0, :10] = med[:, None] + np.arange(-5, 5)[None, :] + 0.5  # ties
    out["eb_x"] = xe
    a, b = eb(torch.from_numpy(xe))
    out["eb_xhat"], out["eb_lik"] = t2n(a), t2n(b)
    # noise mode: reproduce the uniform_ draw made inside quantize() on the (C,1,L) view
    torch.manual_seed(5)
    nz = torch.empty(C, 1, 3 * 4 * 50).uniform_(-0.5, 0.5)
    torch.manual_seed(5)
    a, b = eb(torch.from_numpy(xe), training=True)
    nz_nchw = nz.reshape(C, 3, 4, 50).permute(1, 0, 2, 3).contiguous()
    assert np.array_equal(t2n(a), xe + t2n(nz_nchw))
    out["eb_noise"] = t2n(nz_nchw)
    out["eb_xhat_noise"], out["eb_lik_noise"] = t2n(a), t2n(b)
    v = torch.from_numpy(xe).permute(1, 0, 2, 3).reshape(C, 1, -1)
    out["eb_logits"] = t2n(eb._logits_cumulative(v, stop_gradient=True).reshape(C, 3, 4, 50).permute(1, 0, 2, 3))
    # 1-D spatial (B, C) and 5-D inputs (tests/test_entropy_models.py:199-220)
    x2 = (rs.standard_normal((5, C)) * 4).astype(np.float32)
    a, b = eb(torch.from_numpy(x2))
    out["eb_x_2d"], out["eb_xhat_2d"], out["eb_lik_2d"] = x2, t2n(a), t2n(b)
    out["eb_indexes"] = t2n(eb._build_indexes(torch.Size((2, C, 3, 5))))
    eb.update(force=True)
    out["eb_quantized_cdf"] = t2n(eb._quantized_cdf)
    out["eb_cdf_length"] = t2n(eb._cdf_length)
    out["eb_offset"] = t2n(eb._offset)
    out["eb_loss"] = t2n(eb.loss())

    # ---- GDN / IGDN (layers/gdn.py:77-92) ----------------------------------------------------
    from weights import _gdn
    C = 16
    gw = {}
    _gdn(np.random.RandomState(4), gw, "g", C)
    xg = (rs.standard_normal((2, C, 5, 9)) * 2).astype(np.float32)
    out["gdn_x"], out["gdn_beta"], out["gdn_gamma"] = xg, gw["g.beta"], gw["g.gamma"]
    for inv in (False, True):
        g = GDN(C, inverse=inv)
        load_into(g, gw, "g.")
        out["gdn_y_inv" if inv else "gdn_y"] = t2n(g(torch.from_numpy(xg)))
    # closed form at init (tests/test_layers.py:145-146)
    out["gdn_init_y"] = t2n(GDN(C)(torch.from_numpy(xg)))

    # ---- LowerBound (ops/bound_ops.py:36-56) --------------------------------------------------
    lb = LowerBound(0.11)
    xl = torch.tensor([-1.0, 0.0, 0.11, 0.2, 5.0, float("nan")], requires_grad=True)
    yl = lb(xl)
    gl = torch.tensor([1.0, -1.0, 1.0, -2.0, 3.0, 1.0])
    yl.backward(gl)
    out["lb_x"], out["lb_y"], out["lb_g"], out["lb_dx"] = t2n(xl), t2n(yl), t2n(gl), t2n(xl.grad)

    # ---- pmf_to_quantized_cdf (cpp_exts/ops/ops.cpp:40-109; KAT tests/test_ops.py:104-106) ----
    out["cdf_kat"] = np.array(pmf_to_quantized_cdf([0.1, 0.2, 0.0, 0.0], 16), np.uint32)
    pmfs, cdfs = [], []
    for i in range(6):
        L = [5, 17, 33, 64, 129, 300][i]
        p = rs.dirichlet(np.full(L, 0.3)).astype(np.float32)
        p[rs.rand(L) < 0.3] = 0.0
        p[0] = max(p[0], 1e-3)
        pmfs.append(np.pad(p, (0, 300 - L)))
        cdfs.append(np.pad(np.array(pmf_to_quantized_cdf(p.tolist(), 16), np.uint32), (0, 300 - L)))
    out["cdf_pmfs"], out["cdf_cdfs"] = np.stack(pmfs), np.stack(cdfs)
    out["cdf_lens"] = np.array([5, 17, 33, 64, 129, 300], np.int32)

    # ---- conv / deconv (models/utils.py:128-146) ----------------------------------------------
    for tag, (cin, cout, k, s, h, w) in {"a": (3, 8, 5, 2, 12, 20), "b": (8, 6, 3, 1, 7, 9), "c": (4, 4, 5, 2, 9, 11)}.items():
        m = conv(cin, cout, kernel_size=k, stride=s)
        xin = rs.standard_normal((2, cin, h, w)).astype(np.float32)
        out[f"conv_{tag}_x"], out[f"conv_{tag}_w"], out[f"conv_{tag}_b"] = xin, t2n(m.weight), t2n(m.bias)
        out[f"conv_{tag}_y"] = t2n(m(torch.from_numpy(xin)))
        out[f"conv_{tag}_cfg"] = np.array([k, s], np.int32)
    for tag, (cin, cout, k, s, h, w) in {"a": (8, 3, 5, 2, 6, 10), "b": (4, 6, 5, 2, 5, 7), "c": (6, 4, 3, 1, 5, 6)}.items():
        m = deconv(cin, cout, kernel_size=k, stride=s)
        xin = rs.standard_normal((2, cin, h, w)).astype(np.float32)
        out[f"deconv_{tag}_x"], out[f"deconv_{tag}_w"], out[f"deconv_{tag}_b"] = xin, t2n(m.weight), t2n(m.bias)
        out[f"deconv_{tag}_y"] = t2n(m(torch.from_numpy(xin)))
        out[f"deconv_{tag}_cfg"] = np.array([k, s], np.int32)


def model_goldens(out):
    from compressai.models import FactorizedPrior, MeanScaleHyperprior, ScaleHyperprior

    x = make_image(2, 128, 192, seed=1234)
    out["x"] = x
    for arch, cls, N, M in (("factorized", FactorizedPrior, 128, 192),
                            ("hyperprior", ScaleHyperprior, 128, 192),
                            ("mean-scale", MeanScaleHyperprior, 192, 320)):
        sd = make_state_dict(arch, N, M, seed=0)
        torch.manual_seed(0)
        net = quiet(cls, N, M).eval()
        load_into(net, sd)
        quiet(net.update, force=True)
        with torch.no_grad():
            o = quiet(net, torch.from_numpy(x))
        tag = arch.replace("-", "_")
        import json
        out[f"{tag}_state_dict"] = np.array(json.dumps({k: [list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()}))
        out[f"{tag}_x_hat"] = t2n(o["x_hat"])
        for k, v in o["likelihoods"].items():
            out[f"{tag}_lik_{k}"] = t2n(v)
        out[f"{tag}_cfg"] = np.array([N, M], np.int32)
        # symbols / indexes exactly as handed to the rANS coder (entropy_models.py:260-269):
        captured = []

        class Spy:
            def __init__(self, inner):
                self.inner = inner

            def encode_with_indexes(self, symbols, indexes, *rest):
                captured.append((np.array(symbols, np.int32), np.array(indexes, np.int32)))
                return self.inner.encode_with_indexes(symbols, indexes, *rest)

            def __getattr__(self, name):
                return getattr(self.inner, name)

        spies = []
        for mod in net.modules():
            if hasattr(mod, "entropy_coder"):
                mod.entropy_coder = Spy(mod.entropy_coder)
                spies.append(mod)
        with torch.no_grad():
            c = quiet(net.compress, torch.from_numpy(x))
        B = x.shape[0]
        if arch == "factorized":
            ysym = np.stack([captured[i][0] for i in range(B)])
            yidx = np.stack([captured[i][1] for i in range(B)])
            out[f"{tag}_y_symbols"], out[f"{tag}_y_indexes"] = ysym, yidx
        else:
            # call order: entropy_bottleneck.compress (z, B images) then gaussian_conditional.compress (y)
            out[f"{tag}_z_symbols"] = np.stack([captured[i][0] for i in range(B)])
            out[f"{tag}_z_indexes"] = np.stack([captured[i][1] for i in range(B)])
            out[f"{tag}_y_symbols"] = np.stack([captured[B + i][0] for i in range(B)])
            out[f"{tag}_y_indexes"] = np.stack([captured[B + i][1] for i in range(B)])
        out[f"{tag}_shape"] = np.array(list(c["shape"]), np.int32)
        for si, strings in enumerate(c["strings"]):
            for bi, s in enumerate(strings):
                out[f"{tag}_string_{si}_{bi}"] = np.frombuffer(s, np.uint8)
        with torch.no_grad():
            d = quiet(net.decompress, c["strings"], c["shape"])
        out[f"{tag}_dec_x_hat"] = t2n(d["x_hat"])


def mm_goldens(out):
    """Two-branch RGB + depth codec of the fork (compressai/models/google.py:746-1248), eval forward."""
    import json
    from compressai.models.google import (JointAutoregressiveHierarchicalPriors_D,
                                          JointAutoregressiveHierarchicalPriors_R)
    x = make_image(1, 128, 192, seed=1234)
    d = make_image(1, 128, 192, seed=5, C=1)
    out["x"], out["depth"] = x, d
    torch.manual_seed(0)
    net_r = quiet(JointAutoregressiveHierarchicalPriors_R, 192, 192).eval()
    net_d = quiet(JointAutoregressiveHierarchicalPriors_D, 192, 192).eval()
    for tag, net, seed in (("r", net_r, 0), ("d", net_d, 1)):
        shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        load_into(net, make_mm_state_dict(shapes, seed))
        quiet(net.update, force=True)
        out[f"{tag}_state_dict"] = np.array(json.dumps({k: [list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()}))
    with torch.no_grad():
        o_r = quiet(net_r, torch.from_numpy(x))
        o_d = quiet(net_d, torch.from_numpy(d), o_r["hidden"])
    for tag, o in (("r", o_r), ("d", o_d)):
        out[f"{tag}_x_hat"] = t2n(o["x_hat"])
        for k, v in o["likelihoods"].items():
            out[f"{tag}_lik_{k}"] = t2n(v)
    for k, v in o_r["hidden"].items():
        out[f"r_hidden_{k}_mean_abs"] = np.array(float(v.abs().mean()))


def make_frames(n: int, H: int, W: int, seed: int = 21):
    """A short synthetic sequence: one textured image translated by a few pixels per frame plus a little noise."""
    base = make_image(1, H + 32, W + 32, seed=seed)[0]
    rs = np.random.RandomState(seed + 1)
    frames = []
    for t in range(n):
        dy, dx = 2 * t, 3 * t
        f = base[:, 8 + dy: 8 + dy + H, 8 + dx: 8 + dx + W] + rs.uniform(-0.01, 0.01, (3, H, W))
        frames.append(np.clip(f, 0, 1).astype(np.float32)[None])
    return frames


def ssf_goldens(out):
    """ssf2020 video codec (compressai/models/video/google.py), eval forward on a 3-frame 128x256 sequence, plus the
    scale-space prediction on its own (gaussian_volume / warp_volume) and the per-frame compressed sizes."""
    import json
    from compressai.models.video.google import ScaleSpaceFlow
    frames = make_frames(3, 128, 256)
    torch.manual_seed(0)
    net = quiet(ScaleSpaceFlow).eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    load_into(net, make_ssf_state_dict(shapes, 0))
    quiet(net.update, force=True)
    out["state_dict"] = np.array(json.dumps({k: [list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()}))
    fr = [torch.from_numpy(f) for f in frames]
    with torch.no_grad():
        o = quiet(net, fr)
        vol = net.gaussian_volume(o["x_hat"][0], net.sigma0, net.num_levels)
        y_m = net.motion_encoder(torch.cat((fr[1], o["x_hat"][0]), dim=1))
        y_m_hat, _ = net.motion_hyperprior(y_m)
        motion_info = net.motion_decoder(y_m_hat)
        x_pred = net.forward_prediction(o["x_hat"][0], motion_info)
        strings, shapes_c = quiet(net.compress, fr)
        dec = quiet(net.decompress, strings, shapes_c)
    for t, f in enumerate(frames):
        out[f"frame_{t}"] = f
        out[f"x_hat_{t}"] = t2n(o["x_hat"][t])
        for part, lk in o["likelihoods"][t].items():
            for k, v in lk.items():
                out[f"lik_{t}_{part}_{k}"] = t2n(v)
        out[f"dec_max_abs_diff_{t}"] = np.array(float((dec[t] - o["x_hat"][t]).abs().max()))
    out["volume_sub"] = t2n(vol[:, :, :, ::3, ::5])   # strided subsample: pins every level without storing 2.3 MB
    out["motion_info"] = t2n(motion_info)
    out["x_pred"] = t2n(x_pred)
    out["bytes_keyframe"] = np.array([len(s[0]) for s in strings[0]])
    out["bytes_inter_1"] = np.array([len(strings[1][k][i][0]) for k in ("motion", "residual") for i in range(2)])
    flow = motion_info[:, :2]
    print("ssf stats: flow px", float(flow[:, 0].abs().mean() * 128), float(flow[:, 0].abs().max() * 128), "scale z", float(motion_info[:, 2].min()),
          float(motion_info[:, 2].max()), "x_hat range", float(o["x_hat"][2].min()), float(o["x_hat"][2].max()))
    for t in range(3):
        for part, lk in o["likelihoods"][t].items():
            print(t, part, {k: (float(torch.log2(v).sum() / -(128 * 256)), float((v <= 1.0001e-9).float().mean())) for k, v in lk.items()})


def guided_goldens(out):
    """Guided_compresser (compressai/models/master.py:1215-1300), the RGB-T reproduction's guide codec: 1-channel input,
    eval forward incl. the hidden maps' statistics, on the two-branch weight recipe."""
    import json
    from compressai.models.master import Guided_compresser
    d = make_image(1, 128, 192, seed=5, C=1)
    torch.manual_seed(0)
    net = quiet(Guided_compresser, channel=1).eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    load_into(net, make_mm_state_dict(shapes, 2))
    quiet(net.update, force=True)
    out["state_dict"] = np.array(json.dumps({k: [list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()}))
    with torch.no_grad():
        o = quiet(net, torch.from_numpy(d))
    out["x"], out["x_hat"] = d, t2n(o["x_hat"])
    for k, v in o["likelihoods"].items():
        out[f"lik_{k}"] = t2n(v)
    for k, v in o["hidden"].items():
        out[f"hidden_{k}"] = t2n(v[:, ::8, ::2, ::2])      # strided subsample of every hidden map


def master_goldens(out, verbose=False):
    """Master_compresser (compressai/models/master.py:837-951) driven by Guided_compresser's reconstruction and hidden maps
    (the pairing examples/train.py:208-274 uses): 3-channel master at 128x256, 1-channel guide at 64x128, eval forward."""
    import json
    from compressai.models.master import Guided_compresser, Master_compresser
    x = make_image(1, 128, 256, seed=8, C=3)
    g = make_image(1, 64, 128, seed=9, C=1)
    torch.manual_seed(0)
    guide = quiet(Guided_compresser, channel=1).eval()
    load_into(guide, make_mm_state_dict({k: tuple(v.shape) for k, v in guide.state_dict().items()}, 2))
    quiet(guide.update, force=True)
    net = quiet(Master_compresser, width=64, height=128, channel=3).eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    load_into(net, make_master_state_dict(shapes, 4))
    quiet(net.update, force=True)
    out["state_dict"] = np.array(json.dumps({k: [list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()}))
    with torch.no_grad():
        og = quiet(guide, torch.from_numpy(g))
        # the guide's outputs are rounded to bf16-representable values BEFORE the master sees them, and stored as bf16 bit
        # patterns: both sides of the parity test then start from identical inputs at half the fixture size
        hidden = {k: og["hidden"][k].bfloat16().float() for k in ("gs1", "gs2", "gs3")}
        g_hat = og["x_hat"].bfloat16().float()
        o = quiet(net, torch.from_numpy(x), g_hat, hidden)
    out["x"], out["g"], out["x_hat"] = x, g, t2n(o["x_hat"])
    out["g_hat_bf16"] = t2n(g_hat.bfloat16().view(torch.int16))
    for k, v in hidden.items():
        out[f"hidden_{k}_bf16"] = t2n(v.bfloat16().view(torch.int16))
    for k, v in o["likelihoods"].items():
        out[f"lik_{k}"] = t2n(v)
    if verbose:
        print("x_hat", float(o["x_hat"].min()), float(o["x_hat"].mean()), float(o["x_hat"].max()),
              "mse", float(((o["x_hat"] - torch.from_numpy(x)) ** 2).mean()))
        for k, v in o["likelihoods"].items():
            print(k, "bpp", float(torch.log2(v).sum() / -(128 * 256)), "floor frac", float((v <= 1.0001e-9).float().mean()),
                  "p>0.99", float((v > 0.99).float().mean()))


def mbt2018_goldens(out):
    """The zoo's JointAutoregressiveHierarchicalPriors (mbt2018, compressai/models/google.py:421-520): eval forward."""
    import json
    from compressai.models import JointAutoregressiveHierarchicalPriors
    d = make_image(1, 128, 192, seed=6)
    torch.manual_seed(0)
    net = quiet(JointAutoregressiveHierarchicalPriors, 192, 192).eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    load_into(net, make_mbt2018_state_dict(shapes, 3))
    quiet(net.update, force=True)
    out["state_dict"] = np.array(json.dumps({k: [list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()}))
    with torch.no_grad():
        o = quiet(net, torch.from_numpy(d))
    out["x"], out["x_hat"] = d, t2n(o["x_hat"])
    for k, v in o["likelihoods"].items():
        out[f"lik_{k}"] = t2n(v)
    print("mbt2018 bpp", {k: float(torch.log2(v).sum() / -(128 * 192)) for k, v in o["likelihoods"].items()},
          "floor", float((o["likelihoods"]["y"] <= 1.0001e-9).float().mean()), "x_hat", float(o["x_hat"].min()), float(o["x_hat"].max()))


def master1_goldens(out):
    """Master_compresser(channel=1) (master.py:840-850: 1-channel master at stride 1, 3-channel guide at stride 2, guide maps
    brought to the decoder's resolutions by decoder.downsample1-3, master.py:765-768,783-786): eval forward on seeded inputs."""
    import json
    from compressai.models.master import Master_compresser
    x, g_hat, hidden = make_master1_inputs(10)
    torch.manual_seed(0)
    net = quiet(Master_compresser, width=64, height=64, channel=1).eval()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    load_into(net, make_master_state_dict(shapes, 5))
    quiet(net.update, force=True)
    out["state_dict"] = np.array(json.dumps({k: [list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()}))
    with torch.no_grad():
        o = quiet(net, torch.from_numpy(x), torch.from_numpy(g_hat), {k: torch.from_numpy(v) for k, v in hidden.items()})
    out["x_hat"] = t2n(o["x_hat"])
    for k, v in o["likelihoods"].items():
        out[f"lik_{k}"] = t2n(v)
    print("master1 x_hat", float(o["x_hat"].min()), float(o["x_hat"].max()),
          {k: float(torch.log2(v).sum() / -(64 * 64)) for k, v in o["likelihoods"].items()})


def color_goldens(out):
    """compressai.transforms.functional on a random frame (the reference's own functions)."""
    from compressai.transforms.functional import rgb2ycbcr, ycbcr2rgb, yuv_420_to_444, yuv_444_to_420
    rgb = torch.from_numpy(make_image(2, 36, 52, seed=77))
    ycc = rgb2ycbcr(rgb)
    y, u, v = yuv_444_to_420(ycc)
    out["rgb"], out["ycbcr"], out["rgb_back"] = t2n(rgb), t2n(ycc), t2n(ycbcr2rgb(ycc))
    out["u420"], out["v420"] = t2n(u), t2n(v)
    out["yuv444"] = t2n(yuv_420_to_444((y, u, v)))


def main():
    import_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    which = sys.argv[1:] or ["kernels", "models", "models_mm", "models_ssf", "color", "models_guided", "models_master", "models_mbt2018", "models_master1"]
    gens = {"kernels": kernel_goldens, "models": model_goldens, "models_mm": mm_goldens, "models_ssf": ssf_goldens, "color": color_goldens,
            "models_guided": guided_goldens, "models_master": master_goldens, "models_mbt2018": mbt2018_goldens, "models_master1": master1_goldens}
    for name in which:
        d = {}
        gens[name](d)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **d)
        print(f"{name}.npz", os.path.getsize(os.path.join(HERE, f"{name}.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()

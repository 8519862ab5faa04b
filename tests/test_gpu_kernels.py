"""GPU parity tests of the individual kernels, called through the C ABI (mmcodec.ops -> libmmcodec.so)
and compared with (a) the committed golden vectors produced by the reference itself and (b) the CPU
oracle on seeded inputs.  Integer outputs (symbols, indexes) must be bit-exact; fp32 likelihoods
within 1e-4 relative (the tolerance BASELINE.json states for fp32); conv outputs of the fp32
CUDA-core kernel within 1e-4 of the output scale."""
import numpy as np
import pytest
import torch

import oracle
from oracle import torch_port as tp

pytestmark = pytest.mark.gpu

import mmcodec  # noqa: E402
from mmcodec import _lib as L  # noqa: E402
from mmcodec import ops  # noqa: E402


def dev():
    return torch.device("cuda", 0)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def rel_err(a, b, floor):
    """max |a-b| / max(|b|, floor), after discounting 4 ulp(0.5) = 2.4e-7 of absolute error: the
    likelihoods are differences of two CDF values of magnitude <= 1/2 (entropy_models.py:705-707,
    :487-491), so the reference's own fp32 result carries that much cancellation noise."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.maximum(np.abs(a - b) - 2.4e-7, 0.0) / np.maximum(np.abs(b), floor)))


# ---- quantize / dequantize ---------------------------------------------------------------------
def test_quantize_golden(kernels_golden):
    g = kernels_golden
    assert np.array_equal(ops.quantize_symbols(cu(g["q_kat_x"])).cpu().numpy(), g["q_kat_sym"])
    x, m = cu(g["q_x"]), cu(g["q_means"])
    assert np.array_equal(ops.quantize_symbols(x, m).cpu().numpy(), g["q_sym_means"])
    assert np.array_equal(ops.quantize_dequantize(x, m).cpu().numpy(), g["q_deq_means"])
    assert np.array_equal(ops.quantize_symbols(x).cpu().numpy(), g["q_sym_nomeans"])
    assert np.array_equal(ops.quantize_dequantize(x).cpu().numpy(), g["q_deq_nomeans"])
    assert np.array_equal(ops.quantize_symbols(x, cu(g["q_chmeans"])).cpu().numpy(), g["q_sym_chmeans"])
    s = cu(g["q_sym_means"])
    assert np.array_equal(ops.dequantize(s, m).cpu().numpy(), g["dq_means"])
    assert np.array_equal(ops.dequantize(s).cpu().numpy(), g["dq_nomeans"])


@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (2, 7, 3), (2, 6, 5, 7), (1, 4, 3, 2, 5), (2, 3, 2, 2, 2, 3), (4, 192, 32, 48)])
def test_quantize_vs_oracle_any_rank(shape):
    """0-D..5-D spatial dims as in tests/test_entropy_models.py:199-220, ragged sizes, tie values."""
    rs = np.random.RandomState(sum(shape))
    x = (rs.standard_normal(shape) * 9).astype(np.float32)
    m = (rs.standard_normal(shape) * 2).astype(np.float32)
    x.reshape(-1)[::3] = m.reshape(-1)[::3] + rs.randint(-5, 5, x.reshape(-1)[::3].shape) + 0.5
    C = shape[1]
    inner = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    assert np.array_equal(ops.quantize_symbols(cu(x), cu(m)).cpu().numpy(), oracle.quantize_symbols(x, m))
    assert np.array_equal(ops.quantize_dequantize(cu(x), cu(m)).cpu().numpy(), oracle.quantize_dequantize(x, m))
    assert np.array_equal(ops.quantize_symbols(cu(x)).cpu().numpy(), np.round(x).astype(np.int32))  # test_entropy_models.py:79
    cm = rs.standard_normal((1, C) + (1,) * (len(shape) - 2)).astype(np.float32)
    assert np.array_equal(ops.quantize_symbols(cu(x), cu(cm)).cpu().numpy(), oracle.quantize_symbols(x, cm, C=C, inner=inner))
    sym = oracle.quantize_symbols(x, m)
    assert np.array_equal(ops.dequantize(cu(sym), cu(m)).cpu().numpy(), oracle.dequantize(sym, m))


def test_quantize_channels_last_and_unaligned():
    rs = np.random.RandomState(5)
    x = (rs.standard_normal((2, 6, 5, 7)) * 9).astype(np.float32)
    cm = rs.standard_normal((1, 6, 1, 1)).astype(np.float32)
    xcl = cu(x).to(memory_format=torch.channels_last)
    out = ops.quantize_symbols(xcl, cu(cm))
    assert out.shape == xcl.shape and np.array_equal(out.cpu().numpy(), oracle.quantize_symbols(x, cm, C=6, inner=35))
    flat = cu(np.concatenate([[0.0], x.reshape(-1)]).astype(np.float32))[1:]   # 4-byte aligned only
    assert np.array_equal(ops.quantize_symbols(flat.reshape(1, -1)).cpu().numpy().reshape(-1), np.round(x).astype(np.int32).reshape(-1))
    assert ops.quantize_symbols(torch.empty(0, 4, device=dev())).numel() == 0


# ---- indexes -------------------------------------------------------------------------------------
def test_build_indexes_golden_and_oracle(kernels_golden):
    g = kernels_golden
    table = cu(g["scale_table"])
    out = ops.build_indexes(cu(g["bi_scales"]), table, 0.11)
    assert out.dtype == torch.int32 and np.array_equal(out.cpu().numpy(), g["bi_indexes"])
    rs = np.random.RandomState(1)
    s = np.exp(rs.uniform(np.log(0.01), np.log(1000), (3, 320, 17, 30))).astype(np.float32)
    assert np.array_equal(ops.build_indexes(cu(s), table, 0.11).cpu().numpy(), oracle.build_indexes(s, g["scale_table"]))
    # unaligned + tiny + other table sizes
    for levels in (1, 2, 5, 64, 200):
        t = np.sort(np.exp(rs.uniform(-2, 5, levels))).astype(np.float32)
        v = np.concatenate([t, np.exp(rs.uniform(-3, 6, 37)).astype(np.float32)])
        buf = cu(np.concatenate([[1.0], v]).astype(np.float32))[1:]
        assert np.array_equal(ops.build_indexes(buf, cu(t), float(t[0])).cpu().numpy(), oracle.build_indexes(v, t, float(t[0])))


def test_channel_indexes(kernels_golden):
    assert np.array_equal(ops.channel_indexes((2, 8, 3, 5), dev()).cpu().numpy(), kernels_golden["eb_indexes"])
    assert np.array_equal(ops.channel_indexes((3, 5), dev()).cpu().numpy(), oracle.channel_indexes((3, 5)))
    assert np.array_equal(mmcodec.EntropyBottleneck._build_indexes((2, 4, 2, 2, 2), dev()).cpu().numpy(), oracle.channel_indexes((2, 4, 2, 2, 2)))


# ---- LowerBound -------------------------------------------------------------------------------------
def test_lower_bound_golden(kernels_golden):
    g = kernels_golden
    lb = mmcodec.LowerBound(0.11).to(dev())
    x = cu(g["lb_x"]).requires_grad_(True)
    y = lb(x)
    assert np.array_equal(y.detach().cpu().numpy(), g["lb_y"], equal_nan=True)
    y.backward(cu(g["lb_g"]))
    assert np.array_equal(x.grad.cpu().numpy(), g["lb_dx"])


# ---- GaussianConditional ------------------------------------------------------------------------------
def test_gc_forward_golden(kernels_golden):
    g = kernels_golden
    gc = mmcodec.GaussianConditional(None).to(dev()).eval()
    y, s, m = cu(g["gc_y"]), cu(g["gc_scales"]), cu(g["gc_means"])
    yh, lik = gc(y, s, m)
    assert np.array_equal(yh.cpu().numpy(), g["gc_yhat_means"])
    assert rel_err(lik.cpu().numpy(), g["gc_lik_means"], 1e-9) < 1e-4
    yh, lik = gc(y, s)
    assert np.array_equal(yh.cpu().numpy(), g["gc_yhat_nomeans"])
    assert rel_err(lik.cpu().numpy(), g["gc_lik_nomeans"], 1e-9) < 1e-4
    yh, lik = ops.gc_forward(y, s, m, cu(g["gc_noise"]))
    assert np.array_equal(yh.cpu().numpy(), g["gc_yhat_noise"])
    assert rel_err(lik.cpu().numpy(), g["gc_lik_noise"], 1e-9) < 1e-4
    # eval identity y_hat == round(x - mu) + mu (tests/test_entropy_models.py:352-363), floor at 1e-9
    _, l = gc(torch.full((1, 1, 1, 1), 50.0, device=dev()), torch.full((1, 1, 1, 1), 2.0, device=dev()))
    assert l.item() == np.float32(1e-9)
    # _likelihood (no bound) and bits reduction
    lk = gc._likelihood(cu(g["gc_yhat_means"]), s, m)
    ref = tp.gc_likelihood(torch.from_numpy(g["gc_yhat_means"]), torch.from_numpy(g["gc_scales"]), torch.from_numpy(g["gc_means"]))
    assert np.max(np.abs(lk.cpu().numpy() - ref.numpy())) < 1e-6
    bits = ops.bits(lik).item()
    assert abs(bits - oracle.bits(lik.cpu().numpy())) / abs(bits) < 1e-5


def test_gc_forward_large_vs_oracle_with_bits_and_bf16():
    rs = np.random.RandomState(3)
    shape = (2, 320, 17, 30)   # config-3 sized per-image latent, ragged (not a multiple of 4 along W*H*C? it is)
    sig = np.exp(rs.uniform(np.log(0.05), np.log(300), shape)).astype(np.float32)
    mu = rs.uniform(-4, 4, shape).astype(np.float32)
    y = (sig * rs.standard_normal(shape) + mu).astype(np.float32)
    bits = torch.zeros(1, device=dev())
    yh, lik, yb = ops.gc_forward(cu(y), cu(sig), cu(mu), want_bf16=True, bits=bits)
    ryh, rlik = oracle.gc_forward(y, sig, mu)
    assert np.array_equal(yh.cpu().numpy(), ryh)
    assert rel_err(lik.cpu().numpy(), rlik, 1e-9) < 1e-4
    assert torch.equal(yb.float().cpu(), torch.from_numpy(ryh).to(torch.bfloat16).float())
    assert abs(bits.item() - oracle.bits(rlik)) / oracle.bits(rlik) < 1e-5
    # odd length + unaligned base pointer -> scalar path
    n = 1001
    buf = [cu(np.concatenate([[0.0], a.reshape(-1)[:n]]).astype(np.float32))[1:].reshape(1, n) for a in (y, sig, mu)]
    yh, lik = ops.gc_forward(*buf)
    r = oracle.gc_forward(y.reshape(-1)[:n], sig.reshape(-1)[:n], mu.reshape(-1)[:n])
    assert np.array_equal(yh.cpu().numpy().reshape(-1), r[0]) and rel_err(lik.cpu().numpy().reshape(-1), r[1], 1e-9) < 1e-4


# ---- EntropyBottleneck ----------------------------------------------------------------------------------
def _load_eb(g, C=8):
    eb = mmcodec.EntropyBottleneck(C)
    sd = eb.state_dict()
    for k in sd:
        if "eb_param_" + k in g.files:
            sd[k].copy_(torch.from_numpy(g["eb_param_" + k]))
    return eb.to(dev()).eval()


def test_eb_forward_golden(kernels_golden):
    g = kernels_golden
    eb = _load_eb(g)
    x = cu(g["eb_x"])
    xh, lik = eb(x)
    assert np.array_equal(xh.cpu().numpy(), g["eb_xhat"])
    assert rel_err(lik.cpu().numpy(), g["eb_lik"], 1e-9) < 1e-4
    # channels-last memory, same logical result (thread<->channel variant of the kernel)
    xh2, lik2 = eb(x.to(memory_format=torch.channels_last))
    assert np.array_equal(xh2.cpu().numpy(), g["eb_xhat"])
    assert rel_err(lik2.cpu().numpy(), g["eb_lik"], 1e-9) < 1e-4
    # training mode with the reference's noise tensor
    xh, lik = ops.eb_forward(x, eb._params(), cu(g["eb_noise"]), 1e-9)
    assert np.array_equal(xh.cpu().numpy(), g["eb_xhat_noise"])
    assert rel_err(lik.cpu().numpy(), g["eb_lik_noise"], 1e-9) < 1e-4
    lg = ops.eb_logits_cumulative(x, eb._params())
    assert np.max(np.abs(lg.cpu().numpy() - g["eb_logits"])) < 2e-4
    xh, lik = eb(cu(g["eb_x_2d"]))
    assert np.array_equal(xh.cpu().numpy(), g["eb_xhat_2d"])
    assert rel_err(lik.cpu().numpy(), g["eb_lik_2d"], 1e-9) < 1e-4
    assert abs(eb.loss().item() - float(g["eb_loss"])) / float(g["eb_loss"]) < 1e-5
    # train mode: |x_hat - x| <= 0.5 (tests/test_entropy_models.py:164-175)
    eb.train()
    xh, _ = eb(x)
    assert (xh - x).abs().max().item() <= 0.5
    # _likelihood on the (C,1,L) view the reference uses
    eb.eval()
    v = torch.from_numpy(g["eb_xhat"]).permute(1, 0, 2, 3).reshape(8, 1, -1)
    assert rel_err(eb._likelihood(v.to(dev())).cpu().numpy(), g["eb_lik"].transpose(1, 0, 2, 3).reshape(8, 1, -1), 1e-9) < 2e-4


@pytest.mark.parametrize("shape", [(4, 8), (2, 8, 7), (2, 8, 3, 5), (1, 8, 2, 3, 4), (1, 8, 2, 2, 3, 2)])
def test_eb_forward_any_rank_vs_oracle(kernels_golden, shape):
    """EB eval: x_hat == round(x - median) + median for 0-D..5-D spatial (tests/test_entropy_models.py:177-220)."""
    g = kernels_golden
    eb = _load_eb(g)
    rs = np.random.RandomState(len(shape))
    x = (rs.standard_normal(shape) * 6).astype(np.float32)
    mats = [g[f"eb_param__matrix{i}"] for i in range(5)]
    bias = [g[f"eb_param__bias{i}"] for i in range(5)]
    fac = [g[f"eb_param__factor{i}"] for i in range(4)]
    med = g["eb_param_quantiles"][:, 0, 1]
    inner = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    rxh, rlik = oracle.eb_forward(x, mats, bias, fac, med, 8, inner)
    xh, lik = eb(cu(x))
    assert xh.shape == tuple(shape)
    assert np.array_equal(xh.cpu().numpy(), rxh)
    assert rel_err(lik.cpu().numpy(), rlik, 1e-9) < 1e-4


def test_eb_forward_wide_channels_last_with_bits():
    """C=192 channels-last (the layout the model path uses), total threads % C handled."""
    C = 192
    from weights import _entropy_bottleneck
    w = {}
    _entropy_bottleneck(np.random.RandomState(9), w, "eb", C)
    eb = mmcodec.EntropyBottleneck(C)
    sd = eb.state_dict()
    for k, v in w.items():
        sd[k[3:]].copy_(torch.from_numpy(v))
    eb = eb.to(dev()).eval()
    rs = np.random.RandomState(2)
    x = (rs.standard_normal((3, C, 9, 11)) * 5).astype(np.float32)
    mats = [w[f"eb._matrix{i}"] for i in range(5)]
    bias = [w[f"eb._bias{i}"] for i in range(5)]
    fac = [w[f"eb._factor{i}"] for i in range(4)]
    rxh, rlik = oracle.eb_forward(x, mats, bias, fac, w["eb.quantiles"][:, 0, 1], C, 99)
    bits = torch.zeros(1, device=dev())
    xh, lik, xb = ops.eb_forward(cu(x).to(memory_format=torch.channels_last), eb._params(), None, 1e-9, want_bf16=True, bits=bits)
    assert np.array_equal(xh.cpu().numpy(), rxh)
    assert rel_err(lik.cpu().numpy(), rlik, 1e-9) < 1e-4
    assert torch.equal(xb.float().cpu(), torch.from_numpy(rxh).to(torch.bfloat16).float())
    assert abs(bits.item() - oracle.bits(rlik)) / oracle.bits(rlik) < 1e-5


# ---- GDN ---------------------------------------------------------------------------------------------------
def test_gdn_golden(kernels_golden):
    g = kernels_golden
    x = cu(g["gdn_x"])
    for inv, key in ((False, "gdn_y"), (True, "gdn_y_inv")):
        m = mmcodec.GDN(16, inverse=inv)
        m.beta.data.copy_(torch.from_numpy(g["gdn_beta"]))
        m.gamma.data.copy_(torch.from_numpy(g["gdn_gamma"]))
        m = m.to(dev())
        y = m(x)
        assert y.grad_fn is not None          # differentiable module call, as in the reference (parameters require grad)
        assert rel_err(y.detach().cpu().numpy(), g[key], 1.0) < 1e-5
        with torch.no_grad():
            ycl = m(x.to(memory_format=torch.channels_last))
        assert rel_err(ycl.cpu().numpy(), g[key], 1.0) < 1e-5
    # closed form at init: y = x / sqrt(1 + 0.1 x^2) (tests/test_layers.py:145-146,158-159)
    xx = cu(g["gdn_x"])
    with torch.no_grad():
        y0 = mmcodec.GDN(16).to(dev())(xx)
    assert torch.allclose(y0, xx / torch.sqrt(1 + 0.1 * xx ** 2), atol=1e-5)
    y1 = mmcodec.GDN(16, inverse=True).to(dev())(xx).detach()
    assert torch.allclose(y1, xx * torch.sqrt(1 + 0.1 * xx ** 2), atol=1e-5)


def test_gdn_wide_vs_oracle():
    from weights import _gdn
    C = 192
    w = {}
    _gdn(np.random.RandomState(1), w, "g", C)
    x = (np.random.RandomState(2).standard_normal((2, C, 7, 13)) * 2).astype(np.float32)
    m = mmcodec.GDN(C)
    m.beta.data.copy_(torch.from_numpy(w["g.beta"]))
    m.gamma.data.copy_(torch.from_numpy(w["g.gamma"]))
    y = m.to(dev())(cu(x)).detach().cpu().numpy()
    assert rel_err(y, oracle.gdn_forward(x, w["g.beta"], w["g.gamma"]), 1e-2) < 1e-4


# ---- conv / deconv (CUDA-core kernel, fp32) ------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_conv_deconv_direct_golden(kernels_golden, tag):
    g = kernels_golden
    k, s = (int(v) for v in g[f"conv_{tag}_cfg"])
    w = g[f"conv_{tag}_w"]
    m = mmcodec.conv(w.shape[1], w.shape[0], kernel_size=k, stride=s)
    m.weight.data.copy_(torch.from_numpy(w)); m.bias.data.copy_(torch.from_numpy(g[f"conv_{tag}_b"]))
    with torch.no_grad():     # outside no_grad the layer records an autograd graph (training path), as the reference's does
        y = m.to(dev())(cu(g[f"conv_{tag}_x"]))
    assert tuple(y.shape) == g[f"conv_{tag}_y"].shape
    assert np.max(np.abs(y.cpu().numpy() - g[f"conv_{tag}_y"])) < 1e-4
    k, s = (int(v) for v in g[f"deconv_{tag}_cfg"])
    w = g[f"deconv_{tag}_w"]
    m = mmcodec.deconv(w.shape[0], w.shape[1], kernel_size=k, stride=s)
    m.weight.data.copy_(torch.from_numpy(w)); m.bias.data.copy_(torch.from_numpy(g[f"deconv_{tag}_b"]))
    with torch.no_grad():
        y = m.to(dev())(cu(g[f"deconv_{tag}_x"]))
    assert tuple(y.shape) == g[f"deconv_{tag}_y"].shape
    assert np.max(np.abs(y.cpu().numpy() - g[f"deconv_{tag}_y"])) < 1e-4


@pytest.mark.parametrize("transposed,cin,cout,k,s,h,w,act,gdn", [
    (False, 3, 24, 5, 2, 19, 23, L.ACT_NONE, L.GDN_FORWARD),     # image-edge layer with fused GDN, ragged size
    (False, 20, 70, 3, 1, 9, 11, L.ACT_RELU, L.GDN_NONE),        # Cout spanning two 64-wide chunks
    (False, 1, 8, 5, 2, 16, 16, L.ACT_LEAKY_RELU, L.GDN_NONE),   # 1-channel (depth/IR) input
    (True, 24, 3, 5, 2, 9, 7, L.ACT_NONE, L.GDN_NONE),           # narrow-output deconv (x_hat)
    (True, 16, 24, 5, 2, 6, 5, L.ACT_NONE, L.GDN_INVERSE),       # deconv + IGDN
    (True, 8, 8, 3, 1, 5, 6, L.ACT_RELU, L.GDN_NONE),
])
def test_conv_direct_vs_oracle(transposed, cin, cout, k, s, h, w, act, gdn):
    from weights import _gdn
    rs = np.random.RandomState(cin * 100 + cout)
    x = rs.standard_normal((2, cin, h, w)).astype(np.float32)
    wt = (rs.standard_normal((cin, cout, k, k) if transposed else (cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32)
    b = rs.standard_normal(cout).astype(np.float32)
    actname = {L.ACT_NONE: None, L.ACT_RELU: "relu", L.ACT_LEAKY_RELU: "leaky_relu"}[act]
    ref = (oracle.conv_transpose2d if transposed else oracle.conv2d)(x, wt, b, stride=s, act=actname)
    beta_eff = gamma_eff = None
    if gdn != L.GDN_NONE:
        gw = {}
        _gdn(rs, gw, "g", cout)
        ref = oracle.gdn_forward(ref, gw["g.beta"], gw["g.gamma"], inverse=(gdn == L.GDN_INVERSE))
        be, ge = oracle.gdn_reparam(gw["g.beta"], gw["g.gamma"])
        beta_eff, gamma_eff = cu(be), cu(ge)
    for in_fmt in ("nchw_f32", "nhwc_bf16"):
        for out_fmt in ("nchw_f32", "nhwc_f32", "nhwc_bf16"):
            xin = cu(x) if in_fmt == "nchw_f32" else cu(x).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
            d = ops.conv_desc(transposed, 2, h, w, cin, cout, k, s, L.F32 if in_fmt == "nchw_f32" else L.BF16,
                              L.NCHW if in_fmt == "nchw_f32" else L.NHWC, L.BF16 if out_fmt.endswith("bf16") else L.F32,
                              L.NCHW if out_fmt == "nchw_f32" else L.NHWC, act=act, gdn=gdn, out2=1)
            y, y2 = ops.conv_forward_direct(d, xin, cu(wt), cu(b), beta_eff, gamma_eff)
            y = y.float() if out_fmt == "nchw_f32" else y.float().permute(0, 3, 1, 2)
            tol = 1e-4 if (in_fmt == "nchw_f32" and out_fmt != "nhwc_bf16") else 2e-2
            scale = float(np.abs(ref).max())
            assert tuple(y.shape) == ref.shape
            assert np.max(np.abs(y.cpu().numpy() - ref)) < tol * scale, (in_fmt, out_fmt)
            assert np.max(np.abs(y2.float().permute(0, 3, 1, 2).cpu().numpy() - np.abs(ref))) < 2e-2 * scale


def test_layout_round_trip():
    x = torch.randn(3, 37, 11, 13, device=dev())
    nhwc = ops.nchw_to_nhwc_bf16(x)
    assert torch.equal(nhwc, x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))
    assert torch.equal(ops.nhwc_bf16_to_nchw(nhwc), x.to(torch.bfloat16).float())
    assert torch.equal(ops.to_bf16(x), x.to(torch.bfloat16))


def test_eb_eval_lut_equals_direct_evaluation(kernels_golden):
    """Eval fast path (per-(channel, symbol) likelihood table) vs direct evaluation of the 5-layer CDF model: identical
    x_hat and likelihoods, including symbols outside the +-64 table (direct fallback), NaN, both memory layouts."""
    eb = _load_eb(kernels_golden)
    rs = np.random.RandomState(4)
    x = (rs.standard_normal((3, 8, 6, 10)) * 30).astype(np.float32)
    x[0, :, 0, 0] = 500.0
    x[1, :, 1, 1] = -1e4
    x[2, 0, 0, 0] = np.nan
    for xt in (cu(x), cu(x).to(memory_format=torch.channels_last)):
        a_hat, a_lik = ops.eb_forward(xt, eb._params(), None, 1e-9)
        bits = torch.zeros(1, device=dev())
        b_hat, b_lik, b_bf = ops.eb_forward(xt, eb._params(), None, 1e-9, want_bf16=True, bits=bits, lut=eb._eval_lut())
        assert torch.equal(a_hat.nan_to_num(7.0), b_hat.nan_to_num(7.0))
        assert torch.equal(a_lik.nan_to_num(7.0), b_lik.nan_to_num(7.0))
        assert torch.equal(b_bf.float().nan_to_num(7.0), b_hat.to(torch.bfloat16).float().nan_to_num(7.0))
    # the table is rebuilt when a parameter changes
    lut0 = eb._eval_lut().clone()
    with torch.no_grad():
        eb._bias0.add_(0.25)
    assert not torch.equal(eb._eval_lut(), lut0)


def test_pmf_to_quantized_cdf_device_bit_exact(kernels_golden):
    """Device CDF construction (one CTA per table row) vs the reference's own outputs and vs the host implementation:
    KAT of tests/test_ops.py:104-106, the GaussianConditional table of the zoo models (64 x 3133, most symbols need a stolen
    count), random pmfs with zeros; domain errors as in ops.cpp:46-52."""
    g = kernels_golden
    rs = np.random.RandomState(4)
    rows = [np.array([0.1, 0.2, 0.0, 0.0], dtype=np.float32)] + [np.asarray(p_[:n_], dtype=np.float32) for p_, n_ in zip(g["cdf_pmfs"], g["cdf_lens"])]
    for n in (3, 17, 200, 1500):
        p = rs.dirichlet(np.full(n, 0.05)).astype(np.float32)
        p[rs.rand(n) < 0.4] = 0.0
        p[rs.randint(n)] = max(float(p.max()), 0.5)
        rows.append(p)
    max_len = max(len(r) for r in rows) - 1
    pmf = np.zeros((len(rows), max_len), np.float32)
    tail = np.zeros(len(rows), np.float32)
    lens = np.zeros(len(rows), np.int32)
    for i, r in enumerate(rows):
        pmf[i, : len(r) - 1], tail[i], lens[i] = r[:-1], r[-1], len(r) - 1
    got = ops.pmf_to_quantized_cdf_device(cu(pmf), cu(tail), torch.from_numpy(lens).to(dev()), max_len, 16).cpu().numpy()
    for i, r in enumerate(rows):
        want = np.array(ops.pmf_to_quantized_cdf(r, 16), dtype=np.int64)
        assert np.array_equal(got[i, : len(want)], want), i
        assert not got[i, len(want):].any()
    assert np.array_equal(got[0, :5], g["cdf_kat"].astype(np.int64))                     # the reference's known-answer test
    for i, (c_, n_) in enumerate(zip(g["cdf_cdfs"], g["cdf_lens"])):                     # the reference's own outputs
        assert np.array_equal(got[1 + i, : n_ + 1], c_[: n_ + 1].astype(np.int64)), i
    # the zoo's GaussianConditional tables: update() on the device reproduces the CPU model's tables bit for bit
    gc_cpu = mmcodec.GaussianConditional(None)
    gc_cpu.update_scale_table(mmcodec.get_scale_table())
    gc_dev = mmcodec.GaussianConditional(None).to(dev())
    gc_dev.update_scale_table(mmcodec.get_scale_table())
    assert torch.equal(gc_dev._quantized_cdf.cpu(), gc_cpu._quantized_cdf) and torch.equal(gc_dev._cdf_length.cpu(), gc_cpu._cdf_length)
    with pytest.raises(ValueError):
        ops.pmf_to_quantized_cdf_device(cu(np.array([[0.5, -0.1]], np.float32)), cu(np.array([0.1], np.float32)),
                                        torch.tensor([2], dtype=torch.int32, device=dev()), 2, 16)
    with pytest.raises(ValueError):
        ops.pmf_to_quantized_cdf_device(cu(np.zeros((1, 3), np.float32)), cu(np.zeros(1, np.float32)),
                                        torch.tensor([3], dtype=torch.int32, device=dev()), 3, 16)


def test_colour_transforms_vs_reference_golden():
    """mmcodec.transforms_functional (compressai/transforms/functional.py:26-137) vs the reference's own outputs; fp32, 1e-6."""
    import os
    from mmcodec import transforms_functional as TF
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "color.npz"))
    rgb = cu(g["rgb"])
    ycc = TF.rgb2ycbcr(rgb)
    assert np.abs(ycc.cpu().numpy() - g["ycbcr"]).max() < 1e-6
    assert np.abs(TF.ycbcr2rgb(cu(g["ycbcr"])).cpu().numpy() - g["rgb_back"]).max() < 1e-6
    y, u, v = TF.yuv_444_to_420(cu(g["ycbcr"]))
    assert np.abs(u.cpu().numpy() - g["u420"]).max() < 1e-6 and np.abs(v.cpu().numpy() - g["v420"]).max() < 1e-6
    assert np.array_equal(y.cpu().numpy(), g["ycbcr"][:, :1])
    back = TF.yuv_420_to_444((y, cu(g["u420"]), cu(g["v420"])))
    assert np.abs(back.cpu().numpy() - g["yuv444"]).max() < 1e-6
    t3 = TF.yuv_420_to_444((y, cu(g["u420"]), cu(g["v420"])), return_tuple=True)
    assert len(t3) == 3 and tuple(t3[1].shape) == tuple(y.shape)
    with pytest.raises(ValueError):
        TF.rgb2ycbcr(torch.zeros(2, 4, 8, 8, device=dev()))
    with pytest.raises(ValueError):
        TF.yuv_444_to_420(cu(g["ycbcr"]), mode="nearest")
    with pytest.raises(NotImplementedError):
        TF.yuv_420_to_444((y, u, v), mode="bicubic")


def test_esa_glue_kernels_vs_torch():
    """csrc/esa.cu against the torch ops of ESA.forward (google.py:1449-1459) on the same bf16 maps: max_pool2d(7, 3) is exact;
    bilinear upsample + add and the sigmoid gate are computed in fp32 and rounded once."""
    import torch.nn.functional as F
    from mmcodec import ops
    gen = torch.Generator().manual_seed(12)
    dev = torch.device("cuda", 0)
    for (B, H, W, C) in ((2, 31, 47, 48), (1, 7, 7, 2), (3, 64, 96, 48)):
        x = torch.randn(B, H, W, C, generator=gen).to(dev).bfloat16()
        ref = F.max_pool2d(x.permute(0, 3, 1, 2).float(), kernel_size=7, stride=3).permute(0, 2, 3, 1)
        out = ops.maxpool_nhwc_bf16(x, 7, 3)
        assert tuple(out.shape) == tuple(ref.shape) and torch.equal(out.float(), ref)
    xs = torch.randn(1, 9, 9, 4, generator=gen).to(dev).bfloat16()
    xs[0, 3, 3, 1] = float("nan")
    assert bool(torch.isnan(ops.maxpool_nhwc_bf16(xs, 7, 3)[0, 0, 0, 1])) and not bool(torch.isnan(ops.maxpool_nhwc_bf16(xs, 7, 3)[0, 0, 0, 0]))
    with pytest.raises((ValueError, mmcodec.MmcodecError)):
        ops.maxpool_nhwc_bf16(torch.zeros(1, 5, 9, 4, device=dev, dtype=torch.bfloat16), 7, 3)
    for (B, hs, ws, H, W, C) in ((2, 5, 8, 64, 96, 48), (1, 1, 1, 16, 16, 48), (1, 13, 20, 128, 192, 48), (1, 4, 4, 9, 11, 6)):
        small = torch.randn(B, hs, ws, C, generator=gen).to(dev).bfloat16()
        add = torch.randn(B, H, W, C, generator=gen).to(dev).bfloat16()
        ref = F.interpolate(small.permute(0, 3, 1, 2).float(), (H, W), mode="bilinear", align_corners=False).permute(0, 2, 3, 1) + add.float()
        out = ops.upsample_bilinear_add_bf16(small, add)
        assert float((out.float() - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max()) + 1e-6
    x = (3 * torch.randn(2, 17, 19, 192, generator=gen)).to(dev).bfloat16()
    g = (4 * torch.randn(2, 17, 19, 192, generator=gen)).to(dev).bfloat16()
    ref = x.float() * torch.sigmoid(g.float())
    out = ops.sigmoid_gate_bf16(x, g)
    assert float((out.float() - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max())
    xt, gt = x.flatten()[:13].contiguous(), g.flatten()[:13].contiguous()            # ragged tail (n % 8 != 0)
    assert float((ops.sigmoid_gate_bf16(xt, gt).float() - xt.float() * torch.sigmoid(gt.float())).abs().max()) <= 2 ** -8 * float(ref.abs().max())

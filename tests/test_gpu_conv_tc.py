"""GPU parity tests of the tcgen05 implicit-GEMM conv / deconv kernel (mmc_conv_forward_tc) against the
CPU oracle.  Inputs and weights are rounded to bf16 first, so the only differences left are fp32
accumulation order, the bf16 x^2 / gamma of the fused GDN, and output rounding: tolerance 2e-3 of
the output scale for fp32 outputs, 1e-2 for bf16 outputs (the bf16 tolerance BASELINE.json states)."""
import numpy as np
import pytest
import torch

import oracle
from weights import _gdn

pytestmark = pytest.mark.gpu

from mmcodec import _lib as L  # noqa: E402
from mmcodec import ops  # noqa: E402


def dev():
    return torch.device("cuda", 0)


def bf16_round(a):
    return torch.from_numpy(a).to(torch.bfloat16).float().numpy()


CASES = [
    # transposed, cin, cout, k, s, h, w, act, gdn, B
    (False, 64, 64, 1, 1, 8, 16, L.ACT_NONE, L.GDN_NONE, 1),        # plain GEMM, exactly one tile
    (False, 64, 128, 3, 1, 16, 24, L.ACT_RELU, L.GDN_NONE, 2),      # h_a.0 / h_s.4 shape class
    (False, 128, 128, 5, 2, 32, 48, L.ACT_NONE, L.GDN_NONE, 2),     # g_a.2 class (stride-2 element-strided TMA box)
    (False, 128, 128, 5, 2, 32, 48, L.ACT_NONE, L.GDN_FORWARD, 2),  # + fused GDN
    (False, 128, 192, 5, 2, 16, 24, L.ACT_NONE, L.GDN_NONE, 1),     # g_a.6 class (N=192)
    (False, 192, 320, 5, 2, 34, 60, L.ACT_NONE, L.GDN_NONE, 1),     # mbt2018-mean g_a.6: two N blocks, ragged 17x30 output
    (False, 192, 192, 5, 2, 20, 28, L.ACT_NONE, L.GDN_FORWARD, 1),  # C=192 GDN (two 96-column norm chunks)
    (False, 64, 64, 5, 2, 20, 28, L.ACT_NONE, L.GDN_FORWARD, 2),    # C=64 GDN (x^2 operand through TMEM, 32 columns)
    (True, 64, 64, 3, 1, 9, 11, L.ACT_NONE, L.GDN_INVERSE, 1),      # C=64 IGDN on a stride-1 transposed conv
    (False, 128, 128, 5, 2, 17, 23, L.ACT_LEAKY_RELU, L.GDN_NONE, 2),   # odd input size
    (True, 192, 128, 5, 2, 8, 12, L.ACT_NONE, L.GDN_INVERSE, 2),    # g_s.0 class: 4-phase deconv + IGDN
    (True, 128, 128, 5, 2, 16, 24, L.ACT_RELU, L.GDN_NONE, 1),      # h_s class
    (True, 128, 128, 5, 2, 9, 7, L.ACT_NONE, L.GDN_INVERSE, 1),     # ragged deconv
    (True, 64, 64, 3, 1, 6, 10, L.ACT_NONE, L.GDN_NONE, 1),         # stride-1 transposed conv
    (False, 480, 640, 3, 1, 17, 30, L.ACT_NONE, L.GDN_NONE, 1),     # mbt2018-mean h_s.4: 4 N blocks of 160
]


@pytest.mark.parametrize("transposed,cin,cout,k,s,h,w,act,gdn,B", CASES)
@pytest.mark.parametrize("out_f32", [True, False])
def test_conv_tc_vs_oracle(transposed, cin, cout, k, s, h, w, act, gdn, B, out_f32):
    rs = np.random.RandomState(cin + 7 * cout + k + h)
    x = bf16_round(rs.standard_normal((B, cin, h, w)).astype(np.float32))
    fan = cin * k * k / (s * s if transposed else 1)
    wt = bf16_round((rs.standard_normal((cin, cout, k, k) if transposed else (cout, cin, k, k)) * (2.0 / np.sqrt(fan))).astype(np.float32))
    b = rs.standard_normal(cout).astype(np.float32)
    actname = {L.ACT_NONE: None, L.ACT_RELU: "relu", L.ACT_LEAKY_RELU: "leaky_relu"}[act]
    ref = (oracle.conv_transpose2d if transposed else oracle.conv2d)(x, wt, b, stride=s, act=actname)
    beta_eff = gamma_bf16 = None
    if gdn != L.GDN_NONE:
        gw = {}
        _gdn(rs, gw, "g", cout)
        ref = oracle.gdn_forward(ref, gw["g.beta"], gw["g.gamma"], inverse=(gdn == L.GDN_INVERSE))
        beta_eff, _, gamma_bf16 = ops.gdn_reparam(torch.from_numpy(gw["g.beta"]).to(dev()), torch.from_numpy(gw["g.gamma"]).to(dev()),
                                                  oracle.gdn_beta_bound(), oracle.GDN_GAMMA_BOUND, oracle.GDN_PEDESTAL, want_bf16=True)
    xin = torch.from_numpy(x).to(dev()).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    d = ops.conv_desc(transposed, B, h, w, cin, cout, k, s, L.BF16, L.NHWC, L.F32 if out_f32 else L.BF16, L.NHWC,
                      act=act, gdn=gdn, out2=1)
    packed = ops.conv_pack_weights(d, torch.from_numpy(wt).to(dev()))
    y, y2 = ops.conv_forward_tc(d, xin, packed, torch.from_numpy(b).to(dev()), beta_eff, gamma_bf16)
    torch.cuda.synchronize()
    y = y.float().permute(0, 3, 1, 2).cpu().numpy()
    assert y.shape == ref.shape
    scale = float(np.abs(ref).max())
    err = np.abs(y - ref)
    tol = (2e-3 if gdn == L.GDN_NONE else 6e-3) if out_f32 else 1e-2
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() < tol * scale, f"max err {err.max():.4g} (scale {scale:.4g}) at {worst}: got {y[worst]:.5g} want {ref[worst]:.5g}; " \
                                    f"frac bad {(err > tol * scale).mean():.4f}"
    e2 = np.abs(y2.float().permute(0, 3, 1, 2).cpu().numpy() - np.abs(ref))
    assert e2.max() < 1e-2 * scale


def test_conv_tc_matches_direct_kernel_at_model_scale():
    """g_a.2-sized layer on a batch: tensor-core kernel vs the CUDA-core kernel on the same bf16 data."""
    torch.manual_seed(0)
    B, cin, cout, h, w = 4, 128, 128, 128, 192
    x = torch.randn(B, h, w, cin, device=dev()).to(torch.bfloat16)
    wt = (torch.randn(cout, cin, 5, 5, device=dev()) * 0.02).to(torch.bfloat16).float()
    b = torch.randn(cout, device=dev())
    d = ops.conv_desc(False, B, h, w, cin, cout, 5, 2, L.BF16, L.NHWC, L.F32, L.NHWC)
    y_tc = ops.conv_forward_tc(d, x, ops.conv_pack_weights(d, wt), b)
    y_dc = ops.conv_forward_direct(d, x, wt, b)
    torch.cuda.synchronize()
    assert float((y_tc - y_dc).abs().max()) < 2e-3 * float(y_dc.abs().max())


def test_conv_tc_rejects_unsupported_shapes():
    d = ops.conv_desc(False, 1, 8, 8, 3, 128, 5, 2, L.BF16, L.NHWC, L.BF16, L.NHWC)
    with pytest.raises(NotImplementedError):
        ops.conv_pack_weights(d, torch.zeros(128, 3, 5, 5, device=dev()))


@pytest.mark.parametrize("cin,cout,k,s,h,w,gdn,B", [
    (3, 128, 5, 2, 64, 96, L.GDN_FORWARD, 2),    # g_a.0 of every image model (+GDN)
    (3, 192, 5, 2, 37, 53, L.GDN_FORWARD, 1),    # N=192, odd image size
    (1, 64, 5, 2, 32, 48, L.GDN_NONE, 2),        # 1-channel depth / IR input
    (6, 128, 5, 2, 32, 64, L.GDN_NONE, 1),       # ssf2020 motion encoder input (two stacked frames)
    (3, 64, 3, 1, 20, 28, L.GDN_NONE, 1),        # 3x3 stride-1 edge conv
])
def test_conv_tc_image_edge_input(cin, cout, k, s, h, w, gdn, B):
    """Cin <= 8 convolution on the tensor cores through the zero-padded NHWC8 staging buffer."""
    rs = np.random.RandomState(cin + cout + h)
    x = bf16_round(rs.uniform(0, 1, (B, cin, h, w)).astype(np.float32))
    wt = bf16_round((rs.standard_normal((cout, cin, k, k)) * (2.0 / np.sqrt(cin * k * k))).astype(np.float32))
    b = rs.standard_normal(cout).astype(np.float32)
    ref = oracle.conv2d(x, wt, b, stride=s)
    beta_eff = gamma_bf16 = None
    if gdn != L.GDN_NONE:
        gw = {}
        _gdn(rs, gw, "g", cout)
        ref = oracle.gdn_forward(ref, gw["g.beta"], gw["g.gamma"])
        beta_eff, _, gamma_bf16 = ops.gdn_reparam(torch.from_numpy(gw["g.beta"]).to(dev()), torch.from_numpy(gw["g.gamma"]).to(dev()),
                                                  oracle.gdn_beta_bound(), oracle.GDN_GAMMA_BOUND, oracle.GDN_PEDESTAL, want_bf16=True)
    d = ops.conv_desc(False, B, h, w, cin, cout, k, s, L.BF16, L.NHWC_PAD8, L.F32, L.NHWC, gdn=gdn)
    xp = ops.pad_to_nhwc8(torch.from_numpy(x).to(dev()), d)
    pad = k // 2
    assert xp.shape[3] == 8 and torch.equal(xp[:, pad:pad + h, pad:pad + w, :cin].float().cpu(), torch.from_numpy(x).permute(0, 2, 3, 1))
    assert float(xp[:, :pad].abs().max()) == 0 and float(xp[..., cin:].abs().max() if cin < 8 else 0) == 0
    y = ops.conv_forward_tc(d, xp, ops.conv_pack_weights(d, torch.from_numpy(wt).to(dev())), torch.from_numpy(b).to(dev()), beta_eff, gamma_bf16)
    torch.cuda.synchronize()
    y = y.permute(0, 3, 1, 2).cpu().numpy()
    assert y.shape == ref.shape
    scale = float(np.abs(ref).max())
    err = np.abs(y - ref)
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() < (2e-3 if gdn == L.GDN_NONE else 6e-3) * scale, f"max err {err.max():.4g} (scale {scale:.4g}) at {worst}; frac bad {(err > 6e-3 * scale).mean():.4f}"


@pytest.mark.parametrize("cin,cout,k,h,w,B", [(128, 3, 5, 16, 24, 2), (192, 3, 5, 9, 7, 1), (64, 1, 5, 8, 16, 1), (128, 3, 3, 10, 12, 1)])
def test_conv_tc_narrow_transposed_output(cin, cout, k, h, w, B):
    """Reconstruction layer (deconv N -> 3, stride 2): the four output phases stacked along N, planar fp32 out."""
    rs = np.random.RandomState(cin + cout + h)
    x = bf16_round(rs.standard_normal((B, cin, h, w)).astype(np.float32))
    wt = bf16_round((rs.standard_normal((cin, cout, k, k)) * (2.0 / np.sqrt(cin * k * k / 4))).astype(np.float32))
    b = rs.standard_normal(cout).astype(np.float32)
    ref = oracle.conv_transpose2d(x, wt, b, stride=2)
    xin = torch.from_numpy(x).to(dev()).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    d = ops.conv_desc(True, B, h, w, cin, cout, k, 2, L.BF16, L.NHWC, L.F32, L.NCHW)
    y = ops.conv_forward_tc(d, xin, ops.conv_pack_weights(d, torch.from_numpy(wt).to(dev())), torch.from_numpy(b).to(dev()))
    torch.cuda.synchronize()
    y = y.cpu().numpy()
    assert y.shape == ref.shape
    scale = float(np.abs(ref).max())
    err = np.abs(y - ref)
    worst = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() < 2e-3 * scale, f"max err {err.max():.4g} (scale {scale:.4g}) at {worst}; frac bad {(err > 2e-3 * scale).mean():.4f}"


@pytest.mark.parametrize("transposed,cin,h,w,B,gdn", [
    (False, 128, 200, 312, 3, L.GDN_FORWARD),    # g_a.2 class: 100 x 156 outputs per image -> 7 x 10 tiles x 3 images (odd tile count per phase)
    (True, 192, 50, 78, 2, L.GDN_INVERSE),       # g_s.0 class: 4 deconv phases
])
def test_conv_tc_cta_pair_vs_single(transposed, cin, h, w, B, gdn):
    """CTA-pair kernel (cta_group::2, two tiles per MMA, weight / gamma rows split across the pair) against the single-CTA
    kernel on the same data (bit-identical: same products, same accumulation order) and against the CPU oracle."""
    import os
    rs = np.random.RandomState(h + w)
    cout, k, s = 128, 5, 2
    x = bf16_round(rs.standard_normal((B, cin, h, w)).astype(np.float32))
    fan = cin * k * k / (s * s if transposed else 1)
    wt = bf16_round((rs.standard_normal((cin, cout, k, k) if transposed else (cout, cin, k, k)) * (2.0 / np.sqrt(fan))).astype(np.float32))
    b = rs.standard_normal(cout).astype(np.float32)
    gw = {}
    _gdn(rs, gw, "g", cout)
    beta_eff, _, gamma_bf16 = ops.gdn_reparam(torch.from_numpy(gw["g.beta"]).to(dev()), torch.from_numpy(gw["g.gamma"]).to(dev()),
                                              oracle.gdn_beta_bound(), oracle.GDN_GAMMA_BOUND, oracle.GDN_PEDESTAL, want_bf16=True)
    xin = torch.from_numpy(x).to(dev()).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    d = ops.conv_desc(transposed, B, h, w, cin, cout, k, s, L.BF16, L.NHWC, L.BF16, L.NHWC, gdn=gdn)
    packed = ops.conv_pack_weights(d, torch.from_numpy(wt).to(dev()))
    outs = {}
    for mode in ("2", "0"):      # 2 = pair kernel whenever the shape allows it, 0 = single-CTA kernel
        os.environ["MMC_TC_PAIR"] = mode
        try:
            outs[mode] = ops.conv_forward_tc(d, xin, packed, torch.from_numpy(b).to(dev()), beta_eff, gamma_bf16)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("MMC_TC_PAIR", None)
    assert torch.equal(outs["2"], outs["0"])
    torch.set_num_threads(8)
    ref = (oracle.conv_transpose2d if transposed else oracle.conv2d)(x, wt, b, stride=s, act=None)
    ref = oracle.gdn_forward(ref, gw["g.beta"], gw["g.gamma"], inverse=(gdn == L.GDN_INVERSE))
    y = outs["2"].float().permute(0, 3, 1, 2).cpu().numpy()
    assert np.abs(y - ref).max() < 1e-2 * float(np.abs(ref).max())


@pytest.mark.parametrize("transposed,cin,cout,k,s,gdn,h,w", [
    (True, 128, 128, 5, 2, L.GDN_INVERSE, 20, 28),    # g_s.2 / g_s.4 class: four phases, 3 / 2 / 3 / 2 groups, teams epilogue
    (True, 192, 128, 5, 2, L.GDN_INVERSE, 9, 13),     # g_s.0 class: three K chunks, ragged tiles
    (False, 128, 128, 5, 2, L.GDN_FORWARD, 40, 56),   # g_a.2 class: strided conv, parity groups (pair kernel when forced)
    (False, 192, 128, 3, 1, L.GDN_NONE, 17, 24),      # h_a.0 class: 3x3 stride 1, plain epilogue
    (False, 64, 64, 5, 1, L.GDN_FORWARD, 24, 24),     # C = 64 teams, five-tap groups (20-row patches)
])
def test_opt_in_main_loop_and_epilogue_variants(transposed, cin, cout, k, s, gdn, h, w):
    """The opt-in round-2 paths -- tap groups (MMC_TC_GROUPED=1: one A patch per group and chunk) and the two-team GDN epilogue
    (MMC_TC_TEAMS=2: norm in place, x in registers) -- against the default kernel path and the oracle.  Teams are bit-identical
    to the single-team epilogue; groups change only the fp32 accumulation order of the taps."""
    import os
    rs = np.random.RandomState(cin + cout + k)
    B = 3
    x = bf16_round(rs.standard_normal((B, cin, h, w)).astype(np.float32))
    fan = cin * k * k / (s * s if transposed else 1)
    wt = bf16_round((rs.standard_normal((cin, cout, k, k) if transposed else (cout, cin, k, k)) * (2.0 / np.sqrt(fan))).astype(np.float32))
    b = rs.standard_normal(cout).astype(np.float32)
    gw = {}
    _gdn(rs, gw, "g", cout)
    beta_eff = gamma_bf16 = None
    if gdn != L.GDN_NONE:
        beta_eff, _, gamma_bf16 = ops.gdn_reparam(torch.from_numpy(gw["g.beta"]).to(dev()), torch.from_numpy(gw["g.gamma"]).to(dev()),
                                                  oracle.gdn_beta_bound(), oracle.GDN_GAMMA_BOUND, oracle.GDN_PEDESTAL, want_bf16=True)
    xin = torch.from_numpy(x).to(dev()).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    d = ops.conv_desc(transposed, B, h, w, cin, cout, k, s, L.BF16, L.NHWC, L.BF16, L.NHWC, gdn=gdn)
    packed = ops.conv_pack_weights(d, torch.from_numpy(wt).to(dev()))
    outs = {}
    # "default" here = the single-team epilogue WITHOUT the two-tile software pipeline (MMC_TC_EPI_PIPE=0), which keeps x in fp32
    # registers like the teams epilogue does; "pipe" = the round-end default (x held as bf16 pairs across the norm contraction)
    nopipe = {"MMC_TC_EPI_PIPE": "0"}
    for name, env in (("default", dict(nopipe)), ("teams", {"MMC_TC_TEAMS": "2", **nopipe}), ("grouped", {"MMC_TC_GROUPED": "1", **nopipe}),
                      ("both", {"MMC_TC_GROUPED": "1", "MMC_TC_TEAMS": "2", **nopipe}),
                      ("both+pair", {"MMC_TC_GROUPED": "1", "MMC_TC_TEAMS": "2", "MMC_TC_PAIR": "2", **nopipe}),
                      ("pipe", {"MMC_TC_EPI_PIPE": "1"}), ("pipe+pair", {"MMC_TC_EPI_PIPE": "1", "MMC_TC_PAIR": "2"}),
                      ("pipe+grouped", {"MMC_TC_EPI_PIPE": "1", "MMC_TC_GROUPED": "1"})):
        os.environ.update(env)
        try:
            outs[name] = ops.conv_forward_tc(d, xin, packed, torch.from_numpy(b).to(dev()), beta_eff, gamma_bf16).float()
            torch.cuda.synchronize()
        finally:
            for k_ in env:
                os.environ.pop(k_, None)
    assert torch.equal(outs["teams"], outs["default"])
    ref = (oracle.conv_transpose2d if transposed else oracle.conv2d)(x, wt, b, stride=s, act=None)
    if gdn != L.GDN_NONE:
        ref = oracle.gdn_forward(ref, gw["g.beta"], gw["g.gamma"], inverse=(gdn == L.GDN_INVERSE))
    for name, y in outs.items():
        yy = y.permute(0, 3, 1, 2).cpu().numpy()
        assert np.abs(yy - ref).max() < 1e-2 * float(np.abs(ref).max()), name
    assert float((outs["grouped"] - outs["default"]).abs().max()) <= 2e-2 * float(outs["default"].abs().max())
    assert torch.equal(outs["both"], outs["grouped"])
    # the pipelined epilogue rounds x to bf16 before the scaling: at most one bf16 step (2^-8 relative) from the fp32-x result
    assert torch.equal(outs["pipe"], outs["pipe+pair"])
    assert float((outs["pipe+grouped"] - outs["pipe"]).abs().max()) <= 2e-2 * float(outs["pipe"].abs().max())   # tap order differs
    if gdn != L.GDN_NONE:
        assert float(((outs["pipe"] - outs["default"]).abs() / outs["default"].abs().clamp_min(1e-3)).max()) <= 2.0 ** -7
    else:
        assert torch.equal(outs["pipe"], outs["default"])


@pytest.mark.parametrize("kind,c1,c2,cout,k,s,gdn", [
    ("conv", 192, 192, 192, 5, 1, L.GDN_NONE),       # tran_conv: eg_ext(own) ++ eg_ext(guide)
    ("conv", 192, 192, 192, 5, 2, L.GDN_FORWARD),    # pic2_g_a_conv2 + GDN on (a ++ fused)
    ("deconv", 192, 192, 192, 5, 2, L.GDN_INVERSE),  # pic2_g_s_conv2 + IGDN
    ("deconv", 192, 192, 1, 5, 2, L.GDN_NONE),       # pic2_g_s_conv4: reconstruction layer (GEMM + col2im)
    ("conv", 384, 384, 640, 1, 1, L.GDN_NONE),       # entropy_parameters.0 on (params ++ ctx)
    ("conv", 64, 40, 48, 3, 1, L.GDN_NONE),          # ragged second source
])
def test_two_source_conv_equals_concatenation(kind, c1, c2, cout, k, s, gdn):
    """mmc_conv_forward_tc2: channels from two NHWC tensors, bit-identical to the convolution of their torch.cat."""
    from mmcodec.layers import GDN, conv, deconv
    from mmcodec.transforms import run_layers
    torch.manual_seed(c1 + cout)
    m = (conv if kind == "conv" else deconv)(c1 + c2, cout, kernel_size=k, stride=s).to(dev())
    layers = [m] + ([GDN(cout, inverse=(gdn == L.GDN_INVERSE)).to(dev())] if gdn != L.GDN_NONE else [])
    x1 = torch.randn(2, 12, 20, c1, device=dev()).to(torch.bfloat16)
    x2 = torch.randn(2, 12, 20, c2, device=dev()).to(torch.bfloat16)
    out_fmt = "nchw_f32" if cout <= 4 else "nhwc_bf16"
    with torch.no_grad():
        want = run_layers(layers, torch.cat((x1, x2), dim=-1), "nhwc_bf16", out_fmt)
        got = run_layers(layers, (x1, x2), "nhwc_bf16", out_fmt)
    assert torch.equal(got, want)


@pytest.mark.parametrize("cin,cout,gdn,h,w,B,pair", [
    (128, 128, L.GDN_INVERSE, 40, 56, 3, "0"),     # single-CTA kernel, ragged tiles
    (32, 128, L.GDN_INVERSE, 64, 96, 4, "2"),      # CTA-pair kernel, 192 spatial tiles: more than one wave of 148
    (192, 128, L.GDN_NONE, 17, 23, 2, "0"),        # plain epilogue, odd sizes
])
def test_phase_inner_tile_order_is_bit_identical(cin, cout, gdn, h, w, B, pair):
    """MMC_TC_PHASE_INNER=1 (the stride^2 output phases of a wave of spatial tiles run back to back: the input is read from HBM once)
    only reorders the tiles of a transposed convolution: outputs are bit-identical to the phase-major order, and within the bf16
    tolerance of the oracle.  The default rule switches it on for layers of many waves only, so it is forced here."""
    import os
    rs = np.random.RandomState(cin + cout + h)
    k, s = 5, 2
    x = bf16_round(rs.standard_normal((B, cin, h, w)).astype(np.float32))
    wt = bf16_round((rs.standard_normal((cin, cout, k, k)) * (2.0 / np.sqrt(cin * k * k / (s * s)))).astype(np.float32))
    b = rs.standard_normal(cout).astype(np.float32)
    gw = {}
    _gdn(rs, gw, "g", cout)
    beta_eff = gamma_bf16 = None
    if gdn != L.GDN_NONE:
        beta_eff, _, gamma_bf16 = ops.gdn_reparam(torch.from_numpy(gw["g.beta"]).to(dev()), torch.from_numpy(gw["g.gamma"]).to(dev()),
                                                  oracle.gdn_beta_bound(), oracle.GDN_GAMMA_BOUND, oracle.GDN_PEDESTAL, want_bf16=True)
    xin = torch.from_numpy(x).to(dev()).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    d = ops.conv_desc(True, B, h, w, cin, cout, k, s, L.BF16, L.NHWC, L.BF16, L.NHWC, gdn=gdn)
    packed = ops.conv_pack_weights(d, torch.from_numpy(wt).to(dev()))
    outs = {}
    for inner in ("0", "1"):
        env = {"MMC_TC_PHASE_INNER": inner, "MMC_TC_PAIR": pair}
        os.environ.update(env)
        try:
            outs[inner] = ops.conv_forward_tc(d, xin, packed, torch.from_numpy(b).to(dev()), beta_eff, gamma_bf16).float()
            torch.cuda.synchronize()
        finally:
            for k_ in env:
                os.environ.pop(k_, None)
    assert torch.equal(outs["0"], outs["1"])
    ref = oracle.conv_transpose2d(x, wt, b, stride=s, act=None)
    if gdn != L.GDN_NONE:
        ref = oracle.gdn_forward(ref, gw["g.beta"], gw["g.gamma"], inverse=True)
    yy = outs["1"].permute(0, 3, 1, 2).cpu().numpy()
    assert np.abs(yy - ref).max() < 1e-2 * float(np.abs(ref).max())


@pytest.mark.parametrize("teams", ["2", "3", "4"])
def test_reconstruction_layer_team_counts(teams):
    """The col2im epilogue of the reconstruction layer with 2 / 3 / 4 teams of four warps (MMC_TC_SCATTER_TEAMS; default 3): every
    team count gives the same bits (each output pixel is summed by one thread in a fixed order), at more than one wave of tiles."""
    import os
    rs = np.random.RandomState(11)
    B, cin, cout, k, s, h, w = 3, 128, 3, 5, 2, 70, 90           # 12 x 7 x 3 = 252 overlapping tiles, ragged right / bottom edges
    x = bf16_round(rs.standard_normal((B, cin, h, w)).astype(np.float32))
    wt = bf16_round((rs.standard_normal((cin, cout, k, k)) * (2.0 / np.sqrt(cin * k * k / 4))).astype(np.float32))
    b = rs.standard_normal(cout).astype(np.float32)
    xin = torch.from_numpy(x).to(dev()).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    d = ops.conv_desc(True, B, h, w, cin, cout, k, s, L.BF16, L.NHWC, L.F32, L.NCHW)
    packed = ops.conv_pack_weights(d, torch.from_numpy(wt).to(dev()))
    outs = {}
    for t in ("default", teams):
        if t != "default":
            os.environ["MMC_TC_SCATTER_TEAMS"] = t
        try:
            outs[t] = ops.conv_forward_tc(d, xin, packed, torch.from_numpy(b).to(dev()))
            torch.cuda.synchronize()
        finally:
            os.environ.pop("MMC_TC_SCATTER_TEAMS", None)
    assert torch.equal(outs["default"], outs[teams])
    ref = oracle.conv_transpose2d(x, wt, b, stride=s, act=None)
    assert np.abs(outs[teams].cpu().numpy() - ref).max() < 2e-3 * float(np.abs(ref).max())

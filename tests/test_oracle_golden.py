"""Pins the CPU oracle (oracle/oracle.c + oracle/torch_port.py) to the reference's own outputs
(tests/golden/*.npz, produced by tests/golden/gen_golden.py from /root/reference/CompressAI) and
to the reference's known-answer tests.  CPU only."""
import numpy as np
import pytest
import torch

import oracle
from oracle import torch_port as tp
from weights import make_state_dict


def rel_err(a, b, floor):
    """max |a-b| / max(|b|, floor), after discounting 4 ulp(0.5) = 2.4e-7 of absolute error: the
    likelihoods are differences of two CDF values of magnitude <= 1/2 (entropy_models.py:705-707,
    :487-491), so the reference's own fp32 result carries that much cancellation noise."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.maximum(np.abs(a - b) - 2.4e-7, 0.0) / np.maximum(np.abs(b), floor)))


def test_quantize_kat(kernels_golden):
    g = kernels_golden
    # reference KAT, SURVEY.md Appendix C / tests/test_entropy_models.py:74-79
    assert np.array_equal(oracle.quantize_symbols(g["q_kat_x"]), g["q_kat_sym"])
    assert np.array_equal(oracle.quantize_symbols(np.array([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, -0.0], np.float32)),
                          np.array([0, 2, 2, 0, -2, -2, 0], np.int32))


def test_quantize_dequantize(kernels_golden):
    g = kernels_golden
    x, m = g["q_x"], g["q_means"]
    assert np.array_equal(oracle.quantize_symbols(x, m), g["q_sym_means"])
    assert np.array_equal(oracle.quantize_dequantize(x, m), g["q_deq_means"])
    assert np.array_equal(oracle.quantize_symbols(x), g["q_sym_nomeans"])
    assert np.array_equal(oracle.quantize_dequantize(x), g["q_deq_nomeans"])
    C, inner = x.shape[1], x.shape[2] * x.shape[3]
    assert np.array_equal(oracle.quantize_symbols(x, g["q_chmeans"], C=C, inner=inner), g["q_sym_chmeans"])
    assert np.array_equal(oracle.dequantize(g["q_sym_means"], m), g["dq_means"])
    assert np.array_equal(oracle.dequantize(g["q_sym_means"]), g["dq_nomeans"])


def test_build_indexes(kernels_golden):
    g = kernels_golden
    assert np.array_equal(oracle.build_indexes(g["bi_scales"], g["scale_table"]), g["bi_indexes"])
    # SURVEY.md Appendix C vector
    t = g["scale_table"]
    s = np.array([-1, 0, 0.11, t[0], t[1], 1.0, t[62], 200, 256, 1e6, np.nan], np.float32)
    assert oracle.build_indexes(s, t).tolist() == [0, 0, 0, 0, 1, 18, 62, 61, 63, 63, 63]
    # torch port agrees too
    assert np.array_equal(tp.build_indexes(torch.from_numpy(g["bi_scales"]), torch.from_numpy(t)).numpy(), g["bi_indexes"])
    assert np.array_equal(tp.get_scale_table().numpy(), t)


def test_channel_indexes(kernels_golden):
    assert np.array_equal(oracle.channel_indexes((2, 8, 3, 5)), kernels_golden["eb_indexes"])


def test_gc_forward(kernels_golden):
    g = kernels_golden
    y, s, m = g["gc_y"], g["gc_scales"], g["gc_means"]
    yh, lik = oracle.gc_forward(y, s, m)
    assert np.array_equal(yh, g["gc_yhat_means"])
    assert rel_err(lik, g["gc_lik_means"], 1e-9) < 1e-4
    yh, lik = oracle.gc_forward(y, s)
    assert np.array_equal(yh, g["gc_yhat_nomeans"])
    assert rel_err(lik, g["gc_lik_nomeans"], 1e-9) < 1e-4
    yh, lik = oracle.gc_forward(y, s, m, noise=g["gc_noise"])
    assert np.array_equal(yh, g["gc_yhat_noise"])
    assert rel_err(lik, g["gc_lik_noise"], 1e-9) < 1e-4
    # floor: forward(y=50, sigma=2) -> 1e-9
    _, l = oracle.gc_forward(np.array([50.0], np.float32), np.array([2.0], np.float32))
    assert l[0] == np.float32(1e-9)


def _eb_lists(g):
    return ([g[f"eb_param__matrix{i}"] for i in range(5)], [g[f"eb_param__bias{i}"] for i in range(5)],
            [g[f"eb_param__factor{i}"] for i in range(4)])


def test_eb_forward(kernels_golden):
    g = kernels_golden
    mats, bias, fac = _eb_lists(g)
    med = g["eb_param_quantiles"][:, 0, 1]
    x = g["eb_x"]
    C, inner = x.shape[1], x.shape[2] * x.shape[3]
    xh, lik = oracle.eb_forward(x, mats, bias, fac, med, C, inner)
    assert np.array_equal(xh, g["eb_xhat"])
    assert rel_err(lik, g["eb_lik"], 1e-9) < 1e-4
    xh, lik = oracle.eb_forward(x, mats, bias, fac, med, C, inner, noise=g["eb_noise"])
    assert np.array_equal(xh, g["eb_xhat_noise"])
    assert rel_err(lik, g["eb_lik_noise"], 1e-9) < 1e-4
    lg = oracle.eb_logits_cumulative(x, mats, bias, fac, C, inner)
    assert np.max(np.abs(lg - g["eb_logits"])) < 2e-4
    xh, lik = oracle.eb_forward(g["eb_x_2d"], mats, bias, fac, med, C, 1)
    assert np.array_equal(xh, g["eb_xhat_2d"])
    assert rel_err(lik, g["eb_lik_2d"], 1e-9) < 1e-4


def test_gdn(kernels_golden):
    g = kernels_golden
    x = g["gdn_x"]
    y = oracle.gdn_forward(x, g["gdn_beta"], g["gdn_gamma"])
    assert np.max(np.abs(y - g["gdn_y"])) < 1e-5
    y = oracle.gdn_forward(x, g["gdn_beta"], g["gdn_gamma"], inverse=True)
    assert np.max(np.abs(y - g["gdn_y_inv"]) / np.maximum(np.abs(g["gdn_y_inv"]), 1.0)) < 1e-5
    # closed form at init, tests/test_layers.py:145-146: y = x / sqrt(1 + 0.1 x^2)
    C = x.shape[1]
    ped = 2.0 ** -36
    beta0 = np.sqrt(np.ones(C) + ped).astype(np.float32)
    gamma0 = np.sqrt(0.1 * np.eye(C) + ped).astype(np.float32)
    y0 = oracle.gdn_forward(x, beta0, gamma0)
    assert np.max(np.abs(y0 - x / np.sqrt(1 + 0.1 * x ** 2))) < 1e-5
    assert np.max(np.abs(y0 - g["gdn_init_y"])) < 1e-5


def test_lower_bound(kernels_golden):
    g = kernels_golden
    y = oracle.lower_bound(g["lb_x"], 0.11)
    assert np.array_equal(y, g["lb_y"], equal_nan=True)
    dx = oracle.lower_bound_bwd(g["lb_x"], g["lb_g"], 0.11)
    assert np.array_equal(dx, g["lb_dx"])


def test_pmf_to_quantized_cdf(kernels_golden):
    g = kernels_golden
    # reference KAT tests/test_ops.py:104-106
    assert oracle.pmf_to_quantized_cdf([0.1, 0.2, 0, 0], 16).tolist() == [0, 21845, 65534, 65535, 65536]
    assert np.array_equal(oracle.pmf_to_quantized_cdf([0.1, 0.2, 0, 0], 16), g["cdf_kat"])
    for p, c, L in zip(g["cdf_pmfs"], g["cdf_cdfs"], g["cdf_lens"]):
        assert np.array_equal(oracle.pmf_to_quantized_cdf(p[:L], 16), c[:L + 1])
    for bad in ([-0.1, 0.5], [float("inf"), 0.5], [float("nan"), 0.5]):  # tests/test_ops.py:108-118
        with pytest.raises(ValueError):
            oracle.pmf_to_quantized_cdf(bad, 16)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_conv_deconv(kernels_golden, tag):
    g = kernels_golden
    k, s = g[f"conv_{tag}_cfg"]
    y = oracle.conv2d(g[f"conv_{tag}_x"], g[f"conv_{tag}_w"], g[f"conv_{tag}_b"], stride=int(s))
    assert y.shape == g[f"conv_{tag}_y"].shape
    assert np.max(np.abs(y - g[f"conv_{tag}_y"])) < 1e-4
    k, s = g[f"deconv_{tag}_cfg"]
    y = oracle.conv_transpose2d(g[f"deconv_{tag}_x"], g[f"deconv_{tag}_w"], g[f"deconv_{tag}_b"], stride=int(s))
    assert y.shape == g[f"deconv_{tag}_y"].shape
    assert np.max(np.abs(y - g[f"deconv_{tag}_y"])) < 1e-4


@pytest.mark.parametrize("arch,N,M", [("factorized", 128, 192), ("hyperprior", 128, 192), ("mean-scale", 192, 320)])
def test_torch_port_models(models_golden, arch, N, M):
    """The torch port reproduces the reference model forward and its compress() symbols/indexes."""
    g = models_golden
    tag = arch.replace("-", "_")
    sd = {k: torch.from_numpy(v) for k, v in make_state_dict(arch, N, M, seed=0).items()}
    x = torch.from_numpy(g["x"])
    torch.set_num_threads(8)
    with torch.no_grad():
        out = tp.FORWARD[arch](sd, x)
    assert np.max(np.abs(out["x_hat"].numpy() - g[f"{tag}_x_hat"])) < 1e-4
    for k, l in out["likelihoods"].items():
        assert rel_err(l.numpy(), g[f"{tag}_lik_{k}"], 1e-9) < 1e-4
    B = x.shape[0]
    if arch == "factorized":
        return
    fn = tp.hyperprior_compress_symbols if arch == "hyperprior" else tp.mean_scale_compress_symbols
    with torch.no_grad():
        c = fn(sd, x, tp.get_scale_table())
    for name in ("y_symbols", "y_indexes", "z_symbols", "z_indexes"):
        assert np.array_equal(c[name].reshape(B, -1).numpy(), g[f"{tag}_{name}"]), name
    assert c["y_symbols"].abs().max() > 5 and c["y_indexes"].unique().numel() > 20  # non-degenerate fixture


@pytest.fixture(scope="module")
def mm_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_mm.npz"))


def mm_state_dict(g, tag, seed):
    """Parameters of the two-branch codec from the key/shape contract the reference run recorded + the mask buffer."""
    import json
    from weights import make_mm_state_dict
    shapes = {k: tuple(v[0]) for k, v in json.loads(str(g[f"{tag}_state_dict"])).items()}
    sd = {k: torch.from_numpy(v) for k, v in make_mm_state_dict(shapes, seed).items()}
    sd["context_prediction.mask"] = tp.masked_conv_mask(shapes["context_prediction.weight"], "A")
    return sd


def test_torch_port_multimodality(mm_golden):
    """RGB guide branch + depth branch with cross-modality fusion (google.py:746-1248) vs the reference run."""
    g = mm_golden
    torch.set_num_threads(8)
    sd_r, sd_d = mm_state_dict(g, "r", 0), mm_state_dict(g, "d", 1)
    with torch.no_grad():
        o_r = tp.mm_r_forward(sd_r, torch.from_numpy(g["x"]))
        o_d = tp.mm_d_forward(sd_d, torch.from_numpy(g["depth"]), o_r["hidden"])
    for tag, o in (("r", o_r), ("d", o_d)):
        assert np.max(np.abs(o["x_hat"].numpy() - g[f"{tag}_x_hat"])) < 2e-4 * max(1.0, float(np.abs(g[f"{tag}_x_hat"]).max()))
        for k, l in o["likelihoods"].items():
            assert rel_err(l.numpy(), g[f"{tag}_lik_{k}"], 1e-9) < 1e-3, (tag, k)
    for k, v in o_r["hidden"].items():
        assert abs(float(v.abs().mean()) - float(g[f"r_hidden_{k}_mean_abs"])) < 1e-4 * float(g[f"r_hidden_{k}_mean_abs"])
    # the fixture is not degenerate: y spans many symbols and scales cover the table
    assert float(o_d["y"].abs().max()) > 5 and float(o_d["scales_hat"].max()) > 2 and float(o_d["scales_hat"].min()) < 0.5


@pytest.fixture(scope="module")
def ssf_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_ssf.npz"))


def ssf_state_dict(g):
    import json
    from weights import make_ssf_state_dict
    shapes = {k: tuple(v[0]) for k, v in json.loads(str(g["state_dict"])).items()}
    return {k: torch.from_numpy(v) for k, v in make_ssf_state_dict(shapes, 0).items()}


def test_torch_port_ssf2020(ssf_golden):
    """ScaleSpaceFlow eval forward (keyframe + 2 inter frames), gaussian_volume and warp_volume vs the reference run."""
    g = ssf_golden
    torch.set_num_threads(8)
    sd = ssf_state_dict(g)
    frames = [torch.from_numpy(g[f"frame_{t}"]) for t in range(3)]
    with torch.no_grad():
        o = tp.ssf_forward(sd, frames)
    for t in range(3):
        assert np.max(np.abs(o["x_hat"][t].numpy() - g[f"x_hat_{t}"])) < 1e-4
        for part, lk in o["likelihoods"][t].items():
            for k, v in lk.items():
                assert rel_err(v.numpy(), g[f"lik_{t}_{part}_{k}"], 1e-9) < 1e-3, (t, part, k)
        assert float(g[f"dec_max_abs_diff_{t}"]) < 1e-4     # the reference's own decompress(compress()) reproduces forward
    T = o["trace"][1]
    assert np.max(np.abs(T["volume"][:, :, :, ::3, ::5].numpy() - g["volume_sub"])) < 1e-5
    assert np.max(np.abs(T["motion_info"].numpy() - g["motion_info"])) < 1e-4
    assert np.max(np.abs(T["x_pred"].numpy() - g["x_pred"])) < 1e-4
    mi = g["motion_info"]
    assert np.abs(mi[:, 0]).max() * 128 > 3 and (mi[:, 2] * 3 + 2.5).min() < 0 and (mi[:, 2] * 3 + 2.5).max() > 5   # non-degenerate


def test_torch_port_colour_transforms():
    """rgb2ycbcr / ycbcr2rgb / yuv_444_to_420 / yuv_420_to_444 vs the reference's own functions (tests/golden/color.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "color.npz"))
    rgb = torch.from_numpy(g["rgb"])
    ycc = tp.rgb2ycbcr(rgb)
    assert np.array_equal(ycc.numpy(), g["ycbcr"]) and np.array_equal(tp.ycbcr2rgb(ycc).numpy(), g["rgb_back"])
    y, u, v = tp.yuv_444_to_420(ycc)
    assert np.array_equal(u.numpy(), g["u420"]) and np.array_equal(v.numpy(), g["v420"])
    assert np.array_equal(tp.yuv_420_to_444((y, u, v)).numpy(), g["yuv444"])


def test_torch_port_guided_compresser():
    """Guided_compresser (master.py:1215-1300) = the _R network on a 1-channel input: the port reproduces the reference run."""
    import json
    import os
    from weights import make_mm_state_dict
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_guided.npz"))
    shapes = {k: tuple(v[0]) for k, v in json.loads(str(g["state_dict"])).items()}
    sd = {k: torch.from_numpy(v) for k, v in make_mm_state_dict(shapes, 2).items()}
    sd["context_prediction.mask"] = tp.masked_conv_mask(shapes["context_prediction.weight"], "A")
    torch.set_num_threads(8)
    with torch.no_grad():
        o = tp.mm_r_forward(sd, torch.from_numpy(g["x"]))
    assert np.max(np.abs(o["x_hat"].numpy() - g["x_hat"])) < 2e-4 * max(1.0, float(np.abs(g["x_hat"]).max()))
    for k, l in o["likelihoods"].items():
        assert rel_err(l.numpy(), g[f"lik_{k}"], 1e-9) < 1e-3, k
    for k, v in o["hidden"].items():
        ref = g[f"hidden_{k}"]
        assert np.max(np.abs(v[:, ::8, ::2, ::2].numpy() - ref)) < 1e-4 * max(1.0, float(np.abs(ref).max())), k


def _master_golden():
    import json
    import os
    from weights import make_master_state_dict
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_master.npz"))
    shapes = {k: tuple(v[0]) for k, v in json.loads(str(g["state_dict"])).items()}
    sd = {k: torch.from_numpy(v) for k, v in make_master_state_dict(shapes, 4).items()}
    sd["context_prediction.mask"] = tp.masked_conv_mask(shapes["context_prediction.weight"], "A")
    bf = lambda a: torch.from_numpy(a).view(torch.bfloat16).float()
    hidden = {k: bf(g[f"hidden_{k}_bf16"]) for k in ("gs1", "gs2", "gs3")}
    return g, sd, bf(g["g_hat_bf16"]), hidden


def test_torch_port_master_compresser():
    """Master_compresser (master.py:837-951: Feature_encoder/decoder, Channel_aligner, Spatial_aligner's window
    cross-attention with the shifted second block, Master_decoder) -- the port reproduces the reference run."""
    g, sd, g_hat, hidden = _master_golden()
    torch.set_num_threads(8)
    with torch.no_grad():
        o = tp.master_forward(sd, torch.from_numpy(g["x"]), g_hat, hidden)
    assert np.max(np.abs(o["x_hat"].numpy() - g["x_hat"])) < 2e-4 * max(1.0, float(np.abs(g["x_hat"]).max()))
    for k, l in o["likelihoods"].items():
        assert rel_err(l.numpy(), g[f"lik_{k}"], 1e-9) < 1e-3, k


def test_master_attention_tables():
    """The port's closed-form relative-position index and shift mask equal the buffers the reference registers
    (master.py:512-522, 625-643): checked on the properties those constructions guarantee."""
    idx = tp.relative_position_index(4)
    assert idx.shape == (16, 16) and int(idx.min()) == 0 and int(idx.max()) == 48
    assert torch.equal(torch.diagonal(idx), torch.full((16,), 24))           # zero offset -> centre of the 7x7 table
    assert torch.equal(idx + idx.t(), torch.full((16, 16), 48))              # offset negation mirrors the index
    m = tp.shift_attention_mask(8, 16, 4, 2)
    assert m.shape == (8, 16, 16) and set(m.unique().tolist()) <= {0.0, -100.0}
    assert torch.equal(m, m.transpose(1, 2)) and float(m[0].abs().sum()) == 0.0   # interior windows are unmasked
    assert float(m[-1].abs().sum()) > 0                                       # the wrap-around corner window is not


def test_torch_port_mbt2018():
    """The zoo's JointAutoregressiveHierarchicalPriors (google.py:421-520): the port reproduces the reference run."""
    import json
    import os
    from weights import make_mbt2018_state_dict
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_mbt2018.npz"))
    shapes = {k: tuple(v[0]) for k, v in json.loads(str(g["state_dict"])).items()}
    sd = {k: torch.from_numpy(v) for k, v in make_mbt2018_state_dict(shapes, 3).items()}
    sd["context_prediction.mask"] = tp.masked_conv_mask(shapes["context_prediction.weight"], "A")
    torch.set_num_threads(8)
    with torch.no_grad():
        o = tp.mbt2018_forward(sd, torch.from_numpy(g["x"]))
    assert np.max(np.abs(o["x_hat"].numpy() - g["x_hat"])) < 2e-4 * max(1.0, float(np.abs(g["x_hat"]).max()))
    for k, l in o["likelihoods"].items():
        assert rel_err(l.numpy(), g[f"lik_{k}"], 1e-9) < 1e-3, k


def test_torch_port_master_compresser_one_channel_variant():
    """Master_compresser(channel=1) (swapped strides, decoder.downsample1-3): the port reproduces the reference run on seeded inputs."""
    import json
    import os
    from weights import make_master1_inputs, make_master_state_dict
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "models_master1.npz"))
    shapes = {k: tuple(v[0]) for k, v in json.loads(str(g["state_dict"])).items()}
    assert "decoder.downsample2.weight" in shapes and shapes["fencoder1.conv1.weight"][1] == 1 and shapes["fencoder2.conv1.weight"][1] == 3
    sd = {k: torch.from_numpy(v) for k, v in make_master_state_dict(shapes, 5).items()}
    sd["context_prediction.mask"] = tp.masked_conv_mask(shapes["context_prediction.weight"], "A")
    x, g_hat, hidden = make_master1_inputs(10)
    torch.set_num_threads(8)
    with torch.no_grad():
        o = tp.master_forward(sd, torch.from_numpy(x), torch.from_numpy(g_hat), {k: torch.from_numpy(v) for k, v in hidden.items()})
    assert o["x_hat"].shape == g["x_hat"].shape
    assert np.max(np.abs(o["x_hat"].numpy() - g["x_hat"])) < 2e-4 * max(1.0, float(np.abs(g["x_hat"]).max()))
    for k, l in o["likelihoods"].items():
        assert rel_err(l.numpy(), g[f"lik_{k}"], 1e-9) < 1e-3, k

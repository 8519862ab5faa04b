"""Deterministic, gain-calibrated weights for parity tests.

The reference's zoo can only give random-init weights offline, and with those every latent is
<< 1, every symbol is 0 and every CDF index is 0 -- parity would be vacuous.
This recipe draws every parameter from ``numpy.random.RandomState`` (frozen MT19937 stream, so
the values are identical wherever numpy runs) with per-layer gains chosen so that symbols span
roughly [-25, 25] and the predicted scales populate most of the 64 scale-table bins.  The same
arrays are loaded into the reference modules (when generating goldens) and into the mmcodec
modules / the oracle (when testing), by ``state_dict`` key, which is the reference's contract
(SURVEY.md Appendix D).
"""
from __future__ import annotations

import numpy as np


def _conv_w(rs, cout, cin, k, gain):
    return (rs.standard_normal((cout, cin, k, k)) * (gain / np.sqrt(cin * k * k))).astype(np.float32)


def _deconv_w(rs, cin, cout, k, gain, stride=2):
    return (rs.standard_normal((cin, cout, k, k)) * (gain / np.sqrt(cin * k * k / stride ** 2))).astype(np.float32)


def _bias(rs, c, scale=0.1):
    return (rs.standard_normal(c) * scale).astype(np.float32)


def _gdn(rs, sd, name, C):
    ped = 2.0 ** -36
    beta = rs.uniform(0.5, 2.0, C)
    gamma = 0.1 * np.eye(C) + np.abs(rs.standard_normal((C, C))) * 0.004
    sd[name + ".beta"] = np.sqrt(beta + ped).astype(np.float32)
    sd[name + ".gamma"] = np.sqrt(gamma + ped).astype(np.float32)


def _entropy_bottleneck(rs, sd, name, C):
    filters = (1, 3, 3, 3, 3, 1)
    scale = 10.0 ** (1 / 5)
    for i in range(5):
        init = np.log(np.expm1(1 / scale / filters[i + 1]))
        sd[f"{name}._matrix{i}"] = (init + rs.standard_normal((C, filters[i + 1], filters[i])) * 0.3).astype(np.float32)
        sd[f"{name}._bias{i}"] = rs.uniform(-0.5, 0.5, (C, filters[i + 1], 1)).astype(np.float32)
        if i < 4:
            sd[f"{name}._factor{i}"] = (rs.standard_normal((C, filters[i + 1], 1)) * 0.5).astype(np.float32)
    med = rs.standard_normal(C)
    q = np.stack([med - rs.uniform(5, 15, C), med, med + rs.uniform(5, 15, C)], axis=1)
    sd[f"{name}.quantiles"] = q.reshape(C, 1, 3).astype(np.float32)


def make_state_dict(arch: str, N: int, M: int, seed: int = 0, in_ch: int = 3):
    """arch in {"factorized", "hyperprior", "mean-scale"}; returns {key: np.ndarray(float32)}."""
    rs = np.random.RandomState(seed)
    sd = {}
    # g_a: conv, GDN, conv, GDN, conv, GDN, conv   (models/google.py:143-151)
    chans = [in_ch, N, N, N, M]
    for li, i in enumerate((0, 2, 4, 6)):
        gain = 3.0 if i < 6 else 5.0
        sd[f"g_a.{i}.weight"] = _conv_w(rs, chans[li + 1], chans[li], 5, gain)
        sd[f"g_a.{i}.bias"] = _bias(rs, chans[li + 1])
        if i < 6:
            _gdn(rs, sd, f"g_a.{i + 1}", N)
    # g_s: deconv, IGDN, ... , deconv   (models/google.py:153-161)
    chans = [M, N, N, N, in_ch]
    for li, i in enumerate((0, 2, 4, 6)):
        gain = 0.12 if i == 0 else (0.5 if i < 6 else 0.3)
        sd[f"g_s.{i}.weight"] = _deconv_w(rs, chans[li], chans[li + 1], 5, gain)
        sd[f"g_s.{i}.bias"] = _bias(rs, chans[li + 1])
        if i < 6:
            _gdn(rs, sd, f"g_s.{i + 1}", N)
    if arch == "factorized":
        _entropy_bottleneck(rs, sd, "entropy_bottleneck", M)
        return sd
    _entropy_bottleneck(rs, sd, "entropy_bottleneck", N)
    # h_a: conv3x3 s1, act, conv5 s2, act, conv5 s2   (models/google.py:254-260,363-369)
    sd["h_a.0.weight"] = _conv_w(rs, N, M, 3, 1.0)
    sd["h_a.0.bias"] = _bias(rs, N)
    sd["h_a.2.weight"] = _conv_w(rs, N, N, 5, 1.0)
    sd["h_a.2.bias"] = _bias(rs, N)
    sd["h_a.4.weight"] = _conv_w(rs, N, N, 5, 1.0)
    sd["h_a.4.bias"] = _bias(rs, N, 0.5)
    if arch == "hyperprior":
        # h_s: deconv(N,N), ReLU, deconv(N,N), ReLU, conv3x3(N,M), ReLU   (models/google.py:262-269)
        sd["h_s.0.weight"] = _deconv_w(rs, N, N, 5, 1.0)
        sd["h_s.0.bias"] = _bias(rs, N)
        sd["h_s.2.weight"] = _deconv_w(rs, N, N, 5, 1.5)
        sd["h_s.2.bias"] = _bias(rs, N)
        sd["h_s.4.weight"] = _conv_w(rs, M, N, 3, 2.0)
        sd["h_s.4.bias"] = (_bias(rs, M, 0.3) + 2.0).astype(np.float32)
    elif arch == "mean-scale":
        # h_s: deconv(N,M), LeakyReLU, deconv(M,3M/2), LeakyReLU, conv3x3(3M/2, 2M)   (models/google.py:371-377)
        M32 = M * 3 // 2
        sd["h_s.0.weight"] = _deconv_w(rs, N, M, 5, 1.0)
        sd["h_s.0.bias"] = _bias(rs, M)
        sd["h_s.2.weight"] = _deconv_w(rs, M, M32, 5, 1.5)
        sd["h_s.2.bias"] = _bias(rs, M32)
        w = _conv_w(rs, 2 * M, M32, 3, 2.0)
        w[M:] *= 0.25  # second half of the channels are the means (models/google.py:384)
        sd["h_s.4.weight"] = w
        b = _bias(rs, 2 * M, 0.3)
        b[:M] += 2.0
        sd["h_s.4.bias"] = b.astype(np.float32)
    else:
        raise ValueError(arch)
    return sd


def make_image(B: int, H: int, W: int, seed: int = 1234, C: int = 3):
    """Smooth-ish synthetic image batch in [0, 1] (low-pass noise + fine noise)."""
    rs = np.random.RandomState(seed)
    coarse = rs.uniform(0, 1, (B, C, (H + 7) // 8, (W + 7) // 8))
    img = np.kron(coarse, np.ones((8, 8)))[:, :, :H, :W] * 0.8 + rs.uniform(0, 0.2, (B, C, H, W))
    return img.astype(np.float32)


def make_state_dict_like(shapes, seed: int = 0, gains=None):
    """Deterministic values for ANY of the codec models, driven by the state_dict key names and shapes
    (``shapes``: {key: shape}).  Conv / deconv weights ~ N(0, gain^2 / fan_in), biases ~ N(0, 0.1^2), GDN beta / gamma and
    EntropyBottleneck parameters as in ``make_state_dict``; buffers (masks, bounds, CDF tables) are left alone.
    ``gains`` maps a key prefix to a gain (longest prefix wins, default 1.0)."""
    gains = gains or {}
    rs = np.random.RandomState(seed)
    ped = 2.0 ** -36
    sd = {}

    def gain_of(key):
        best, g = -1, 1.0
        for p, v in gains.items():
            if key.startswith(p) and len(p) > best:
                best, g = len(p), v
        return g

    done_eb = set()
    for key in sorted(shapes):
        shp = tuple(shapes[key])
        leaf = key.rsplit(".", 1)[-1]
        if leaf in ("mask", "bound", "pedestal", "target", "scale_table", "scale_bound", "_offset", "_quantized_cdf", "_cdf_length"):
            continue
        if leaf.startswith(("_matrix", "_bias", "_factor")) or leaf == "quantiles":
            prefix = key.rsplit(".", 1)[0]
            if prefix not in done_eb:
                done_eb.add(prefix)
                C = shapes[prefix + ".quantiles"][0]
                _entropy_bottleneck(np.random.RandomState(seed + 17 + len(prefix)), sd, prefix, C)
            continue
        if leaf == "beta":
            sd[key] = np.sqrt(rs.uniform(0.5, 2.0, shp) + ped).astype(np.float32)
        elif leaf == "gamma":
            C = shp[0]
            sd[key] = np.sqrt(0.1 * np.eye(C) + np.abs(rs.standard_normal(shp)) * 0.004 + ped).astype(np.float32)
        elif leaf == "weight" and len(shp) == 4:
            is_deconv = ("g_s" in key and "conv" in key or key.startswith("h_s.0") or key.startswith("h_s.2") or "deconv" in key
                         or "decoder" in key)
            fan = shp[0 if is_deconv else 1] * shp[2] * shp[3]
            if is_deconv:
                fan /= 4.0
            sd[key] = (rs.standard_normal(shp) * (gain_of(key) / np.sqrt(fan))).astype(np.float32)
        elif leaf == "bias":
            sd[key] = (rs.standard_normal(shp) * 0.1).astype(np.float32)
    return sd


MM_GAINS = {"enc1.g_a_conv4": 5.0, "enc1.": 3.0, "dec1.g_s_conv1": 0.12, "dec1.": 0.5, "dec1.g_s_conv4": 0.3, "h_s.4": 2.0,
            "pic2_g_a_conv4": 5.0, "pic2_g_a": 3.0, "pic2_g_s_conv1": 0.12, "pic2_g_s": 0.5, "pic2_g_s_conv4": 0.3,
            "tran_conv": 1.5, "eg_ext": 1.5, "context_prediction": 0.5, "entropy_parameters": 1.5}


def make_mm_state_dict(shapes, seed: int = 0):
    """Weights for JointAutoregressiveHierarchicalPriors_R / _D (two-branch RGB + depth codec): generic recipe plus a
    positive offset on the scale half of the entropy-parameter head so that predicted scales cover the scale table."""
    sd = make_state_dict_like(shapes, seed, MM_GAINS)
    M = shapes["entropy_parameters.4.bias"][0] // 2
    sd["entropy_parameters.4.bias"][:M] += 2.0
    sd["entropy_parameters.4.weight"][M:] *= 0.25
    return sd


SSF_GAINS = {"img_encoder.6": 15.0, "res_encoder.6": 25.0, "motion_encoder.6": 16.0,
             "img_decoder.6": 0.1, "res_decoder.6": 0.04, "motion_decoder.6": 0.25,
             "img_hyperprior.hyper_encoder.4": 6.0, "res_hyperprior.hyper_encoder.4": 6.0, "motion_hyperprior.hyper_encoder.4": 6.0,
             "img_hyperprior.hyper_decoder_scale.deconv3": 4.0, "res_hyperprior.hyper_decoder_scale.deconv3": 4.0,
             "motion_hyperprior.hyper_decoder_scale.deconv3": 4.0}


def make_ssf_state_dict(shapes, seed: int = 0):
    """Weights for ScaleSpaceFlow: generic recipe; gains chosen so that latents span tens of symbols, predicted scales
    cover the scale table (a few per cent of the likelihoods on the 1e-9 floor), reconstructions stay image-like over the
    recurrence, the decoded flow is a few pixels and the scale field covers all levels of the volume."""
    sd = make_state_dict_like(shapes, seed, SSF_GAINS)
    # motion decoder output = (flow x, flow y, scale field) in NORMALISED grid units: keep the flow at a few per cent of
    # the frame and spread the scale field over about [-0.9, 0.9] (volume index 3 z + 2.5, 6 levels)
    w = sd["motion_decoder.6.weight"]            # (Cin, 3, 5, 5)
    w[:, :2] *= 0.03
    w[:, 2] *= 1.2
    sd["motion_decoder.6.bias"] = np.array([0.01, -0.015, 0.1], dtype=np.float32)
    for hp in ("img_hyperprior", "res_hyperprior", "motion_hyperprior"):
        sd[f"{hp}.hyper_decoder_scale.deconv3.bias"] += 3.0     # scales are QReLU outputs: keep most of them positive
    sd["img_decoder.6.bias"] = np.full(3, 0.45, dtype=np.float32)
    return sd


MASTER_GAINS = {"g_a.6": 5.0, "g_a.": 3.0, "decoder.g_s_conv1": 0.12, "decoder.g_s_conv": 0.5, "decoder.g_s_conv4": 0.6,
                "decoder.downsample": 1.0, "h_s.4": 2.0, "context_prediction": 0.5, "entropy_parameters": 1.5,
                "fencoder": 1.2, "fencoder1.resblock": 0.5, "fencoder2.resblock": 0.5, "ch_aligner": 1.3,
                "fdecoder.resblock": 0.4, "fdecoder.resblock1.conv1": 0.2, "fdecoder.resblock1.skip": 0.2, "fdecoder.conv": 0.2,
                "fdecoder.deconv1": 0.12}


def make_master_state_dict(shapes, seed: int = 0):
    """Weights for Master_compresser (master.py:837-902): the generic recipe for the conv / GDN / entropy-model tensors plus
    values for what only this model has -- Linear weights ~ N(0, 1/fan_in), LayerNorm weight 1 +- 0.1, relative-position
    tables ~ N(0, 0.5^2) (so the window bias and the shift mask matter), 2x2 patch-embedding / recovery kernels."""
    sd = make_state_dict_like(shapes, seed, MASTER_GAINS)
    rs = np.random.RandomState(seed + 101)
    for key in sorted(shapes):
        shp = tuple(shapes[key])
        leaf = key.rsplit(".", 1)[-1]
        if ".norm" in key and leaf == "weight":
            sd[key] = (1.0 + 0.1 * rs.standard_normal(shp)).astype(np.float32)
        elif ".norm" in key and leaf == "bias":
            sd[key] = (0.05 * rs.standard_normal(shp)).astype(np.float32)
        elif leaf == "relative_position_bias_table":
            sd[key] = (0.5 * rs.standard_normal(shp)).astype(np.float32)
        elif leaf == "weight" and len(shp) == 2:
            g = 2.0 if (".qkv" in key) else 1.0
            sd[key] = (rs.standard_normal(shp) * (g / np.sqrt(shp[1]))).astype(np.float32)
        elif "patch_embeding" in key and leaf == "weight":
            sd[key] = (rs.standard_normal(shp) * (1.0 / np.sqrt(shp[1] * 4))).astype(np.float32)
        elif key.endswith("recovery.weight"):
            sd[key] = (rs.standard_normal(shp) * (1.0 / np.sqrt(shp[0]))).astype(np.float32)
        elif key.endswith("fdecoder.deconv1.weight"):
            sd[key] = (rs.standard_normal(shp) * (MASTER_GAINS["fdecoder.deconv1"] / np.sqrt(shp[0] * 9 / 4.0))).astype(np.float32)
    M = shapes["entropy_parameters.4.bias"][0] // 2
    sd["entropy_parameters.4.bias"][:M] += 2.0
    sd["entropy_parameters.4.weight"][M:] *= 0.25
    sd["fdecoder.deconv1.bias"] = np.full(shapes["fdecoder.deconv1.bias"], 0.45, dtype=np.float32)
    return sd


MBT2018_GAINS = {"g_a.6": 3.0, "g_a.": 3.0, "g_s.0": 0.12, "g_s.": 0.5, "g_s.6": 0.3, "h_s.4": 2.0, "context_prediction": 0.5,
                 "entropy_parameters": 1.5}


def make_mbt2018_state_dict(shapes, seed: int = 0):
    """Weights for the zoo's JointAutoregressiveHierarchicalPriors (mbt2018): the two-branch recipe on the stock g_a / g_s names."""
    sd = make_state_dict_like(shapes, seed, MBT2018_GAINS)
    M = shapes["entropy_parameters.4.bias"][0] // 2
    sd["entropy_parameters.4.bias"][:M] += 2.0
    sd["entropy_parameters.4.weight"][M:] *= 0.25
    return sd


def make_master1_inputs(seed: int = 10):
    """Inputs of the 1-channel-master golden (regenerated from the seed on both sides, not stored): 1 x 64 x 64 master image,
    3 x 128 x 128 decoded guide, and guide hidden maps gs1..gs3 at twice the master decoder's resolutions (bf16-representable)."""
    rs = np.random.RandomState(seed)
    x = make_image(1, 64, 64, seed=seed + 1, C=1)
    g_hat = make_image(1, 128, 128, seed=seed + 2, C=3)
    hidden = {}
    for k, r in (("gs1", 16), ("gs2", 32), ("gs3", 64)):
        v = (rs.standard_normal((1, 192, r, r)) * 0.5).astype(np.float32)
        hidden[k] = (v.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)      # truncate to bf16-representable values
    return x, g_hat, hidden

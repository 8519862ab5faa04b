"""Host-side mirror of ``compressai.models.google`` for the hot path: CompressionModel,
FactorizedPrior, ScaleHyperprior, MeanScaleHyperprior with the reference's constructor
arguments, sub-module names (hence ``state_dict`` keys) and return structures
(compressai/models/google.py:58-416).  ``forward`` keeps every intermediate on the device in the
kernels' native layouts: NHWC bf16 between transform layers, NHWC fp32 for the latents that feed
the entropy stage; tensors returned to the caller have the reference's logical (N, C, H, W) shape
(channels-last strided for the latent-sized ones, planar for x_hat).
"""
from __future__ import annotations

import math
import warnings

import torch
import torch.nn as nn
from torch import Tensor

from . import autograd as AG
from . import ops
from .entropy_models import EntropyBottleneck, GaussianConditional
from .layers import GDN, conv, deconv
from .transforms import TransformStack, current_precision, run_layers

__all__ = ["CompressionModel", "FactorizedPrior", "ScaleHyperprior", "MeanScaleHyperprior", "get_scale_table",
           "SCALES_MIN", "SCALES_MAX", "SCALES_LEVELS", "MODELS", "CFGS", "build_model", "set_entropy_coder"]

SCALES_MIN = 0.11
SCALES_MAX = 256
SCALES_LEVELS = 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    """compressai/models/google.py:208-214 (computed on the host by torch, then kept as a buffer)"""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


def _resize_registered_buffers(module, module_name, buffer_names, state_dict):
    """Resize variable-length CDF buffers before load_state_dict (compressai/models/utils.py:90-125)."""
    for name in buffer_names:
        key = f"{module_name}.{name}"
        if key not in state_dict:
            raise RuntimeError(f'Missing key "{key}" in state_dict')
        new = state_dict[key]
        cur = getattr(module, name)
        if cur.shape != new.shape:
            setattr(module, name, torch.empty(new.shape, dtype=cur.dtype, device=cur.device))


def _nhwc_to_logical(t: Tensor) -> Tensor:
    return t.permute(0, 3, 1, 2)


class CompressionModel(nn.Module):
    """compressai/models/google.py:58-123"""

    def __init__(self, entropy_bottleneck_channels, init_weights=None):
        super().__init__()
        self.entropy_bottleneck = EntropyBottleneck(entropy_bottleneck_channels)
        if init_weights is not None:
            warnings.warn("init_weights was removed as it was never functional", DeprecationWarning)

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def _bottleneck(self, v: Tensor):
        """entropy_bottleneck(v) -> (v_hat as bf16 for the next transform, likelihoods); training mode draws the uniform
        noise with torch's generator and records the backward (``_noise_override`` lets the parity tests inject the oracle's)."""
        eb = self.entropy_bottleneck
        if self.training:
            v_hat, lik = AG.eb_forward(v, eb, self._draw("z", v))
            return AG.cast_bf16(v_hat), lik
        self.begin_forward()                    # the bottleneck is the first entropy stage of every zoo forward: new sum
        acc = self._bits_accumulator(v)
        v_hat, lik, v_hat_bf16 = ops.eb_forward(v, eb._params(), None, eb._lik_bound(), want_bf16=True, lut=eb._eval_lut(), bits=acc)
        lik._mmc_bits_total = acc
        # precision("fp32"): the next transform takes the fp32 values (medians are not integers: bf16 would round them)
        return (v_hat if current_precision() == "fp32" else v_hat_bf16), lik

    def _conditional(self, y: Tensor, scales: Tensor, means):
        """gaussian_conditional(y, scales, means) -> (y_hat as bf16, likelihoods)"""
        gc = self.gaussian_conditional
        bound, lb = gc.lower_bound_scale._sync_bound(), gc._lik_bound()
        if self.training:
            y_hat, lik = AG.gc_forward(y, scales, means, self._draw("y", y), bound, lb)
            return AG.cast_bf16(y_hat), lik
        acc = self._bits_accumulator(y)
        y_hat, lik, y_hat_bf16 = ops.gc_forward(y, scales, means, None, bound, lb, want_bf16=True, bits=acc)
        lik._mmc_bits_total = acc
        return (y_hat if current_precision() == "fp32" else y_hat_bf16), lik

    def _bits_accumulator(self, like: Tensor) -> Tensor:
        """Eval forward: -sum(log2 likelihood) of ALL likelihood tensors of one forward accumulates into one fp32 scalar inside the
        entropy kernels (warp-shuffle + one atomic per CTA) and travels with the returned tensors as ``_mmc_bits_total``;
        ``bpp(out)`` then needs no second pass over them.  ``_bottleneck`` (the first entropy stage of a forward) starts a new sum."""
        acc = getattr(self, "_bits_acc", None)
        if acc is None or acc.device != like.device:
            acc = torch.zeros(1, dtype=torch.float32, device=like.device)
            object.__setattr__(self, "_bits_acc", acc)
        return acc

    def begin_forward(self):
        object.__setattr__(self, "_bits_acc", None)

    def _draw(self, key: str, like: Tensor) -> Tensor:
        noise = getattr(self, "_noise_override", None) or {}
        return noise[key].to(like.device) if key in noise else torch.empty_like(like).uniform_(-0.5, 0.5)

    def _tag_layer_names(self):
        """Give every conv its state_dict prefix (g_a.0, h_s.4, ...) for profiling labels."""
        for name, m in self.named_modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                m._mmc_name = name

    def forward(self, *args):
        raise NotImplementedError()

    def update(self, force=False):
        updated = False
        for m in self.children():
            if not isinstance(m, EntropyBottleneck):
                continue
            updated |= m.update(force=force)
        return updated

    def load_state_dict(self, state_dict, strict: bool = True):
        _resize_registered_buffers(self.entropy_bottleneck, "entropy_bottleneck",
                                   ["_quantized_cdf", "_offset", "_cdf_length"], state_dict)
        return super().load_state_dict(state_dict, strict=strict)

    @staticmethod
    def bpp(out, num_pixels=None) -> float:
        """sum over likelihood tensors of log(l).sum() / (-ln2 * pixels)
        (compressai/utils/eval_model/__main__t.py:197-200), reduced on the device."""
        liks = list(out["likelihoods"].values())
        fused = [getattr(lk, "_mmc_bits_total", None) for lk in liks]
        if liks and fused[0] is not None and all(f is fused[0] for f in fused):
            acc = fused[0]                      # summed inside the entropy kernels of the forward that produced `out`
        else:
            acc = None
            for lk in liks:
                acc = ops.bits(lk, acc)
        if num_pixels is None:
            x = out["x_hat"]
            num_pixels = x.size(0) * x.size(2) * x.size(3)
        return float(acc.item()) / num_pixels


def _stack_input(t: Tensor) -> Tensor:
    """logical (B, C, H, W) fp32 tensor -> the (B, H, W, C) input of the next transform stack: bf16 on the fast path, the fp32
    values themselves under precision("fp32")"""
    if current_precision() == "fp32":
        return t.float().permute(0, 2, 3, 1).contiguous()
    return ops.to_bf16(t).permute(0, 2, 3, 1).contiguous()


def _g_a(N, M, channel):
    return TransformStack(conv(channel, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))


def _g_s(N, M, channel):
    return TransformStack(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True), deconv(N, N),
                          GDN(N, inverse=True), deconv(N, channel))


class FactorizedPrior(CompressionModel):
    """compressai/models/google.py:127-204"""

    def __init__(self, N, M, channel=3, **kwargs):
        super().__init__(entropy_bottleneck_channels=M, **kwargs)
        self.g_a = _g_a(N, M, channel)
        self.g_s = _g_s(N, M, channel)
        self.N = N
        self.M = M
        self._tag_layer_names()

    @property
    def downsampling_factor(self) -> int:
        return 2 ** 4

    def forward(self, x):
        """models/google.py:172-182"""
        y = run_layers(list(self.g_a), x, "nchw_f32", "nhwc_f32")
        y_l = _nhwc_to_logical(y)
        y_hat_bf16, y_lik = self._bottleneck(y_l)
        x_hat = run_layers(list(self.g_s), y_hat_bf16.permute(0, 2, 3, 1), "nhwc_bf16", "nchw_f32")
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik}}

    @classmethod
    def from_state_dict(cls, state_dict, channel=3):
        N = state_dict["g_a.0.weight"].size(0)
        M = state_dict["g_a.6.weight"].size(0)
        net = cls(N, M, channel=channel)
        net.load_state_dict(state_dict)
        return net

    def symbols_and_indexes(self, x):
        """compress() up to the rANS call (models/google.py:196-199)."""
        y = _nhwc_to_logical(run_layers(list(self.g_a), x, "nchw_f32", "nhwc_f32"))
        y_symbols, y_indexes = self.entropy_bottleneck.symbols_and_indexes(y)
        return {"y_symbols": y_symbols, "y_indexes": y_indexes, "shape": y.size()[-2:]}

    def compress(self, x):
        """models/google.py:196-199"""
        y = _nhwc_to_logical(run_layers(list(self.g_a), x, "nchw_f32", "nhwc_f32"))
        y_strings = self.entropy_bottleneck.compress(y)
        return {"strings": [y_strings], "shape": y.size()[-2:]}

    def decompress(self, strings, shape):
        """models/google.py:201-205"""
        assert isinstance(strings, list) and len(strings) == 1
        y_hat = self.entropy_bottleneck.decompress(strings[0], shape)
        x_hat = run_layers(list(self.g_s), _stack_input(y_hat), "nhwc_bf16", "nchw_f32")
        return {"x_hat": x_hat.clamp_(0, 1)}


class ScaleHyperprior(CompressionModel):
    """compressai/models/google.py:218-344"""

    def __init__(self, N, M, channel=3, **kwargs):
        super().__init__(entropy_bottleneck_channels=N, **kwargs)
        self.g_a = _g_a(N, M, channel)
        self.g_s = _g_s(N, M, channel)
        self.h_a = TransformStack(conv(M, N, stride=1, kernel_size=3), nn.ReLU(inplace=True), conv(N, N),
                                  nn.ReLU(inplace=True), conv(N, N))
        self.h_s = TransformStack(deconv(N, N), nn.ReLU(inplace=True), deconv(N, N), nn.ReLU(inplace=True),
                                  conv(N, M, stride=1, kernel_size=3), nn.ReLU(inplace=True))
        self.gaussian_conditional = GaussianConditional(None)
        self.N = int(N)
        self.M = int(M)
        self._tag_layer_names()

    @property
    def downsampling_factor(self) -> int:
        return 2 ** (4 + 2)

    # -- shared pieces -------------------------------------------------------------------------
    def _analysis(self, x):
        """y = g_a(x) (fp32 NHWC) and z = h_a(|y|) (fp32 NHWC); |y| is written by g_a's last kernel."""
        y, y_abs = run_layers(list(self.g_a), x, "nchw_f32", "nhwc_f32", out2=1)
        z = run_layers(list(self.h_a), y_abs, "nhwc_bf16", "nhwc_f32")
        return y, z

    def forward(self, x):
        """models/google.py:281-295"""
        y, z = self._analysis(x)
        z_hat_bf16, z_lik = self._bottleneck(_nhwc_to_logical(z))
        scales_hat = run_layers(list(self.h_s), z_hat_bf16.permute(0, 2, 3, 1), "nhwc_bf16", "nhwc_f32")
        y_hat_bf16, y_lik = self._conditional(_nhwc_to_logical(y), _nhwc_to_logical(scales_hat), None)
        x_hat = run_layers(list(self.g_s), y_hat_bf16.permute(0, 2, 3, 1), "nhwc_bf16", "nchw_f32")
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}}

    def load_state_dict(self, state_dict, strict: bool = True):
        _resize_registered_buffers(self.gaussian_conditional, "gaussian_conditional",
                                   ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
        return super().load_state_dict(state_dict, strict=strict)

    @classmethod
    def from_state_dict(cls, state_dict, channel=3):
        N = state_dict["g_a.0.weight"].size(0)
        M = state_dict["g_a.6.weight"].size(0)
        net = cls(N, M, channel=channel)
        net.load_state_dict(state_dict)
        return net

    def update(self, scale_table=None, force=False):
        if scale_table is None:
            scale_table = get_scale_table()
        updated = self.gaussian_conditional.update_scale_table(scale_table, force=force)
        updated |= super().update(force=force)
        return updated

    def _z_path(self, z):
        """z symbols/indexes and z_hat = dequantize(symbols, medians), i.e. what
        entropy_bottleneck.decompress(entropy_bottleneck.compress(z)) returns (models/google.py:327-328)."""
        eb = self.entropy_bottleneck
        z_l = _nhwc_to_logical(z)
        z_symbols, z_indexes = eb.symbols_and_indexes(z_l)
        medians = eb._get_medians().detach().reshape(1, -1, 1, 1)
        z_hat = eb.dequantize(z_symbols, medians)
        return z_symbols, z_indexes, z_hat

    def symbols_and_indexes(self, x):
        """compress() up to the two rANS calls (models/google.py:324-332)."""
        gc = self.gaussian_conditional
        y, z = self._analysis(x)
        z_symbols, z_indexes, z_hat = self._z_path(z)
        z_hat_bf16 = _stack_input(z_hat)
        scales_hat = _nhwc_to_logical(run_layers(list(self.h_s), z_hat_bf16, "nhwc_bf16", "nhwc_f32"))
        y_indexes = gc.build_indexes(scales_hat)
        y_symbols, y_indexes = gc.symbols_and_indexes(_nhwc_to_logical(y), y_indexes)
        return {"y_symbols": y_symbols, "y_indexes": y_indexes, "z_symbols": z_symbols, "z_indexes": z_indexes,
                "shape": z_symbols.size()[-2:]}

    def _coder_tables(self, em):
        return em._quantized_cdf, em._cdf_length, em._offset

    def compress(self, x):
        """models/google.py:324-332 (mean-scale: :393-404): symbols / indexes on the GPU; rANS streams on the host ("ans", the
        reference's byte-identical stream) or on the device ("ans-lanes", see EntropyModel.CODERS / set_entropy_coder)."""
        c = self.symbols_and_indexes(x)
        y_strings = self.gaussian_conditional.encode_symbols(c["y_symbols"], c["y_indexes"])
        z_strings = self.entropy_bottleneck.encode_symbols(c["z_symbols"], c["z_indexes"])
        return {"strings": [y_strings, z_strings], "shape": c["shape"]}

    def _synthesis_from(self, y_hat):
        x_hat = run_layers(list(self.g_s), _stack_input(y_hat), "nhwc_bf16", "nchw_f32")
        return {"x_hat": x_hat.clamp_(0, 1)}

    def decompress(self, strings, shape):
        """models/google.py:334-344"""
        assert isinstance(strings, list) and len(strings) == 2
        gc = self.gaussian_conditional
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape)
        z_hat_bf16 = _stack_input(z_hat)
        scales_hat = _nhwc_to_logical(run_layers(list(self.h_s), z_hat_bf16, "nhwc_bf16", "nhwc_f32"))
        indexes = gc.build_indexes(scales_hat)
        y_hat = gc.decompress(strings[0], indexes, z_hat.dtype)
        return self._synthesis_from(y_hat)


class MeanScaleHyperprior(ScaleHyperprior):
    """compressai/models/google.py:348-416"""

    def __init__(self, N, M, channel=3, **kwargs):
        super().__init__(N, M, channel, **kwargs)
        self.h_a = TransformStack(conv(M, N, stride=1, kernel_size=3), nn.LeakyReLU(inplace=True), conv(N, N),
                                  nn.LeakyReLU(inplace=True), conv(N, N))
        self.h_s = TransformStack(deconv(N, M), nn.LeakyReLU(inplace=True), deconv(M, M * 3 // 2),
                                  nn.LeakyReLU(inplace=True), conv(M * 3 // 2, M * 2, stride=1, kernel_size=3))
        self._tag_layer_names()

    def _analysis(self, x):
        """No abs() in the mean-scale model (models/google.py:381): h_a reads y itself."""
        y, y_bf16 = run_layers(list(self.g_a), x, "nchw_f32", "nhwc_f32", out2=2)
        z = run_layers(list(self.h_a), y_bf16, "nhwc_bf16", "nhwc_f32")
        return y, z

    def _gaussian_params(self, z_hat_bf16_nhwc):
        """scales_hat, means_hat = h_s(z_hat).chunk(2, 1) (models/google.py:383-384) as NHWC fp32 views."""
        params = run_layers(list(self.h_s), z_hat_bf16_nhwc, "nhwc_bf16", "nhwc_f32")   # (B, H, W, 2M)
        return params[..., : self.M], params[..., self.M:]

    def forward(self, x):
        """models/google.py:379-391"""
        y, z = self._analysis(x)
        z_hat_bf16, z_lik = self._bottleneck(_nhwc_to_logical(z))
        scales_hat, means_hat = self._gaussian_params(z_hat_bf16.permute(0, 2, 3, 1))
        y_hat_bf16, y_lik = self._conditional(_nhwc_to_logical(y), _nhwc_to_logical(scales_hat), _nhwc_to_logical(means_hat))
        x_hat = run_layers(list(self.g_s), y_hat_bf16.permute(0, 2, 3, 1), "nhwc_bf16", "nchw_f32")
        return {"x_hat": x_hat, "likelihoods": {"y": y_lik, "z": z_lik}}

    def symbols_and_indexes(self, x):
        """compress() up to the two rANS calls (models/google.py:393-404)."""
        gc = self.gaussian_conditional
        y, z = self._analysis(x)
        z_symbols, z_indexes, z_hat = self._z_path(z)
        scales_hat, means_hat = self._gaussian_params(_stack_input(z_hat))
        y_indexes = gc.build_indexes(_nhwc_to_logical(scales_hat))
        y_symbols, y_indexes = gc.symbols_and_indexes(_nhwc_to_logical(y), y_indexes, means=_nhwc_to_logical(means_hat))
        return {"y_symbols": y_symbols, "y_indexes": y_indexes, "z_symbols": z_symbols, "z_indexes": z_indexes,
                "shape": z_symbols.size()[-2:]}

    def decompress(self, strings, shape):
        """models/google.py:406-416"""
        assert isinstance(strings, list) and len(strings) == 2
        gc = self.gaussian_conditional
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape)
        scales_hat, means_hat = self._gaussian_params(_stack_input(z_hat))
        scales_hat, means_hat = _nhwc_to_logical(scales_hat), _nhwc_to_logical(means_hat)
        indexes = gc.build_indexes(scales_hat)
        y_hat = gc.decompress(strings[0], indexes, means=means_hat)
        return self._synthesis_from(y_hat)


MODELS = {"bmshj2018-factorized": FactorizedPrior, "bmshj2018-hyperprior": ScaleHyperprior, "mbt2018-mean": MeanScaleHyperprior}

# (N, M) per quality, compressai/zoo/image.py:189-220
CFGS = {
    "bmshj2018-factorized": {q: ((128, 192) if q <= 5 else (192, 320)) for q in range(1, 9)},
    "bmshj2018-hyperprior": {q: ((128, 192) if q <= 5 else (192, 320)) for q in range(1, 9)},
    "mbt2018-mean": {q: ((128, 192) if q <= 4 else (192, 320)) for q in range(1, 9)},
}


def set_entropy_coder(net: nn.Module, name: str) -> nn.Module:
    """Select the entropy coder of every entropy model of ``net``: "ans" (reference-compatible streams, host threads) or
    "ans-lanes" (libmmcodec's lane container, coded on the device).  compress() and decompress() must use the same coder."""
    from .entropy_models import EntropyModel
    if name not in EntropyModel.CODERS:
        raise ValueError(f'Invalid entropy coder "{name}" (available: {EntropyModel.CODERS})')
    for m in net.modules():
        if isinstance(m, EntropyModel):
            m.entropy_coder_name = name
    return net


def build_model(architecture: str, quality: int, channel: int = 3, **kwargs):
    """Random-init model by zoo name (compressai/zoo/image.py:249-273 without the pretrained download)."""
    if architecture not in MODELS:
        raise ValueError(f'Invalid architecture name "{architecture}"')
    if quality not in CFGS[architecture]:
        raise ValueError(f'Invalid quality value "{quality}"')
    return MODELS[architecture](*CFGS[architecture][quality], channel=channel, **kwargs)

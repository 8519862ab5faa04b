"""Host-side mirror of ``compressai.entropy_models`` (EntropyModel, EntropyBottleneck,
GaussianConditional): same constructor arguments, parameter / buffer names and exceptions as
compressai/entropy_models/entropy_models.py; the per-element work runs in libmmcodec kernels.

On the hot path (forward / quantize / dequantize / build_indexes / _build_indexes / _likelihood)
everything is one fused kernel per call.  ``update()`` builds the CDF tables once per model
(off the per-image path: SURVEY.md section 8f row 2) with a handful of tiny torch ops plus the
library's host-side ``pmf_to_quantized_cdf``.  ``compress`` / ``decompress`` hand the int32 symbols / indexes
(``symbols_and_indexes``, exactly what the reference passes to ``encode_with_indexes``) to the library's host rANS
coder, which is bitstream-compatible with ``compressai.ans``.
"""
from __future__ import annotations

import warnings
from typing import Any, List, Optional, Tuple, Union

import numpy as np
import scipy.stats
import torch
import torch.nn as nn
from torch import Tensor

from . import ops
from .layers import LowerBound

__all__ = ["EntropyModel", "EntropyBottleneck", "GaussianConditional"]


class EntropyModel(nn.Module):
    """Entropy model base class (entropy_models.py:101-327)."""

    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder: Optional[str] = None,
                 entropy_coder_precision: int = 16):
        super().__init__()
        self.entropy_coder_name = entropy_coder or "ans"
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.likelihood_bound_value = float(likelihood_bound)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        # filled by update()
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    @property
    def offset(self):
        return self._offset

    @property
    def quantized_cdf(self):
        return self._quantized_cdf

    @property
    def cdf_length(self):
        return self._cdf_length

    def _lik_bound(self) -> float:
        return self.likelihood_lower_bound._sync_bound() if self.use_likelihood_bound else 0.0

    def quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None) -> Tensor:
        """entropy_models.py:157-182"""
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            noise = torch.empty_like(inputs).uniform_(-0.5, 0.5)  # drawn by torch so both sides share the stream
            return ops.quantize_noise(inputs, noise)
        if mode == "dequantize":
            return ops.quantize_dequantize(inputs, means)
        return ops.quantize_symbols(inputs, means)

    def _quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None) -> Tensor:
        warnings.warn("_quantize is deprecated. Use quantize instead.")
        return self.quantize(inputs, mode, means)

    @staticmethod
    def dequantize(inputs: Tensor, means: Optional[Tensor] = None, dtype: torch.dtype = torch.float) -> Tensor:
        """entropy_models.py:190-199"""
        out = ops.dequantize(inputs, means)
        want = means.dtype if means is not None else dtype
        return out if out.dtype == want else out.to(want)

    @classmethod
    def _dequantize(cls, inputs: Tensor, means: Optional[Tensor] = None) -> Tensor:
        warnings.warn("_dequantize. Use dequantize instead.")
        return cls.dequantize(inputs, means)

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        """entropy_models.py:206-214 (on the device when the model lives there: one kernel for the whole table)"""
        if pmf.is_cuda:
            return ops.pmf_to_quantized_cdf_device(pmf[:, :max_length], tail_mass, pmf_length.to(pmf.device), max_length,
                                                   self.entropy_coder_precision)
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        pmf, tail_mass = pmf.detach().cpu(), tail_mass.detach().cpu()
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            _cdf = torch.IntTensor(ops.pmf_to_quantized_cdf(prob.numpy(), self.entropy_coder_precision))
            cdf[i, : _cdf.size(0)] = _cdf
        return cdf.to(self._quantized_cdf.device)

    def _check_cdf_size(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if len(self._quantized_cdf.size()) != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")

    def _check_offsets_size(self):
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if len(self._offset.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")

    def _check_cdf_length(self):
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if len(self._cdf_length.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    def symbols_and_indexes(self, inputs: Tensor, indexes: Tensor, means: Optional[Tensor] = None):
        """The validated int32 (symbols, indexes) pair that ``compress`` hands to the entropy coder
        (entropy_models.py:237-258): everything of ``compress`` that runs per element."""
        symbols = self.quantize(inputs, "symbols", means)
        if len(inputs.size()) < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        return symbols, indexes.int()

    def compress(self, inputs, indexes, means=None):
        """entropy_models.py:237-270: one rANS stream per image, byte-identical to the reference coder."""
        symbols, indexes = self.symbols_and_indexes(inputs, indexes, means)
        return self.encode_symbols(symbols, indexes)

    # Entropy coders: "ans" = the reference's stream (compressai.ans, one serial rANS chain per image; coded on host threads,
    # byte-identical); "ans-lanes" = the same symbols / tables / escape scheme in libmmcodec's own lane container, coded ON THE
    # DEVICE (csrc/rans_device.cu) -- not readable by the reference, chosen explicitly: EntropyModel(entropy_coder="ans-lanes") or
    # mmcodec.set_entropy_coder(net, "ans-lanes").
    CODERS = ("ans", "ans-lanes")

    def _coder(self) -> str:
        if self.entropy_coder_name not in self.CODERS:
            raise NotImplementedError(f'entropy coder "{self.entropy_coder_name}" is not available; libmmcodec implements {self.CODERS}')
        return self.entropy_coder_name

    def encode_symbols(self, symbols, indexes):
        """int32 symbols / indexes (B, ...) -> list of B byte strings with this model's tables and coder."""
        if self._coder() == "ans-lanes":
            return ops.rans_encode_device(symbols, indexes, self._quantized_cdf, self._cdf_length, self._offset)
        return ops.rans_encode(symbols, indexes, self._quantized_cdf, self._cdf_length, self._offset)

    def decode_symbols(self, strings, indexes):
        if self._coder() == "ans-lanes":
            return ops.rans_decode_device(list(strings), indexes.int(), self._quantized_cdf, self._cdf_length, self._offset)
        return ops.rans_decode(list(strings), indexes.int(), self._quantized_cdf, self._cdf_length, self._offset)

    def decompress(self, strings, indexes, dtype: torch.dtype = torch.float, means: Tensor = None):
        """entropy_models.py:272-327"""
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if len(indexes.size()) < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        if means is not None:
            if means.size()[:2] != indexes.size()[:2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size():
                for i in range(2, len(indexes.size())):
                    if means.size(i) != 1:
                        raise ValueError("Invalid means parameters")
        symbols = self.decode_symbols(strings, indexes)
        return self.dequantize(symbols, means, dtype)


class EntropyBottleneck(EntropyModel):
    """Factorised-prior entropy bottleneck (entropy_models.py:330-574)."""

    _offset: Tensor

    def __init__(self, channels: int, *args: Any, tail_mass: float = 1e-9, init_scale: float = 10,
                 filters: Tuple[int, ...] = (3, 3, 3, 3), **kwargs: Any):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        if self.filters != (3, 3, 3, 3):
            raise NotImplementedError("libmmcodec implements the default EntropyBottleneck filters (3, 3, 3, 3) only")

        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        channels = self.channels
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            matrix = torch.Tensor(channels, filters[i + 1], filters[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))
            bias = torch.Tensor(channels, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))
            if i < len(self.filters):
                factor = torch.Tensor(channels, filters[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))

        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2]

    def _params(self):
        """Device pointers of the raw parameter blocks (no packing pass: the kernel applies
        softplus / tanh itself, so there is nothing to invalidate when the parameters change)."""
        return ops.make_eb_params([getattr(self, f"_matrix{i}") for i in range(5)],
                                  [getattr(self, f"_bias{i}") for i in range(5)],
                                  [getattr(self, f"_factor{i}") for i in range(4)],
                                  self.quantiles[:, 0, 1])

    def _eval_lut(self):
        """Likelihood table for the eval fast path, rebuilt whenever a parameter (or its storage / device) changes."""
        ps = [getattr(self, f"_matrix{i}") for i in range(5)] + [getattr(self, f"_bias{i}") for i in range(5)] + \
             [getattr(self, f"_factor{i}") for i in range(4)] + [self.quantiles]
        key = tuple((t._version, t.data_ptr()) for t in ps) + (self._lik_bound(),)
        if getattr(self, "_lut_key", None) != key:
            self._lut = ops.eb_build_lut(self._params(), self.channels, self._lik_bound(), self.quantiles.device)
            self._lut_key = key
        return self._lut

    # ---- once-per-model table construction (entropy_models.py:396-441) -------------------------
    def _logits_cumulative_host(self, inputs: Tensor) -> Tensor:
        logits = inputs
        for i in range(len(self.filters) + 1):
            logits = torch.matmul(torch.nn.functional.softplus(getattr(self, f"_matrix{i}").detach()), logits)
            logits = logits + getattr(self, f"_bias{i}").detach()
            if i < len(self.filters):
                logits = logits + torch.tanh(getattr(self, f"_factor{i}").detach()) * torch.tanh(logits)
        return logits

    def update(self, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        medians = self.quantiles[:, 0, 1].detach()
        minima = torch.clamp(torch.ceil(medians - self.quantiles[:, 0, 0].detach()).int(), min=0)
        maxima = torch.clamp(torch.ceil(self.quantiles[:, 0, 2].detach() - medians).int(), min=0)
        self._offset = -minima
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max().item())
        samples = torch.arange(max_length, device=pmf_start.device)
        samples = samples[None, :] + pmf_start[:, None, None]
        lower = self._logits_cumulative_host(samples - 0.5)
        upper = self._logits_cumulative_host(samples + 0.5)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._cdf_length = pmf_length + 2
        return True

    def loss(self) -> Tensor:
        """entropy_models.py:450-454: gradient flows to ``quantiles`` only (the reference evaluates the cumulative with
        stop_gradient=True on the density parameters)."""
        if torch.is_grad_enabled() and self.quantiles.requires_grad:
            # (C, 1, 3) values: host-side glue on 3 numbers per channel, same torch ops as the reference
            logits = self._logits_cumulative_host(self.quantiles)
            return torch.abs(logits - self.target).sum()
        q = self.quantiles.detach()                                  # (C, 1, 3)
        logits = ops.eb_logits_cumulative(q.permute(1, 0, 2).contiguous(), self._params())   # (1, C, 3)
        return torch.abs(logits.permute(1, 0, 2) - self.target).sum()

    def _likelihood(self, inputs: Tensor) -> Tensor:
        """entropy_models.py:480-492 on a (C, 1, L) tensor, as in the reference."""
        x = inputs.permute(1, 0, 2)                                   # (1, C, L): channel is dim 1
        p = self._params()
        lower = ops.eb_logits_cumulative(x - 0.5, p)
        upper = ops.eb_logits_cumulative(x + 0.5, p)
        sign = -torch.sign(lower + upper)
        lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
        return lik.permute(1, 0, 2)

    def forward(self, x: Tensor, training: Optional[bool] = None) -> Tuple[Tensor, Tensor]:
        """entropy_models.py:495-540.  One kernel; no permute copies (the kernel indexes channels in
        place for both NCHW and channels-last memory)."""
        if training is None:
            training = self.training
        if x.dim() < 2:
            raise ValueError("EntropyBottleneck expects (N, C, ...) inputs")
        if x.shape[1] != self.channels:
            raise ValueError(f"expected {self.channels} channels, got {x.shape[1]}")
        if training:
            # differentiable like the reference's module (rate gradient to the caller's transforms and to the density
            # parameters): recorded by mmcodec.autograd._EbFn when autograd is on and anything requires grad
            from . import autograd as AG
            with torch.no_grad():
                noise = torch.empty_like(x, dtype=torch.float32).uniform_(-0.5, 0.5)
            return AG.eb_forward(x, self, noise)
        # eval mode ("dequantize"): round() has a zero gradient, so nothing flows to the input in the reference either; the
        # density parameters are not trained in eval mode -- no graph is recorded
        with torch.no_grad():
            return ops.eb_forward(x, self._params(), None, self._lik_bound(), lut=self._eval_lut())

    @staticmethod
    def _build_indexes(size, device=None):
        """entropy_models.py:542-553"""
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        return ops.channel_indexes(size, device)

    @staticmethod
    def _extend_ndims(tensor, n):
        return tensor.reshape(-1, *([1] * n)) if n > 0 else tensor.reshape(-1)

    def symbols_and_indexes(self, x: Tensor, indexes: Tensor = None, means: Tensor = None):  # type: ignore[override]
        """EntropyBottleneck.compress up to the coder call (entropy_models.py:559-566)."""
        if indexes is None:
            indexes = self._build_indexes(x.size(), x.device)
        if means is None:
            spatial_dims = len(x.size()) - 2
            medians = self._extend_ndims(self._get_medians().detach(), spatial_dims)
            means = medians.expand(x.size(0), *([-1] * (spatial_dims + 1)))
        return super().symbols_and_indexes(x, indexes, means)

    def _expanded_medians(self, batch: int, spatial_dims: int) -> Tensor:
        medians = self._extend_ndims(self._get_medians().detach(), spatial_dims)
        return medians.expand(batch, *([-1] * (spatial_dims + 1)))

    def compress(self, x):
        """entropy_models.py:559-566"""
        indexes = self._build_indexes(x.size(), x.device)
        return super().compress(x, indexes, self._expanded_medians(x.size(0), len(x.size()) - 2))

    def decompress(self, strings, size):
        """entropy_models.py:568-574"""
        output_size = (len(strings), self._quantized_cdf.size(0), *size)
        indexes = self._build_indexes(output_size, self._quantized_cdf.device)
        medians = self._expanded_medians(len(strings), len(size))
        return super().decompress(strings, indexes, medians.dtype, medians)


class GaussianConditional(EntropyModel):
    """Gaussian conditional layer (entropy_models.py:577-740)."""

    def __init__(self, scale_table: Optional[Union[List, Tuple]], *args: Any, scale_bound: float = 0.11,
                 tail_mass: float = 1e-9, **kwargs: Any):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = self.scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]) if scale_bound is not None else None)

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    @staticmethod
    def _standardized_quantile(quantile):
        return scipy.stats.norm.ppf(quantile)

    def update_scale_table(self, scale_table, force=False):
        """entropy_models.py:643-652"""
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    def update(self):
        """entropy_models.py:655-679: pmf over |k| <= ceil(6.11 sigma) per scale -> quantized CDF tables."""
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        table = self.scale_table.detach().cpu()
        pmf_center = torch.ceil(table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(torch.max(pmf_length).item())
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        samples_scale = table.unsqueeze(1).float()
        const = float(-(2 ** -0.5))
        upper = 0.5 * torch.erfc(const * ((0.5 - samples) / samples_scale))
        lower = 0.5 * torch.erfc(const * ((-0.5 - samples) / samples_scale))
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        dev = self.scale_table.device
        # the pmf itself is evaluated with the reference's CPU ops (bit-identical tables); the integer CDF construction runs on
        # the device when the model lives there
        self._quantized_cdf = self._pmf_to_cdf(pmf.to(dev), tail_mass.to(dev), pmf_length, max_length).to(dev)
        self._offset = (-pmf_center).to(dev)
        self._cdf_length = (pmf_length + 2).to(dev)

    def _likelihood(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None) -> Tensor:
        """entropy_models.py:692-709: likelihood AT the given values (the reference does not re-quantise here: a caller may pass
        noisy values), no lower bound.  The fused kernel's additive-noise mode with a zero noise tensor evaluates exactly that:
        v = inputs + 0.  (The forward kernel uses the fast erfc of csrc/entropy.cu, fractional error < 1.2e-7; gc_backward
        differentiates the exact erfcf -- the two agree to that error.)"""
        zero = torch.zeros_like(inputs, dtype=torch.float32)
        _, lik = ops.gc_forward(inputs, scales, means, zero, self.lower_bound_scale._sync_bound(), 0.0)
        return lik

    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                training: Optional[bool] = None) -> Tuple[Tensor, Tensor]:
        """entropy_models.py:715-731, one kernel."""
        if training is None:
            training = self.training
        if inputs.shape != scales.shape:
            raise ValueError("`inputs` and `scales` should have the same size.")
        from . import autograd as AG
        with torch.no_grad():
            noise = torch.empty_like(inputs, dtype=torch.float32).uniform_(-0.5, 0.5) if training else None
        # differentiable like the reference's module: mmcodec.autograd._GcFn records the call when autograd is on and inputs /
        # scales / means require grad (noise mode: gradients to all three; eval mode: to the scales only, round() has none)
        return AG.gc_forward(inputs, scales, means, noise, self.lower_bound_scale._sync_bound(), self._lik_bound())

    def build_indexes(self, scales: Tensor) -> Tensor:
        """entropy_models.py:735-740, one kernel instead of 63 x 3 launches."""
        if self.scale_table.numel() == 0:
            # the reference returns all -1 here (len(table) - 1 with an empty table); same value, no kernel
            return torch.full(scales.shape, -1, dtype=torch.int32, device=scales.device)
        key = (self.scale_table._version, self.scale_table.data_ptr())
        if getattr(self, "_table_checked", None) != key:
            t = self.scale_table.detach().cpu()
            if t.numel() > 256 or bool((t[1:] < t[:-1]).any()):
                raise ValueError("scale_table must be sorted ascending with at most 256 levels")
            self._table_checked = key
        return ops.build_indexes(scales, self.scale_table, self.lower_bound_scale._sync_bound())

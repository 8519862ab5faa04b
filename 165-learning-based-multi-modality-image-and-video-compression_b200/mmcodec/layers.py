"""Host-side mirror of ``compressai.layers`` / ``compressai.ops`` / ``compressai.models.utils`` for
the hot path: same class names, constructor arguments, parameter / buffer names (``state_dict``
contract) and error behaviour; ``forward`` runs libmmcodec kernels.

Round-1 scope is the inference path: forwards run without recording autograd graphs, except
``LowerBound`` whose custom gradient (compressai/ops/bound_ops.py:40-56) is also a kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch import Tensor

from . import _lib as L
from . import ops

__all__ = ["LowerBound", "NonNegativeParametrizer", "GDN", "conv", "deconv", "Conv2d", "ConvTranspose2d"]


class _LowerBoundFunction(torch.autograd.Function):
    """compressai/ops/bound_ops.py:45-56"""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x)
        ctx.bound = float(bound)
        return ops.lower_bound(x, ctx.bound)

    @staticmethod
    def backward(ctx, grad_output):
        (x,) = ctx.saved_tensors
        return ops.lower_bound_bwd(x, grad_output, ctx.bound), None


class LowerBound(nn.Module):
    """``torch.max(x, bound)`` with pass-through gradient towards the bound (bound_ops.py:59-80)."""

    bound: Tensor

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))
        self._bound_value = float(self.bound.item())
        self._bound_seen = self.bound._version

    def forward(self, x: Tensor) -> Tensor:
        return _LowerBoundFunction.apply(x, self._sync_bound())

    def __prepare_scriptable__(self):
        # torch.jit.script compiles a stand-in whose forward is the registered op mmcodec::lower_bound (mmcodec/library.py)
        from .library import ScriptableLowerBound
        return ScriptableLowerBound(self)

    def _sync_bound(self) -> float:
        # the buffer may have been overwritten by load_state_dict; read it back once per version
        v = (self.bound._version, self.bound.data_ptr())
        if self._bound_seen != v:
            self._bound_value = float(self.bound.item())
            self._bound_seen = v
        return self._bound_value


class NonNegativeParametrizer(nn.Module):
    """Non-negative re-parametrisation (compressai/ops/parametrizers.py:38-64)."""

    pedestal: Tensor

    def __init__(self, minimum: float = 0, reparam_offset: float = 2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        bound = (self.minimum + self.reparam_offset ** 2) ** 0.5
        self.lower_bound = LowerBound(bound)

    def init(self, x: Tensor) -> Tensor:
        # construction-time initialisation of a parameter (host side, parametrizers.py:58-59)
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x: Tensor) -> Tensor:
        out = self.lower_bound(x)
        return out ** 2 - self.pedestal

    def __prepare_scriptable__(self):
        from .library import ScriptableNonNegativeParametrizer
        return ScriptableNonNegativeParametrizer(self)


class GDN(nn.Module):
    r"""Generalized Divisive Normalization (compressai/layers/gdn.py:40-92).

    y[i] = x[i] / sqrt(beta[i] + sum_j gamma[i, j] * x[j]^2)      (inverse: multiply by the sqrt)
    """

    def __init__(self, in_channels: int, inverse: bool = False, beta_min: float = 1e-6, gamma_init: float = 0.1):
        super().__init__()
        beta_min = float(beta_min)
        gamma_init = float(gamma_init)
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=beta_min)
        beta = torch.ones(in_channels)
        self.beta = nn.Parameter(self.beta_reparam.init(beta))
        self.gamma_reparam = NonNegativeParametrizer()
        gamma = gamma_init * torch.eye(in_channels)
        self.gamma = nn.Parameter(self.gamma_reparam.init(gamma))
        self._cache_key = None
        self._cache = None

    def effective_params(self, want_bf16: bool = True):
        """(beta_eff f32 [C], gamma_eff f32 [C][C], gamma_eff bf16) cached per parameter version."""
        key = (self.beta._version, self.gamma._version, self.beta.data_ptr(), self.gamma.data_ptr())
        if self._cache_key != key:
            br, gr = self.beta_reparam, self.gamma_reparam
            self._cache = ops.gdn_reparam(self.beta, self.gamma,
                                          (br.minimum + br.reparam_offset ** 2) ** 0.5,
                                          (gr.minimum + gr.reparam_offset ** 2) ** 0.5,
                                          br.reparam_offset ** 2, want_bf16=True)
            self._cache_key = key
        return self._cache

    def forward(self, x: Tensor) -> Tensor:
        if x.dim() != 4:
            raise ValueError("GDN expects a 4-D (B, C, H, W) tensor")  # `_, C, _, _ = x.size()` in gdn.py:78
        # differentiable like the reference's module: with autograd on and a parameter or the input requiring grad the call is
        # recorded (mmcodec.autograd._GdnFn); otherwise it is the plain forward kernel
        from . import autograd as AG
        return AG.gdn_forward(self, x)

    def __prepare_scriptable__(self):
        # compressai tests/test_scripting.py:37-59 scripts GDN: the scripted module is one call of the registered op mmcodec::gdn
        # on the same Parameters (mmcodec/library.py)
        from .library import ScriptableGDN
        return ScriptableGDN(self)


class _PackedWeightMixin:
    """Caches the bf16 tap-major weight pack the tensor-core kernels stream (derived data, rebuilt
    whenever the Parameter's version or storage changes, e.g. after load_state_dict / optimizer step)."""

    def packed_weight(self, d: L.ConvDesc):
        key = (self.weight._version, self.weight.data_ptr(), d.in_layout, d.out_layout, d.transposed)
        if getattr(self, "_pack_key", None) != key:
            self._pack = ops.conv_pack_weights(d, self.weight)
            self._pack_key = key
        return self._pack

    def f32_weight(self):
        w = self.weight.detach()
        return w if (w.dtype == torch.float32 and w.is_contiguous()) else w.float().contiguous()


class Conv2d(nn.Conv2d, _PackedWeightMixin):
    """nn.Conv2d whose forward runs the mmcodec kernels (NCHW fp32 in -> logical NCHW fp32 out)."""

    def forward(self, x: Tensor) -> Tensor:
        from .transforms import run_layers
        return run_layers([self], x, "nchw_f32", "nchw_f32")


class ConvTranspose2d(nn.ConvTranspose2d, _PackedWeightMixin):
    def forward(self, x: Tensor, output_size=None) -> Tensor:
        from .transforms import run_layers
        return run_layers([self], x, "nchw_f32", "nchw_f32")


def conv(in_channels, out_channels, kernel_size=5, stride=2):
    """compressai/models/utils.py:128-135"""
    return Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(in_channels, out_channels, kernel_size=5, stride=2):
    """compressai/models/utils.py:138-146"""
    return ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                           output_padding=stride - 1, padding=kernel_size // 2)

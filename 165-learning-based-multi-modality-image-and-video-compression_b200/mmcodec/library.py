"""``torch.library`` registration of the stand-alone layer ops, so that TorchScript survives the drop-in.

The reference scripts its layers (``torch.jit.script(GDN(128))``, compressai tests/test_scripting.py:37-59).  A kernel-backed
``forward`` that marshals pointers through ctypes is not TorchScript; an operator with a registered schema is.  This module

  * defines ``mmcodec::gdn`` and ``mmcodec::lower_bound`` (schemas below) with CUDA implementations on libmmcodec, fake (meta)
    implementations for tracing / ``torch.compile`` graph capture by callers, and autograd formulas on the same backward kernels the
    eager modules use (compressai/layers/gdn.py:77-92, compressai/ops/bound_ops.py:45-56);
  * gives ``mmcodec.GDN`` / ``mmcodec.LowerBound`` / ``mmcodec.NonNegativeParametrizer`` a ``__prepare_scriptable__`` hook:
    ``torch.jit.script(module)`` scripts a small stand-in that SHARES the module's Parameters and buffers (same ``state_dict``
    keys as the reference's scripted module) and whose ``forward`` is one call of the registered op.

There is no CPU kernel behind the ops: a CPU tensor raises ``NotImplementedError`` from the dispatcher (no CPU path, by design).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn
from torch import Tensor

from . import ops

__all__ = ["ScriptableGDN", "ScriptableLowerBound", "ScriptableNonNegativeParametrizer"]


# ---------------------------------------------------------------------------------------------
# operators
# ---------------------------------------------------------------------------------------------
@torch.library.custom_op("mmcodec::lower_bound", mutates_args=(), device_types="cuda")
def _lower_bound_op(x: Tensor, bound: float) -> Tensor:
    return ops.lower_bound(x, bound).reshape(x.shape)


@_lower_bound_op.register_fake
def _(x: Tensor, bound: float) -> Tensor:
    return torch.empty(x.shape, dtype=torch.float32, device=x.device)


@torch.library.custom_op("mmcodec::lower_bound_bwd", mutates_args=(), device_types="cuda")
def _lower_bound_bwd_op(x: Tensor, grad_out: Tensor, bound: float) -> Tensor:
    return ops.lower_bound_bwd(x, grad_out, bound).reshape(x.shape)


@_lower_bound_bwd_op.register_fake
def _(x: Tensor, grad_out: Tensor, bound: float) -> Tensor:
    return torch.empty(x.shape, dtype=torch.float32, device=x.device)


def _lb_setup(ctx, inputs, output):
    x, bound = inputs
    ctx.save_for_backward(x)
    ctx.bound = float(bound)


def _lb_backward(ctx, grad_out):
    (x,) = ctx.saved_tensors
    return torch.ops.mmcodec.lower_bound_bwd(x, grad_out.contiguous(), ctx.bound), None


_lower_bound_op.register_autograd(_lb_backward, setup_context=_lb_setup)


@torch.library.custom_op("mmcodec::gdn", mutates_args=(), device_types="cuda")
def _gdn_op(x: Tensor, beta: Tensor, gamma: Tensor, beta_bound: float, gamma_bound: float, pedestal: float, inverse: bool) -> Tensor:
    """y = x / sqrt(beta' + gamma' x^2) (inverse: x * sqrt) with beta' = max(beta, beta_bound)^2 - pedestal, gamma' likewise: the raw
    Parameters go in, the non-negative re-parametrisation (parametrizers.py:61-64) runs on the device like the rest."""
    beta_eff, gamma_eff, _ = ops.gdn_reparam(beta, gamma, beta_bound, gamma_bound, pedestal)
    return ops.gdn_forward(x, beta_eff, gamma_eff, inverse)


@_gdn_op.register_fake
def _(x, beta, gamma, beta_bound, gamma_bound, pedestal, inverse):
    return torch.empty_like(x, dtype=torch.float32)


@torch.library.custom_op("mmcodec::gdn_bwd", mutates_args=(), device_types="cuda")
def _gdn_bwd_op(x: Tensor, grad_out: Tensor, beta: Tensor, gamma: Tensor, beta_bound: float, gamma_bound: float, pedestal: float,
                inverse: bool) -> Tuple[Tensor, Tensor, Tensor]:
    # the eager module's backward (mmcodec.autograd._GdnFn) on a parameter-holding stand-in for the module
    from . import autograd as AG
    from .layers import GDN, NonNegativeParametrizer
    g = GDN.__new__(GDN)
    nn.Module.__init__(g)
    g.inverse = bool(inverse)
    offset = float(pedestal) ** 0.5
    g.beta_reparam = NonNegativeParametrizer(minimum=float(beta_bound) ** 2 - float(pedestal), reparam_offset=offset)
    g.gamma_reparam = NonNegativeParametrizer(minimum=float(gamma_bound) ** 2 - float(pedestal), reparam_offset=offset)
    g.__dict__["beta"], g.__dict__["gamma"] = beta.detach(), gamma.detach()
    g._cache_key = g._cache = None
    x_pre = ops.nchw_to_nhwc_bf16(x.detach().float().contiguous())
    gy = ops.nchw_to_nhwc_bf16(grad_out.detach().float().contiguous())
    dx, dbeta, dgamma = AG._gdn_backward(g, x_pre, gy)
    return dx.permute(0, 3, 1, 2).float().contiguous(), dbeta.to(beta.dtype), dgamma.to(gamma.dtype)


@_gdn_bwd_op.register_fake
def _(x, grad_out, beta, gamma, beta_bound, gamma_bound, pedestal, inverse):
    return torch.empty_like(x, dtype=torch.float32), torch.empty_like(beta), torch.empty_like(gamma)


def _gdn_setup(ctx, inputs, output):
    x, beta, gamma, beta_bound, gamma_bound, pedestal, inverse = inputs
    ctx.save_for_backward(x, beta, gamma)
    ctx.consts = (float(beta_bound), float(gamma_bound), float(pedestal), bool(inverse))


def _gdn_backward_formula(ctx, grad_out):
    x, beta, gamma = ctx.saved_tensors
    dx, dbeta, dgamma = torch.ops.mmcodec.gdn_bwd(x, grad_out.contiguous(), beta, gamma, *ctx.consts)
    return dx, dbeta, dgamma, None, None, None, None


_gdn_op.register_autograd(_gdn_backward_formula, setup_context=_gdn_setup)


# ---------------------------------------------------------------------------------------------
# scriptable stand-ins (what torch.jit.script compiles in place of the ctypes-driven modules)
# ---------------------------------------------------------------------------------------------
class ScriptableLowerBound(nn.Module):
    """compressai/ops/bound_ops.py:59-80 as one registered op; shares the ``bound`` buffer."""

    def __init__(self, m: nn.Module):
        super().__init__()
        self._buffers["bound"] = m.bound
        self.bound_value: float = float(m.bound.item())

    def forward(self, x: Tensor) -> Tensor:
        return torch.ops.mmcodec.lower_bound(x, self.bound_value)


class ScriptableNonNegativeParametrizer(nn.Module):
    """compressai/ops/parametrizers.py:38-64; shares ``pedestal`` and ``lower_bound.bound``."""

    def __init__(self, m: nn.Module):
        super().__init__()
        self._buffers["pedestal"] = m.pedestal
        self.lower_bound = ScriptableLowerBound(m.lower_bound)
        self.pedestal_value: float = float(m.pedestal.item())

    def forward(self, x: Tensor) -> Tensor:
        out = self.lower_bound(x)
        return out * out - self.pedestal_value


class ScriptableGDN(nn.Module):
    """compressai/layers/gdn.py:40-92 as one registered op; shares ``beta`` / ``gamma`` (the same Parameter objects) and the
    re-parametrisation buffers, so ``state_dict()`` of the scripted module has the reference's keys."""

    def __init__(self, m: nn.Module):
        super().__init__()
        self._parameters["beta"] = m.beta
        self._parameters["gamma"] = m.gamma
        self.beta_reparam = ScriptableNonNegativeParametrizer(m.beta_reparam)
        self.gamma_reparam = ScriptableNonNegativeParametrizer(m.gamma_reparam)
        self.inverse: bool = bool(m.inverse)
        self.beta_bound: float = float(m.beta_reparam.lower_bound.bound.item())
        self.gamma_bound: float = float(m.gamma_reparam.lower_bound.bound.item())
        self.pedestal: float = float(m.beta_reparam.pedestal.item())

    def forward(self, x: Tensor) -> Tensor:
        return torch.ops.mmcodec.gdn(x, self.beta, self.gamma, self.beta_bound, self.gamma_bound, self.pedestal, self.inverse)

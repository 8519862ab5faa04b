"""``mmcodec.accelerate(model)`` -- the drop-in behind LIVE ``compressai`` modules.

The mirror classes of this package are complete models; this function is for a model object that was built by the reference
itself (``compressai.zoo`` / ``compressai.models.*`` or a user's subclass): it swaps, in place, the hot-path submodules for
their kernel-backed counterparts and leaves everything else -- the model class, its ``forward`` / ``compress`` / ``decompress``
Python code, ``state_dict`` keys, optimizers, checkpoints -- untouched:

  * ``compressai.layers.GDN``                                -> ``mmcodec.GDN``
  * ``compressai.entropy_models.EntropyBottleneck``          -> ``mmcodec.EntropyBottleneck``
  * ``compressai.entropy_models.GaussianConditional``        -> ``mmcodec.GaussianConditional``
  * ``compressai.ops.LowerBound`` / ``NonNegativeParametrizer`` inside them -> the mirrors (same buffers)
  * ``nn.Sequential`` made only of conv()/deconv()-style layers, GDN, ReLU, LeakyReLU (g_a, g_s, h_a, h_s)
                                                              -> ``mmcodec.transforms.TransformStack`` (fused launches)
  * stand-alone conv()/deconv()-style ``nn.Conv2d`` / ``nn.ConvTranspose2d`` -> ``mmcodec.layers.Conv2d`` / ``ConvTranspose2d``

Every replacement ADOPTS the original's ``__dict__``: the same ``Parameter`` and buffer objects (so an optimizer built before the
call keeps working and ``state_dict()`` is bit-identical), the same hyper-parameters.  The replacement's class is created on the
fly as ``type(name, (MirrorClass, OriginalClass), {})``, so ``isinstance(m, compressai.entropy_models.EntropyBottleneck)`` --
which ``CompressionModel.update()`` and ``aux_loss()`` rely on (compressai/models/google.py:100-125) -- still holds, while method
lookup finds the mirror first.  Modules are matched structurally (class name + the attributes the mirror needs), so no
``compressai`` import is required here.
"""
from __future__ import annotations

from typing import Dict, Tuple, Type

import torch.nn as nn

from . import entropy_models as EM
from . import layers as LY
from .transforms import TransformStack, parse_layers

__all__ = ["accelerate"]

_CLASS_CACHE: Dict[Tuple[type, type], type] = {}


def _hybrid(mirror: Type[nn.Module], original: Type[nn.Module]) -> Type[nn.Module]:
    """mirror-first subclass of both: mirror methods, ``isinstance`` of the original"""
    if issubclass(mirror, original):
        return mirror
    key = (mirror, original)
    cls = _CLASS_CACHE.get(key)
    if cls is None:
        cls = _CLASS_CACHE[key] = type(original.__name__, (mirror, original), {"__module__": mirror.__module__, "_mmc_accelerated": True})
    return cls


def _adopt(module: nn.Module, mirror: Type[nn.Module]) -> nn.Module:
    new = object.__new__(_hybrid(mirror, type(module)))
    new.__dict__.update(module.__dict__)          # shares _parameters / _buffers and every hyper-parameter
    new.__dict__["_modules"] = dict(module._modules)     # own child table: nested replacements must not touch the original object
    return new


def _is(module: nn.Module, name: str, *attrs: str) -> bool:
    return type(module).__name__ == name and not getattr(type(module), "_mmc_accelerated", False) \
        and not type(module).__module__.startswith("mmcodec") and all(hasattr(module, a) for a in attrs)


def _lower_bound(m: nn.Module) -> nn.Module:
    new = _adopt(m, LY.LowerBound)
    new._bound_value = float(m.bound.item())
    new._bound_seen = None
    return new


def _parametrizer(m: nn.Module) -> nn.Module:
    new = _adopt(m, LY.NonNegativeParametrizer)
    if not hasattr(new, "minimum"):
        # reference revisions that keep only the bound: minimum = bound^2 - pedestal
        new.reparam_offset = float(m.pedestal.item()) ** 0.5
        new.minimum = float(m.lower_bound.bound.item()) ** 2 - float(m.pedestal.item())
    new._modules["lower_bound"] = _lower_bound(m.lower_bound)
    return new


def _gdn(m: nn.Module) -> nn.Module:
    new = _adopt(m, LY.GDN)
    new._modules["beta_reparam"] = _parametrizer(m.beta_reparam)
    new._modules["gamma_reparam"] = _parametrizer(m.gamma_reparam)
    new._cache_key, new._cache = None, None
    return new


def _entropy_common(new: nn.Module, m: nn.Module) -> None:
    coder = getattr(m, "entropy_coder", None)
    new.entropy_coder_name = getattr(coder, "name", None) or getattr(m, "entropy_coder_name", "ans")
    new.use_likelihood_bound = bool(getattr(m, "use_likelihood_bound", hasattr(m, "likelihood_lower_bound")))
    if new.use_likelihood_bound:
        new._modules["likelihood_lower_bound"] = _lower_bound(m.likelihood_lower_bound)
        new.likelihood_bound_value = float(m.likelihood_lower_bound.bound.item())
    else:
        new.likelihood_bound_value = 0.0


def _entropy_bottleneck(m: nn.Module) -> nn.Module:
    if tuple(m.filters) != (3, 3, 3, 3):
        raise NotImplementedError("libmmcodec implements the default EntropyBottleneck filters (3, 3, 3, 3) only")
    new = _adopt(m, EM.EntropyBottleneck)
    _entropy_common(new, m)
    return new


def _gaussian_conditional(m: nn.Module) -> nn.Module:
    new = _adopt(m, EM.GaussianConditional)
    _entropy_common(new, m)
    new._modules["lower_bound_scale"] = _lower_bound(m.lower_bound_scale)
    return new


def _conv_like(m: nn.Module) -> bool:
    """conv() / deconv() of compressai/models/utils.py:128-146 (what parse_layers accepts)"""
    if type(m) not in (nn.Conv2d, nn.ConvTranspose2d):
        return False
    try:
        parse_layers([m])
        return True
    except NotImplementedError:
        return False


def _conv(m: nn.Module) -> nn.Module:
    return _adopt(m, LY.ConvTranspose2d if isinstance(m, nn.ConvTranspose2d) else LY.Conv2d)


def _convert(m: nn.Module):
    """replacement for ``m`` or None"""
    if _is(m, "GDN", "beta", "gamma", "inverse", "beta_reparam", "gamma_reparam"):
        return _gdn(m)
    if _is(m, "EntropyBottleneck", "_matrix0", "_bias0", "_factor0", "quantiles", "filters", "channels"):
        return _entropy_bottleneck(m)
    if _is(m, "GaussianConditional", "scale_table", "lower_bound_scale", "tail_mass"):
        return _gaussian_conditional(m)
    if _is(m, "LowerBound", "bound"):
        return _lower_bound(m)
    if _conv_like(m):
        return _conv(m)
    if type(m) is nn.Sequential and len(m) > 0:
        kids = [(_convert(c) or c) for c in m]
        try:
            parse_layers(kids)
        except NotImplementedError:
            return None                              # not a pure transform stack: handled child by child by the caller
        return TransformStack(*kids)
    return None


def accelerate(model: nn.Module) -> nn.Module:
    """Swap the hot-path submodules of a live reference model for their libmmcodec counterparts, in place; returns ``model``.
    The swapped modules compute on CUDA tensors only -- there is no CPU path."""
    replaced = 0

    def walk(parent: nn.Module):
        nonlocal replaced
        for name, child in list(parent._modules.items()):
            if child is None:
                continue
            new = _convert(child)
            if new is not None:
                parent._modules[name] = new
                replaced += 1
            else:
                walk(child)

    top = _convert(model)
    if top is not None:
        return top
    walk(model)
    model._mmc_accelerated_modules = replaced
    return model

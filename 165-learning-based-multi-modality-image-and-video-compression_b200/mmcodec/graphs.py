"""CUDA-graph capture of a whole forward pass.

Small-batch and sequential workloads (one 1080p frame at a time in ssf2020: ~450 kernel launches per GOP, each a few tens
of microseconds of device work) are bound by host-side launch cost, not by the kernels.  ``GraphedForward`` runs a callable
once under ``torch.cuda.graph`` -- every libmmcodec launch (tensor maps included: they are kernel parameters) is recorded
on the capturing stream -- and replays it with one ``cudaGraphLaunch`` per call.  Inputs are copied into the captured
input buffers; outputs are the captured output tensors (overwritten by the next call).

The graph reads the packed weights, GDN re-parametrisations and likelihood tables that existed at capture time.  Pass the modules
whose parameters matter as ``modules=`` and the graph is re-captured automatically after any in-place parameter update.
"""
from __future__ import annotations

from typing import Any, Callable

import torch
from torch.utils import _pytree as pytree

__all__ = ["GraphedForward"]


def module_signature(modules):
    """What a captured graph of these modules bakes in: version AND storage of every parameter and buffer (an optimizer step bumps
    the version; ``net.to(...)``, ``p.data = ...`` or ``load_state_dict(assign=True)`` change the storage without touching it; the
    scale table, the LowerBound bounds and the MaskedConv mask are buffers) plus the training flags.  The same key the per-module
    caches (packed weights, LUTs, GDN re-parametrisation) use."""
    sig = []
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            sig.append((t._version, t.data_ptr()))
        sig.append(m.training)
    return tuple(sig)


class GraphedForward:
    def __init__(self, fn: Callable[..., Any], *example_inputs: Any, warmup: int = 2, modules=None):
        if modules is None and isinstance(fn, torch.nn.Module):
            modules = [fn]
        self._modules = list(modules or [])
        self._warmup = warmup
        self._capture(fn, example_inputs)

    def _signature(self):
        return module_signature(self._modules)

    def _capture(self, fn, example_inputs):
        warmup = self._warmup
        self._sig = self._signature()
        flat, self._spec = pytree.tree_flatten(example_inputs)
        for t in flat:
            if torch.is_tensor(t) and not t.is_cuda:
                raise RuntimeError("GraphedForward captures CUDA work only; got a CPU tensor (there is no CPU path)")
        self._static_in = [t.clone() if torch.is_tensor(t) else t for t in flat]
        self._fn = fn
        args = pytree.tree_unflatten(self._static_in, self._spec)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):      # fills every per-parameter cache (packed weights, LUTs, GDN reparam) before capture
                fn(*args)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self._static_out = fn(*args)

    def __call__(self, *inputs: Any) -> Any:
        if self._modules and self._signature() != self._sig:
            self._capture(self._fn, inputs)        # parameters changed since the capture: the cached packs / tables are stale
        flat, spec = pytree.tree_flatten(inputs)
        if spec != self._spec:
            raise ValueError("GraphedForward: input structure differs from the captured one")
        for dst, src in zip(self._static_in, flat):
            if torch.is_tensor(dst):
                if dst.shape != src.shape or dst.dtype != src.dtype:
                    raise ValueError("GraphedForward: input shape / dtype differs from the captured one")
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self._static_out

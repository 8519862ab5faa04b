"""Training step of the two-branch codec -- mirror of the loop body in examples/train.py (RateDistortionLoss :59-82,
configure_optimizers :111-142, train_one_epoch_master :208-260) with data-parallel gradient averaging.

One process per GPU, batch shards, replicas of the weights.  The only exchange step is the gradient all-reduce:
``GradBucketReducer`` flattens gradients into fixed-size fp32 buckets as the backward pass produces them (post-accumulate
hooks), launches one asynchronous NCCL all-reduce per full bucket on a side stream so that it overlaps the rest of the
backward, and copies the averaged values back before the optimizer step.  Parameters that never receive a gradient (the
unused g_a / g_s / h_a / h_s inherited from MeanScaleHyperprior) are left out -- consistently on every
rank, because the graph is the same everywhere.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn
from torch import Tensor

__all__ = ["RateDistortionLoss", "configure_optimizers", "GradBucketReducer", "TrainStep", "GraphedTrainStep"]


class RateDistortionLoss(nn.Module):
    """examples/train.py:59-82: loss = lmbda[q] * MSE + bpp."""

    def __init__(self, q: int):
        super().__init__()
        self.mse = nn.MSELoss()
        self.lmbda = [256, 512, 1024, 2048, 4096, 8192, 10240]
        self.q = q

    def forward(self, output, target):
        N, _, H, W = target.size()
        num_pixels = N * H * W
        out = {}
        out["bpp_loss"] = sum((torch.log(lk).sum() / (-math.log(2) * num_pixels)) for lk in output["likelihoods"].values())
        out["mse_loss"] = self.mse(output["x_hat"], target)
        out["loss"] = self.lmbda[self.q] * out["mse_loss"] + out["bpp_loss"]
        return out


def configure_optimizers(net: nn.Module, learning_rate: float = 1e-4, aux_learning_rate: float = 1e-3, capturable: bool = False):
    """examples/train.py:111-142: Adam on everything but the EntropyBottleneck quantiles, a second Adam on those.
    ``capturable``: keep the step counters on the device so that ``step()`` can be recorded in a CUDA graph."""
    params = dict(net.named_parameters())
    main = sorted(n for n, p in params.items() if not n.endswith(".quantiles") and p.requires_grad)
    aux = sorted(n for n, p in params.items() if n.endswith(".quantiles") and p.requires_grad)
    assert not set(main) & set(aux) and len(main) + len(aux) == sum(p.requires_grad for p in params.values())
    extra = {"capturable": True} if capturable else {}
    return (torch.optim.Adam((params[n] for n in main), lr=learning_rate, **extra),
            torch.optim.Adam((params[n] for n in aux), lr=aux_learning_rate, **extra))


class GradBucketReducer:
    """Bucketed, overlapped gradient averaging over a process group (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, params: Iterable[nn.Parameter], bucket_bytes: int = 32 << 20, group=None, flat: bool = False):
        """``flat``: every parameter's ``.grad`` is a VIEW into a pre-allocated fp32 bucket buffer (call ``zero_()`` instead of
        ``zero_grad(set_to_none=True)``); the collectives then run in place on the buckets -- no ``torch.cat`` flattening and no
        copy-back (three passes over ~160 MB per step in the round-1 graphed data-parallel step).  Parameters that never receive
        a gradient keep a zero gradient instead of ``None`` (Adam then leaves them unchanged, as it does for ``None``)."""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        # reverse registration order ~ the order in which the backward pass produces gradients
        self.params: List[nn.Parameter] = [p for p in params if p.requires_grad][::-1]
        self.buckets: List[List[int]] = []
        cur, size = [], 0
        for i, p in enumerate(self.params):
            cur.append(i)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {i: b for b, idxs in enumerate(self.buckets) for i in idxs}
        self.flat: Optional[List[Tensor]] = None
        if flat and self.params:
            self.flat = []
            for idxs in self.buckets:
                n = sum(self.params[i].numel() for i in idxs)
                buf = torch.zeros(n, dtype=torch.float32, device=self.params[idxs[0]].device)
                off = 0
                for i in idxs:
                    p = self.params[i]
                    if p.dtype != torch.float32:
                        raise TypeError("flat gradient buckets need fp32 parameters")
                    p.grad = buf[off:off + p.numel()].view_as(p)
                    off += p.numel()
                self.flat.append(buf)
        self._ready = [0] * len(self.buckets)
        self._work: Dict[int, tuple] = {}
        self._stream = torch.cuda.Stream() if (self.params and self.params[0].is_cuda) else None
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]
        self.enabled = True

    def _make_hook(self, i: int):
        def hook(_p):
            if not self.enabled or self.world == 1:
                return
            b = self.bucket_of[i]
            self._ready[b] += 1
            if self._ready[b] == len(self.buckets[b]):
                self._launch(b)
        return hook

    def zero_(self):
        """flat mode: the replacement of ``zero_grad`` (one memset per bucket; the ``.grad`` views stay in place)"""
        for buf in self.flat or []:
            buf.zero_()

    def _avg_op(self):
        # NCCL averages inside the collective; gloo has no AVG: sum, then one in-place scale
        return dist.ReduceOp.AVG if dist.get_backend(self.group) == "nccl" else dist.ReduceOp.SUM

    def _launch(self, b: int):
        if self.flat is not None:
            buf, op = self.flat[b], self._avg_op()
            if self._stream is not None:
                self._stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._stream):
                    work = dist.all_reduce(buf, op=op, group=self.group, async_op=True)
            else:
                work = dist.all_reduce(buf, op=op, group=self.group, async_op=True)
            self._work[b] = (work, buf, None if op == dist.ReduceOp.AVG else "scale")
            return
        grads = [self.params[i].grad for i in self.buckets[b] if self.params[i].grad is not None]
        if not grads:
            return
        if self._stream is not None:
            self._stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._stream):
                flat = torch.cat([g.reshape(-1).float() for g in grads])
                work = dist.all_reduce(flat, group=self.group, async_op=True)
        else:
            flat = torch.cat([g.reshape(-1).float() for g in grads])
            work = dist.all_reduce(flat, group=self.group, async_op=True)
        self._work[b] = (work, flat, grads)

    def finish(self):
        """Call after backward(): reduces the buckets that never filled up (parameters without gradient), waits for all
        collectives and writes the averaged gradients back."""
        if self.world == 1:
            self._ready = [0] * len(self.buckets)
            return
        for b in range(len(self.buckets)):
            if b not in self._work:
                self._launch(b)
        for b, (work, flat, grads) in sorted(self._work.items()):
            work.wait()
            if self._stream is not None:
                torch.cuda.current_stream().wait_stream(self._stream)
            if self.flat is not None:
                if grads == "scale":
                    flat.div_(self.world)
                continue
            flat.div_(self.world)
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        self._work.clear()
        self._ready = [0] * len(self.buckets)

    def reduce_now(self):
        """Average every existing gradient across the group, bucket by bucket, on the current stream (no overlap; used when the
        backward pass ran inside a CUDA graph and the hooks did not launch anything)."""
        if self.world == 1:
            return
        if self.flat is not None:
            op = self._avg_op()
            for buf in self.flat:
                dist.all_reduce(buf, op=op, group=self.group)
                if op != dist.ReduceOp.AVG:
                    buf.div_(self.world)
            return
        for idxs in self.buckets:
            grads = [self.params[i].grad for i in idxs if self.params[i].grad is not None]
            if not grads:
                continue
            flat = torch.cat([g.reshape(-1).float() for g in grads])
            dist.all_reduce(flat, group=self.group)
            flat.div_(self.world)
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()

    def remove(self):
        for h in self._hooks:
            h.remove()


class TrainStep:
    """One optimisation step of the second-modality branch with a frozen guide branch (examples/train.py:216-253):
    hidden = guide(rgb) under no_grad; out = net(x, hidden); loss.backward(); clip; Adam; aux loss; aux Adam.
    A net whose ``forward_takes_guide_image`` is set (Master_compresser) is called as net(x, guided, hidden), the pairing of
    train_one_epoch_master (examples/train.py:208-233)."""

    def __init__(self, net: nn.Module, guide: Optional[nn.Module], quality: int = 3, learning_rate: float = 1e-4,
                 aux_learning_rate: float = 1e-3, clip_max_norm: float = 1.0, bucket_bytes: int = 32 << 20, group=None,
                 capturable: bool = False):
        self.net, self.guide = net, guide
        self.criterion = RateDistortionLoss(quality)
        self.optimizer, self.aux_optimizer = configure_optimizers(net, learning_rate, aux_learning_rate, capturable)
        self.clip_max_norm = clip_max_norm
        # gradients live in flat fp32 buckets on the GPU: in-place collectives, one memset per bucket instead of zero_grad
        self.reducer = GradBucketReducer(net.parameters(), bucket_bytes, group, flat=next(net.parameters()).is_cuda)

    def forward_backward(self, x: Tensor, guided: Optional[Tensor] = None) -> Dict[str, Tensor]:
        """Guide forward (no grad), forward, rate-distortion loss backward, aux loss backward; gradients are left in ``.grad``
        (bucket all-reduces are launched by the hooks as the gradients appear, unless the reducer is disabled)."""
        self.net.train()
        hidden = None
        if self.guide is not None:
            with torch.no_grad():
                hidden = self.guide(guided)["hidden"]
        if self.reducer.flat is not None:
            self.reducer.zero_()
        else:
            self.optimizer.zero_grad(set_to_none=True)
            self.aux_optimizer.zero_grad(set_to_none=True)
        if hidden is not None and getattr(self.net, "forward_takes_guide_image", False):
            out_net = self.net(x, guided, hidden)
        else:
            out_net = self.net(x, hidden) if hidden is not None else self.net(x)
        out = self.criterion(out_net, x)
        out["loss"].backward()
        aux_loss = self.net.aux_loss()
        aux_loss.backward()
        out["aux_loss"] = aux_loss.detach()
        return {k: v.detach() for k, v in out.items()}

    def update(self) -> None:
        """Gradient clipping (main parameters) and the two Adam steps."""
        if self.clip_max_norm > 0:
            torch.nn.utils.clip_grad_norm_((p for g in self.optimizer.param_groups for p in g["params"]), self.clip_max_norm)
        self.optimizer.step()
        self.aux_optimizer.step()

    def __call__(self, x: Tensor, guided: Optional[Tensor] = None) -> Dict[str, Tensor]:
        out = self.forward_backward(x, guided)
        self.reducer.finish()
        self.update()
        return out


class GraphedTrainStep:
    """``TrainStep`` replayed as ONE CUDA graph.

    The training step of the fusion models is a few hundred small launches (forward, dgrad / wgrad per layer, the attention
    blocks' elementwise ops, bucket flattening, clipping, two Adam updates): on a B200 the device finishes them faster than
    Python can issue them.  The first ``warmup`` calls run eagerly (real optimisation steps; they also fill every per-shape
    cache); the next call records one whole step -- guide forward, forward, loss, backward, NCCL bucket all-reduces on the side
    stream, clipping, both Adam updates (``capturable=True``) -- under ``torch.cuda.graph`` and every call from then on copies
    its batch into the captured input buffers and launches the graph.  With more than one rank the collectives stay outside:
    graph 1 = forward + backward into static gradient buffers, then the bucketed NCCL all-reduce (eager, current stream), then
    graph 2 = clipping + both Adam updates.  Everything that depends on the parameters (bf16 weight
    packs, GDN re-parametrisations) is recomputed INSIDE the graph, because the optimizer step of the previous replay changed
    them; after each replay the parameters' version counters are bumped so that caches used by later eager calls are rebuilt.

    Inputs must keep their shape and dtype.  The returned tensors are the captured outputs (overwritten by the next call)."""

    def __init__(self, net: nn.Module, guide: Optional[nn.Module], warmup: int = 3, esa_on_kernels: bool = True, **kwargs):
        self.step = TrainStep(net, guide, capturable=True, **kwargs)
        if esa_on_kernels:
            # the ESA gates' 42 small convolutions (forward + backward) on the libmmcodec kernels: slower than the library path when every
            # launch is issued from Python (48.4 vs 39.3 ms per step), faster once the step is a graph (36.2 vs 37.1 ms)
            from .models_mm import ESA
            for m in net.modules():
                if isinstance(m, ESA):
                    m.train_on_kernels = True
        self.warmup = max(1, warmup)
        self.calls = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None

    def _copy_in(self, x: Tensor, guided: Optional[Tensor]):
        for dst, src in ((self._x, x), (self._g, guided)):
            if dst is None:
                continue
            if src is None or dst.shape != src.shape or dst.dtype != src.dtype:
                raise ValueError("GraphedTrainStep: input shape / dtype differs from the captured one")
            dst.copy_(src, non_blocking=True)

    def __call__(self, x: Tensor, guided: Optional[Tensor] = None) -> Dict[str, Tensor]:
        if not x.is_cuda:
            raise RuntimeError("GraphedTrainStep captures CUDA work only (there is no CPU path)")
        self.calls += 1
        if self.graph is None and self.calls <= self.warmup:
            with torch.enable_grad():
                return self.step(x, guided)
        if self.graph is None:
            self._x, self._g = x.clone(), (guided.clone() if guided is not None else None)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            if self.step.reducer.world == 1:
                with torch.cuda.graph(self.graph), torch.enable_grad():
                    self._out = self.step(self._x, self._g)
            else:
                # data parallel: the collectives stay OUTSIDE the graphs (graph 1: forward + backward into static gradient
                # buffers; eager bucketed NCCL all-reduce; graph 2: clip + Adam x2)
                self.step.reducer.enabled = False
                with torch.cuda.graph(self.graph), torch.enable_grad():
                    self._out = self.step.forward_backward(self._x, self._g)
                self.graph.replay()                      # capture does not execute: materialise the gradients once
                self.step.reducer.reduce_now()
                self.update_graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.update_graph):
                    self.step.update()
                self.update_graph.replay()
                self._bump_versions()
                return self._out
        else:
            self._copy_in(x, guided)
        self.graph.replay()
        if self.step.reducer.world > 1:
            self.step.reducer.reduce_now()
            self.update_graph.replay()
        self._bump_versions()
        return self._out

    def _bump_versions(self):
        for p in self.step.net.parameters():
            if p.grad is not None:
                torch.autograd.graph.increment_version(p)

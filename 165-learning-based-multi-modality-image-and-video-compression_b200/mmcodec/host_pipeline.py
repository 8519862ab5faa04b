"""Host-buffer entry point: ``CompressionModel.forward`` for callers whose images and results live in
(pinned) host memory, which is how the reference's evaluation loop uses the model
(compressai/utils/eval_model/__main__t.py:149-211: load image -> ``.to(device)`` -> forward -> metrics
on the host).

The batch is cut into micro-batches; the host->device copy of micro-batch i+1, the kernels of
micro-batch i and the device->host copy of the results of micro-batch i-1 run concurrently on three
CUDA streams (PCIe is full duplex, the copy engines are independent of the SMs).  Images are
independent, so the results are bit-identical to one big forward.  On return the caller's current
stream is ordered after every copy: ``torch.cuda.current_stream().synchronize()`` (or an event
recorded on it) makes the host buffers valid.

``outputs="metrics"`` keeps what that evaluation loop keeps of a forward (``inference_entropy_estimation``,
__main__t.py:151-173): per-image bpp and MSE (PSNR = -10 log10 MSE), reduced on the device by ``mmc_image_bits`` /
``mmc_image_sse`` inside the same CUDA graph, so the device->host traffic is 8 bytes per image instead of the reconstruction
and the likelihood tensors and the pipeline is bound by the host->device copy of the images alone.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor


def _pinned_like(shape, channels_last: bool) -> Tensor:
    """Pinned fp32 host tensor of logical `shape`; channels-last memory when the device tensor is, so that the
    device->host copy is one flat memcpy per micro-batch."""
    if channels_last and len(shape) == 4:
        b, c, h, w = shape
        return torch.empty((b, h, w, c), dtype=torch.float32).pin_memory().permute(0, 3, 1, 2)
    return torch.empty(shape, dtype=torch.float32).pin_memory()


class HostPipeline:
    def __init__(self, net, micro_batch: int = 8, device: Optional[torch.device] = None, use_graphs: bool = True, outputs: str = "full",
                 concurrent_slots: Optional[bool] = None, n_slots: int = 2):
        if outputs not in ("full", "metrics"):
            raise ValueError('outputs must be "full" (x_hat + likelihoods) or "metrics" (per-image bpp and mse)')
        self.outputs = outputs
        self.net = net
        self.micro_batch = int(micro_batch)
        self.device = device or next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs the model on a CUDA device (no CPU path)")
        self.use_graphs = bool(use_graphs)
        self.s_h2d = torch.cuda.Stream(self.device)
        self.s_run = torch.cuda.Stream(self.device)
        self.s_d2h = torch.cuda.Stream(self.device)
        # optional second run stream: consecutive micro-batches may then overlap on the device, so the under-filled grids of the
        # low-resolution layers and every kernel's tail wave of one micro-batch are filled by the other's CTAs (each slot's graph has its
        # own memory pool).  Measured: +7 % when the copy is cheap (uint8 input: 8,993 -> 9,599 img/s), -3 % when the pipeline is bound
        # by the fp32 host->device copy (8,478 -> 8,274 img/s) -- so the default (None) enables it for uint8 inputs only.
        self.concurrent_slots = concurrent_slots
        # input slots (device staging buffers, one captured graph each): the host->device copy of micro-batch i + n_slots - 1 may run
        # while micro-batch i computes.  Even, so that a slot always replays on the same run stream.
        self.n_slots = max(2, int(n_slots) + (int(n_slots) & 1))
        self.s_run2 = torch.cuda.Stream(self.device)
        self._slots = None
        self._u8_slots = None
        self._graphs = None      # per slot: (CUDAGraph, captured output dict) -- no allocator traffic, one launch per micro-batch
        self._out: Optional[Dict[str, Tensor]] = None

    def _buffers(self, x_host: Tensor):
        mb = min(self.micro_batch, x_host.shape[0])
        shape = (mb,) + tuple(x_host.shape[1:])
        # the captured graphs read the packed weights / tables that existed at capture time: re-capture after any parameter update
        from .graphs import module_signature
        sig = module_signature([self.net])        # parameter AND buffer versions + storages (not just parameter versions)
        if sig != getattr(self, "_param_sig", None):
            self._graphs = None
            self._param_sig = sig
        if self._slots is None or tuple(self._slots[0].shape) != shape:
            self._slots = [torch.zeros(shape, dtype=torch.float32, device=self.device) for _ in range(self.n_slots)]
            self._u8_slots = None
            self._graphs = None
        if x_host.dtype == torch.uint8 and self._u8_slots is None:
            # 8-bit images (as decoded from PNG): the copy carries one byte per sample, ToTensor's /255 runs on the device
            self._u8_slots = [torch.zeros(shape, dtype=torch.uint8, device=self.device) for _ in range(self.n_slots)]
        if self.use_graphs and self._graphs is None:
            graphs = []
            with torch.cuda.stream(self.s_run), torch.no_grad():
                self._forward(self._slots[0])     # fills the per-parameter caches (packed weights, LUTs) outside the capture
                torch.cuda.synchronize(self.device)
                for k in range(self.n_slots):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.s_run):
                        o = self._forward(self._slots[k])
                    graphs.append((g, o))
            torch.cuda.synchronize(self.device)
            self._graphs = graphs
        return self._slots, mb

    def _forward(self, x: Tensor):
        o = self.net(x)
        if self.outputs == "full":
            return o
        from . import ops
        n = x.shape[0]
        m = torch.zeros((2, n), dtype=torch.float32, device=x.device)        # row 0: bpp, row 1: mse
        for lk in o["likelihoods"].values():
            ops.image_bits(lk, m[0], 1.0 / (x.shape[2] * x.shape[3]))
        ops.image_mse(x, o["x_hat"], m[1])
        return {"metrics": m}

    def _run(self, k: int, n: int):
        """Forward of the first n images of slot k on the current (run) stream."""
        if self.use_graphs and n == self._slots[k].shape[0]:
            g, o = self._graphs[k]
            g.replay()
            return o, True
        with torch.no_grad():
            return self._forward(self._slots[k][:n]), False       # ragged tail micro-batch: eager launches

    def __call__(self, x_host: Tensor, out: Optional[Dict[str, Tensor]] = None) -> Dict[str, Tensor]:
        """x_host: (B, C, H, W) fp32 host tensor (pinned for asynchronous copies), or uint8 (8-bit samples, converted on the
        device exactly as torchvision's ToTensor does on the host).  Returns / fills
        {"x_hat": (B,C,H,W), "likelihoods": {name: (B,C',H',W')}} pinned host tensors, or, with outputs="metrics",
        {"bpp": (B,), "mse": (B,)} pinned host tensors."""
        if x_host.is_cuda:
            raise ValueError("HostPipeline takes host tensors; call the model directly for device tensors")
        if x_host.dtype not in (torch.float32, torch.uint8):
            raise TypeError("HostPipeline takes fp32 images in [0, 1] or uint8 images (converted as ToTensor does: / 255)")
        from . import ops
        as_u8 = x_host.dtype == torch.uint8
        concurrent = as_u8 if self.concurrent_slots is None else bool(self.concurrent_slots)
        B = x_host.shape[0]
        slots, mb = self._buffers(x_host)
        caller = torch.cuda.current_stream(self.device)
        start = torch.cuda.Event()
        start.record(caller)
        for s in (self.s_h2d, self.s_run, self.s_run2, self.s_d2h):
            s.wait_event(start)
        slot_free = [None] * self.n_slots      # event: kernels that read slot k have finished
        out_free = [None] * self.n_slots       # event: the captured outputs of slot k have been copied to the host
        last_d2h = None
        result = out if out is not None else self._out
        if result is not None and next(iter(result.values())).shape[0] != B:
            result = None
        i = 0
        for lo in range(0, B, mb):
            hi = min(lo + mb, B)
            k = i % self.n_slots
            with torch.cuda.stream(self.s_h2d):
                if slot_free[k] is not None:
                    self.s_h2d.wait_event(slot_free[k])
                (self._u8_slots if as_u8 else slots)[k][: hi - lo].copy_(x_host[lo:hi], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_h2d)
            s_run = self.s_run2 if ((k & 1) and concurrent) else self.s_run
            with torch.cuda.stream(s_run):
                s_run.wait_event(ev_in)
                if out_free[k] is not None:
                    s_run.wait_event(out_free[k])
                if as_u8:
                    ops.u8_to_f32(self._u8_slots[k][: hi - lo], out=slots[k][: hi - lo])
                o, static = self._run(k, hi - lo)
                ev_run = torch.cuda.Event()
                ev_run.record(s_run)
                slot_free[k] = ev_run
            if self.outputs == "metrics":
                if result is None:
                    result = {"bpp": torch.empty(B, dtype=torch.float32).pin_memory(), "mse": torch.empty(B, dtype=torch.float32).pin_memory()}
                    if out is None:
                        self._out = result
                with torch.cuda.stream(self.s_d2h):
                    self.s_d2h.wait_event(ev_run)
                    result["bpp"][lo:hi].copy_(o["metrics"][0], non_blocking=True)
                    result["mse"][lo:hi].copy_(o["metrics"][1], non_blocking=True)
                    if not static:
                        o["metrics"].record_stream(self.s_d2h)
                    last_d2h = torch.cuda.Event()
                    last_d2h.record(self.s_d2h)
                    out_free[k] = last_d2h
                i += 1
                continue
            if result is None:
                # first call: allocate pinned result buffers with the device tensors' memory formats
                result = {"x_hat": _pinned_like((B,) + tuple(o["x_hat"].shape[1:]), False),
                          "likelihoods": {n: _pinned_like((B,) + tuple(t.shape[1:]), not t.is_contiguous())
                                          for n, t in o["likelihoods"].items()}}
                if out is None:
                    self._out = result
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(ev_run)
                result["x_hat"][lo:hi].copy_(o["x_hat"], non_blocking=True)
                for n, t in o["likelihoods"].items():
                    result["likelihoods"][n][lo:hi].copy_(t, non_blocking=True)
                if not static:
                    o["x_hat"].record_stream(self.s_d2h)
                    for t in o["likelihoods"].values():
                        t.record_stream(self.s_d2h)
                last_d2h = torch.cuda.Event()
                last_d2h.record(self.s_d2h)
                out_free[k] = last_d2h
            i += 1
        if last_d2h is not None:
            caller.wait_event(last_d2h)
        return result

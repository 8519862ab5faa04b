"""``CompressionModel.compress`` for a stream of batches with the host entropy coding of batch i overlapped with the GPU work of
batch i + 1 (compressai/models/google.py:324-332,393-404 + entropy_models.py:237-270).

compress() is two stages with different owners: the transforms, quantisation and CDF-index kernels on the GPU (~5.5 ms per eight
1088 x 1920 images) and the rANS coder on host threads (one serial stream per image, ~9 ms per image, images in parallel).  Called
in a loop they run back to back; here ``submit(x)`` enqueues the GPU stage, stages the int32 symbols / indexes into one of
``depth`` sets of pinned host buffers and hands them to a coding worker, so the next ``submit`` can start its kernels while the
previous batch is being coded (the C coder releases the GIL).  The byte strings are those of ``net.compress(x)``.
With the device coder selected (``mmcodec.set_entropy_coder(net, "ans-lanes")``) the symbols never leave the GPU: the coding kernels
follow the transforms on the same stream and only the containers are copied back; the worker just slices the staged bytes.

    pipe = mmcodec.CompressPipeline(net)
    futures = [pipe.submit(x) for x in batches]          # at most `depth` batches in flight: submit blocks on the oldest
    results = [f.result() for f in futures]              # {"strings": [[bytes] * B, [bytes] * B], "shape": (h, w)}
"""
from __future__ import annotations

import collections
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Dict, List

import torch

from . import ops


class CompressPipeline:
    def __init__(self, net, depth: int = 2):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        if not hasattr(net, "symbols_and_indexes"):
            raise TypeError("CompressPipeline needs a model with symbols_and_indexes() (Factorized / ScaleHyperprior / MeanScaleHyperprior)")
        self.net = net
        self.depth = int(depth)
        self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="mmcodec-rans")
        self._sets: List[Dict[str, torch.Tensor]] = [dict() for _ in range(self.depth)]
        self._inflight = collections.deque()
        self._n = 0
        self._code_streams = None
        self._copy_stream = None

    def _stage(self, bufs: Dict[str, torch.Tensor], name: str, t: torch.Tensor) -> torch.Tensor:
        """logical-order int32 copy of a device tensor in this set's pinned buffer (asynchronous on the current stream)"""
        t = t.detach()
        t = (t if t.dtype == torch.int32 else t.int()).contiguous()
        buf = bufs.get(name)
        if buf is None or buf.numel() < t.numel():
            buf = bufs[name] = torch.empty(max(t.numel(), 1), dtype=torch.int32).pin_memory()
        view = buf[: t.numel()].view(t.shape)
        view.copy_(t, non_blocking=True)
        return view

    def _upload(self, bufs, x_host: torch.Tensor) -> torch.Tensor:
        """Host images (pinned memory for a truly asynchronous copy) -> this set's device input buffer on a COPY stream: the
        host->device copy of batch i + 1 runs under the kernels of batch i instead of in front of its own."""
        dev = next(self.net.parameters()).device
        buf = bufs.get("x_dev")
        if buf is None or buf.shape != x_host.shape or buf.dtype != x_host.dtype:
            buf = bufs["x_dev"] = torch.empty(x_host.shape, dtype=x_host.dtype, device=dev)
            bufs.pop("x_free", None)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(self._copy_stream):
            if "x_free" in bufs:
                self._copy_stream.wait_event(bufs["x_free"])
            buf.copy_(x_host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        torch.cuda.current_stream(dev).wait_event(ev)
        return buf

    def submit(self, x: torch.Tensor) -> Future:
        """Enqueue one batch: a CUDA tensor, or host images (see _upload).  Returns a Future of compress()'s result."""
        if len(self._inflight) >= self.depth:
            self._inflight.popleft().result()          # the buffer set about to be reused has been coded
        bufs = self._sets[self._n % self.depth]
        self._n += 1
        if not x.is_cuda:
            x = self._upload(bufs, x)
        with torch.no_grad():
            c = self.net.symbols_and_indexes(x)
        if "x_dev" in bufs:                             # this set's input buffer may be overwritten once the GPU stage has read it
            bufs["x_free"] = torch.cuda.Event()
            bufs["x_free"].record(torch.cuda.current_stream(x.device))
        names = [k for k in ("y", "z") if f"{k}_symbols" in c]
        net = self.net

        def tabs(em):
            return em._quantized_cdf, em._cdf_length, em._offset
        # hyperprior models: y is coded with the Gaussian conditional's tables, z with the bottleneck's; factorized: y with the bottleneck's
        models = {"y": net.gaussian_conditional if "z" in names else net.entropy_bottleneck}
        if "z" in names:
            models["z"] = net.entropy_bottleneck
        tables = {k: tabs(m) for k, m in models.items()}
        shape = c.get("shape")
        if all(m._coder() == "ans-lanes" for m in models.values()):
            # device coder: the symbols stay on the GPU; only the containers (and their sizes) travel, staged in this set's pinned buffers
            # The coding kernels are a few hundred serial chains (one warp per SM, no shared memory): they run on side streams -- y and
            # z each on its own -- next to the NEXT batch's transforms instead of in front of them.  The handles keep the symbol /
            # index tensors alive until the side streams are done with them (collect() waits for that).
            pinned = bufs.setdefault("lane", {})
            main = torch.cuda.current_stream(x.device)
            ready = torch.cuda.Event()
            ready.record(main)
            if self._code_streams is None:
                self._code_streams = [torch.cuda.Stream(device=x.device) for _ in range(2)]
            handles, done = [], []
            for k, st in zip(names, self._code_streams):
                with torch.cuda.stream(st):
                    st.wait_event(ready)
                    handles.append(ops.rans_encode_device_launch(c[f"{k}_symbols"], c[f"{k}_indexes"], *tables[k], pinned=pinned))
                    ev = torch.cuda.Event()
                    ev.record(st)
                    done.append(ev)

            def collect():
                for ev_ in done:
                    ev_.synchronize()
                return {"strings": [h.collect() for h in handles], "shape": shape}

            fut = self._pool.submit(collect)
            self._inflight.append(fut)
            return fut
        staged = {k: (self._stage(bufs, k + "s", c[f"{k}_symbols"]), self._stage(bufs, k + "i", c[f"{k}_indexes"])) for k in names}
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(x.device))

        def code():
            ev.synchronize()
            strings = [ops.rans_encode(staged[k][0], staged[k][1], *tables[k]) for k in names]
            return {"strings": strings, "shape": shape}

        fut = self._pool.submit(code)
        self._inflight.append(fut)
        return fut

    def close(self):
        self._pool.shutdown(wait=True)

"""Functional front-end of the C ABI on torch CUDA tensors.

torch is plumbing here (device memory, streams); every function enqueues hand-written kernels of
libmmcodec.so on the current CUDA stream.  CPU tensors are rejected: there is no CPU path.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib as L

Tensor = torch.Tensor


def _require_cuda(*tensors) -> None:
    """Every kernel is enqueued on the CURRENT device's current stream (see _stream): a tensor that lives on another device
    would be read through a foreign pointer on the wrong stream, so it is rejected here instead (use torch.cuda.device(...)
    or torch.cuda.set_device around the call, as for any stream-ordered CUDA library)."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mmcodec runs on CUDA tensors only (B200 / sm_100a); got a CPU tensor. "
                               "There is no CPU fallback.")
        if cur is None:
            cur = torch._C._cuda_getDevice()
        if t.device.index != cur:
            raise RuntimeError(f"mmcodec: tensor on cuda:{t.device.index} but the current device is cuda:{cur}; "
                               "wrap the call in torch.cuda.device(tensor.device)")


def _ptr(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    # raw cudaStream_t of torch's current stream on the current device; the Python-level torch.cuda.current_stream() costs
    # ~10 us per call (device-index resolution + Stream object), which was a third of the per-launch host overhead
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _f32c(t: Tensor) -> Tensor:
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


def _is_channels_last(t: Tensor) -> bool:
    return t.dim() >= 3 and not t.is_contiguous() and t.permute(0, *range(2, t.dim()), 1).is_contiguous()


def _view_oci(t: Tensor) -> Tuple[Tensor, int, int, int]:
    """Return (tensor, outer, C, inner) for a logical (N, C, *spatial) tensor, keeping a channels-last
    memory layout if it has one, otherwise forcing NCHW-contiguous."""
    if t.dim() < 2:
        raise ValueError("expected a tensor with at least 2 dimensions (N, C, ...)")
    C = t.shape[1]
    if _is_channels_last(t):
        return t, t.numel() // C if C else 0, C, 1
    t = t.contiguous()
    inner = int(np.prod(t.shape[2:])) if t.dim() > 2 else 1
    return t, t.shape[0], C, inner


def _like_layout(t: Tensor, ref: Tensor) -> Tensor:
    """Make t (same logical shape as ref) share ref's memory layout."""
    t = t.expand_as(ref) if t.shape != ref.shape else t
    if t.stride() == ref.stride() and t.dtype == torch.float32:
        return t
    out = torch.empty_like(ref, dtype=torch.float32)
    out.copy_(t)
    return out


def _means_arg(x: Tensor, means: Optional[Tensor], C: int):
    """Classify a broadcastable means tensor: none / per-channel / full."""
    if means is None:
        return L.MEANS_NONE, None
    if means.dim() == x.dim() and means.shape[1] == C and means.numel() == C:
        return L.MEANS_PER_CHANNEL, _f32c(means.reshape(-1))
    return L.MEANS_FULL, _like_layout(means, x)


# ---- quantize / dequantize ---------------------------------------------------------------------
def quantize_symbols(x: Tensor, means: Optional[Tensor] = None) -> Tensor:
    """EntropyModel.quantize(x, "symbols", means) -> int32 (entropy_models.py:157-182)."""
    _require_cuda(x, means)
    x = x.float()
    if x.dim() < 2:
        xv, outer, C, inner = x.contiguous(), 1, 1, x.numel()
    else:
        xv, outer, C, inner = _view_oci(x)
    mode, m = _means_arg(xv, means, C)
    out = torch.empty_like(xv, dtype=torch.int32)
    with _Timed("quantize_symbols|entropy", xv.numel() * (8.0 + (4.0 if mode == L.MEANS_FULL else 0.0))):
        L.check(L.lib().mmc_quantize_symbols(_ptr(xv), _ptr(m), mode, outer, C, inner, _ptr(out), _stream()))
    return out


def quantize_dequantize(x: Tensor, means: Optional[Tensor] = None) -> Tensor:
    """EntropyModel.quantize(x, "dequantize", means) (entropy_models.py:169-178)."""
    _require_cuda(x, means)
    x = x.float()
    if x.dim() < 2:
        xv, outer, C, inner = x.contiguous(), 1, 1, x.numel()
    else:
        xv, outer, C, inner = _view_oci(x)
    mode, m = _means_arg(xv, means, C)
    out = torch.empty_like(xv)
    L.check(L.lib().mmc_quantize_dequantize(_ptr(xv), _ptr(m), mode, outer, C, inner, _ptr(out), _stream()))
    return out


def quantize_noise(x: Tensor, noise: Tensor) -> Tensor:
    _require_cuda(x, noise)
    x = _f32c(x)
    noise = _f32c(noise)
    out = torch.empty_like(x)
    L.check(L.lib().mmc_quantize_noise(_ptr(x), _ptr(noise), x.numel(), _ptr(out), _stream()))
    return out


def dequantize(symbols: Tensor, means: Optional[Tensor] = None) -> Tensor:
    """EntropyModel.dequantize (entropy_models.py:190-199)."""
    _require_cuda(symbols, means)
    s = symbols.int()
    if s.dim() < 2:
        sv, outer, C, inner = s.contiguous(), 1, 1, s.numel()
    else:
        sv, outer, C, inner = _view_oci(s)
    mode, m = _means_arg(sv, means, C)
    out = torch.empty_like(sv, dtype=torch.float32)
    with _Timed("dequantize|entropy", sv.numel() * (8.0 + (4.0 if mode == L.MEANS_FULL else 0.0))):
        L.check(L.lib().mmc_dequantize(_ptr(sv), _ptr(m), mode, outer, C, inner, _ptr(out), _stream()))
    return out


def lower_bound(x: Tensor, bound: float) -> Tensor:
    _require_cuda(x)
    x = _f32c(x)
    out = torch.empty_like(x)
    L.check(L.lib().mmc_lower_bound(_ptr(x), float(bound), x.numel(), _ptr(out), _stream()))
    return out


def lower_bound_bwd(x: Tensor, grad_out: Tensor, bound: float) -> Tensor:
    _require_cuda(x, grad_out)
    x, g = _f32c(x), _f32c(grad_out)
    out = torch.empty_like(x)
    L.check(L.lib().mmc_lower_bound_bwd(_ptr(x), _ptr(g), float(bound), x.numel(), _ptr(out), _stream()))
    return out


def build_indexes(scales: Tensor, scale_table: Tensor, bound: float) -> Tensor:
    """GaussianConditional.build_indexes (entropy_models.py:735-740)."""
    _require_cuda(scales, scale_table)
    s = scales.float()
    if not (s.is_contiguous() or _is_channels_last(s)):
        s = s.contiguous()
    table = _f32c(scale_table)
    out = torch.empty_like(s, dtype=torch.int32)
    with _Timed("build_indexes|entropy", s.numel() * 8.0):
        L.check(L.lib().mmc_build_indexes(_ptr(s), _ptr(table), table.numel(), float(bound), s.numel(), _ptr(out), _stream()))
    return out


def channel_indexes(size, device) -> Tensor:
    """EntropyBottleneck._build_indexes (entropy_models.py:542-553)."""
    size = tuple(int(s) for s in size)
    out = torch.empty(size, dtype=torch.int32, device=device)
    _require_cuda(out)
    C = size[1]
    inner = int(np.prod(size[2:])) if len(size) > 2 else 1
    with _Timed("channel_indexes|entropy", out.numel() * 4.0):
        L.check(L.lib().mmc_channel_indexes(size[0], C, inner, _ptr(out), _stream()))
    return out


# ---- entropy models ----------------------------------------------------------------------------
def make_eb_params(matrices, biases, factors, medians: Optional[Tensor]):
    keep = [_f32c(t.detach()) for t in list(matrices) + list(biases) + list(factors)]
    med = _f32c(medians.detach().reshape(-1)) if medians is not None else None
    p = L.EbParams()
    for k in range(5):
        p.matrix[k] = keep[k].data_ptr()
        p.bias[k] = keep[5 + k].data_ptr()
    for k in range(4):
        p.factor[k] = keep[10 + k].data_ptr()
    p.medians = med.data_ptr() if med is not None else None
    keep.append(med)
    return p, keep


EB_LUT_HALF_WIDTH = 64   # symbols tabulated per channel on the eval fast path: -64 .. 64


def eb_build_lut(params, C: int, likelihood_bound: float, device) -> Tensor:
    """[C][129] table of bound(likelihood(k + median_c)), |k| <= 64 (see mmc_eb_build_lut)."""
    p, _keep = params
    lut = torch.empty((C, 2 * EB_LUT_HALF_WIDTH + 1), dtype=torch.float32, device=device)
    _require_cuda(lut)
    L.check(L.lib().mmc_eb_build_lut(ctypes.byref(p), float(likelihood_bound), C, EB_LUT_HALF_WIDTH, _ptr(lut), _stream()))
    return lut


def eb_forward(x: Tensor, params, noise: Optional[Tensor] = None, likelihood_bound: float = 1e-9,
               want_bf16: bool = False, bits: Optional[Tensor] = None, lut: Optional[Tensor] = None):
    """EntropyBottleneck.forward on a logical (N, C, *spatial) tensor (entropy_models.py:495-540).
    Returns (x_hat, likelihood[, x_hat_bf16]) in x's memory layout.  With `lut` (eval mode only) the likelihood
    comes from the per-(channel, symbol) table built by eb_build_lut."""
    _require_cuda(x, noise)
    p, _keep = params
    xv, outer, C, inner = _view_oci(x.float())
    nz = _like_layout(noise, xv) if noise is not None else None
    x_hat = torch.empty_like(xv)
    lik = torch.empty_like(xv)
    xb = torch.empty_like(xv, dtype=torch.bfloat16) if want_bf16 else None
    if lut is not None and nz is None:
        with _Timed("entropy_bottleneck|eb_lut", xv.numel() * (12.0 + (2.0 if want_bf16 else 0.0))):
            L.check(L.lib().mmc_eb_forward_lut(_ptr(xv), ctypes.byref(p), _ptr(lut), EB_LUT_HALF_WIDTH, float(likelihood_bound),
                                               outer, C, inner, _ptr(x_hat), _ptr(xb), _ptr(lik), _ptr(bits), _stream()))
        return (x_hat, lik, xb) if want_bf16 else (x_hat, lik)
    with _Timed("entropy_bottleneck|eb", xv.numel() * (12.0 + (4.0 if nz is not None else 0.0) + (2.0 if want_bf16 else 0.0))):
        L.check(L.lib().mmc_eb_forward(_ptr(xv), _ptr(nz), ctypes.byref(p), float(likelihood_bound), outer, C, inner,
                                       _ptr(x_hat), _ptr(xb), _ptr(lik), _ptr(bits), _stream()))
    return (x_hat, lik, xb) if want_bf16 else (x_hat, lik)


def eb_logits_cumulative(x: Tensor, params) -> Tensor:
    """_logits_cumulative on a logical (N, C, *spatial) tensor (entropy_models.py:457-477)."""
    _require_cuda(x)
    p, _keep = params
    xv, outer, C, inner = _view_oci(x.float())
    out = torch.empty_like(xv)
    L.check(L.lib().mmc_eb_logits_cumulative(_ptr(xv), ctypes.byref(p), outer, C, inner, _ptr(out), _stream()))
    return out


def gc_forward(x: Tensor, scales: Tensor, means: Optional[Tensor] = None, noise: Optional[Tensor] = None,
               scale_bound: float = 0.11, likelihood_bound: float = 1e-9, want_bf16: bool = False,
               bits: Optional[Tensor] = None):
    """GaussianConditional.forward (entropy_models.py:715-731); elementwise, any common layout."""
    _require_cuda(x, scales, means, noise)
    xv = x.float()
    if not (xv.is_contiguous() or _is_channels_last(xv)):
        xv = xv.contiguous()
    s = _like_layout(scales, xv)
    m = _like_layout(means, xv) if means is not None else None
    nz = _like_layout(noise, xv) if noise is not None else None
    x_hat = torch.empty_like(xv)
    lik = torch.empty_like(xv)
    xb = torch.empty_like(xv, dtype=torch.bfloat16) if want_bf16 else None
    with _Timed("gaussian_conditional|gc", xv.numel() * (16.0 + (4.0 if m is not None else 0.0) + (4.0 if nz is not None else 0.0) + (2.0 if want_bf16 else 0.0))):
        L.check(L.lib().mmc_gc_forward(_ptr(xv), _ptr(s), _ptr(m), _ptr(nz), float(scale_bound), float(likelihood_bound),
                                       xv.numel(), _ptr(x_hat), _ptr(xb), _ptr(lik), _ptr(bits), _stream()))
    return (x_hat, lik, xb) if want_bf16 else (x_hat, lik)


def bits(likelihood: Tensor, accum: Optional[Tensor] = None) -> Tensor:
    """-sum(log2(likelihood)) accumulated into a 1-element fp32 tensor (examples/train.py:74-77)."""
    _require_cuda(likelihood)
    lk = likelihood.float()
    if not (lk.is_contiguous() or _is_channels_last(lk)):
        lk = lk.contiguous()
    if accum is None:
        accum = torch.zeros(1, dtype=torch.float32, device=lk.device)
    L.check(L.lib().mmc_bits(_ptr(lk), lk.numel(), _ptr(accum), _stream()))
    return accum


def pmf_to_quantized_cdf(pmf, precision: int = 16):
    """compressai._CXX.pmf_to_quantized_cdf (cpp_exts/ops/ops.cpp:40-109); host-side, returns list[int]."""
    arr = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32).reshape(-1))
    cdf = np.empty(arr.size + 1, np.uint32)
    L.check(L.lib().mmc_pmf_to_quantized_cdf_host(arr.ctypes.data, arr.size, int(precision), cdf.ctypes.data))
    return cdf.tolist()


def pmf_to_quantized_cdf_device(pmf: Tensor, tail_mass: Tensor, pmf_length: Tensor, max_length: int, precision: int = 16) -> Tensor:
    """EntropyModel._pmf_to_cdf for all rows at once on the device: int32 (rows, max_length + 2), bit-exact with the reference."""
    _require_cuda(pmf, tail_mass, pmf_length)
    pmf = _f32c(pmf.detach())
    rows = pmf.shape[0]
    tail = _f32c(tail_mass.detach().reshape(-1))
    lens = pmf_length.detach().to(torch.int32).contiguous()
    cdf = torch.empty((rows, max_length + 2), dtype=torch.int32, device=pmf.device)
    status = torch.empty(rows, dtype=torch.int32, device=pmf.device)
    L.check(L.lib().mmc_pmf_to_quantized_cdf(_ptr(pmf), pmf.shape[1], _ptr(tail), _ptr(lens), rows, int(max_length), int(precision),
                                             _ptr(cdf), _ptr(status), _stream()))
    if int(status.min().item()) != 0:
        raise ValueError("Invalid `pmf`: negative, non-finite or all-zero row, or no symbol can donate frequency")
    return cdf


def _i32np(t):
    a = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    return np.ascontiguousarray(a, dtype=np.int32)


_host_stage = __import__("threading").local()


def _stage_i32(tensors):
    """int32 host arrays in LOGICAL (N, C, ...) order for the coder.  CUDA tensors are put in that order on the device (the
    symbol / index tensors are channels-last there; transposing 21 MB per image batch on the host cost more than coding it),
    copied asynchronously into cached pinned buffers and waited for once."""
    cache = getattr(_host_stage, "bufs", None)
    if cache is None:
        cache = _host_stage.bufs = {}
    out, pending = [], False
    for slot, t in enumerate(tensors):
        if isinstance(t, torch.Tensor) and t.is_cuda:
            t = t.detach()
            t = (t if t.dtype == torch.int32 else t.int()).contiguous()
            buf = cache.get(slot)
            if buf is None or buf.numel() < t.numel():
                buf = cache[slot] = torch.empty(max(t.numel(), 1), dtype=torch.int32).pin_memory()
            view = buf[: t.numel()].view(t.shape)
            view.copy_(t, non_blocking=True)
            out.append(view)
            pending = True
        else:
            out.append(t)
    if pending:
        torch.cuda.current_stream().synchronize()
    return [_i32np(t) for t in out]


def rans_encode(symbols: Tensor, indexes: Tensor, cdf: Tensor, cdf_lengths: Tensor, offsets: Tensor):
    """compressai.ans.RansEncoder.encode_with_indexes for every image of a batch (rans_interface.cpp:108-213);
    symbols / indexes are (B, ...) int32 tensors (any device), returns a list of B byte strings."""
    sym, idx = _stage_i32((symbols, indexes))
    B = sym.shape[0]
    n = sym[0].size if B else 0
    tab, lens, offs = _coder_tables_host(cdf, cdf_lengths, offsets)
    nbytes = np.zeros(max(B, 1), dtype=np.uint64)
    cap = 4 * n + 64
    for _ in range(2):
        out = np.empty((max(B, 1), cap), dtype=np.uint8)
        rc = L.lib().mmc_rans_encode_batch_host(sym.ctypes.data, idx.ctypes.data, B, n, tab.ctypes.data, tab.shape[0], tab.shape[1],
                                                lens.ctypes.data, offs.ctypes.data, out.ctypes.data, cap, nbytes.ctypes.data)
        if rc == L.MMC_EINVAL and B and int(nbytes.max()) > cap:
            cap = int(nbytes.max())     # bypass-heavy stream: retry with the size the coder reported
            continue
        L.check(rc)
        break
    return [out[b, : int(nbytes[b])].tobytes() for b in range(B)]


_tables_host_cache = {}


def _coder_tables_host(cdf: Tensor, cdf_lengths: Tensor, offsets: Tensor):
    """Host int32 copies of the CDF tables, cached per (storage, version): three small device->host copies per call otherwise."""
    if not isinstance(cdf, torch.Tensor):
        return _i32np(cdf), _i32np(cdf_lengths).reshape(-1), _i32np(offsets).reshape(-1)
    key = (cdf.data_ptr(), cdf._version, cdf_lengths.data_ptr(), cdf_lengths._version, offsets.data_ptr(), offsets._version, tuple(cdf.shape))
    hit = _tables_host_cache.get(key)
    if hit is None:
        if len(_tables_host_cache) > 32:
            _tables_host_cache.clear()
        hit = _tables_host_cache[key] = (_i32np(cdf), _i32np(cdf_lengths).reshape(-1), _i32np(offsets).reshape(-1))
    return hit


def rans_decode(strings, indexes: Tensor, cdf: Tensor, cdf_lengths: Tensor, offsets: Tensor) -> Tensor:
    """compressai.ans.RansDecoder.decode_with_indexes for a batch (rans_interface.cpp:215-284); returns int32 symbols
    shaped like `indexes`, on `indexes`' device."""
    (idx,) = _stage_i32((indexes,))
    B = idx.shape[0]
    n = idx[0].size if B else 0
    tab, lens, offs = _coder_tables_host(cdf, cdf_lengths, offsets)
    blob = np.frombuffer(b"".join(strings), dtype=np.uint8) if strings else np.zeros(0, np.uint8)
    blob = np.ascontiguousarray(blob)
    sizes = np.array([len(s) for s in strings], dtype=np.uint64)
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64) if B else np.zeros(0, np.uint64)
    out = np.empty(idx.shape, dtype=np.int32)
    L.check(L.lib().mmc_rans_decode_batch_host(blob.ctypes.data, starts.ctypes.data, sizes.ctypes.data, idx.ctypes.data, B, n,
                                               tab.ctypes.data, tab.shape[0], tab.shape[1], lens.ctypes.data, offs.ctypes.data,
                                               out.ctypes.data))
    res = torch.from_numpy(out)
    return res.to(indexes.device) if isinstance(indexes, torch.Tensor) else res


# ---- device coder (csrc/rans_device.cu): the "lane container", symbols never leave the GPU --------------------------
LANE_MAGIC = b"MMCL"
_LANE_STATUS = {1: "an index is outside the CDF table", 2: "a CDF row is not strictly increasing within 2^16", 4: "output capacity too small",
                8: "stream is malformed, truncated or does not match the CDF tables"}


def _lane_status_error(name: str, st: int):
    msg = "; ".join(v for k, v in _LANE_STATUS.items() if st & k) or f"status {st}"
    return ValueError(f"{name}: {msg}")


def _i32dev(t: Tensor, device) -> Tensor:
    t = t.detach()
    if t.device != device:
        t = t.to(device)
    return (t if t.dtype == torch.int32 else t.int()).contiguous()


def rans_lanes_default(n: int) -> int:
    return int(L.lib().mmc_rans_lanes_default(int(n)))


class LaneEncodeHandle:
    """An enqueued device encode: the containers sit in ``out`` (uint8 [B][cap], device), their sizes and the status word in ``meta``;
    the first ``head`` bytes of every row and ``meta`` are on their way to pinned host buffers on the launching stream.
    ``collect()`` (after that stream -- or an event recorded on it -- has been waited for) returns the B byte strings."""

    def __init__(self, args, out, meta, out_h, meta_h, B, head):
        self.args, self.out, self.meta, self.out_h, self.meta_h, self.B, self.head = args, out, meta, out_h, meta_h, B, head

    def collect(self):
        B = self.B
        st = int(self.meta_h[B + 1].item()) & 0xffffffff
        if st == 4 and B:
            # capacity too small (escape-heavy data): re-run blocking with the exact size the header pass reported
            return rans_encode_device(*self.args, cap=(int(self.meta_h[:B].max().item()) + 3) & ~3)
        if st:
            raise _lane_status_error("rans_encode_device", st)
        sizes = [int(v) for v in self.meta_h[:B].tolist()]
        host = self.out_h.numpy()
        res = []
        for b, nb in enumerate(sizes):
            if nb <= self.head:
                res.append(host[b, :nb].tobytes())
            else:       # longer than the staged head: fetch the row's tail (synchronous, rare)
                res.append(self.out[b, :nb].cpu().numpy().tobytes())
        return res


_lane_pinned = __import__("threading").local()


def rans_encode_device_launch(symbols: Tensor, indexes: Tensor, cdf: Tensor, cdf_lengths: Tensor, offsets: Tensor,
                              lanes: Optional[int] = None, cap: Optional[int] = None, pinned: Optional[dict] = None) -> LaneEncodeHandle:
    """Enqueue the device encode of a batch and the device->host copies of its result on the current stream; no synchronisation.
    ``pinned``: a dict the caller owns for the staging buffers (one per in-flight batch); default: a per-thread cache."""
    _require_cuda(symbols, indexes)
    dev = symbols.device
    sym, idx = _i32dev(symbols, dev), _i32dev(indexes, dev)
    if sym.shape != idx.shape:
        raise ValueError("rans_encode_device: symbols / indexes shape mismatch")
    B = sym.shape[0]
    n = sym[0].numel() if B else 0
    tab, lens, offs = _i32dev(cdf, dev), _i32dev(cdf_lengths, dev).reshape(-1), _i32dev(offsets, dev).reshape(-1)
    S = int(lanes) if lanes is not None else rans_lanes_default(n)
    nb = ctypes.c_size_t()
    L.check(L.lib().mmc_rans_device_workspace(B, S, tab.shape[0], tab.shape[1], ctypes.byref(nb)))
    ws = torch.empty(max(nb.value, 16), dtype=torch.uint8, device=dev)
    meta = torch.zeros(B + 2, dtype=torch.int64, device=dev)          # nbytes[B], pad, status (int32 in the last slot)
    if cap is None:
        cap = (16 + 8 * S + 2 * n + 64 + 3) & ~3                         # one word per symbol: enough unless escapes dominate
    out = torch.empty((max(B, 1), cap), dtype=torch.uint8, device=dev)
    with _Timed("rans_encode|coder"):
        L.check(L.lib().mmc_rans_encode_device(_ptr(sym), _ptr(idx), B, n, _ptr(tab), tab.shape[0], tab.shape[1], _ptr(lens), _ptr(offs), S,
                                               _ptr(out), cap, _ptr(meta), _ptr(ws), _ptr(meta[B + 1:]), _stream()))
    # stage the result: sizes + status, and the head of every container (2 bits per symbol covers typical rates; longer rows are
    # fetched in collect())
    head = min(cap, (16 + 8 * S + n // 4 + 1024 + 3) & ~3)
    if pinned is None:
        pinned = getattr(_lane_pinned, "bufs", None)
        if pinned is None:
            pinned = _lane_pinned.bufs = {}
    key = ("lane", id(cdf) if not isinstance(cdf, torch.Tensor) else cdf.data_ptr())
    buf = pinned.get(key)
    if buf is None or buf[0].shape[0] < max(B, 1) or buf[0].shape[1] < head or buf[1].numel() < B + 2:
        buf = pinned[key] = (torch.empty((max(B, 1), head), dtype=torch.uint8).pin_memory(), torch.empty(B + 2, dtype=torch.int64).pin_memory())
    out_h, meta_h = buf[0][: max(B, 1), :head], buf[1][: B + 2]
    out_h.copy_(out[:, :head], non_blocking=True)
    meta_h.copy_(meta, non_blocking=True)
    return LaneEncodeHandle((symbols, indexes, cdf, cdf_lengths, offsets, lanes), out, meta, out_h, meta_h, B, head)


def rans_encode_device(symbols: Tensor, indexes: Tensor, cdf: Tensor, cdf_lengths: Tensor, offsets: Tensor, lanes: Optional[int] = None,
                       cap: Optional[int] = None):
    """Entropy-code every image of a batch ON THE DEVICE into lane containers (see csrc/rans_device.cu): same symbols, CDF tables,
    indexes and escape scheme as ``rans_encode``, a different (not reference-compatible) container.  symbols / indexes: (B, ...)
    CUDA tensors.  Returns a list of B byte strings (blocking; ``rans_encode_device_launch`` is the asynchronous form)."""
    h = rans_encode_device_launch(symbols, indexes, cdf, cdf_lengths, offsets, lanes, cap)
    torch.cuda.current_stream().synchronize()
    return h.collect()


def rans_decode_device(strings, indexes: Tensor, cdf: Tensor, cdf_lengths: Tensor, offsets: Tensor) -> Tensor:
    """Decode lane containers produced by ``rans_encode_device``; returns int32 symbols shaped like ``indexes`` on its device."""
    _require_cuda(indexes)
    dev = indexes.device
    idx = _i32dev(indexes, dev)
    B = idx.shape[0]
    if len(strings) != B:
        raise ValueError("rans_decode_device: one stream per image expected")
    n = idx[0].numel() if B else 0
    max_lanes = 1
    for s_ in strings:
        if len(s_) < 16 or len(s_) % 4 or s_[:4] != LANE_MAGIC:
            raise ValueError("rans_decode_device: not a lane container (these streams come from the host / reference coder?)")
        max_lanes = max(max_lanes, int.from_bytes(s_[8:12], "little"))
    if max_lanes > 1024:
        raise ValueError("rans_decode_device: malformed header (lane count)")
    tab, lens, offs = _i32dev(cdf, dev), _i32dev(cdf_lengths, dev).reshape(-1), _i32dev(offsets, dev).reshape(-1)
    sizes = np.array([len(s_) for s_ in strings], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64) if B else np.zeros(0, np.int64)
    blob = torch.from_numpy(np.frombuffer(b"".join(strings), dtype=np.uint8).copy() if B else np.zeros(4, np.uint8)).to(dev)
    meta = torch.from_numpy(np.concatenate([starts, sizes, [0]]).astype(np.int64)).to(dev)
    out = torch.empty(idx.shape, dtype=torch.int32, device=dev)
    with _Timed("rans_decode|coder"):
        L.check(L.lib().mmc_rans_decode_device(_ptr(blob), _ptr(meta), _ptr(meta[B:]), _ptr(idx), B, n, max_lanes, _ptr(tab), tab.shape[0], tab.shape[1],
                                               _ptr(lens), _ptr(offs), _ptr(out), _ptr(meta[2 * B:]), _stream()))
    st = int(meta[2 * B].item()) & 0xffffffff
    if st:
        raise _lane_status_error("rans_decode_device", st)
    return out


# ---- GDN -----------------------------------------------------------------------------------------
def gdn_reparam(beta: Tensor, gamma: Tensor, beta_bound: float, gamma_bound: float, pedestal: float,
                want_bf16: bool = False):
    _require_cuda(beta, gamma)
    beta, gamma = _f32c(beta.detach()), _f32c(gamma.detach())
    C = beta.numel()
    be = torch.empty_like(beta)
    ge = torch.empty_like(gamma)
    gb = torch.empty_like(gamma, dtype=torch.bfloat16) if want_bf16 else None
    L.check(L.lib().mmc_gdn_reparam(_ptr(beta), _ptr(gamma), C, float(beta_bound), float(gamma_bound), float(pedestal),
                                    _ptr(be), _ptr(ge), _ptr(gb), _stream()))
    return be, ge, gb


def gdn_forward(x: Tensor, beta_eff: Tensor, gamma_eff: Tensor, inverse: bool) -> Tensor:
    """GDN.forward on a logical (B, C, H, W) fp32 tensor (layers/gdn.py:77-92)."""
    _require_cuda(x)
    if x.dim() != 4:
        raise ValueError("GDN expects a 4-D (B, C, H, W) tensor")
    xv = x.float()
    B, C, H, W = xv.shape
    if _is_channels_last(xv):
        layout = L.NHWC
    else:
        xv, layout = xv.contiguous(), L.NCHW
    y = torch.empty_like(xv)
    L.check(L.lib().mmc_gdn_forward(_ptr(xv), _ptr(beta_eff), _ptr(gamma_eff), int(bool(inverse)), B, C, H * W, layout,
                                    _ptr(y), _stream()))
    return y


# ---- convolutions ----------------------------------------------------------------------------------
def conv_desc(transposed, B, H, W, Cin, Cout, k, stride, in_dtype, in_layout, out_dtype, out_layout, act=L.ACT_NONE,
              gdn=L.GDN_NONE, out2=0) -> L.ConvDesc:
    return L.ConvDesc(int(bool(transposed)), B, H, W, Cin, Cout, k, stride, in_dtype, in_layout, out_dtype, out_layout,
                      act, gdn, int(out2))


def conv_out_size(d: L.ConvDesc) -> Tuple[int, int]:
    ho, wo = ctypes.c_int(), ctypes.c_int()
    L.check(L.lib().mmc_conv_out_size(ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo)))
    return ho.value, wo.value


def _alloc_out(d: L.ConvDesc, device):
    # mmc_conv_out_size without the call: conv ceil(H / stride), deconv H * stride
    Ho, Wo = (d.H * d.stride, d.W * d.stride) if d.transposed else (-(-d.H // d.stride), -(-d.W // d.stride))
    dt = torch.float32 if d.out_dtype == L.F32 else torch.bfloat16
    shape = (d.B, d.Cout, Ho, Wo) if d.out_layout == L.NCHW else (d.B, Ho, Wo, d.Cout)
    y = torch.empty(shape, dtype=dt, device=device)
    y2 = torch.empty((d.B, Ho, Wo, d.Cout), dtype=torch.bfloat16, device=device) if d.out2_bf16 else None
    return y, y2


def conv_forward_direct(d: L.ConvDesc, x: Tensor, w: Tensor, bias: Optional[Tensor], beta_eff=None, gamma_eff=None,
                        name: str = "conv"):
    """CUDA-core conv / deconv.  x is the raw buffer in the layout/dtype the descriptor names."""
    _require_cuda(x, w)
    y, y2 = _alloc_out(d, x.device)
    with _Timed(name + "|direct", conv_flops(d) if _profile is not None else 0.0):
        L.check(L.lib().mmc_conv_forward_direct(ctypes.byref(d), _ptr(x), _ptr(w), _ptr(bias), _ptr(beta_eff),
                                                _ptr(gamma_eff), _ptr(y), _ptr(y2), _stream()))
    return (y, y2) if d.out2_bf16 else y


def conv_pack_weights(d: L.ConvDesc, w: Tensor) -> Tensor:
    _require_cuda(w)
    nbytes = ctypes.c_size_t()
    L.check(L.lib().mmc_conv_pack_weights(ctypes.byref(d), None, None, ctypes.byref(nbytes), None))
    packed = torch.empty(nbytes.value, dtype=torch.uint8, device=w.device)
    wf = _f32c(w.detach())
    L.check(L.lib().mmc_conv_pack_weights(ctypes.byref(d), _ptr(wf), _ptr(packed), ctypes.byref(nbytes), _stream()))
    return packed


def conv_forward_tc(d: L.ConvDesc, x: Tensor, w_packed: Tensor, bias: Optional[Tensor], beta_eff=None,
                    gamma_bf16=None, name: str = "conv"):
    """Tensor-core (tcgen05) implicit-GEMM conv / deconv on NHWC bf16 input.  ``x`` may be a pair (x1, x2) of NHWC tensors whose
    channels are concatenated by the kernel's K loop (no torch.cat copy)."""
    if isinstance(x, (tuple, list)):
        x1, x2 = x
        _require_cuda(x1, x2, w_packed)
        y, y2 = _alloc_out(d, x1.device)
        with _Timed(name + "|tc", conv_flops(d) if _profile is not None else 0.0):
            L.check(L.lib().mmc_conv_forward_tc2(ctypes.byref(d), _ptr(x1), int(x1.shape[-1]), _ptr(x2), _ptr(w_packed), _ptr(bias),
                                                 _ptr(beta_eff), _ptr(gamma_bf16), _ptr(y), _ptr(y2), _stream()))
        return (y, y2) if d.out2_bf16 else y
    _require_cuda(x, w_packed)
    y, y2 = _alloc_out(d, x.device)
    with _Timed(name + "|tc", conv_flops(d) if _profile is not None else 0.0):
        try:
            L.check(L.lib().mmc_conv_forward_tc(ctypes.byref(d), _ptr(x), _ptr(w_packed), _ptr(bias), _ptr(beta_eff),
                                                _ptr(gamma_bf16), _ptr(y), _ptr(y2), _stream()))
        except L.MmcodecError as e:
            raise L.MmcodecError(f"{e} [layer {name}: B={d.B} {d.H}x{d.W} {d.Cin}->{d.Cout} k{d.k} s{d.stride} "
                                 f"{'deconv' if d.transposed else 'conv'} gdn={d.gdn}]") from None
    return (y, y2) if d.out2_bf16 else y


def pad_to_nhwc8(x: Tensor, d: L.ConvDesc) -> Tensor:
    """fp32 NCHW image (<= 8 channels) -> zero-padded bf16 [B][Hp][Wp][8] staging buffer for a NHWC_PAD8 conv."""
    _require_cuda(x)
    x = _f32c(x)
    B, C, H, W = x.shape
    hp, wp = ctypes.c_int(), ctypes.c_int()
    L.check(L.lib().mmc_conv_pad8_size(ctypes.byref(d), ctypes.byref(hp), ctypes.byref(wp)))
    out = torch.empty((B, hp.value, wp.value, 8), dtype=torch.bfloat16, device=x.device)
    with _Timed("pad8|layout", float(B) * (C * H * W * 4.0 + hp.value * wp.value * 16.0)):
        L.check(L.lib().mmc_pad_nchw_to_nhwc8(_ptr(x), B, C, H, W, d.k // 2, hp.value, wp.value, _ptr(out), _stream()))
    return out


def nchw_to_nhwc_bf16(x: Tensor) -> Tensor:
    _require_cuda(x)
    x = _f32c(x)
    B, C, H, W = x.shape
    y = torch.empty((B, H, W, C), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().mmc_nchw_f32_to_nhwc_bf16(_ptr(x), B, C, H * W, _ptr(y), _stream()))
    return y


def to_bf16(x: Tensor) -> Tensor:
    """fp32 -> bf16 copy of a dense tensor, keeping its memory layout."""
    _require_cuda(x)
    x = x.float()
    if not (x.is_contiguous() or _is_channels_last(x)):
        x = x.contiguous()
    y = torch.empty_like(x, dtype=torch.bfloat16)
    L.check(L.lib().mmc_f32_to_bf16(_ptr(x), x.numel(), _ptr(y), _stream()))
    return y


def split_bf16x3(x_nhwc: Tensor, square: bool = False) -> Tensor:
    """(B, H, W, C) fp32 -> (B, H, W, 3C) bf16 [hi | lo | hi] of x (or of x^2) (fp32 precision mode, see mmc_split_f32_bf16x3)."""
    _require_cuda(x_nhwc)
    x = _f32c(x_nhwc)
    B, H, W, C = x.shape
    y = torch.empty((B, H, W, 3 * C), dtype=torch.bfloat16, device=x.device)
    with _Timed("split_bf16x3|layout", x.numel() * 10.0):
        L.check(L.lib().mmc_split_f32_bf16x3(_ptr(x), B * H * W, C, int(bool(square)), _ptr(y), _stream()))
    return y


def gdn_apply_f32(x: Tensor, norm: Tensor, inverse: bool) -> Tensor:
    """y = x * rsqrt(norm) (inverse: x * sqrt(norm)) on fp32 tensors of identical layout (fp32-mode GDN, see mmc_gdn_apply_f32)."""
    _require_cuda(x, norm)
    x, norm = _f32c(x), _f32c(norm)
    y = torch.empty_like(x)
    with _Timed("gdn_apply|layout", x.numel() * 12.0):
        L.check(L.lib().mmc_gdn_apply_f32(_ptr(x), _ptr(norm), int(bool(inverse)), x.numel(), _ptr(y), _stream()))
    return y


def nhwc_bf16_to_nchw(x: Tensor) -> Tensor:
    _require_cuda(x)
    B, H, W, C = x.shape
    y = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    L.check(L.lib().mmc_nhwc_bf16_to_nchw_f32(_ptr(x.contiguous()), B, C, H * W, _ptr(y), _stream()))
    return y


# ---- scale-space flow (ssf2020) ----------------------------------------------------------------------
def gaussian_volume(x: Tensor, kernel1d: Tensor, num_levels: int) -> Tensor:
    """ScaleSpaceFlow.gaussian_volume (models/video/google.py:331-355): (N, C, H, W) fp32 -> (N, C, num_levels + 1, H, W).
    `kernel1d` is gaussian_kernel1d(k, sigma) (any device; its host copy parametrises the blur kernel)."""
    _require_cuda(x)
    x = _f32c(x)
    N, C, H, W = x.shape
    k = np.ascontiguousarray(kernel1d.detach().cpu().numpy(), dtype=np.float32)
    nb = ctypes.c_size_t()
    L.check(L.lib().mmc_gaussian_volume_workspace(N * C, H, W, ctypes.byref(nb)))
    ws = torch.empty(max(nb.value, 16), dtype=torch.uint8, device=x.device)
    vol = torch.empty((N, C, num_levels + 1, H, W), dtype=torch.float32, device=x.device)
    with _Timed("gaussian_volume|scale_space"):
        L.check(L.lib().mmc_gaussian_volume(_ptr(x), N * C, H, W, k.ctypes.data, k.size, int(num_levels), _ptr(ws), nb.value,
                                            _ptr(vol), _stream()))
    return vol


def scale_space_warp(volume: Tensor, motion_info: Tensor, base_x: Tensor, base_y: Tensor, x_cur: Optional[Tensor] = None):
    """ScaleSpaceFlow.warp_volume (models/video/google.py:357-375) -> x_pred [, x_cur - x_pred]."""
    _require_cuda(volume, motion_info, base_x, base_y, x_cur)
    if volume.dim() != 5:
        raise ValueError(f"Invalid number of dimensions for volume {volume.dim()}")
    volume, motion_info = _f32c(volume), _f32c(motion_info)
    N, C, D, H, W = volume.shape
    if tuple(motion_info.shape) != (N, 3, H, W):
        raise ValueError("motion_info must be (N, 3, H, W): flow x, flow y, scale field")
    x_pred = torch.empty((N, C, H, W), dtype=torch.float32, device=volume.device)
    x_res = None
    if x_cur is not None:
        x_cur = _f32c(x_cur)
        if x_cur.shape != x_pred.shape:
            raise ValueError("x_cur must have the shape of the prediction")
        x_res = torch.empty_like(x_pred)
    with _Timed("warp_volume|scale_space"):
        L.check(L.lib().mmc_scale_space_warp(_ptr(volume), _ptr(motion_info), _ptr(_f32c(base_x)), _ptr(_f32c(base_y)), N, C, D, H, W,
                                             _ptr(x_cur), _ptr(x_pred), _ptr(x_res), _stream()))
    return (x_pred, x_res) if x_cur is not None else x_pred


def add(a: Tensor, b: Tensor) -> Tensor:
    _require_cuda(a, b)
    a, b = _f32c(a), _f32c(b)
    if a.shape != b.shape:
        raise ValueError("add: shape mismatch")
    out = torch.empty_like(a)
    L.check(L.lib().mmc_add(_ptr(a), _ptr(b), a.numel(), _ptr(out), _stream()))
    return out


# ---- token-side kernels of the window cross-attention (master.py:484-742) -------------------------------------
def _bf16c(t: Tensor) -> Tensor:
    if t.dtype != torch.bfloat16:
        raise TypeError("expected a bf16 tensor")
    return t if t.is_contiguous() else t.contiguous()


def layernorm_bf16(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-5, delta: Optional[Tensor] = None, want_sum: bool = False):
    """LayerNorm over the last dimension of a bf16 tensor; with ``delta`` computes LayerNorm(x + delta) and, with ``want_sum``,
    also returns the bf16 sum (the updated residual stream): (sum, normed)."""
    _require_cuda(x, weight, bias)
    x = _bf16c(x)
    C = x.shape[-1]
    if delta is not None:
        delta = _bf16c(delta)
        if delta.shape != x.shape:
            raise ValueError("layernorm_bf16: delta shape mismatch")
    y = torch.empty_like(x)
    s = torch.empty_like(x) if (want_sum and delta is not None) else None
    with _Timed("layernorm|attn"):
        L.check(L.lib().mmc_layernorm_bf16(_ptr(x), _ptr(delta), _ptr(_f32c(weight.detach())), _ptr(_f32c(bias.detach())), x.numel() // C, C,
                                           float(eps), _ptr(s), _ptr(y), _stream()))
    return (s, y) if want_sum else y


def gelu_bf16(x: Tensor) -> Tensor:
    _require_cuda(x)
    x = _bf16c(x)
    y = torch.empty_like(x)
    with _Timed("gelu|attn"):
        L.check(L.lib().mmc_gelu_bf16(_ptr(x), x.numel(), _ptr(y), _stream()))
    return y


def window_attention(q: Tensor, kv: Tensor, bias_table: Tensor, window: int, shift: int, heads: int, scale: float) -> Tensor:
    """Windowed multi-head cross-attention on (B, H, W, C) bf16 token grids (see mmc_window_attention)."""
    _require_cuda(q, kv, bias_table)
    q, kv = _bf16c(q), _bf16c(kv)
    B, H, W, C = q.shape
    if kv.shape != (B, H, W, 2 * C) or C % heads:
        raise ValueError("window_attention: kv must be (B, H, W, 2C) and C a multiple of heads")
    if tuple(bias_table.shape) != ((2 * window - 1) ** 2, heads):
        raise ValueError("window_attention: bias table must be ((2 ws - 1)^2, heads)")
    out = torch.empty_like(q)
    with _Timed("window_attention|attn"):
        L.check(L.lib().mmc_window_attention(_ptr(q), _ptr(kv), _ptr(_f32c(bias_table.detach())), B, H, W, heads, C // heads, window, shift,
                                             float(scale), _ptr(out), _stream()))
    return out


def channel_mean(x: Tensor) -> Tensor:
    """(B, H, W, C) fp32 NHWC -> (B, C) spatial mean (AdaptiveAvgPool2d(1)); per-sample result independent of the batch size."""
    _require_cuda(x)
    x = _f32c(x)
    B, H, W, C = x.shape
    out = torch.empty((B, C), dtype=torch.float32, device=x.device)
    nbytes = ctypes.c_size_t()
    L.check(L.lib().mmc_channel_mean_workspace(B, H * W, C, ctypes.byref(nbytes)))
    ws = torch.empty(max(1, nbytes.value), dtype=torch.uint8, device=x.device)
    with _Timed("channel_mean|attn"):
        L.check(L.lib().mmc_channel_mean(_ptr(x), B, H * W, C, _ptr(ws), _ptr(out), _stream()))
    return out


def channel_affine_bf16(x: Tensor, gamma: Tensor, beta: Tensor) -> Tensor:
    """y = gamma[b, c] * x + beta[b, c] on a (B, H, W, C) bf16 map, gamma / beta (B, C) fp32."""
    _require_cuda(x, gamma, beta)
    x = _bf16c(x)
    B, H, W, C = x.shape
    if tuple(gamma.shape) != (B, C) or tuple(beta.shape) != (B, C):
        raise ValueError("channel_affine_bf16: gamma / beta must be (B, C)")
    y = torch.empty_like(x)
    with _Timed("channel_affine|attn"):
        L.check(L.lib().mmc_channel_affine_bf16(_ptr(x), _ptr(_f32c(gamma)), _ptr(_f32c(beta)), B, H * W, C, _ptr(y), _stream()))
    return y


# ---- non-convolution steps of the ESA gate (google.py:1445-1459) ------------------------------------------------
def maxpool_nhwc_bf16(x: Tensor, k: int, stride: int) -> Tensor:
    _require_cuda(x)
    x = _bf16c(x)
    B, H, W, C = x.shape
    if H < k or W < k:
        raise ValueError(f"max-pool window {k} does not fit the {H}x{W} map")
    y = torch.empty((B, (H - k) // stride + 1, (W - k) // stride + 1, C), dtype=torch.bfloat16, device=x.device)
    with _Timed("maxpool|esa"):
        L.check(L.lib().mmc_maxpool_nhwc_bf16(_ptr(x), B, H, W, C, k, stride, _ptr(y), _stream()))
    return y


def upsample_bilinear_add_bf16(small: Tensor, add: Tensor) -> Tensor:
    """F.interpolate(small, add's size, mode="bilinear", align_corners=False) + add, NHWC bf16."""
    _require_cuda(small, add)
    small, add = _bf16c(small), _bf16c(add)
    B, hs, ws, C = small.shape
    if add.shape[0] != B or add.shape[3] != C:
        raise ValueError("upsample_bilinear_add_bf16: batch / channel mismatch")
    y = torch.empty_like(add)
    with _Timed("upsample_add|esa"):
        L.check(L.lib().mmc_upsample_bilinear_add_bf16(_ptr(small), B, hs, ws, C, _ptr(add), add.shape[1], add.shape[2], _ptr(y), _stream()))
    return y


def sigmoid_gate_bf16(x: Tensor, gate: Tensor) -> Tensor:
    _require_cuda(x, gate)
    x, gate = _bf16c(x), _bf16c(gate)
    if x.shape != gate.shape:
        raise ValueError("sigmoid_gate_bf16: shape mismatch")
    y = torch.empty_like(x)
    with _Timed("sigmoid_gate|esa"):
        L.check(L.lib().mmc_sigmoid_gate_bf16(_ptr(x), _ptr(gate), x.numel(), _ptr(y), _stream()))
    return y


# ---- backward passes of the fusion layers' non-convolution steps (csrc/fusion_bwd.cu) ---------------------------------
def maxpool_nhwc_bf16_idx(x: Tensor, k: int, stride: int):
    """Max-pool that also returns the arg-max map (uint8, window-local index) for ``maxpool_nhwc_bf16_bwd``."""
    _require_cuda(x)
    x = _bf16c(x)
    B, H, W, C = x.shape
    if H < k or W < k:
        raise ValueError(f"max-pool window {k} does not fit the {H}x{W} map")
    shape = (B, (H - k) // stride + 1, (W - k) // stride + 1, C)
    y = torch.empty(shape, dtype=torch.bfloat16, device=x.device)
    idx = torch.empty(shape, dtype=torch.uint8, device=x.device)
    with _Timed("maxpool|esa"):
        L.check(L.lib().mmc_maxpool_nhwc_bf16_idx(_ptr(x), B, H, W, C, k, stride, _ptr(y), _ptr(idx), _stream()))
    return y, idx


def maxpool_nhwc_bf16_bwd(gy: Tensor, idx: Tensor, in_shape, k: int, stride: int) -> Tensor:
    _require_cuda(gy, idx)
    gy = _bf16c(gy)
    B, H, W, C = in_shape
    if tuple(gy.shape) != tuple(idx.shape) or idx.dtype != torch.uint8:
        raise ValueError("maxpool_nhwc_bf16_bwd: gy / idx mismatch")
    dx = torch.empty((B, H, W, C), dtype=torch.bfloat16, device=gy.device)
    with _Timed("maxpool_bwd|esa"):
        L.check(L.lib().mmc_maxpool_nhwc_bf16_bwd(_ptr(gy), _ptr(idx.contiguous()), B, H, W, C, k, stride, _ptr(dx), _stream()))
    return dx


def upsample_bilinear_bwd_bf16(g: Tensor, hs: int, ws: int) -> Tensor:
    """Adjoint of the bilinear upsampling (hs, ws) -> g's size: (B, H, W, C) bf16 -> (B, hs, ws, C) bf16."""
    _require_cuda(g)
    g = _bf16c(g)
    B, H, W, C = g.shape
    out = torch.empty((B, hs, ws, C), dtype=torch.bfloat16, device=g.device)
    with _Timed("upsample_bwd|esa"):
        L.check(L.lib().mmc_upsample_bilinear_bwd_bf16(_ptr(g), B, H, W, C, hs, ws, _ptr(out), _stream()))
    return out


def sigmoid_gate_bwd_bf16(g: Tensor, x: Tensor, gate: Tensor):
    """(dx, dgate) of y = x * sigmoid(gate)."""
    _require_cuda(g, x, gate)
    g, x, gate = _bf16c(g), _bf16c(x), _bf16c(gate)
    if not (g.shape == x.shape == gate.shape):
        raise ValueError("sigmoid_gate_bwd_bf16: shape mismatch")
    dx, dgate = torch.empty_like(x), torch.empty_like(x)
    with _Timed("sigmoid_gate_bwd|esa"):
        L.check(L.lib().mmc_sigmoid_gate_bwd_bf16(_ptr(g), _ptr(x), _ptr(gate), x.numel(), _ptr(dx), _ptr(dgate), _stream()))
    return dx, dgate


def gelu_bwd_bf16(g: Tensor, x: Tensor) -> Tensor:
    _require_cuda(g, x)
    g, x = _bf16c(g), _bf16c(x)
    dx = torch.empty_like(x)
    with _Timed("gelu_bwd|attn"):
        L.check(L.lib().mmc_gelu_bwd_bf16(_ptr(g), _ptr(x), x.numel(), _ptr(dx), _stream()))
    return dx


def layernorm_bwd_bf16(g: Tensor, v: Tensor, weight: Tensor, eps: float = 1e-5, g_sum: Optional[Tensor] = None):
    """(dv bf16, dweight fp32 [C], dbias fp32 [C]) of y = LayerNorm(v); ``g_sum`` is added to dv (residual-stream gradient)."""
    _require_cuda(g, v, weight)
    g, v = _bf16c(g), _bf16c(v)
    C = v.shape[-1]
    if g.shape != v.shape or (g_sum is not None and g_sum.shape != v.shape):
        raise ValueError("layernorm_bwd_bf16: shape mismatch")
    if g_sum is not None:
        g_sum = _bf16c(g_sum)
    dv = torch.empty_like(v)
    dwb = torch.zeros((2, C), dtype=torch.float32, device=v.device)
    with _Timed("layernorm_bwd|attn"):
        L.check(L.lib().mmc_layernorm_bwd_bf16(_ptr(g), _ptr(v), _ptr(g_sum), _ptr(_f32c(weight.detach())), v.numel() // C, C, float(eps),
                                               _ptr(dv), _ptr(dwb[0]), _ptr(dwb[1]), _stream()))
    return dv, dwb[0], dwb[1]


def window_attention_bwd(q: Tensor, kv: Tensor, bias_table: Tensor, dout: Tensor, window: int, shift: int, heads: int, scale: float):
    """(dq, dkv, dtable) of ``window_attention``."""
    _require_cuda(q, kv, bias_table, dout)
    q, kv, dout = _bf16c(q), _bf16c(kv), _bf16c(dout)
    B, H, W, C = q.shape
    if kv.shape != (B, H, W, 2 * C) or dout.shape != q.shape or C % heads:
        raise ValueError("window_attention_bwd: shape mismatch")
    dq, dkv = torch.empty_like(q), torch.empty_like(kv)
    dtable = torch.zeros(((2 * window - 1) ** 2, heads), dtype=torch.float32, device=q.device)
    with _Timed("window_attention_bwd|attn"):
        L.check(L.lib().mmc_window_attention_bwd(_ptr(q), _ptr(kv), _ptr(_f32c(bias_table.detach())), _ptr(dout), B, H, W, heads, C // heads,
                                                 window, shift, float(scale), _ptr(dq), _ptr(dkv), _ptr(dtable), _stream()))
    return dq, dkv, dtable


def conv3x3_mean(t: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    """Spatial mean of conv3x3(t) (stride 1, padding 1) per sample and output channel, computed from border-corrected channel sums
    of ``t`` instead of the convolution (see mmc_conv3x3_mean).  t: (B, H, W, C) bf16 NHWC; weight (O, C, 3, 3); returns (B, O) fp32."""
    _require_cuda(t, weight)
    t = _bf16c(t)
    B, H, W, C = t.shape
    O = weight.shape[0]
    if tuple(weight.shape) != (O, C, 3, 3):
        raise ValueError("conv3x3_mean: weight must be (O, C, 3, 3)")
    nbytes = ctypes.c_size_t()
    L.check(L.lib().mmc_conv3x3_mean_workspace(B, H, W, C, ctypes.byref(nbytes)))
    ws = torch.empty(max(1, nbytes.value), dtype=torch.uint8, device=t.device)
    out = torch.empty((B, O), dtype=torch.float32, device=t.device)
    with _Timed("conv3x3_mean|attn"):
        L.check(L.lib().mmc_conv3x3_mean(_ptr(t), B, H, W, C, _ptr(_f32c(weight.detach())), _ptr(_f32c(bias.detach())) if bias is not None else None,
                                         O, _ptr(ws), _ptr(out), _stream()))
    return out


# ---- per-image metrics of a forward pass (eval_model/__main__t.py:151-173) ---------------------------------------
def _sample_contiguous(t: Tensor) -> Tensor:
    t = t if t.dtype == torch.float32 else t.float()
    return t if (t.is_contiguous() or _is_channels_last(t)) else t.contiguous()


def image_bits(likelihood: Tensor, accum: Tensor, scale: float) -> Tensor:
    """accum[b] += scale * -sum(log2(likelihood[b])); ``accum``: zero-initialised (B,) fp32."""
    _require_cuda(likelihood, accum)
    lk = _sample_contiguous(likelihood)
    B = lk.shape[0]
    L.check(L.lib().mmc_image_bits(_ptr(lk), B, lk.numel() // max(B, 1), float(scale), _ptr(accum), _stream()))
    return accum


def image_mse(a: Tensor, b: Tensor, accum: Tensor) -> Tensor:
    """accum[i] += mean((a[i] - b[i])^2); a, b: same shape and memory format."""
    _require_cuda(a, b, accum)
    a, b = _sample_contiguous(a), _sample_contiguous(b)
    if a.shape != b.shape or a.stride() != b.stride():
        b = b.contiguous()
        a = a.contiguous()
    B = a.shape[0]
    n = a.numel() // max(B, 1)
    L.check(L.lib().mmc_image_sse(_ptr(a), _ptr(b), B, n, 1.0 / max(n, 1), _ptr(accum), _stream()))
    return accum


def u8_to_f32(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """uint8 image samples -> fp32 / 255 (ToTensor), into ``out`` if given (contiguous fp32 of the same shape)."""
    _require_cuda(x)
    if x.dtype != torch.uint8 or not x.is_contiguous():
        raise TypeError("u8_to_f32: expected a contiguous uint8 tensor")
    if out is None:
        out = torch.empty(x.shape, dtype=torch.float32, device=x.device)
    elif out.dtype != torch.float32 or not out.is_contiguous() or out.shape != x.shape:
        raise TypeError("u8_to_f32: out must be a contiguous fp32 tensor of the same shape")
    L.check(L.lib().mmc_u8_to_f32(_ptr(x), x.numel(), _ptr(out), _stream()))
    return out


# ---- backward of the transforms ------------------------------------------------------------------------
def wgrad(s_nhwc: Tensor, l_nhwc: Tensor, k: int, stride: int, scale: float = 1.0, mask: Optional[Tensor] = None,
          name: str = "conv") -> Tensor:
    """Weight gradient (Cs, Cl, k, k) fp32 of a conv / deconv layer from its low-resolution side S and high-resolution side
    L, both NHWC bf16 with channel counts that are multiples of 8 (see mmc_wgrad_tc)."""
    _require_cuda(s_nhwc, l_nhwc)
    if s_nhwc.dtype != torch.bfloat16 or l_nhwc.dtype != torch.bfloat16:
        raise ValueError("wgrad expects NHWC bf16 operands")
    s_nhwc, l_nhwc = s_nhwc.contiguous(), l_nhwc.contiguous()
    B, Hs, Ws, Cs = s_nhwc.shape
    _, Hl, Wl, Cl = l_nhwc.shape
    if stride == 1 and mask is None and Cs % 128 != 0 and Cl % 128 == 0 and Cs <= 256 and (Hs, Ws) == (Hl, Wl):
        # Stride 1: the two sides are interchangeable (dW[cs][cl][k] = R[cl][cs][K-1-k] with the roles swapped).  Put the side
        # whose channel count fills the 128-row M tiles on M (tran_conv: 384 x 192 instead of a half-empty second tile of 192).
        r = wgrad(l_nhwc, s_nhwc, k, 1, scale, None, name)
        return r.permute(1, 0, 2, 3).flip(2, 3).contiguous()
    ws = torch.zeros((k * k, Cs, Cl), dtype=torch.float32, device=s_nhwc.device)
    dw = torch.empty((Cs, Cl, k, k), dtype=torch.float32, device=s_nhwc.device)
    with _Timed(name + "|wgrad", 2.0 * Cs * Cl * k * k * B * Hs * Ws if _profile is not None else 0.0):
        L.check(L.lib().mmc_wgrad_tc(_ptr(s_nhwc), _ptr(l_nhwc), B, Cs, Cl, Hs, Ws, Hl, Wl, k, stride, _ptr(ws), _stream()))
    m = _f32c(mask) if mask is not None else None
    L.check(L.lib().mmc_wgrad_finalize(_ptr(ws), k, Cs, Cl, float(scale), _ptr(m), 0, _ptr(dw), _stream()))
    return dw


def wgrad_edge(s_nhwc: Tensor, l_nhwc8: Tensor, k: int, stride: int, name: str = "conv") -> Tensor:
    """Weight gradient (Cs, 8, k, k) of an edge layer whose high-resolution side has <= 8 channels (padded to 8): the k x k
    neighbourhoods are gathered into k*k*8 channels and the gradient is one k = 1 GEMM (see mmc_im2col8)."""
    _require_cuda(s_nhwc, l_nhwc8)
    s_nhwc, l_nhwc8 = s_nhwc.contiguous(), l_nhwc8.contiguous()
    B, Hs, Ws, Cs = s_nhwc.shape
    _, H, W, c8 = l_nhwc8.shape
    if c8 != 8:
        raise ValueError("wgrad_edge expects the narrow tensor padded to 8 channels")
    col = torch.empty((B, Hs, Ws, k * k * 8), dtype=torch.bfloat16, device=s_nhwc.device)
    L.check(L.lib().mmc_im2col8(_ptr(l_nhwc8), B, H, W, k, stride, Hs, Ws, _ptr(col), _stream()))
    dw = wgrad(s_nhwc, col, 1, 1, name=name)                       # (Cs, k*k*8, 1, 1)
    return dw.reshape(Cs, k, k, 8).permute(0, 3, 1, 2)


def nchw_to_nhwc8(x: Tensor) -> Tensor:
    """fp32 NCHW tensor with <= 8 channels -> bf16 NHWC with the channels zero-padded to 8 (no spatial padding)."""
    _require_cuda(x)
    x = _f32c(x)
    B, C, H, W = x.shape
    out = torch.empty((B, H, W, 8), dtype=torch.bfloat16, device=x.device)
    L.check(L.lib().mmc_pad_nchw_to_nhwc8(_ptr(x), B, C, H, W, 0, H, W, _ptr(out), _stream()))
    return out


def act_bwd(grad_out: Tensor, y: Tensor, act: int) -> Tensor:
    _require_cuda(grad_out, y)
    g, y = grad_out.contiguous(), y.contiguous()
    out = torch.empty_like(g)
    L.check(L.lib().mmc_act_bwd(_ptr(g), _ptr(y), int(act), g.numel(), _ptr(out), _stream()))
    return out


def colsum(g: Tensor, scale: float = 1.0) -> Tensor:
    """fp32 [C] = scale * sum over all leading dims of a (..., C) bf16 tensor (bias / beta gradients)."""
    _require_cuda(g)
    g = g.contiguous()
    C = g.shape[-1]
    out = torch.zeros(C, dtype=torch.float32, device=g.device)
    L.check(L.lib().mmc_colsum_bf16(_ptr(g), g.numel() // C, C, float(scale), _ptr(out), _stream()))
    return out


def square(x: Tensor) -> Tensor:
    _require_cuda(x)
    x = x.contiguous()
    out = torch.empty_like(x)
    L.check(L.lib().mmc_square_bf16(_ptr(x), x.numel(), _ptr(out), _stream()))
    return out


def gdn_bwd_t(grad_out: Tensor, x: Tensor, norm: Tensor, inverse: bool) -> Tensor:
    _require_cuda(grad_out, x, norm)
    t = torch.empty_like(x)
    L.check(L.lib().mmc_gdn_bwd_t(_ptr(grad_out.contiguous()), _ptr(x), _ptr(norm), int(bool(inverse)), x.numel(), _ptr(t), _stream()))
    return t


def gdn_bwd_dx(grad_out: Tensor, x: Tensor, norm: Tensor, u: Tensor, inverse: bool) -> Tensor:
    _require_cuda(grad_out, x, norm, u)
    dx = torch.empty_like(x)
    L.check(L.lib().mmc_gdn_bwd_dx(_ptr(grad_out.contiguous()), _ptr(x), _ptr(norm), _ptr(u), int(bool(inverse)), x.numel(), _ptr(dx), _stream()))
    return dx


def reparam_bwd(p: Tensor, dp_eff: Tensor, bound: float) -> Tensor:
    _require_cuda(p, dp_eff)
    p, d = _f32c(p.detach()), _f32c(dp_eff)
    out = torch.empty_like(p)
    L.check(L.lib().mmc_reparam_bwd(_ptr(p), _ptr(d), float(bound), p.numel(), _ptr(out), _stream()))
    return out


def gc_backward(xv: Tensor, s: Tensor, m: Optional[Tensor], nz: Optional[Tensor], g: Tensor, scale_bound: float, likelihood_bound: float):
    """Gradients of GaussianConditional.forward's likelihood; all tensors share xv's memory layout."""
    _require_cuda(xv, s, m, nz, g)
    dx = torch.empty_like(xv) if nz is not None else None
    ds = torch.empty_like(xv)
    dm = torch.empty_like(xv) if (m is not None and nz is not None) else None
    L.check(L.lib().mmc_gc_backward(_ptr(xv), _ptr(s), _ptr(m), _ptr(nz), _ptr(g), float(scale_bound), float(likelihood_bound), xv.numel(),
                                    _ptr(dx), _ptr(ds), _ptr(dm), _stream()))
    return dx, ds, dm


def eb_backward(xv: Tensor, nz: Tensor, g: Tensor, params, likelihood_bound: float, outer: int, C: int, inner: int):
    """dx and packed parameter gradients [C][58] of EntropyBottleneck.forward (noise mode)."""
    _require_cuda(xv, nz, g)
    p, _keep = params
    dx = torch.empty_like(xv)
    dparams = torch.zeros((C, 58), dtype=torch.float32, device=xv.device)
    L.check(L.lib().mmc_eb_backward(_ptr(xv), _ptr(nz), _ptr(g), ctypes.byref(p), float(likelihood_bound), outer, C, inner, _ptr(dx),
                                    _ptr(dparams), _stream()))
    return dx, dparams


# ---- optional per-launch device timing (bench.py roofline pass; off by default) ---------------------
_profile = None


def start_profile():
    global _profile
    _profile = []
    return _profile


def stop_profile(with_work: bool = False):
    """Returns {name: mean ms per launch} for everything recorded since start_profile()
    (with_work: {name: (ms, algorithmic FLOPs per launch)})."""
    global _profile
    rec, _profile = _profile or [], None
    torch.cuda.synchronize()
    acc, work = {}, {}
    for name, e0, e1, flops in rec:
        acc.setdefault(name, []).append(e0.elapsed_time(e1))
        work[name] = work.get(name, 0.0) + flops
    if with_work == "total":    # {name: (total ms, total work, launches)} over everything recorded
        return {k: (sum(v), work[k], len(v)) for k, v in acc.items()}
    if with_work:
        return {k: (sum(v) / len(v), work[k] / len(v)) for k, v in acc.items()}
    return {k: sum(v) / len(v) for k, v in acc.items()}


# profile entries whose `work` is algorithmic BYTES (HBM-bound kernels), not FLOPs
HBM_KERNEL_SUFFIXES = ("|entropy", "|gc", "|eb", "|eb_lut", "|layout")


def conv_flops(d: L.ConvDesc) -> float:
    """Algorithmic FLOPs of one conv / deconv launch: 2*Cin*Cout*k*k*B*Hout*Wout for a conv,
    2*Cin*Cout*k*k*B*Hin*Win for a transposed conv, plus 2*C*C*B*Hout*Wout when GDN / IGDN is fused."""
    Ho, Wo = conv_out_size(d)
    pix = d.H * d.W if d.transposed else Ho * Wo
    f = 2.0 * d.Cin * d.Cout * d.k * d.k * d.B * pix
    if d.gdn != L.GDN_NONE:
        f += 2.0 * d.Cout * d.Cout * d.B * Ho * Wo
    return f


class _Timed:
    def __init__(self, name, flops: float = 0.0):
        self.name = name
        self.flops = flops

    def __enter__(self):
        if _profile is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if _profile is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _profile.append((self.name, self.e0, e1, self.flops))


def launch_count() -> int:
    return int(L.lib().mmc_launch_count())


def reset_launch_count() -> None:
    L.lib().mmc_reset_launch_count()

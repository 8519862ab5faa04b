"""GPU tests of the device coder (csrc/rans_device.cu).

* byte-exact against the CPU restatement of the container (oracle/lane_rans.py) on small cases: ragged sizes, lane counts that do
  not divide the symbol count, fewer symbols than lanes, empty input, escape (bypass) symbols at the int32 extremes;
* at BASELINE sizes (1088 x 1920 latents: 1.57 M symbols per image) through size-independent properties: decode(encode(s)) == s,
  the stream carries the same symbols as the reference-compatible host coder's, and costs at most the lane headers more;
* malformed input fails loudly; model-level compress / decompress with the coder selected reproduces the host path's x_hat bit for
  bit (same symbols -> same reconstruction)."""
import numpy as np
import pytest
import torch

from oracle import lane_rans

pytestmark = pytest.mark.gpu

import mmcodec  # noqa: E402
from mmcodec import ops  # noqa: E402


def dev():
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def gc_tables():
    gc = mmcodec.GaussianConditional(None)
    gc.update_scale_table(mmcodec.models.get_scale_table())
    return gc._quantized_cdf, gc._cdf_length, gc._offset


def _draw(B, n, gen, escapes=True):
    idx = torch.randint(0, 64, (B, n), generator=gen, dtype=torch.int32)
    scale = torch.as_tensor(mmcodec.models.get_scale_table())[idx.long()]
    sym = torch.round(torch.randn(B, n, generator=gen) * scale).to(torch.int32)
    if escapes and n >= 8:
        sym[:, ::5] = torch.randint(-70000, 70000, sym[:, ::5].shape, generator=gen, dtype=torch.int32)
        sym[0, 1], sym[0, 2], sym[0, 3] = 2 ** 30, -2 ** 30, 12345678
    return sym, idx


@pytest.mark.parametrize("n,lanes", [(0, 4), (1, 4), (3, 8), (100, 1), (257, 7), (1000, 32), (4099, 33), (3000, None)])
def test_container_bytes_equal_the_cpu_restatement(gc_tables, n, lanes):
    gen = torch.Generator().manual_seed(n + (lanes or 0))
    sym, idx = _draw(2, n, gen)
    streams = ops.rans_encode_device(sym.to(dev()), idx.to(dev()), *gc_tables, lanes=lanes)
    cdfs, sizes, offs = (t.numpy() for t in gc_tables)
    for b in range(2):
        want = lane_rans.encode(sym[b].tolist(), idx[b].tolist(), cdfs, sizes, offs, lanes)
        assert streams[b] == want, (b, len(streams[b]), len(want))
    out = ops.rans_decode_device(streams, idx.to(dev()), *gc_tables)
    assert out.is_cuda and torch.equal(out.cpu(), sym)


def test_full_size_round_trip_and_rate_vs_host_coder(gc_tables):
    """cfg 3 latents: 192 x 68 x 120 symbols per 1088 x 1920 image, batch 4."""
    gen = torch.Generator().manual_seed(77)
    n = 192 * 68 * 120
    sym, idx = _draw(4, n, gen, escapes=False)
    sym[1, ::997] = 4000                                   # a sprinkle of escapes in one image
    sd, idd = sym.to(dev()), idx.to(dev())
    streams = ops.rans_encode_device(sd, idd, *gc_tables)
    assert all(s[:4] == ops.LANE_MAGIC for s in streams)
    S = ops.rans_lanes_default(n)
    assert S == lane_rans.lanes_default(n) == 256
    assert torch.equal(ops.rans_decode_device(streams, idd, *gc_tables).cpu(), sym)
    host = ops.rans_encode(sym, idx, *gc_tables)            # the reference-compatible stream of the same symbols
    assert torch.equal(ops.rans_decode(host, idx, *gc_tables), sym)
    for b in range(4):
        extra = len(streams[b]) - len(host[b])
        assert extra <= 16 + 14 * S + 4, (b, extra)      # 8 bytes of header + <= 6 bytes of flush / rounding per lane
        assert extra / len(host[b]) < 0.01                  # < 1 % rate overhead at this size and rate (~5 bits per symbol)
    # a second batch through the asynchronous form, two launches in flight
    h1 = ops.rans_encode_device_launch(sd[:2], idd[:2], *gc_tables, pinned={})
    h2 = ops.rans_encode_device_launch(sd[2:], idd[2:], *gc_tables, pinned={})
    torch.cuda.synchronize()
    assert h1.collect() + h2.collect() == streams


def test_escape_heavy_data_grows_the_output_and_round_trips(gc_tables):
    gen = torch.Generator().manual_seed(5)
    idx = torch.randint(0, 64, (2, 20000), generator=gen, dtype=torch.int32)
    sym = torch.randint(-2 ** 30, 2 ** 30, (2, 20000), generator=gen, dtype=torch.int32)      # every symbol escapes with 8 nibbles
    streams = ops.rans_encode_device(sym.to(dev()), idx.to(dev()), *gc_tables, lanes=16)
    assert min(len(s) for s in streams) > 2 * 20000 + 64     # beyond the first-guess capacity: the retry path ran
    assert torch.equal(ops.rans_decode_device(streams, idx.to(dev()), *gc_tables).cpu(), sym)


def test_malformed_input_fails_loudly(gc_tables):
    gen = torch.Generator().manual_seed(9)
    sym, idx = _draw(2, 5000, gen)
    sd, idd = sym.to(dev()), idx.to(dev())
    streams = ops.rans_encode_device(sd, idd, *gc_tables, lanes=8)
    with pytest.raises(ValueError):
        ops.rans_encode_device(sd, idd + 64, *gc_tables)                                   # index outside the table
    with pytest.raises(ValueError):
        ops.rans_decode_device([s[: len(s) // 2 // 4 * 4] for s in streams], idd, *gc_tables)   # truncated
    with pytest.raises(ValueError):
        ops.rans_decode_device(ops.rans_encode(sym, idx, *gc_tables), idd, *gc_tables)    # a host / reference stream
    with pytest.raises(ValueError):
        ops.rans_decode_device(streams, idd[:, :-1].contiguous(), *gc_tables)             # symbol count mismatch
    bad = bytearray(streams[0])
    bad[16] ^= 0x40                                                                        # corrupt lane 0's start state
    with pytest.raises(ValueError):
        ops.rans_decode_device([bytes(bad), streams[1]], idd, *gc_tables)
    with pytest.raises(RuntimeError):
        ops.rans_encode_device(sym, idx, *gc_tables)                                       # CPU tensors: no fallback


@pytest.mark.parametrize("arch", ["bmshj2018-factorized", "bmshj2018-hyperprior", "mbt2018-mean"])
def test_models_compress_with_the_device_coder(arch):
    torch.manual_seed(1)
    net = mmcodec.build_model(arch, 3).eval()
    # spread the latents so that the streams are not trivial
    with torch.no_grad():
        for p in net.g_a.parameters():
            if p.dim() == 4:
                p.mul_(3.0)
    net.update(force=True)
    net = net.to(dev())
    x = torch.rand(3, 3, 128, 192, generator=torch.Generator().manual_seed(2)).to(dev())
    with torch.no_grad():
        ref = net.compress(x)
        ref_hat = net.decompress(ref["strings"], ref["shape"])["x_hat"]
        mmcodec.set_entropy_coder(net, "ans-lanes")
        try:
            out = net.compress(x)
            hat = net.decompress(out["strings"], out["shape"])["x_hat"]
            pipe = mmcodec.CompressPipeline(net)
            x_pinned = x.cpu().pin_memory()                           # host input: uploaded on the pipeline's copy stream
            piped = [f.result() for f in [pipe.submit(x), pipe.submit(x_pinned), pipe.submit(x_pinned)]]
            pipe.close()
            with pytest.raises(ValueError):
                net.decompress(ref["strings"], ref["shape"])          # reference-format streams into the lane decoder
        finally:
            mmcodec.set_entropy_coder(net, "ans")
    assert all(s[:4] == ops.LANE_MAGIC for group in out["strings"] for s in group)
    assert torch.equal(hat, ref_hat)                                  # same symbols -> same reconstruction, bit for bit
    assert all(p["strings"] == out["strings"] for p in piped) and len(piped) == 3
    total = lambda strings: sum(len(s) for group in strings for s in group)
    # tiny images: the lane headers (<= 8 lanes per tensor: 80 bytes) are visible, the payload is not larger
    assert total(out["strings"]) <= total(ref["strings"]) + 3 * 2 * (16 + 14 * 8 + 4)
    with pytest.raises(ValueError):
        mmcodec.set_entropy_coder(net, "rangecoder")

"""mmcodec's coder and CDF construction against the REFERENCE'S OWN C++ (oracle/_ref).

``oracle/_ref`` holds the reference's two pybind11 extensions -- ``ans`` (compressai/cpp_exts/rans/rans_interface.cpp:108-359 on
third_party/ryg_rans) and ``_CXX`` (compressai/cpp_exts/ops/ops.cpp:40-109, ``pmf_to_quantized_cdf``) -- compiled by
``make -C oracle ref`` from the sources where they lie under /root/reference (``__graft_entry__.build()`` does it when that tree
exists; the binaries are git-ignored and travel to the GPU box with the snapshot).  They are the checker here, never the product:
  * the host rANS coder of libmmcodec (``ops.rans_encode`` / ``rans_decode``, csrc/rans.cu) must emit byte-identical streams and
    decode the reference's streams, on random streams that cover in-table symbols, escapes (bypass) and every table of the scale
    grid -- not only on the goldens captured from a model run;
  * ``ops.pmf_to_quantized_cdf`` (host) and the C oracle's restatement must equal ``_CXX.pmf_to_quantized_cdf`` on random pmfs,
    including the frequency-stealing loop (more symbols than precision headroom) and the reference's error cases.
The tests skip when the binaries are absent (a checkout that was never built next to the reference)."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
HAVE = bool(glob.glob(os.path.join(REF_DIR, "ans.*so"))) and bool(glob.glob(os.path.join(REF_DIR, "_CXX.*so")))
pytestmark = pytest.mark.skipif(not HAVE, reason="oracle/_ref not built (make -C oracle ref next to /root/reference)")

import mmcodec  # noqa: E402
from mmcodec import ops  # noqa: E402


def _ref():
    # pybind11 registers the coder classes process-wide: if the reference package itself was imported earlier in this process
    # (tests/test_accelerate.py builds it in a scratch directory), its `compressai.ans` IS the same code -- reuse it; otherwise load
    # oracle/_ref and publish it under the reference's module names so that a later import of the package reuses ours
    ans, cxx = sys.modules.get("compressai.ans"), sys.modules.get("compressai._CXX")
    if ans is None or cxx is None:
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import _CXX as cxx
        import ans
        sys.modules.setdefault("compressai.ans", ans)
        sys.modules.setdefault("compressai._CXX", cxx)
    return ans, cxx


def _tables():
    gc = mmcodec.GaussianConditional(None)
    gc.update_scale_table(mmcodec.models.get_scale_table())
    eb = mmcodec.EntropyBottleneck(8)
    eb.update()
    return gc, eb


# (streams of fewer than ~4 symbols are left out: the reference sizes its output buffer as one 32-bit word per symbol and the final
#  state flush writes two, rans_interface.cpp:176-200 -- a heap overrun in the reference itself for 1-symbol streams)
@pytest.mark.parametrize("seed,n,escape_every", [(1, 17, 0), (2, 5000, 0), (3, 5000, 7), (4, 60000, 3)])
def test_host_rans_coder_equals_reference_binary(seed, n, escape_every):
    ans, _ = _ref()
    gc, _eb = _tables()
    cdf, lens, offs = gc._quantized_cdf, gc._cdf_length, gc._offset
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, cdf.shape[0], (2, n), generator=g, dtype=torch.int32)
    scale = torch.tensor(mmcodec.models.get_scale_table())[idx.long()]
    sym = torch.round(torch.randn(2, n, generator=g) * scale).to(torch.int32)
    if escape_every:
        k = len(sym[1, ::escape_every])
        sym[1, ::escape_every] = torch.randint(-70000, 70000, (k,), generator=g, dtype=torch.int32)     # far outside the tables: bypass nibbles
    ours = ops.rans_encode(sym, idx, cdf, lens, offs)
    enc, dec = ans.RansEncoder(), ans.RansDecoder()
    cdf_l, lens_l, offs_l = cdf.tolist(), lens.reshape(-1).int().tolist(), offs.reshape(-1).int().tolist()
    for b in range(2):
        theirs = enc.encode_with_indexes(sym[b].tolist(), idx[b].tolist(), cdf_l, lens_l, offs_l)      # entropy_models.py:260-269
        assert ours[b] == theirs, (b, len(ours[b]), len(theirs))
        # cross decode: the reference decodes our stream (same bytes, but through its own reader) and we decode its stream
        assert dec.decode_with_indexes(ours[b], idx[b].tolist(), cdf_l, lens_l, offs_l) == sym[b].tolist()
    assert torch.equal(ops.rans_decode([enc.encode_with_indexes(sym[b].tolist(), idx[b].tolist(), cdf_l, lens_l, offs_l) for b in range(2)],
                                       idx, cdf, lens, offs), sym)


def test_host_rans_coder_equals_reference_binary_on_bottleneck_tables():
    ans, _ = _ref()
    _gc, eb = _tables()
    cdf, lens, offs = eb._quantized_cdf, eb._cdf_length, eb._offset
    g = torch.Generator().manual_seed(11)
    n = 4096
    idx = torch.arange(cdf.shape[0], dtype=torch.int32).repeat(n // cdf.shape[0]).reshape(1, n)
    sym = torch.randint(-12, 13, (1, n), generator=g, dtype=torch.int32)
    ours = ops.rans_encode(sym, idx, cdf, lens, offs)
    theirs = ans.RansEncoder().encode_with_indexes(sym[0].tolist(), idx[0].tolist(), cdf.tolist(), lens.reshape(-1).int().tolist(),
                                                   offs.reshape(-1).int().tolist())
    assert ours[0] == theirs


def test_pmf_to_quantized_cdf_equals_reference_binary():
    _, cxx = _ref()
    import oracle
    rs = np.random.RandomState(5)
    cases = []
    for L in (1, 2, 3, 17, 64, 300, 2000):
        p = rs.rand(L).astype(np.float32) ** 4 + 1e-9
        cases.append((p / p.sum()).astype(np.float32))
    cases.append(np.full(65536, 1.0 / 65536, np.float32))                   # as many symbols as 2^16: every frequency exactly 1
    cases.append(np.full(60000, 1.0 / 60000, np.float32))                   # rounding to 1 each leaves 5536 units to distribute
    cases.append(np.concatenate([np.full(65000, 1e-7, np.float32), np.array([0.9935], np.float32)]))    # stealing from one heavy symbol
    cases.append(np.array([0.1, 0.2, 0.0, 0.0], np.float32))                # tests/test_ops.py:104-106
    for p in cases:
        for precision in (16, 12) if len(p) < 2000 else (16,):
            want = cxx.pmf_to_quantized_cdf(p.tolist(), precision)
            assert ops.pmf_to_quantized_cdf(p.tolist(), precision) == want, (len(p), precision)
            got_c = oracle.pmf_to_quantized_cdf(p, precision)
            assert list(np.asarray(got_c).astype(np.int64)) == want, (len(p), precision)
    # (more symbols than 2^precision cannot all get a frequency >= 1: the reference then runs past `assert(best_steal != -1)`,
    #  ops.cpp:86-97, which NDEBUG removes -- undefined behaviour there, a loud ValueError here)
    with pytest.raises(ValueError, match="donate"):
        ops.pmf_to_quantized_cdf([1.0 / 70000] * 70000, 16)
    for bad in ([-0.1, 0.5], [float("inf"), 0.5], [float("nan"), 0.5]):    # ops.cpp:44-52 -> ValueError on both sides
        with pytest.raises(ValueError):
            cxx.pmf_to_quantized_cdf(bad, 16)
        with pytest.raises(ValueError):
            ops.pmf_to_quantized_cdf(bad, 16)


def test_entropy_model_tables_equal_reference_binary_row_by_row():
    """GaussianConditional.update() / EntropyBottleneck.update() build their int32 tables through mmcodec's CDF construction: every
    row must equal what the reference's _CXX produces from the same pmf (entropy_models.py:206-214, _pmf_to_cdf)."""
    _, cxx = _ref()
    gc, _eb = _tables()
    table = torch.tensor(mmcodec.models.get_scale_table())
    # restate GaussianConditional.update's pmf (entropy_models.py:655-689) with torch CPU ops, one row at a time
    from scipy.stats import norm
    multiplier = -norm.ppf(1e-9 / 2)
    pmf_center = torch.ceil(table * multiplier).int()
    for r in (0, 1, 17, 40, 63):
        L = int(2 * pmf_center[r] + 1)
        samples = torch.abs(torch.arange(L).int() - pmf_center[r]).float()
        s = table[r]
        const = -(2 ** -0.5)
        upper = 0.5 * torch.erfc(const * (0.5 - samples) / s)
        lower = 0.5 * torch.erfc(const * (-0.5 - samples) / s)
        pmf = upper - lower
        tail = 2 * lower[:1]
        want = cxx.pmf_to_quantized_cdf(pmf.tolist() + tail.tolist(), 16)
        assert int(gc._cdf_length[r]) == L + 2
        assert gc._quantized_cdf[r, :L + 2].tolist() == want


# ---- on the GPU box: the DEVICE CDF construction and the full compress() path against the reference binaries ------------------
@pytest.mark.gpu
def test_device_cdf_construction_equals_reference_binary():
    _, cxx = _ref()
    dev = torch.device("cuda", 0)
    rs = np.random.RandomState(9)
    rows = []
    for n in (2, 5, 33, 257, 1200, 3133):
        p = rs.dirichlet(np.full(n, 0.03)).astype(np.float32)
        p[rs.rand(n) < 0.5] = 0.0                       # zeros: these symbols get a stolen count
        p[rs.randint(n)] = max(float(p.max()), 0.4)
        rows.append(p)
    max_len = max(len(r) for r in rows) - 1
    pmf = np.zeros((len(rows), max_len), np.float32)
    tail = np.zeros(len(rows), np.float32)
    lens = np.zeros(len(rows), np.int32)
    for i, r in enumerate(rows):
        pmf[i, : len(r) - 1], tail[i], lens[i] = r[:-1], r[-1], len(r) - 1
    got = ops.pmf_to_quantized_cdf_device(torch.from_numpy(pmf).to(dev), torch.from_numpy(tail).to(dev), torch.from_numpy(lens).to(dev),
                                          max_len, 16).cpu().numpy()
    for i, r in enumerate(rows):
        want = np.array(cxx.pmf_to_quantized_cdf(r.tolist(), 16), dtype=np.int64)
        assert np.array_equal(got[i, : len(want)], want), i


@pytest.mark.gpu
def test_compress_strings_equal_reference_coder_on_device_symbols():
    """models/google.py:393-404 end to end: the symbols / indexes the kernels produce, coded by the REFERENCE's RansEncoder with the
    tables update() built, give the bytes compress() returns; the reference's decoder restores the symbols from them."""
    ans, _ = _ref()
    from weights import make_image, make_state_dict
    dev = torch.device("cuda", 0)
    sd = {k: torch.from_numpy(v) for k, v in make_state_dict("mean-scale", 192, 320, seed=0).items()}
    net = mmcodec.MeanScaleHyperprior(192, 320).eval()
    net.update()
    net.load_state_dict({**net.state_dict(), **sd})
    net.update(force=True)
    net = net.to(dev)
    x = torch.from_numpy(make_image(2, 128, 192)).to(dev)
    with torch.no_grad():
        out = net.compress(x)
        si = net.symbols_and_indexes(x)
    enc, dec = ans.RansEncoder(), ans.RansDecoder()
    for name, em, strings in (("y", net.gaussian_conditional, out["strings"][0]), ("z", net.entropy_bottleneck, out["strings"][1])):
        cdf = em._quantized_cdf.cpu().tolist()
        lens = em._cdf_length.reshape(-1).int().cpu().tolist()
        offs = em._offset.reshape(-1).int().cpu().tolist()
        sym, idx = si[f"{name}_symbols"].cpu(), si[f"{name}_indexes"].cpu()
        for b in range(x.shape[0]):
            s_b, i_b = sym[b].reshape(-1).tolist(), idx[b].reshape(-1).tolist()
            assert enc.encode_with_indexes(s_b, i_b, cdf, lens, offs) == strings[b], (name, b)
            assert dec.decode_with_indexes(strings[b], i_b, cdf, lens, offs) == s_b

"""CPU-only tests of the host side: the C-ABI library loads and exports every declared symbol, the
module classes keep the reference's state_dict contract and error behaviour, update() builds the
reference's CDF tables, and nothing silently falls back to the CPU.  No kernel is launched."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

# (not `from conftest import ...`: spawned gloo workers inherit sys.path, and a test that imported the reference put ITS
#  conftest.py in front of ours)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "165-learning-based-multi-modality-image-and-video-compression_b200")
for _p in (ROOT, PKG_DIR, os.path.join(ROOT, "tests", "golden")):
    if _p not in __import__("sys").path:
        __import__("sys").path.insert(0, _p)

import mmcodec
from mmcodec import _lib
from weights import _entropy_bottleneck, make_state_dict


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mmcodec.h")).read()
    declared = set(re.findall(r"MMC_API[^;(]*?\b(mmc_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 25
    assert os.path.exists(_lib.LIB_PATH), "libmmcodec.so not built: run python __graft_entry__.py"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/mmcodec.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert _lib.lib().mmc_version() == 100


def test_library_links_no_torch():
    """The drop-in boundary is a plain C ABI: no torch / python symbols in the shared object."""
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "python" not in out, out


def test_no_cpu_fallback():
    net = mmcodec.ScaleHyperprior(128, 192).eval()
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        net(torch.rand(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        mmcodec.GDN(8)(torch.rand(1, 8, 4, 4))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        mmcodec.EntropyBottleneck(8)(torch.rand(1, 8, 4, 4))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        mmcodec.conv(3, 8)(torch.rand(1, 3, 8, 8))


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(PKG_DIR, "mmcodec")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, re.M), f
    for f in os.listdir(os.path.join(PKG_DIR, "csrc")):
        assert "oracle" not in open(os.path.join(PKG_DIR, "csrc", f)).read(), f


@pytest.mark.parametrize("arch,cls,N,M", [("factorized", mmcodec.FactorizedPrior, 128, 192),
                                          ("hyperprior", mmcodec.ScaleHyperprior, 128, 192),
                                          ("mean_scale", mmcodec.MeanScaleHyperprior, 192, 320)])
def test_state_dict_contract(models_golden, arch, cls, N, M):
    """Same keys, shapes and dtypes as the reference module after update()."""
    ref = json.loads(str(models_golden[f"{arch}_state_dict"]))
    net = cls(N, M).eval()
    net.update(force=True)
    mine = {k: [list(v.shape), str(v.dtype)] for k, v in net.state_dict().items()}
    assert set(mine) == set(ref), set(mine) ^ set(ref)
    for k in ref:
        assert mine[k][1] == ref[k][1], k
        if not k.endswith("_quantized_cdf") and not k.endswith("_cdf_length") and not k.endswith("_offset"):
            assert mine[k][0] == ref[k][0], k
    # load_state_dict round trip incl. variable-size CDF buffers (models/utils.py:90-125)
    net2 = cls(N, M)
    net2.load_state_dict(net.state_dict())
    assert torch.equal(net2.entropy_bottleneck._quantized_cdf, net.entropy_bottleneck._quantized_cdf)
    net3 = cls.from_state_dict(net.state_dict())
    assert net3.N == N and net3.M == M


def test_compression_model_has_15_params():
    """tests/test_models.py:54-58"""
    assert len(list(mmcodec.CompressionModel(32).parameters())) == 15


def test_gc_update_tables_match_reference(kernels_golden):
    g = kernels_golden
    gc = mmcodec.GaussianConditional(None)
    assert gc.update_scale_table(mmcodec.get_scale_table())
    assert np.array_equal(gc.scale_table.numpy(), g["scale_table"])
    assert np.array_equal(gc._quantized_cdf.numpy(), g["gc_quantized_cdf"])
    assert np.array_equal(gc._cdf_length.numpy(), g["gc_cdf_length"])
    assert np.array_equal(gc._offset.numpy(), g["gc_offset"])
    assert not gc.update_scale_table(mmcodec.get_scale_table())  # already initialised
    t = mmcodec.get_scale_table()
    assert t[0] == 0.11 and t[-1] == 256 and len(t) == 64  # tests/test_models.py:242-258


def test_eb_update_tables_match_reference(kernels_golden):
    g = kernels_golden
    C = 8
    w = {}
    _entropy_bottleneck(np.random.RandomState(3), w, "eb", C)
    eb = mmcodec.EntropyBottleneck(C)
    sd = eb.state_dict()
    for k, v in w.items():
        sd[k[3:]].copy_(torch.from_numpy(v))
    assert eb.update()
    assert np.array_equal(eb._quantized_cdf.numpy(), g["eb_quantized_cdf"])
    assert np.array_equal(eb._cdf_length.numpy(), g["eb_cdf_length"])
    assert np.array_equal(eb._offset.numpy(), g["eb_offset"])
    assert not eb.update()
    assert eb.update(force=True)


def test_pmf_to_quantized_cdf_host(kernels_golden):
    g = kernels_golden
    from mmcodec.ops import pmf_to_quantized_cdf
    assert pmf_to_quantized_cdf([0.1, 0.2, 0.0, 0.0], 16) == [0, 21845, 65534, 65535, 65536]  # tests/test_ops.py:104-106
    for p, c, L in zip(g["cdf_pmfs"], g["cdf_cdfs"], g["cdf_lens"]):
        assert pmf_to_quantized_cdf(p[:L], 16) == c[:L + 1].tolist()
    for bad in ([-0.1, 0.5], [float("inf"), 0.5], [float("nan"), 0.5]):  # tests/test_ops.py:108-118
        with pytest.raises(ValueError):
            pmf_to_quantized_cdf(bad, 16)


def test_reference_error_behaviour():
    gc = mmcodec.GaussianConditional(None)
    with pytest.raises(ValueError, match="Invalid quantization mode"):
        gc.quantize(torch.rand(1, 2, 3, 3), mode="toto")  # tests/test_entropy_models.py:49-55
    with pytest.raises(ValueError):
        mmcodec.GaussianConditional(1)  # invalid scale_table type, tests/test_entropy_models.py:287-310
    with pytest.raises(ValueError):
        mmcodec.GaussianConditional([])
    with pytest.raises(ValueError):
        mmcodec.GaussianConditional([1, 0.5])
    with pytest.raises(ValueError):
        mmcodec.GaussianConditional([0, 1])
    with pytest.raises(ValueError):
        mmcodec.GaussianConditional(None, scale_bound=-0.1)
    with pytest.raises(ValueError, match="Invalid architecture"):
        mmcodec.build_model("nope", 1)
    with pytest.raises(ValueError, match="Invalid quality"):
        mmcodec.build_model("bmshj2018-factorized", 9)
    with pytest.raises(NotImplementedError):
        mmcodec.EntropyBottleneck(8, filters=(3, 3))


def test_zoo_configs_match_reference():
    """(N, M) per quality, compressai/zoo/image.py:189-220; configs 1-3 of BASELINE.json."""
    assert mmcodec.models.CFGS["bmshj2018-factorized"][1] == (128, 192)
    assert mmcodec.models.CFGS["bmshj2018-hyperprior"][4] == (128, 192)
    assert mmcodec.models.CFGS["bmshj2018-hyperprior"][6] == (192, 320)
    assert mmcodec.models.CFGS["mbt2018-mean"][6] == (192, 320)
    assert mmcodec.models.CFGS["mbt2018-mean"][4] == (128, 192)
    net = mmcodec.build_model("mbt2018-mean", 6)
    assert net.g_a[6].out_channels == 320 and net.h_s[4].out_channels == 640


def test_gdn_init_and_parse():
    from mmcodec.transforms import parse_layers
    net = mmcodec.ScaleHyperprior(128, 192)
    steps = parse_layers(list(net.g_a))
    assert [s.gdn is not None for s in steps] == [True, True, True, False]
    steps = parse_layers(list(net.h_s))
    assert [s.transposed for s in steps] == [True, True, False] and all(s.act == _lib.ACT_RELU for s in steps)
    g = mmcodec.GDN(16)
    # reparametrised init: beta = sqrt(1 + 2^-36), gamma = sqrt(0.1 I + 2^-36)  (layers/gdn.py:66-74)
    assert torch.allclose(g.beta ** 2 - 2.0 ** -36, torch.ones(16))
    assert torch.allclose(g.gamma ** 2 - 2.0 ** -36, 0.1 * torch.eye(16), atol=1e-7)
    with pytest.raises(NotImplementedError):
        parse_layers([torch.nn.Conv2d(3, 8, 7, padding=3)])


def test_synthetic_weights_are_stable():
    """The deterministic weight recipe must be bit-stable: goldens were generated with it."""
    sd = make_state_dict("hyperprior", 128, 192, seed=0)
    assert abs(float(sd["g_a.0.weight"][0, 0, 0, 0]) - 0.61108565) < 1e-6
    assert sd["g_s.6.weight"].shape == (128, 3, 5, 5)


@pytest.mark.parametrize("arch,cls,N,M", [("factorized", mmcodec.FactorizedPrior, 128, 192),
                                          ("hyperprior", mmcodec.ScaleHyperprior, 128, 192),
                                          ("mean_scale", mmcodec.MeanScaleHyperprior, 192, 320)])
def test_rans_streams_byte_identical_to_reference(models_golden, arch, cls, N, M):
    """The host rANS coder on the reference's own symbols / indexes (captured at its encode_with_indexes call) and the
    CDF tables built by our update(): streams byte-identical to the reference's (SURVEY.md 8d parity gate)."""
    from mmcodec import ops
    g = models_golden
    net = cls(N, M).eval()
    sd = {k: torch.from_numpy(v) for k, v in make_state_dict(arch.replace("_", "-"), N, M, seed=0).items()}
    net.update()
    net.load_state_dict({**net.state_dict(), **sd})
    net.update(force=True)
    ems = [("y", net.entropy_bottleneck)] if arch == "factorized" else [("y", net.gaussian_conditional), ("z", net.entropy_bottleneck)]
    for si, (name, em) in enumerate(ems):
        sym, idx = torch.from_numpy(g[f"{arch}_{name}_symbols"]), torch.from_numpy(g[f"{arch}_{name}_indexes"])
        strings = ops.rans_encode(sym, idx, em._quantized_cdf, em._cdf_length, em._offset)
        for b, s in enumerate(strings):
            assert s == g[f"{arch}_string_{si}_{b}"].tobytes(), (name, b)
        assert torch.equal(ops.rans_decode(strings, idx, em._quantized_cdf, em._cdf_length, em._offset), sym)


def test_rans_bypass_and_errors():
    from mmcodec import ops
    eb = mmcodec.EntropyBottleneck(4)
    eb.update()
    tabs = (eb._quantized_cdf, eb._cdf_length, eb._offset)
    idx = torch.arange(4, dtype=torch.int32).repeat(3, 25)                  # (3, 100)
    sym = torch.randint(-200000, 200000, (3, 100), dtype=torch.int32)      # far outside the tables -> bypass nibbles
    sym[0, :10] = torch.tensor([0, 1, -1, 10, -10, 11, -11, 12, 2 ** 30, -2 ** 30])
    strings = ops.rans_encode(sym, idx, *tabs)
    assert torch.equal(ops.rans_decode(strings, idx, *tabs), sym)
    assert ops.rans_encode(sym[:0], idx[:0], *tabs) == []
    with pytest.raises(ValueError):
        ops.rans_encode(sym, idx + 4, *tabs)                                # index outside the table
    with pytest.raises(ValueError):
        ops.rans_decode([s[:8] for s in strings], idx, *tabs)               # truncated stream


def test_rans_fast_path_exact_and_mixed_streams():
    """The encoder's multiply-by-reciprocal state update equals the reference's division for every frequency, and long streams
    that mix regular symbols with escape / bypass symbols (buffer growth path) round-trip."""
    from mmcodec import _lib, ops
    assert _lib.lib().mmc_rans_selftest() == 0
    gc = mmcodec.GaussianConditional(None)
    gc.update_scale_table(mmcodec.models.get_scale_table())
    tabs = (gc._quantized_cdf, gc._cdf_length, gc._offset)
    g = torch.Generator().manual_seed(7)
    n = 200000
    idx = torch.randint(0, 64, (2, n), generator=g, dtype=torch.int32)
    scale = torch.tensor(mmcodec.models.get_scale_table())[idx.long()]
    sym = torch.round(torch.randn(2, n, generator=g) * scale).to(torch.int32)
    sym[1, ::3] = torch.randint(-5000, 5000, (len(sym[1, ::3]),), generator=g, dtype=torch.int32)   # every third symbol escapes
    strings = ops.rans_encode(sym, idx, *tabs)
    assert torch.equal(ops.rans_decode(strings, idx, *tabs), sym)
    assert len(strings[1]) > len(strings[0])


def test_lane_container_oracle_round_trip_and_token_parity():
    """oracle/lane_rans.py (the CPU restatement of the device coder's container, csrc/rans_device.cu): round trips for ragged
    sizes / lane counts / escapes, carries exactly the symbols the reference-compatible host coder carries, and costs at most the
    lane headers more than the reference's stream."""
    from mmcodec import ops
    from oracle import lane_rans
    gc = mmcodec.GaussianConditional(None)
    gc.update_scale_table(mmcodec.models.get_scale_table())
    tabs = (gc._quantized_cdf, gc._cdf_length, gc._offset)
    cdfs, sizes, offs = (t.numpy() for t in tabs)
    g = torch.Generator().manual_seed(3)
    for n, lanes in ((0, 4), (1, 4), (5, 8), (100, 1), (257, 7), (3000, 4), (3000, None)):
        idx = torch.randint(0, 64, (1, n), generator=g, dtype=torch.int32)
        scale = torch.tensor(mmcodec.models.get_scale_table())[idx.long()]
        sym = torch.round(torch.randn(1, n, generator=g) * scale).to(torch.int32)
        if n >= 5:
            sym[0, ::4] = torch.randint(-70000, 70000, (len(sym[0, ::4]),), generator=g, dtype=torch.int32)     # escapes
            sym[0, 1] = 2 ** 30
            sym[0, 2] = -2 ** 30
        stream = lane_rans.encode(sym[0].tolist(), idx[0].tolist(), cdfs, sizes, offs, lanes)
        S = lanes if lanes is not None else lane_rans.lanes_default(n)
        assert stream[:4] == b"MMCL" and len(stream) % 4 == 0
        assert lane_rans.decode(stream, idx[0].tolist(), cdfs, sizes, offs) == sym[0].tolist()
        host = ops.rans_encode(sym, idx, *tabs)
        assert torch.equal(ops.rans_decode(host, idx, *tabs), sym)            # same symbols through the reference-compatible stream
        # rate: the reference's stream + the lane headers (8 bytes per lane + 16) + each lane's own flush (its 32-bit end state carries
        # ~2 bytes less than it occupies) and word rounding: <= 6 bytes per lane
        assert len(stream) <= len(host[0]) + 16 + 14 * S + 4, (n, lanes, len(stream), len(host[0]))
    assert [lane_rans.lanes_default(v) for v in (10, 18_432, 97_920, 295_000, 1_570_000, 2_611_200)] == [4, 8, 32, 64, 256, 512]
    with pytest.raises(ValueError):
        lane_rans.decode(b"XXXX" + stream[4:], idx[0].tolist(), cdfs, sizes, offs)


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    lo, hi = bench.shard_range(64, rank, world)
    t = torch.tensor([float(hi - lo)])
    dist.all_reduce(t)
    mx = torch.tensor([1.0 + rank])
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    q.put((rank, lo, hi, t.item(), mx.item()))
    dist.destroy_process_group()


def test_batch_sharding_world_size_2_gloo():
    """N>1 path of bench.py: contiguous batch shards, no data-path collective, MAX-over-ranks timing."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert res[0][1:3] == (0, 32) and res[1][1:3] == (32, 64)
    assert res[0][3] == 64.0 and res[0][4] == 2.0


def _reducer_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch.nn as nn
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                    # identical replicas
    net = nn.Sequential(nn.Linear(6, 5), nn.Tanh(), nn.Linear(5, 4), nn.Linear(4, 3))
    unused = nn.Parameter(torch.zeros(7))                   # a parameter that never receives a gradient
    params = list(net.parameters()) + [unused]
    red = mmcodec.GradBucketReducer(params, bucket_bytes=64)   # tiny buckets: several collectives, one left partially filled
    x = torch.arange(12, dtype=torch.float32).reshape(2, 6) * (rank + 1)
    net(x).pow(2).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    red.finish()
    first = [p.grad.tolist() for p in net.parameters()]
    # second mode (GraphedTrainStep with several ranks): hooks disabled during backward, one explicit reduction afterwards
    red.enabled = False
    net.zero_grad(set_to_none=True)
    net(x + 1.0).pow(2).sum().backward()
    local2 = [p.grad.clone() for p in net.parameters()]
    red.reduce_now()
    second = [p.grad.tolist() for p in net.parameters()]
    unused_none = unused.grad is None
    # third mode: flat buckets -- .grad are views into per-bucket buffers, collectives in place, zero_() instead of zero_grad
    red.remove()
    net.zero_grad(set_to_none=True)
    red3 = mmcodec.GradBucketReducer(params, bucket_bytes=64, flat=True)
    views_ok = all(p.grad is not None and p.grad.untyped_storage().data_ptr() == red3.flat[red3.bucket_of[i]].untyped_storage().data_ptr()
                   for i, p in enumerate(red3.params))
    for _ in range(2):                                       # second round: zero_() really clears the accumulated gradients
        red3.zero_()
        net(x + 2.0).pow(2).sum().backward()
        for w_, _, _ in red3._work.values():                 # local copy only after the in-flight bucket collectives (they
            w_.wait()                                        # are in place) -- recompute the local gradient instead
        red3.finish()
    red3.enabled = False                                     # the rank-local gradient of the same batch, for the expected average
    red3.zero_()
    net(x + 2.0).pow(2).sum().backward()
    local3 = [p.grad.clone() for p in net.parameters()]
    red3.enabled = True
    red3.zero_()
    net(x + 2.0).pow(2).sum().backward()
    red3.finish()
    third = [p.grad.tolist() for p in net.parameters()]
    q.put((rank, [g.tolist() for g in local], first, unused_none, len(red.buckets), [g.tolist() for g in local2], second,
           views_ok and bool((unused.grad == 0).all()), [g.tolist() for g in local3], third))
    dist.destroy_process_group()


def test_grad_bucket_reducer_world_size_2_gloo():
    """Data-parallel exchange step of the training path: bucketed all-reduce averages the gradients of both ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    ps = [ctx.Process(target=_reducer_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    (_, l0, a0, u0, nb, m0, b0, v0, f0, t0), (_, l1, a1, u1, _, m1, b1, v1, f1, t1) = res
    assert u0 and u1 and nb >= 3 and v0 and v1
    for loc0, loc1, red0, red1 in ((l0, l1, a0, a1), (m0, m1, b0, b1), (f0, f1, t0, t1)):   # hook-driven buckets, reduce_now(), flat buckets
        for g0, g1, r0, r1 in zip(loc0, loc1, red0, red1):
            want = (torch.tensor(g0) + torch.tensor(g1)) / 2
            assert torch.allclose(torch.tensor(r0), want, rtol=1e-6, atol=1e-6) and torch.allclose(torch.tensor(r1), want, rtol=1e-6, atol=1e-6)


def test_training_glue_on_cpu():
    """RateDistortionLoss and configure_optimizers follow examples/train.py:59-82,111-142 (host logic, no kernels)."""
    out = {"x_hat": torch.full((2, 1, 4, 4), 0.5), "likelihoods": {"y": torch.full((2, 3, 2, 2), 0.25), "z": torch.full((2, 3, 1, 1), 0.5)}}
    target = torch.zeros(2, 1, 4, 4)
    res = mmcodec.RateDistortionLoss(2)(out, target)
    assert abs(float(res["bpp_loss"]) - (24 * 2 + 6 * 1) / 32) < 1e-6 and abs(float(res["mse_loss"]) - 0.25) < 1e-7
    assert abs(float(res["loss"]) - (1024 * 0.25 + 54 / 32)) < 1e-4
    net = mmcodec.FactorizedPrior(8, 8)
    opt, aux = mmcodec.configure_optimizers(net)
    n_main = sum(len(g["params"]) for g in opt.param_groups)
    n_aux = sum(len(g["params"]) for g in aux.param_groups)
    assert n_aux == 1 and n_main + n_aux == len(list(net.parameters()))


def test_fusion_model_mirrors_state_dict_contract_on_cpu():
    """Host-side mirrors of the fork's own models are constructible without a GPU and expose exactly the reference's state_dict
    (keys; shapes except the data-dependent CDF buffers) -- compared with the lists the golden generators recorded from the
    reference itself: Master_compresser (master.py:837-902), Guided_compresser (:1215-1268), mbt2018 (google.py:421-497)."""
    from oracle import torch_port as tp
    golden = os.path.join(ROOT, "tests", "golden")
    data_dependent = ("_offset", "_quantized_cdf", "_cdf_length", "scale_table")
    cases = (("models_master.npz", lambda: mmcodec.Master_compresser(width=64, height=128, channel=3)),
             ("models_guided.npz", lambda: mmcodec.Guided_compresser(channel=1)),
             ("models_mbt2018.npz", lambda: mmcodec.build_model("mbt2018", 3)))
    for fname, make in cases:
        ref = {k: tuple(v[0]) for k, v in json.loads(str(np.load(os.path.join(golden, fname))["state_dict"])).items()}
        net = make()
        mine = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert set(mine) == set(ref), (fname, sorted(set(mine) ^ set(ref))[:5])
        assert all(mine[k] == ref[k] for k in ref if not k.endswith(data_dependent)), fname
    # the attention buffers the reference registers (master.py:512-522, 625-643) equal the restatement's closed forms
    net = mmcodec.Master_compresser(width=64, height=128, channel=3)
    blk = net.decoder.sp_aligner3.blocks[1]                      # tokens 16 x 32, window 4, shift 2
    assert blk.shift_size == 2 and torch.equal(blk.attn.relative_position_index, tp.relative_position_index(4))
    assert torch.equal(blk.attn_mask, tp.shift_attention_mask(16, 32, 4, 2))
    assert net.decoder.sp_aligner1.blocks[1].shift_size == 0      # 4 x 8 tokens: min(resolution) <= window -> no shift (master.py:601-603)
    assert "attn_mask" not in dict(net.decoder.sp_aligner1.blocks[1].named_buffers())
    one = mmcodec.Master_compresser(width=64, height=64, channel=1)   # 1-channel master: extra downsample convs, swapped strides
    assert {"decoder.downsample1.weight", "decoder.downsample3.bias"} <= set(one.state_dict())
    assert one.fencoder1.conv1.stride == (1, 1) and one.fencoder2.conv1.stride == (2, 2) and one.fdecoder.deconv1.stride == (1, 1)
    # and there is no CPU execution path
    with pytest.raises((RuntimeError, mmcodec.MmcodecError)):
        net.eval()(torch.zeros(1, 3, 128, 256), torch.zeros(1, 1, 64, 128), {k: torch.zeros(1, 192, 8 * 2 ** i, 16 * 2 ** i) for i, k in enumerate(("gs1", "gs2", "gs3"))})


def test_bench_reference_arm_of_the_pair_workloads():
    """bench.py --impl reference for the RGB-T pair workload: runs the CPU port (no GPU, no /root/reference), prints ONE JSON line
    with the contract's keys; the training workloads report `unavailable` and exit 0."""
    import subprocess
    import sys
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "master-forward", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "img/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "master-train"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "unavailable" in json.loads(r.stdout.strip().splitlines()[-1])


def test_torchscript_stand_ins_and_op_schemas():
    """torch.jit.script of the layer mirrors compiles (compressai tests/test_scripting.py:37-59) into one call of a registered
    torch.library op; the scripted module keeps the reference's state_dict keys and shares the Parameters; no CPU kernel exists."""
    import torch
    import mmcodec
    g = mmcodec.GDN(128)
    m = torch.jit.script(g)
    assert "ops.mmcodec.gdn" in m.code
    assert list(m.state_dict().keys()) == ["beta", "gamma", "beta_reparam.pedestal", "beta_reparam.lower_bound.bound",
                                           "gamma_reparam.pedestal", "gamma_reparam.lower_bound.bound"]
    assert m.state_dict()["beta"].data_ptr() == g.beta.data_ptr()
    assert "ops.mmcodec.lower_bound" in torch.jit.script(mmcodec.LowerBound(0.11)).code
    schema = str(torch.ops.mmcodec.gdn.default._schema)
    assert schema == "mmcodec::gdn(Tensor x, Tensor beta, Tensor gamma, float beta_bound, float gamma_bound, float pedestal, bool inverse) -> Tensor"
    # fake (meta) implementation: shape / dtype propagation without a device
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        y = torch.ops.mmcodec.gdn(torch.empty(2, 128, 4, 4), torch.empty(128), torch.empty(128, 128), 1e-3, 4e-6, 1.5e-11, False)
        assert tuple(y.shape) == (2, 128, 4, 4) and y.dtype == torch.float32
    with pytest.raises((RuntimeError, NotImplementedError)):
        m(torch.rand(1, 128, 1, 1))                      # CPU tensor: loud failure, there is no CPU path

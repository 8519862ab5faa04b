#!/usr/bin/env python
"""bench.py -- headline benchmark: img/s of the codec forward + likelihoods on B200.

    python bench.py --gpus N --steps K --warmup W            (N=1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W   (N>1, one rank per GPU)
    python bench.py --impl reference ...                     (reference CPU path, see below)

Workload (BASELINE.json configs[1]): bmshj2018-hyperprior quality 4 (N=128, M=192), eval forward
(g_a, h_a, EntropyBottleneck, h_s, GaussianConditional, g_s) on a batch of 64 synthetic 768x512 RGB
images per GPU, random-init weights.  A "step" is one forward over one batch.  Images are
independent, so N GPUs run N replicas on disjoint batch shards with no data-path collective
(weak scaling: 64 images per GPU); the only collectives are the timing barrier and a MAX over
ranks of the device time.

Prints ONE JSON line (rank 0).  `value` = images/s with the batch resident in HBM; `e2e` = the
same through the public module API with pinned host buffers, H2D and D2H inside the timed region.
`--impl reference` times the reference's CPU execution path (oracle/torch_port.py, the restated
stock-torch-op sequence the reference runs; the reference itself is Python and cannot travel to
the GPU box) on the host cores.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "165-learning-based-multi-modality-image-and-video-compression_b200")
for p in (ROOT, PKG_DIR, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "img/s (768x512 codec fwd+likelihoods)"
UNIT = "img/s"
# name -> (zoo architecture, quality, H, W, default batch per GPU, call, description).  The default (and the only one the
# driver runs) is configs[1]; the others are BASELINE.json's parity configs, measurable with --workload for the record.
WORKLOADS = {
    "hyperprior": ("bmshj2018-hyperprior", 4, 512, 768, 64, "forward",
                   "bmshj2018-hyperprior q4 eval forward, batch 64 x 768x512 RGB (BASELINE.json configs[1])"),
    "factorized": ("bmshj2018-factorized", 1, 512, 768, 64, "forward",
                   "bmshj2018-factorized q1 eval forward, 768x512 RGB (BASELINE.json configs[0], batched)"),
    "mbt-mean-symbols": ("mbt2018-mean", 6, 1088, 1920, 8, "symbols",
                         "mbt2018-mean q6 compress() symbol/index path, 1920x1080 padded to 1088 (BASELINE.json configs[2])"),
    "mbt-mean-compress": ("mbt2018-mean", 6, 1088, 1920, 8, "compress",
                          "mbt2018-mean q6 compress() incl. host rANS coding, 1920x1080 padded to 1088"),
}
# workloads with their own input structure (a GOP of frames / an RGB + depth pair): measured by bench_other()
OTHER_WORKLOADS = {
    "ssf2020": ("ssf2020 video codec eval forward, one GOP of 8 frames 1920x1152 per GPU, frames in order "
                "(BASELINE.json configs[4]); unit = frames"),
    "mm-train": ("RGB + depth two-branch codec training step (frozen RGB guide forward, depth branch forward + backward, "
                 "bucketed NCCL gradient all-reduce overlapped with backward, grad clip, Adam + aux Adam), 768x512 pairs "
                 "(BASELINE.json configs[3]); unit = pairs"),
    "mm-forward": ("RGB + depth two-branch codec (JointAutoregressiveHierarchicalPriors_R/_D) eval forward, 768x512 pairs "
                   "(forward half of BASELINE.json configs[3]); unit = pairs"),
    "master-train": ("RGB-T reproduction training step (frozen Guided_compresser forward on the 1x256x384 guide, Master_compresser forward + "
                     "backward on the 3x512x768 master, bucketed NCCL gradient all-reduce, grad clip, Adam + aux Adam; examples/train.py:208-233); "
                     "unit = pairs"),
    "master-forward": ("RGB-T reproduction (Guided_compresser on the 1x256x384 guide + Master_compresser on the 3x512x768 master: feature "
                       "codecs, channel aligner, three window cross-attention stages) eval forward; unit = pairs"),
}
ARCH, QUALITY, H, W = "bmshj2018-hyperprior", 4, 512, 768
WORKLOAD = WORKLOADS["hyperprior"][6]


# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the default workload's largest kernels, from the
# committed ncu capture profiles/r02_ncu_conv_phase_inner.csv (batch 64 x 768x512, round-end default path).  With the phase-inner
# tile order the large transposed convolutions read their input from HBM once (g_s.4: 0.404 GB read for a 0.403 GB input,
# 1.553 GB written: 1.96 GB against 1.97 GB algorithmic; before: 1.61 GB read, 3.18 GB in total = 1.6x).
NCU_DRAM_BYTES = {"g_s.4|tc": 0.403608e9 + 1.553492e9, "g_a.2|tc": 1.612e9 + 0.384e9, "g_a.0|tc": 0.408e9 + 1.557e9,
                  "g_s.6|tc": 1.622228e9 + 0.287555e9, "g_s.2|tc": 0.101597e9 + 0.345687e9}
NCU_SOURCE = "profiles/r02_ncu_conv_phase_inner.csv (ncu dram__bytes_read/write.sum per launch, profiles/_fwd_once.py; round-end default path)"


def reference_coder_seconds(entropy_model, symbols, indexes):
    """One image coded by the REFERENCE'S OWN rANS coder (oracle/_ref/ans*.so: compressai/cpp_exts/rans/rans_interface.cpp compiled by
    `make -C oracle ref`), marshalled the way the reference does it (entropy_models.py:260-269: `.tolist()` per image, then
    `encode_with_indexes`).  Returns (seconds, bytes) or raises if the binary was never built.  cpu_baseline leg only."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    ans = sys.modules.get("compressai.ans")
    if ans is None:
        if ref_dir not in sys.path:
            sys.path.insert(0, ref_dir)
        import ans
    cdf = entropy_model._quantized_cdf.cpu().tolist()
    lens = entropy_model._cdf_length.reshape(-1).int().cpu().tolist()
    offs = entropy_model._offset.reshape(-1).int().cpu().tolist()
    sym, idx = symbols.cpu(), indexes.cpu()
    enc = ans.RansEncoder()
    t0 = time.perf_counter()
    out = enc.encode_with_indexes(sym.reshape(-1).int().tolist(), idx.reshape(-1).int().tolist(), cdf, lens, offs)
    return time.perf_counter() - t0, out


def shard_range(total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `total` independent units for `rank` (no collective needed)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bench_config(batch: int, world: int):
    """`config` of the JSON line; identical for both arms (the reference arm runs bounded samples of it)."""
    mb = batch * 3 * H * W * 4 / 1e6
    return {"workload": WORKLOAD, "batch_per_gpu": batch, "image": f"{W}x{H}", "weights": "random init",
            "parallelism": f"batch-sharded replicas x{world}, no data-path collective",
            "l2": f"inputs larger than L2 ({mb:.0f} MB fp32 batch and GB-sized first activations), no flush needed"}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None, "reasons": reasons}


def cpu_reference_throughput(batch: int, steps: int, warmup: int, threads: int = 0):
    """Reference CPU path (oracle/torch_port.py) on `threads` host threads (0 = all cores); returns (img/s, ms_per_step, cores)."""
    import torch
    from oracle import torch_port as tp
    import mmcodec
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    net = mmcodec.build_model(ARCH, QUALITY).eval()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.rand(batch, 3, H, W, generator=torch.Generator().manual_seed(1234))
    if ARCH == "mbt2018-mean":
        table = tp.get_scale_table()
        fn = lambda sd_, x_: tp.mean_scale_compress_symbols(sd_, x_, table)
    else:
        fn = tp.FORWARD["hyperprior" if ARCH == "bmshj2018-hyperprior" else "factorized"]
    with torch.no_grad():
        for _ in range(warmup):
            fn(sd, x)
        t0 = time.perf_counter()
        for _ in range(steps):
            out = fn(sd, x)
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, cores, (tp.bpp(out, batch * H * W) if "likelihoods" in out else None)


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `index` BEFORE the pinned host buffers are allocated, so that
    first-touch places them on the GPU's own NUMA node (round-1 SCALE: with every rank's pinned memory on one node the fp32
    host->device stream of 8 ranks collapsed to 186 GB/s aggregate).  Best effort: containers may restrict the CPU set."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        n = (os.cpu_count() or 64 + 63) // 64 + 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
            return {"bound_cpus": len(use), "allowed_cpus": len(allowed)}
        return {"bound_cpus": 0, "allowed_cpus": len(allowed), "note": "GPU-local CPUs not in this container's CPU set"}
    except Exception as e:
        return {"bound_cpus": 0, "note": repr(e)[:120]}


def copy_ceilings(x_host, dev, barrier, reps: int = 3):
    """Pure-copy ceilings for the e2e numbers, measured live with every rank copying at the same time: the step's input batch
    pinned host -> device, a result-sized buffer device -> pinned host, and both directions at once (GB/s, this rank)."""
    import torch
    out_bytes = int(x_host.numel() * 4 * 1.26)          # x_hat + likelihoods of the headline model: 1.26x the input bytes
    d_in = torch.empty_like(x_host, device=dev)
    d_out = torch.empty(out_bytes // 4, dtype=torch.float32, device=dev)
    h_out = torch.empty(out_bytes // 4, dtype=torch.float32).pin_memory()
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    res = {}

    def timed(fn):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / reps

    ms = timed(lambda: d_in.copy_(x_host, non_blocking=True))
    res["h2d_gbs"] = x_host.numel() * x_host.element_size() / ms / 1e6
    ms = timed(lambda: h_out.copy_(d_out, non_blocking=True))
    res["d2h_gbs"] = out_bytes / ms / 1e6

    def both():
        cur = torch.cuda.current_stream(dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        s1.wait_event(ev); s2.wait_event(ev)
        with torch.cuda.stream(s1):
            d_in.copy_(x_host, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
        cur.wait_stream(s1); cur.wait_stream(s2)
    ms = timed(both)
    res["bidir_h2d_gbs"] = x_host.numel() * x_host.element_size() / ms / 1e6
    res["bidir_d2h_gbs"] = out_bytes / ms / 1e6
    return res


def train_step_record(dev, world, dist, steps: int = 5):
    """Sub-record of the default line so that the driver's 1..8-GPU scaling run also exercises the ONE exchange step of the path
    (BASELINE.json configs[3]): RGB + depth two-branch codec, 4 pairs of 768x512 per GPU, frozen guide forward, depth-branch
    forward + backward, bucketed NCCL all-reduce of the fp32 gradients (in place on flat buckets), clip, Adam x2
    (mmcodec.GraphedTrainStep).  Every rank runs it; device-timed, MAX over ranks."""
    import torch
    import mmcodec
    torch.manual_seed(0)
    net_r = mmcodec.JointAutoregressiveHierarchicalPriors_R(192, 192).eval()
    net_d = mmcodec.JointAutoregressiveHierarchicalPriors_D(192, 192)
    for n in (net_r, net_d):
        n.update()
        n.to(dev)
    units = 4
    gen = torch.Generator().manual_seed(99 + int(os.environ.get("RANK", "0")))
    rgb = torch.rand(units, 3, 512, 768, generator=gen).to(dev)
    depth = torch.rand(units, 1, 512, 768, generator=gen).to(dev)
    step = mmcodec.GraphedTrainStep(net_d, net_r, warmup=2, quality=3)
    for _ in range(4):                      # two eager optimisation steps, the capture, one replay
        out = step(depth, rgb)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step(depth, rgb)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0]) / steps
    red = step.step.reducer
    grad_bytes = sum(b.numel() * 4 for b in (red.flat or []))
    loss = float(out["loss"])
    del step, net_r, net_d
    torch.cuda.empty_cache()
    return {"metric": "pairs/s (RGB + depth 768x512 training step)", "value": units * world / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms,
            "steps": steps, "pairs_per_gpu": units, "scaling": "weak", "loss": loss,
            "all_reduce_bytes_per_step": grad_bytes if world > 1 else 0, "buckets": len(red.buckets),
            "exchange": (f"NCCL all-reduce (AVG) of {len(red.buckets)} flat fp32 gradient buckets over {world} ranks, between the forward+backward "
                         f"graph and the clip+Adam graph") if world > 1 else "none (one rank)",
            "api": "mmcodec.GraphedTrainStep(net_d, net_r)(depth, rgb)"}


def init_nccl_quietly(dist, dev):
    """NCCL prints its version banner on stdout when the communicator is created; keep stdout for the ONE JSON line."""
    import torch
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        warm = torch.zeros(1, device=dev)
        dist.all_reduce(warm)          # forces communicator creation now
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def reference_pair_forward(args, world):
    """Reference arm of the pair workloads: the reference's own CPU execution path (stock torch ops, restated in oracle/torch_port.py
    and pinned to the reference's outputs by tests/golden) on all host threads, ONE pair per step (a bounded sample of the workload)."""
    import torch
    import mmcodec
    from oracle import torch_port as tp
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(1234)
    as_sd = lambda net: {k: v.detach().float() for k, v in net.state_dict().items()}
    if args.workload == "mm-forward":
        sd_r, sd_d = as_sd(mmcodec.JointAutoregressiveHierarchicalPriors_R(192, 192)), as_sd(mmcodec.JointAutoregressiveHierarchicalPriors_D(192, 192))
        x, d = torch.rand(1, 3, 512, 768, generator=gen), torch.rand(1, 1, 512, 768, generator=gen)
        step = lambda: tp.mm_d_forward(sd_d, d, tp.mm_r_forward(sd_r, x)["hidden"])
    else:
        sd_g, sd_m = as_sd(mmcodec.Guided_compresser(channel=1)), as_sd(mmcodec.Master_compresser(width=256, height=384, channel=3))
        x, t = torch.rand(1, 3, 512, 768, generator=gen), torch.rand(1, 1, 256, 384, generator=gen)

        def step():
            og = tp.mm_r_forward(sd_g, t)
            return tp.master_forward(sd_m, x, og["x_hat"], og["hidden"])
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 1))):
            step()
        t0 = time.perf_counter()
        for _ in range(max(1, args.steps)):
            step()
        dt = time.perf_counter() - t0
    steps = max(1, args.steps)
    v = steps / dt
    return {"impl": "reference", "metric": f"pairs/s ({args.workload})", "value": v, "unit": "img/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": OTHER_WORKLOADS[args.workload], "units_per_gpu": 1, "weights": "random init", "parallelism": f"one process, {cores} host threads"},
            "cpu_baseline": {"value": v, "unit": "img/s", "cores": cores, "kind": "port", "sample": f"{steps} steps of ONE pair (same models / resolution), torch CPU ops on {cores} threads"},
            "e2e": {"value": v, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def bench_other(args):
    """For-the-record lines of the GOP / pair workloads (same timing rules; not the driver's headline run)."""
    import torch
    import torch.distributed as dist
    import mmcodec
    from mmcodec import ops
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            if args.workload in ("mm-forward", "master-forward"):
                print(json.dumps(reference_pair_forward(args, world)))
            else:
                print(json.dumps({"impl": "reference", "unavailable": f"reference arm is implemented for the forward workloads only, not {args.workload}"}))
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_nccl_quietly(dist, dev)
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(1234 + rank)
    if args.workload == "ssf2020":
        net = mmcodec.ScaleSpaceFlow().eval()
        net.update()
        net = net.to(dev)
        units = args.batch or 8
        host = [torch.rand(1, 3, 1152, 1920, generator=gen).pin_memory() for _ in range(units)]
        step = lambda xs: net(xs)["x_hat"][-1]
        e2e_out = lambda o: list(o["x_hat"])
        run = lambda xs: net(xs)
    elif args.workload == "mm-train":
        net_r = mmcodec.JointAutoregressiveHierarchicalPriors_R(192, 192).eval()
        net_d = mmcodec.JointAutoregressiveHierarchicalPriors_D(192, 192)
        for n in (net_r, net_d):
            n.update()
            n.to(dev)
        units = args.batch or 4
        host = [torch.rand(units, 3, 512, 768, generator=gen).pin_memory(), torch.rand(units, 1, 512, 768, generator=gen).pin_memory()]
        trainer = mmcodec.TrainStep(net_d, net_r, quality=3)

        def run(xs):
            with torch.enable_grad():
                return trainer(xs[1], xs[0])
        e2e_out = lambda o: [o["loss"]]
    elif args.workload == "master-train":
        guide = mmcodec.Guided_compresser(channel=1).eval()
        master = mmcodec.Master_compresser(width=256, height=384, channel=3)
        for n in (guide, master):
            n.update()
            n.to(dev)
        units = args.batch or 4
        host = [torch.rand(units, 3, 512, 768, generator=gen).pin_memory(), torch.rand(units, 1, 256, 384, generator=gen).pin_memory()]
        trainer = mmcodec.TrainStep(master, guide, quality=3)

        def run(xs):
            with torch.enable_grad():
                return trainer(xs[0], xs[1])
        e2e_out = lambda o: [o["loss"]]
    elif args.workload == "master-forward":
        guide = mmcodec.Guided_compresser(channel=1).eval()
        master = mmcodec.Master_compresser(width=256, height=384, channel=3).eval()
        for n in (guide, master):
            n.update()
            n.to(dev)
        units = args.batch or 8
        host = [torch.rand(units, 3, 512, 768, generator=gen).pin_memory(), torch.rand(units, 1, 256, 384, generator=gen).pin_memory()]

        def run(xs):
            og = guide(xs[1])
            return {"g": og, "m": master(xs[0], og["x_hat"], og["hidden"])}
        e2e_out = lambda o: [o["g"]["x_hat"], o["m"]["x_hat"]]
    else:
        net_r = mmcodec.JointAutoregressiveHierarchicalPriors_R(192, 192).eval()
        net_d = mmcodec.JointAutoregressiveHierarchicalPriors_D(192, 192).eval()
        for n in (net_r, net_d):
            n.update()
            n.to(dev)
        units = args.batch or 8
        host = [torch.rand(units, 3, 512, 768, generator=gen).pin_memory(), torch.rand(units, 1, 512, 768, generator=gen).pin_memory()]

        def run(xs):
            o_r = net_r(xs[0])
            return {"r": o_r, "d": net_d(xs[1], o_r["hidden"])}
        e2e_out = lambda o: [o["r"]["x_hat"], o["d"]["x_hat"]]
    xs = [t.to(dev) for t in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            run(xs)
        barrier()
        ops.reset_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            run(xs)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = ops.launch_count()
        # results land in pinned host buffers (same memory layout as the device tensors), as a serving loop would keep them
        pinned_out = [torch.empty_strided(t.shape, t.stride(), dtype=t.dtype, pin_memory=True) for t in e2e_out(run(xs))]
        def e2e_step():
            outs = e2e_out(run([t.to(dev, non_blocking=True) for t in host]))
            for dst, src in zip(pinned_out, outs):
                dst.copy_(src, non_blocking=True)
        for _ in range(2):          # untimed: lets the caching allocator settle on the e2e allocation pattern (fresh input tensors per step)
            e2e_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        torch.cuda.synchronize()
        ms_e2e = (time.perf_counter() - t0) * 1e3
        ops.start_profile()
        run(xs)
        prof = ops.stop_profile(with_work="total")
        graph_error = None
        if args.workload in ("mm-train", "master-train"):
            # the whole optimisation step (forward, backward, bucket all-reduces, clip, Adam x2) replayed as ONE CUDA graph
            ms_graph = ms
            try:
                trainer.reducer.remove()
                g_net, g_guide = trainer.net, trainer.guide
                graphed = mmcodec.GraphedTrainStep(g_net, g_guide, warmup=2, quality=3)
                run_g = (lambda xs: graphed(xs[1], xs[0])) if args.workload == "mm-train" else (lambda xs: graphed(xs[0], xs[1]))
                for _ in range(4):
                    run_g(xs)
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                for _ in range(args.steps):
                    run_g(xs)
                g1.record()
                barrier()
                ms_graph = g0.elapsed_time(g1)
            except Exception as e:   # keep the eager number, say why
                graph_error = f"{type(e).__name__}: {e}"[:300]
        else:
            # launch-bound workloads: the same forward replayed as ONE CUDA graph (mmcodec.GraphedForward)
            graphed = mmcodec.GraphedForward(run, xs)
            graphed(xs)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(args.steps):
                graphed(xs)
            g1.record()
            barrier()
            ms_graph = g0.elapsed_time(g1)
    t = torch.tensor([ms, ms_e2e, ms_graph], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.destroy_process_group()
    if rank != 0:
        return
    ms_eager, ms_e2e, ms = float(t[0]), float(t[1]), float(t[2])
    total_f = sum(v[1] for v in prof.values())
    kernel_ms = sum(v[0] for v in prof.values())
    top = sorted(prof.items(), key=lambda kv: -kv[1][0])[:14]
    print(json.dumps({"metric": f"{'frames' if args.workload == 'ssf2020' else 'pairs'}/s ({args.workload})", "value": units * world * args.steps / (ms * 1e-3),
                      "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                      "launch_mode": ("eager launches (graph capture failed: " + graph_error + ")") if graph_error else (("one CUDA graph per optimisation step (mmcodec.GraphedTrainStep)" if world == 1 else "mmcodec.GraphedTrainStep: graph 1 = forward + backward, eager bucketed NCCL all-reduce, graph 2 = clip + Adam x2") if args.workload in ("mm-train", "master-train") else "one CUDA graph per step (mmcodec.GraphedForward)"), "ms_per_step_eager": ms_eager / args.steps,
                      "sum_of_kernel_ms": kernel_ms,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": OTHER_WORKLOADS[args.workload], "units_per_gpu": units, "weights": "random init",
                                 "parallelism": f"data parallel x{world}" + (", bucketed NCCL all-reduce of fp32 gradients" if args.workload in ("mm-train", "master-train") else ", no collective")},
                      "gpu_launches": launches,
                      "e2e": {"value": units * world * args.steps / (ms_e2e * 1e-3), "unit": "img/s", "ms_per_step": ms_e2e / args.steps,
                              "h2d_bytes_per_step": sum(t_.numel() * 4 for t_ in host), "d2h_bytes_per_step": sum(t_.numel() * t_.element_size() for t_ in pinned_out)},
                      "roofline": {"bound": "tensor", "step_tflops": total_f / (ms / args.steps * 1e-3) / 1e12, "conv_flops_per_step": total_f,
                                   "top_kernels_total_ms_x_launches": {k: [round(v[0], 4), v[2]] for k, v in top}}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="images per GPU per step (default: the workload's)")
    ap.add_argument("--workload", default="hyperprior", choices=sorted(WORKLOADS) + sorted(OTHER_WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--micro-batch", type=int, default=8, help="images per pipelined micro-batch on the host-buffer path")
    ap.add_argument("--no-train-record", action="store_true", help="skip the training-step sub-record of the default line")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"],
                    help='arithmetic of the transform stacks: bf16 operands (default) or mmcodec.precision("fp32") (three-term bf16 split, ~1e-5)')
    args = ap.parse_args()

    if args.workload in OTHER_WORKLOADS:
        return bench_other(args)
    global ARCH, QUALITY, H, W, WORKLOAD
    ARCH, QUALITY, H, W, default_batch, call, WORKLOAD = WORKLOADS[args.workload]
    if args.batch is None:
        args.batch = default_batch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        b = 2
        v, ms, cores, _ = cpu_reference_throughput(b, max(1, args.steps), max(1, min(args.warmup, 1)))
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": bench_config(args.batch, max(1, args.gpus)),
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{b}-image batches of the same model/resolution, torch CPU ops, {cores} threads"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import mmcodec
    from mmcodec import ops

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_nccl_quietly(dist, dev)

    numa = bind_to_gpu_numa_node(local_rank)
    mmcodec.precision.set(args.precision)
    B = args.batch
    torch.manual_seed(0)
    net = mmcodec.build_model(ARCH, QUALITY).eval()
    net.update()
    net = net.to(dev)
    # every rank draws the global batch stream and keeps its own shard (independent images)
    lo, hi = shard_range(B * world, rank, world)
    gen = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.rand(B, 3, H, W, generator=gen).pin_memory()
    x = x_host.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if call == "forward":
        step_fn = net
    elif call == "symbols":
        step_fn = net.symbols_and_indexes
    else:
        step_fn = net.compress

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            out = step_fn(x)
        barrier()
        ops.reset_launch_count()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            out = step_fn(x)
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        launches = ops.launch_count()
        bpp = net.bpp(out) if call == "forward" else None

        # ---- end to end: the host-buffer API (mmcodec.HostPipeline): pinned host images in; H2D / kernels / D2H of consecutive
        #      micro-batches overlap.  Headline e2e = what the reference's evaluation loop keeps of a forward (per-image bpp and
        #      MSE -> PSNR, eval_model/__main__t.py:151-173), reduced on the device inside the same CUDA graph and read back
        #      (8 bytes per image); `e2e_full_outputs` = the same call returning x_hat + likelihoods to pinned host memory. ----
        e2e_full = e2e_u8 = device_coder_rec = None
        if call == "forward":
            def time_pipe(pipe, x_in=None):
                x_in = x_host if x_in is None else x_in
                pipe(x_in)
                torch.cuda.synchronize()
                for _ in range(2):
                    pipe(x_in)
                barrier()
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for _ in range(args.steps):
                    r_ = pipe(x_in)
                f1.record()
                barrier()
                return f0.elapsed_time(f1), r_
            ms_full, res = time_pipe(mmcodec.HostPipeline(net, micro_batch=args.micro_batch))
            full_bpp = float(sum(torch.log(l).sum() for l in res["likelihoods"].values()) / (-math.log(2) * B * H * W))
            full_mse = float(((res["x_hat"] - x_host) ** 2).mean())
            e2e_full = {"ms_per_step": ms_full / args.steps, "d2h_bytes_per_step": (res["x_hat"].numel() + sum(l.numel() for l in res["likelihoods"].values())) * 4,
                        "bpp": full_bpp, "mse": full_mse, "api": f"mmcodec.HostPipeline(net, micro_batch={args.micro_batch})(x_pinned) -> x_hat, likelihoods"}
            del res
            ms_e2e, res = time_pipe(mmcodec.HostPipeline(net, micro_batch=args.micro_batch, outputs="metrics"))
            e2e_bpp = float(res["bpp"].mean())
            e2e_mse = float(res["mse"].mean())
            if abs(e2e_bpp - full_bpp) > 1e-3 * full_bpp or abs(e2e_mse - full_mse) > 1e-3 * full_mse:
                raise SystemExit(f"metrics-mode e2e disagrees with the full-output e2e: bpp {e2e_bpp} vs {full_bpp}, mse {e2e_mse} vs {full_mse}")
            d2h = (res["bpp"].numel() + res["mse"].numel()) * 4
            e2e_api = f'mmcodec.HostPipeline(net, micro_batch={args.micro_batch}, outputs="metrics")(x_pinned) -> per-image bpp, mse'
            # for the record (NOT the headline: different input data): the same call fed with 8-bit images, as an image loader
            # decodes them -- one byte per sample over PCIe, ToTensor's /255 on the device
            x_u8 = (x_host * 255).round().to(torch.uint8).pin_memory()
            ms_u8, res_u8 = time_pipe(mmcodec.HostPipeline(net, micro_batch=args.micro_batch, outputs="metrics"), x_u8)
            e2e_u8 = {"ms_per_step": ms_u8 / args.steps, "h2d_bytes_per_step": x_u8.numel(), "d2h_bytes_per_step": d2h,
                      "bpp": float(res_u8["bpp"].mean()), "input": "uint8 host images (synthetic images rounded to 8 bits), converted on the device"}
        else:
            # symbol / compress path: pinned host images in, int32 symbols+indexes (or rANS byte strings) on the host out
            x_dev = torch.empty_like(x)
            pinned = {}

            def host_step():
                x_dev.copy_(x_host, non_blocking=True)
                o = step_fn(x_dev)
                if call == "symbols":
                    res_ = {}
                    for k, v in o.items():
                        if torch.is_tensor(v):
                            if k not in pinned:       # pinned int32 result buffers, allocated once (memory format of the device tensor)
                                pinned[k] = torch.empty_like(v, device="cpu").pin_memory()
                            pinned[k].copy_(v, non_blocking=True)
                            res_[k] = pinned[k]
                    torch.cuda.current_stream().synchronize()
                    return res_
                return o
            host_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                res = host_step()
            torch.cuda.synchronize()
            ms_e2e = (time.perf_counter() - t0) * 1e3
            e2e_sync_ms = None
            ref_coder_rec = None
            if call == "compress":
                # reported baseline: the reference's own coder binary on the symbols of image 0 (one thread, as the reference runs it)
                try:
                    si = net.symbols_and_indexes(x_dev)
                    secs, nsym, same = 0.0, 0, True
                    for nm, em, k in (("y", net.gaussian_conditional, 0), ("z", net.entropy_bottleneck, 1)):
                        t_, bytes_ = reference_coder_seconds(em, si[f"{nm}_symbols"][0], si[f"{nm}_indexes"][0])
                        secs += t_
                        nsym += si[f"{nm}_symbols"][0].numel()
                        same = same and bytes_ == res["strings"][k][0]
                    ref_coder_rec = {"kind": "reference", "what": "oracle/_ref ans.RansEncoder (the reference's rans_interface.cpp), image 0, .tolist() marshalling included",
                                     "cores": 1, "value": 1.0 / secs, "unit": "img/s", "ns_per_symbol": secs / nsym * 1e9,
                                     "bytes_identical_to_net_compress": bool(same)}
                except Exception as e:      # binary not built next to the reference: say so, keep the line
                    ref_coder_rec = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
                # the same call through mmcodec.CompressPipeline: host rANS coding of batch i overlaps the GPU work of batch i + 1
                e2e_sync_ms = ms_e2e / args.steps
                pipe = mmcodec.CompressPipeline(net, depth=2)
                x_dev.copy_(x_host, non_blocking=True)
                first = pipe.submit(x_dev).result()
                if first["strings"] != res["strings"]:
                    raise SystemExit("CompressPipeline strings differ from net.compress()")
                barrier()
                t0 = time.perf_counter()
                futs = []
                for _ in range(args.steps):
                    x_dev.copy_(x_host, non_blocking=True)
                    futs.append(pipe.submit(x_dev))
                res = [f.result() for f in futs][-1]
                torch.cuda.synchronize()
                ms_e2e = (time.perf_counter() - t0) * 1e3
                pipe.close()
                # the same pipeline with the DEVICE coder (lane container, csrc/rans_device.cu): symbols never leave the GPU
                host_strings = res["strings"]
                mmcodec.set_entropy_coder(net, "ans-lanes")
                try:
                    pipe = mmcodec.CompressPipeline(net, depth=2)
                    x_dev.copy_(x_host, non_blocking=True)
                    first = pipe.submit(x_dev).result()
                    hat_dev = net.decompress(first["strings"], first["shape"])["x_hat"]
                    ms_variants = {}
                    for variant in ("main_stream_upload", "copy_stream_upload"):
                        # the host->device copy of the batch either in front of its kernels on the main stream, or on the pipeline's
                        # copy stream under the previous batch's kernels (submit() of a pinned host tensor); both are reported
                        barrier()
                        t0 = time.perf_counter()
                        futs = []
                        for _ in range(args.steps):
                            if variant == "main_stream_upload":
                                x_dev.copy_(x_host, non_blocking=True)
                                futs.append(pipe.submit(x_dev))
                            else:
                                futs.append(pipe.submit(x_host))
                        res_dev = [f.result() for f in futs][-1]
                        torch.cuda.synchronize()
                        ms_variants[variant] = (time.perf_counter() - t0) * 1e3 / args.steps
                    ms_dev = min(ms_variants.values())
                    pipe.close()
                    # coding kernels alone, event-timed on the launching stream (y and z of one batch)
                    c_ = net.symbols_and_indexes(x_dev)
                    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize()
                    ev0.record()
                    for _ in range(args.steps):
                        hy = ops.rans_encode_device_launch(c_["y_symbols"], c_["y_indexes"], *net._coder_tables(net.gaussian_conditional))
                        hz = ops.rans_encode_device_launch(c_["z_symbols"], c_["z_indexes"], *net._coder_tables(net.entropy_bottleneck))
                    ev1.record()
                    torch.cuda.synchronize()
                    ms_code = ev0.elapsed_time(ev1) / args.steps
                finally:
                    mmcodec.set_entropy_coder(net, "ans")
                hat_host = net.decompress(host_strings, res["shape"])["x_hat"]
                if not torch.equal(hat_dev, hat_host):
                    raise SystemExit("device-coder round trip differs from the host-coder round trip")
                nb_host = sum(len(s_) for ss in host_strings for s_ in ss)
                nb_dev = sum(len(s_) for ss in res_dev["strings"] for s_ in ss)
                device_coder_rec = {"value": B * world / (ms_dev * 1e-3), "unit": UNIT, "ms_per_step": ms_dev,
                                    "ms_per_step_by_upload": ms_variants, "coder_ms_per_step": ms_code, "bytes_per_step": nb_dev, "bytes_per_step_host_coder": nb_host,
                                    "rate_overhead": nb_dev / nb_host - 1.0,
                                    "lanes_y": ops.rans_lanes_default(c_["y_symbols"][0].numel()),
                                    "round_trip": "decompress(x_hat) bit-identical to the host-coder round trip",
                                    "api": 'mmcodec.set_entropy_coder(net, "ans-lanes"); mmcodec.CompressPipeline(net).submit(x_pinned) -> lane containers '
                                           "(not reference-compatible; same symbols, tables and escape scheme)"}
            e2e_bpp = None
            d2h = (sum(v.numel() * 4 for v in res.values()) if call == "symbols"
                   else sum(len(s_) for ss in res["strings"] for s_ in ss) + 4 * B * (res["shape"][0] * res["shape"][1]) * 0)
            e2e_api = ("net.symbols_and_indexes(x_pinned.to(device)) -> pinned host int32" if call == "symbols" else
                       "mmcodec.CompressPipeline(net, depth=2).submit(x_pinned.to(device)) -> rANS byte strings (byte-identical to net.compress)")
        if rank == 0:
            sampler.stop_flag.set()
            sampler.join(2)
        ceil = copy_ceilings(x_host, dev, barrier) if call == "forward" else None

        # ---- per-layer device times for the roofline (separate pass, events around each launch) ---
        prof = ops.start_profile()
        for _ in range(2):
            step_fn(x)
        torch.cuda.synchronize()
        layer_prof = ops.stop_profile(with_work=True)

    train_rec = None
    if args.workload == "hyperprior" and args.precision == "bf16" and not args.no_train_record:
        del out
        torch.cuda.empty_cache()
        try:
            train_rec = train_step_record(dev, world, dist)
        except Exception as e:       # the headline line must survive
            train_rec = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    t = torch.tensor([ms_total, ms_e2e, e2e_full["ms_per_step"] if e2e_full else 0.0, e2e_u8["ms_per_step"] if e2e_u8 else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    if ceil is not None and world > 1:
        tc = torch.tensor([ceil["h2d_gbs"], ceil["d2h_gbs"], ceil["bidir_h2d_gbs"], ceil["bidir_d2h_gbs"]], device=dev)
        tsum = tc.clone()
        dist.all_reduce(tc, op=dist.ReduceOp.MIN)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ceil = {"h2d_gbs": float(tc[0]), "d2h_gbs": float(tc[1]), "bidir_h2d_gbs": float(tc[2]), "bidir_d2h_gbs": float(tc[3]),
                "aggregate_h2d_gbs": float(tsum[0]), "aggregate_d2h_gbs": float(tsum[1]), "per_rank": "min over ranks, all ranks copying at once"}
    if e2e_full:
        e2e_full["ms_per_step"] = float(t[2])
    if e2e_u8:
        e2e_u8["ms_per_step"] = float(t[3])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    with contextlib.suppress(Exception):
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
    # dominant kernel = the launch with the largest device time in the profiled pass; achieved = its algorithmic FLOPs
    # (SURVEY.md 8d formulas, computed from the launch's descriptor in mmcodec.ops.conv_flops) / its event-timed duration
    roofline = None
    if layer_prof:
        hbm = {k: v for k, v in layer_prof.items() if k.endswith(ops.HBM_KERNEL_SUFFIXES)}
        layer_flops = {k: v for k, v in layer_prof.items() if k not in hbm}
        name, (ms, f) = max(layer_flops.items(), key=lambda kv: kv[1][0])
        achieved = f / (ms * 1e-3) / 1e12
        total_f = sum(v[1] for v in layer_flops.values())
        roofline = {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved / peak_tf,
                    "traffic": NCU_DRAM_BYTES.get(name) if (args.workload == "hyperprior" and B == 64) else None,
                    "traffic_source": NCU_SOURCE, "peak_source": peak_src,
                    "ms_per_launch": ms, "flops_per_launch": f,
                    "step_tflops": total_f / (ms_total / args.steps * 1e-3) / 1e12,
                    "layer_ms": {k: round(v[0], 4) for k, v in layer_prof.items()},
                    # the memory-bound stage (quantise / CDF indexes / likelihoods / layout): algorithmic bytes per launch
                    #  over the event-timed duration, against the measured HBM copy peak
                    "hbm_kernels": {k: {"ms": round(v[0], 4), "algorithmic_mb": round(v[1] / 1e6, 2),
                                        "gbs": round(v[1] / (v[0] * 1e-3) / 1e9, 1) if v[0] > 0 else None,
                                        "frac_of_hbm_peak": round(v[1] / (v[0] * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)), 3) if v[0] > 0 else None}
                                    for k, v in hbm.items()}}

    value = B * world * args.steps / (ms_total * 1e-3)
    e2e_value = B * world * args.steps / (ms_e2e * 1e-3)
    h2d = x_host.numel() * 4
    metric = METRIC if args.workload == "hyperprior" else f"img/s ({W}x{H} {ARCH} {call})"
    line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "fp32 (bf16 x3 operand split on the bf16 tensor cores, fp32 accumulate and activations)",
            "data": "synthetic",
            "config": bench_config(B, world),
            "bpp": bpp, "gpu_launches": launches, "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "bpp": e2e_bpp, "api": e2e_api},
            "roofline": roofline}
    if os.environ.get("MMC_BENCH_RETRIED"):
        line["retried_after_fault"] = os.environ["MMC_BENCH_RETRIED"]
    if train_rec is not None:
        line["train_step"] = train_rec
    if ceil is not None:
        # the e2e legs as fractions of the measured pure-copy ceilings: metrics mode is bound by the input copy alone, the
        # full-output call by both directions at once
        line["copy_ceiling"] = {k: (round(v, 2) if isinstance(v, float) else v) for k, v in ceil.items()}
        line["copy_ceiling"]["numa"] = numa
        line["e2e"]["h2d_gbs"] = h2d / (ms_e2e / args.steps) / 1e6
        line["e2e"]["frac_of_h2d_ceiling"] = line["e2e"]["h2d_gbs"] / ceil["h2d_gbs"]
        line["e2e"]["note"] = ("the step's result read back is the per-image metric (bpp, MSE) the reference's evaluation loop keeps; "
                               "the call that returns x_hat + likelihoods to the host is e2e_full_outputs")
        if e2e_full:
            e2e_full["d2h_gbs"] = e2e_full["d2h_bytes_per_step"] / e2e_full["ms_per_step"] / 1e6
            e2e_full["frac_of_bidir_d2h_ceiling"] = e2e_full["d2h_gbs"] / ceil["bidir_d2h_gbs"]
    if e2e_full:
        e2e_full["value"] = B * world / (e2e_full["ms_per_step"] * 1e-3)
        e2e_full["unit"] = UNIT
        line["e2e_full_outputs"] = e2e_full
    if call == "compress" and device_coder_rec:
        line["e2e_device_coder"] = device_coder_rec
    if call == "compress" and ref_coder_rec:
        line["reference_coder_baseline"] = ref_coder_rec
    if call == "compress" and e2e_sync_ms:
        line["e2e_sync_compress"] = {"value": B * world / (e2e_sync_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_sync_ms,
                                     "api": "net.compress(x_pinned.to(device)) called in a loop (GPU stage and host coding back to back)"}
    if e2e_u8:
        e2e_u8["value"] = B * world / (e2e_u8["ms_per_step"] * 1e-3)
        e2e_u8["unit"] = UNIT
        line["e2e_uint8_input"] = e2e_u8
    if world == 1 and not args.no_cpu_baseline:
        with contextlib.redirect_stdout(io.StringIO()):
            v, ms, cores, cpu_bpp = cpu_reference_throughput(2, 3, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"3 steps of 2 images (same model/resolution) on {cores} host threads, torch CPU ops",
                                "bpp": cpu_bpp}
        # the reference's own evaluation setting is ONE thread (utils/eval_model/__main__t.py:61 torch.set_num_threads(1))
        with contextlib.redirect_stdout(io.StringIO()):
            v1, ms1, _, _ = cpu_reference_throughput(1, 2, 1, threads=1)
        line["cpu_baseline_1thread"] = {"value": v1, "unit": UNIT, "cores": 1, "kind": "port",
                                        "sample": "2 steps of 1 image, torch CPU ops, torch.set_num_threads(1) as in eval_model/__main__t.py:61"}
        if args.workload == "hyperprior":
            # the reference's GPU execution path: the same op sequence on stock PyTorch (cuDNN / ATen) on THIS B200 -- fp32 NCHW as
            # the reference runs it, TF32 allowed, and bf16 autocast + channels_last, its fastest stock configuration
            # (profiles/probe_torch_cuda.py: torch.nn.functional only, random weights; no kernel of this repository involved)
            try:
                sys.path.insert(0, os.path.join(ROOT, "profiles"))
                import probe_torch_cuda
                tg = probe_torch_cuda.measure(B, steps=5, dev=dev)
                best = min(tg["fp32_nchw_ms"], tg["tf32_nchw_ms"], tg["bf16_autocast_channels_last_ms"])
                line["torch_gpu_baseline"] = {
                    "unit": UNIT, "batch": B,
                    "fp32_nchw": B / tg["fp32_nchw_ms"] * 1e3, "tf32_nchw": B / tg["tf32_nchw_ms"] * 1e3,
                    "bf16_autocast_channels_last": B / tg["bf16_autocast_channels_last_ms"] * 1e3,
                    "value": B / best * 1e3, "speedup_device_resident": (ms_total / args.steps) and best / (ms_total / args.steps),
                    "what": "stock torch 2.11 ops (cuDNN convolutions, ATen elementwise) on the same GPU, device-resident batch, 5 timed steps after 3 warm-ups"}
            except Exception as e:  # the record must not take the headline line down
                line["torch_gpu_baseline"] = {"unavailable": repr(e)[:200]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _main_with_one_retry():
    """Single-process runs: a sticky CUDA fault (seen in ~3 of ~70 fresh-process runs of round 2 as an "illegal memory access" in the
    first forward of the synthesis stack on some boxes, not reproducible on others) kills the context, so the
    whole measurement is re-executed ONCE in a fresh process and the line carries "retried_after_fault".  Multi-rank runs are
    left to the launcher."""
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 or os.environ.get("MMC_BENCH_RETRIED"):
        return main()
    try:
        return main()
    except Exception as e:
        msg = f"{type(e).__name__}: {e}"
        if "illegal memory access" not in msg and "CUDA error" not in msg:
            raise
        sys.stderr.write(f"bench.py: CUDA fault in the first attempt ({msg[:200]}); re-executing once\n")
        sys.stderr.flush()
        os.environ["MMC_BENCH_RETRIED"] = msg[:200]
        os.execv(sys.executable, [sys.executable] + sys.argv)


if __name__ == "__main__":
    _main_with_one_retry()

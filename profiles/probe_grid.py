"""Profiling aid: time one g_a.2-class conv (128->128, 5x5 s2, batch 64 @ 256x384 in) with the persistent grid
restricted to G CTAs (env MMC_TC_GRID).  If the kernel were SM-bound, time would scale as 148/G; if it is bound by
the chip-wide L2 -> SM bandwidth, time stays flat until G gets small."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.join(ROOT, "165-learning-based-multi-modality-image-and-video-compression_b200"))
    import torch
    from mmcodec import ops, _lib as L
    dev = torch.device("cuda", 0)
    B, cin, cout, h, w = 64, 128, int(os.environ.get("PROBE_COUT", "128")), 256, 384
    x = torch.randn(B, h, w, cin, device=dev).to(torch.bfloat16)
    wt = torch.randn(cout, cin, 5, 5, device=dev) * 0.02
    b = torch.randn(cout, device=dev)
    d = ops.conv_desc(False, B, h, w, cin, cout, 5, 2, L.BF16, L.NHWC, L.BF16, L.NHWC)
    pk = ops.conv_pack_weights(d, wt)
    for _ in range(3):
        ops.conv_forward_tc(d, x, pk, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.conv_forward_tc(d, x, pk, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    flops = 2.0 * cin * cout * 25 * B * (h // 2) * (w // 2)
    print(json.dumps({"grid": os.environ.get("MMC_TC_GRID", "148"), "cout": cout, "ms": ms, "tflops": flops / ms / 1e9}))
else:
    for g in (148, 74):
        env = dict(os.environ, MMC_TC_GRID=str(g))
        print(subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True).stdout.strip())
    # operand-bandwidth check: N = 64 / 128 / 256 with and without TMA traffic
    for cout in (64, 128, 256):
        for dbg in (0, 1):
            env = dict(os.environ, PROBE_COUT=str(cout), MMC_TC_DEBUG=str(dbg))
            print("cout", cout, "debug", dbg, subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True).stdout.strip())
    # MMC_TC_DEBUG=1: TMA traffic off after priming (MMA + smem operand reads only); =2: MMAs off (TMA streaming only)
    for dbg in (1, 2):
        env = dict(os.environ, MMC_TC_DEBUG=str(dbg))
        print("debug", dbg, subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True).stdout.strip())

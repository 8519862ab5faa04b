"""Profiling aid: fused-GDN layers (g_a.0-class image-edge conv, g_s.4-class deconv + IGDN) under the MMC_TC_DEBUG modes."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.join(ROOT, "165-learning-based-multi-modality-image-and-video-compression_b200"))
    import torch
    from mmcodec import ops, _lib as L
    dev = torch.device("cuda", 0)
    B, C = 64, 128
    beta = torch.ones(C, device=dev) * 1.0
    gamma = (0.1 * torch.eye(C, device=dev) + 0.001).sqrt()
    be, ge, gb = ops.gdn_reparam(beta, gamma, 1e-3, 2 ** -18, 2 ** -36, want_bf16=True)
    res = {}
    # g_a.0: 3 -> 128, 5x5 s2 + GDN on 512x768
    x = torch.rand(B, 3, 512, 768, device=dev)
    d = ops.conv_desc(False, B, 512, 768, 3, C, 5, 2, L.BF16, L.NHWC_PAD8, L.BF16, L.NHWC, gdn=L.GDN_FORWARD)
    xp = ops.pad_to_nhwc8(x, d)
    wt = torch.randn(C, 3, 5, 5, device=dev) * 0.1
    pk = ops.conv_pack_weights(d, wt)
    bias = torch.randn(C, device=dev)
    def timeit(fn):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10
    res["g_a.0"] = timeit(lambda: ops.conv_forward_tc(d, xp, pk, bias, be, gb))
    d0 = ops.conv_desc(False, B, 512, 768, 3, C, 5, 2, L.BF16, L.NHWC_PAD8, L.BF16, L.NHWC)
    res["g_a.0 no gdn"] = timeit(lambda: ops.conv_forward_tc(d0, xp, pk, bias))
    # g_s.4: deconv 128 -> 128 + IGDN, 128x192 -> 256x384
    xi = torch.randn(B, 128, 192, C, device=dev).to(torch.bfloat16)
    d2 = ops.conv_desc(True, B, 128, 192, C, C, 5, 2, L.BF16, L.NHWC, L.BF16, L.NHWC, gdn=L.GDN_INVERSE)
    w2 = torch.randn(C, C, 5, 5, device=dev) * 0.02
    pk2 = ops.conv_pack_weights(d2, w2)
    res["g_s.4"] = timeit(lambda: ops.conv_forward_tc(d2, xi, pk2, bias, be, gb))
    d3 = ops.conv_desc(True, B, 128, 192, C, C, 5, 2, L.BF16, L.NHWC, L.BF16, L.NHWC)
    res["g_s.4 no gdn"] = timeit(lambda: ops.conv_forward_tc(d3, xi, pk2, bias))
    print(json.dumps({k: round(v, 4) for k, v in res.items()}))
else:
    for dbg in (0, 1, 2, 3):
        env = dict(os.environ, MMC_TC_DEBUG=str(dbg))
        print("debug", dbg, subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True).stdout.strip())

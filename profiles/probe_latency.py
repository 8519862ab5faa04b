"""Batch-1 latency of the hyperprior forward: eager launches vs one CUDA graph (host launch overhead vs device time)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "165-learning-based-multi-modality-image-and-video-compression_b200")]
import torch
import mmcodec
from mmcodec import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net = mmcodec.build_model("bmshj2018-hyperprior", 4).eval()
net.update(); net.to(dev)
for B in (1, 8):
    x = torch.rand(B, 3, 512, 768, device=dev)
    with torch.no_grad():
        for _ in range(5):
            net(x)
        torch.cuda.synchronize()
        ops.reset_launch_count()
        t0 = time.perf_counter()
        for _ in range(50):
            net(x)
        host = (time.perf_counter() - t0) / 50 * 1e3          # host time to ENQUEUE (no sync inside)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 50 * 1e3
        n = ops.launch_count() / 50
        g = mmcodec.GraphedForward(net, x)
        for _ in range(3):
            g(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            g(x)
        torch.cuda.synchronize()
        graph = (time.perf_counter() - t0) / 50 * 1e3
    print(f"B={B}: eager {wall:.3f} ms per forward (host enqueue {host:.3f} ms, {n:.0f} launches -> {host / n * 1e3:.1f} us per launch), graph replay {graph:.3f} ms")

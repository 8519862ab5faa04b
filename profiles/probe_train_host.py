"""How much of the training step is device time?  torch.profiler totals vs wall clock (profiles/README.md)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "165-learning-based-multi-modality-image-and-video-compression_b200")]
import torch
import mmcodec
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
torch.manual_seed(0)
net_r = mmcodec.JointAutoregressiveHierarchicalPriors_R(192, 192).eval()
net_d = mmcodec.JointAutoregressiveHierarchicalPriors_D(192, 192)
for n in (net_r, net_d):
    n.update(); n.to(dev)
x, d = torch.rand(4, 3, 512, 768, device=dev), torch.rand(4, 1, 512, 768, device=dev)
step = mmcodec.TrainStep(net_d, net_r, quality=3)
for _ in range(3):
    step(d, x)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    step(d, x)
torch.cuda.synchronize()
print("wall ms per step", (time.perf_counter() - t0) / 3 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(d, x)
    torch.cuda.synchronize()
ev = prof.key_averages()
cuda_total = sum(e.self_device_time_total for e in ev) / 1e3
print("device ms (all kernels, incl. torch's)", cuda_total)
rows = sorted(ev, key=lambda e: -e.self_device_time_total)[:18]
for e in rows:
    print(f"{e.self_device_time_total / 1e3:8.3f} ms  x{e.count:4d}  {e.key[:90]}")

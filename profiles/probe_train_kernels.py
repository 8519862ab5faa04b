"""Device-time breakdown of one training step (all kernels, ours and the library's) with torch.profiler.
Usage: python profiles/probe_train_kernels.py [mm|master]   -> prints the top kernels by total CUDA time."""
import glob
import os
import sys

import torch

sys.path.insert(0, glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "165-*"))[0])
import mmcodec  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "mm"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
gen = torch.Generator().manual_seed(1)
if which == "mm":
    guide = mmcodec.JointAutoregressiveHierarchicalPriors_R(192, 192).eval()
    net = mmcodec.JointAutoregressiveHierarchicalPriors_D(192, 192)
    x, g = torch.rand(4, 1, 512, 768, generator=gen).to(dev), torch.rand(4, 3, 512, 768, generator=gen).to(dev)
else:
    guide = mmcodec.Guided_compresser(channel=1).eval()
    net = mmcodec.Master_compresser(width=256, height=384, channel=3)
    x, g = torch.rand(4, 3, 512, 768, generator=gen).to(dev), torch.rand(4, 1, 256, 384, generator=gen).to(dev)
for n in (guide, net):
    n.update()
    n.to(dev)
step = mmcodec.TrainStep(net, guide, quality=3)
for _ in range(3):
    step(x, g)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    step(x, g)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in rows)
print(f"total device time {total / 1e3:.2f} ms over {sum(e.count for e in rows)} kernels")
for e in rows[:45]:
    print(f"{e.device_time_total / 1e3:8.3f} ms  x{e.count:4d}  {e.key[:150]}")

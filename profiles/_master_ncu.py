"""ncu target: one eval forward of the RGB-T pair (Guided_compresser + Master_compresser, 8 pairs of 3x512x768 / 1x256x384)
after two warm-up forwards; every launch of the third forward is what the launch list / the full capture look at."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "165-learning-based-multi-modality-image-and-video-compression_b200")]
import torch
import mmcodec
dev = torch.device("cuda", 0)
torch.manual_seed(0)
guide = mmcodec.Guided_compresser(channel=1).eval()
master = mmcodec.Master_compresser(width=256, height=384, channel=3).eval()
for n in (guide, master):
    n.update()
    n.to(dev)
gen = torch.Generator().manual_seed(1)
x = torch.rand(8, 3, 512, 768, generator=gen).to(dev)
t = torch.rand(8, 1, 256, 384, generator=gen).to(dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
with torch.no_grad():
    for _ in range(reps):
        og = guide(t)
        o = master(x, og["x_hat"], og["hidden"])
torch.cuda.synchronize()
print("ok", float(o["x_hat"].mean()))

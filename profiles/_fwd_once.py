"""Two eval forwards of the headline workload (bmshj2018-hyperprior q4, 64 x 768x512) for ncu captures: the first packs the weights
and warms up, the second is the one to profile (ncu: -k regex:conv_tc -s 14 -c 14).  MMC_FWD_BATCH overrides the batch."""
import glob, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, glob.glob(os.path.join(ROOT, "165-*"))[0]]
import torch, mmcodec
torch.manual_seed(0)
net = mmcodec.build_model("bmshj2018-hyperprior", 4).eval()
net.update()
net = net.cuda()
x = torch.rand(int(os.environ.get("MMC_FWD_BATCH", "64")), 3, 512, 768, device="cuda")
with torch.no_grad():
    for _ in range(2):
        out = net(x)
torch.cuda.synchronize()
print("bpp", net.bpp(out))

"""Per-layer device times of the headline forward under different MMC_TC_* environment settings (set by the caller)."""
import glob, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, glob.glob(os.path.join(ROOT, "165-*"))[0]]
import torch, mmcodec
from mmcodec import ops
torch.manual_seed(0)
net = mmcodec.build_model("bmshj2018-hyperprior", 4).eval()
net.update()
net = net.cuda()
x = torch.rand(64, 3, 512, 768, device="cuda")
with torch.no_grad():
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    ops.start_profile()
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    prof = ops.stop_profile(with_work=True)
keep = ("g_a.0", "g_a.2", "g_a.4", "g_s.0", "g_s.2", "g_s.4", "g_s.6")
print(os.environ.get("TAG", ""), {k.split("|")[0]: round(v[0], 3) for k, v in prof.items() if k.split("|")[0] in keep}, "total", round(sum(v[0] for v in prof.values()), 3))

"""Where does a micro-batch's time go?  For batch 8 / 16 / 64 of the headline model: sum of the per-layer kernel times (event pairs
around every launch), the eager forward, and the CUDA-graph replay of the same forward (+ the metrics reduction kernels)."""
import glob, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, glob.glob(os.path.join(ROOT, "165-*"))[0]]
import torch, mmcodec
from mmcodec import ops
torch.manual_seed(0)
net = mmcodec.build_model("bmshj2018-hyperprior", 4).eval()
net.update()
net = net.cuda()

def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for B in (8, 16, 64):
    x = torch.rand(B, 3, 512, 768, device="cuda")
    with torch.no_grad():
        for _ in range(3):
            net(x)
        torch.cuda.synchronize()
        ops.start_profile()
        for _ in range(3):
            net(x)
        torch.cuda.synchronize()
        prof = ops.stop_profile(with_work=True)
        layers = {k.split("|")[0]: round(v[0], 4) for k, v in prof.items()}
        eager = timed(lambda: net(x))
        g = mmcodec.GraphedForward(net, x)
        graphed = timed(lambda: g(x))
    print(f"batch={B} sum_of_layers={sum(layers.values()):.3f} eager={eager:.3f} graph={graphed:.3f} ms  per-image graph={graphed / B * 1e3:.1f} us", flush=True)
    print("   ", layers, flush=True)

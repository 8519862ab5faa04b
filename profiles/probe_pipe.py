"""Host-buffer pipeline variants of the headline workload (64 x 768x512 fp32 pinned images in, per-image metrics out):
micro-batch size x concurrent run streams.  Prints ms per 64-image step (CUDA events, 5 steps after 3 warm-ups)."""
import glob, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, glob.glob(os.path.join(ROOT, "165-*"))[0]]
import torch, mmcodec
torch.manual_seed(0)
net = mmcodec.build_model("bmshj2018-hyperprior", 4).eval()
net.update()
net = net.cuda()
x = torch.rand(64, 3, 512, 768).pin_memory()
variants = [(8, False), (8, True), (16, False), (16, True), (4, True), (32, False)]
if len(sys.argv) > 1:
    variants = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for var in variants:
    mb, conc = var[0], bool(var[1])
    slots = var[2] if len(var) > 2 else 2
    pipe = mmcodec.HostPipeline(net, micro_batch=mb, outputs="metrics", concurrent_slots=conc, n_slots=slots)
    for _ in range(3):
        r = pipe(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        r = pipe(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"micro_batch={mb} concurrent={int(conc)} slots={slots} ms_per_step={ms:.3f} img/s={64e3 / ms:.0f} bpp={float(r['bpp'].mean()):.6f}", flush=True)
    del pipe

"""Where does compress() (mbt2018-mean q6, 8 x 1088x1920) spend its time?  Phase timings on the box."""
import glob, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, glob.glob(os.path.join(ROOT, "165-*"))[0]]
import numpy as np, torch, mmcodec
from mmcodec import ops
torch.manual_seed(0)
net = mmcodec.build_model("mbt2018-mean", 6).eval()
net.update()
net = net.cuda()
x = torch.rand(8, 3, 1088, 1920, device="cuda")
def t():
    torch.cuda.synchronize(); return time.perf_counter()
with torch.no_grad():
    for it in range(3):
        t0 = t(); c = net.symbols_and_indexes(x); t1 = t()
        ys, yi = ops._stage_i32((c["y_symbols"], c["y_indexes"])); t2 = t()
        tabs = ops._coder_tables_host(*net._coder_tables(net.gaussian_conditional)); t3 = t()
        s1 = ops.rans_encode(torch.from_numpy(ys), torch.from_numpy(yi), *net._coder_tables(net.gaussian_conditional)); t4 = t()
        s2 = ops.rans_encode(c["z_symbols"], c["z_indexes"], *net._coder_tables(net.entropy_bottleneck)); t5 = t()
        r = net.compress(x); t6 = t()
        print(f"gpu symbols {1e3*(t1-t0):.1f} | stage y {1e3*(t2-t1):.1f} | tables {1e3*(t3-t2):.2f} | encode y {1e3*(t4-t3):.1f} | z total {1e3*(t5-t4):.1f} | compress() {1e3*(t6-t5):.1f} ms; y bytes {sum(map(len, s1))} nonzero {float((c['y_symbols'] != 0).float().mean()):.3f}")
    pipe = mmcodec.CompressPipeline(net, depth=2)
    for _ in range(3):
        pipe.submit(x).result()
    t0 = t()
    futs = [pipe.submit(x) for _ in range(6)]
    res = [f.result() for f in futs]
    t1 = t()
    print(f"pipeline: {1e3*(t1-t0)/6:.1f} ms per batch of 8 = {48/(t1-t0):.0f} img/s; identical {res[-1]['strings'] == r['strings']}")

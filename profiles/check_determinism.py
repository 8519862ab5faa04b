"""Debug aid: run every transform stack / model several times on the same input and report bitwise mismatches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "165-learning-based-multi-modality-image-and-video-compression_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, mmcodec
from mmcodec.transforms import run_layers
from weights import make_state_dict, make_image
dev = torch.device("cuda", 0)
for arch, cls, N, M in (("factorized", mmcodec.FactorizedPrior, 128, 192), ("hyperprior", mmcodec.ScaleHyperprior, 128, 192)):
    sd = {k: torch.from_numpy(v) for k, v in make_state_dict(arch, N, M, seed=0).items()}
    net = cls(N, M).eval(); net.update(); net.load_state_dict({**net.state_dict(), **sd}); net = net.to(dev)
    x = torch.from_numpy(make_image(2, 128, 192)).to(dev)
    with torch.no_grad():
        layers = list(net.g_a)
        cur = x
        # layer by layer through g_a
        for i in range(0, 7, 2):
            sub = layers[i:i + 2] if i < 6 else layers[i:i + 1]
            fmt_in = "nchw_f32" if i == 0 else "nhwc_bf16"
            outs = [run_layers(sub, cur, fmt_in, "nhwc_bf16" if i < 6 else "nhwc_f32") for _ in range(8)]
            torch.cuda.synchronize()
            bad = sum(int(not torch.equal(outs[0], o)) for o in outs[1:])
            print(arch, "g_a layer", i, "mismatching repeats:", bad, "maxdiff", max(float((outs[0].float() - o.float()).abs().max()) for o in outs[1:]))
            cur = outs[0]
        y = cur
        ys = [net(x)["x_hat"] for _ in range(6)]
        torch.cuda.synchronize()
        print(arch, "forward x_hat mismatching repeats:", sum(int(not torch.equal(ys[0], o)) for o in ys[1:]))

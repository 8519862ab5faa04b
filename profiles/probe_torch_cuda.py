"""What stock PyTorch (cuDNN / ATen) does for the same forward on the same B200 -- the reference's actual GPU execution path
("the real existing GPU path the kernels must beat").  Self-contained: the bmshj2018-hyperprior q4 op sequence of
compressai/models/google.py:281-295 written with torch.nn.functional on random weights (no repo code involved), fp32 NCHW as
the reference runs it, and bf16 autocast + channels_last as its fastest stock configuration."""
import math, sys, time
import torch
import torch.nn.functional as F

N, M, H, W = 128, 192, 512, 768
P, eb_m, eb_b, eb_f = {}, [], [], []


def build(dev):
    """random weights of the architecture (no checkpoint, no repo code)"""
    g = torch.Generator(device="cpu").manual_seed(0)
    def w(*s): return (torch.randn(*s, generator=g) / math.sqrt(s[1] * s[2] * s[3])).to(dev)
    P.clear()
    for i, (ci, co) in enumerate([(3, N), (N, N), (N, N), (N, M)]):
        P[f"ga{i}"] = (w(co, ci, 5, 5), torch.zeros(co, device=dev))
    for i, (ci, co) in enumerate([(M, N), (N, N), (N, N), (N, 3)]):
        P[f"gs{i}"] = (w(ci, co, 5, 5), torch.zeros(co, device=dev))
    for i in range(3):
        P[f"gdn_a{i}"] = (torch.ones(N, device=dev), 0.1 * torch.eye(N, device=dev).reshape(N, N, 1, 1))
        P[f"gdn_s{i}"] = (torch.ones(N, device=dev), 0.1 * torch.eye(N, device=dev).reshape(N, N, 1, 1))
    P["ha0"] = (w(N, M, 3, 3), torch.zeros(N, device=dev)); P["ha1"] = (w(N, N, 5, 5), torch.zeros(N, device=dev)); P["ha2"] = (w(N, N, 5, 5), torch.zeros(N, device=dev))
    P["hs0"] = (w(N, N, 5, 5), torch.zeros(N, device=dev)); P["hs1"] = (w(N, N, 5, 5), torch.zeros(N, device=dev)); P["hs2"] = (w(M, N, 3, 3), torch.zeros(M, device=dev))
    eb_m[:] = [torch.randn(N, 3, 1, device=dev), torch.randn(N, 3, 3, device=dev), torch.randn(N, 3, 3, device=dev), torch.randn(N, 3, 3, device=dev), torch.randn(N, 1, 3, device=dev)]
    eb_b[:] = [torch.randn(N, 3, 1, device=dev) for _ in range(4)] + [torch.randn(N, 1, 1, device=dev)]
    eb_f[:] = [torch.randn(N, 3, 1, device=dev) for _ in range(4)]


def gdn(x, p, inverse=False):
    beta, gamma = p
    n = F.conv2d(x * x, gamma.to(x.dtype), beta.to(x.dtype))
    return x * (torch.sqrt(n) if inverse else torch.rsqrt(n))

def eb_logits(v):
    for i in range(5):
        v = torch.matmul(F.softplus(eb_m[i]), v) + eb_b[i]
        if i < 4:
            v = v + torch.tanh(eb_f[i]) * torch.tanh(v)
    return v

def forward(x):
    y = x
    for i in range(4):
        y = F.conv2d(y, *P[f"ga{i}"], stride=2, padding=2)
        if i < 3: y = gdn(y, P[f"gdn_a{i}"])
    z = F.relu(F.conv2d(torch.abs(y), *P["ha0"], padding=1))
    z = F.relu(F.conv2d(z, *P["ha1"], stride=2, padding=2))
    z = F.conv2d(z, *P["ha2"], stride=2, padding=2)
    zf = z.float()
    v = zf.permute(1, 0, 2, 3).reshape(N, 1, -1)
    vq = torch.round(v)
    lo, up = eb_logits(vq - 0.5), eb_logits(vq + 0.5)
    sgn = -torch.sign(lo + up)
    z_lik = torch.abs(torch.sigmoid(sgn * up) - torch.sigmoid(sgn * lo)).clamp_min(1e-9)
    z_hat = vq.reshape(N, -1, z.shape[2], z.shape[3]).permute(1, 0, 2, 3).to(z.dtype)
    s = F.relu(F.conv_transpose2d(z_hat, *P["hs0"], stride=2, padding=2, output_padding=1))
    s = F.relu(F.conv_transpose2d(s, *P["hs1"], stride=2, padding=2, output_padding=1))
    s = F.relu(F.conv2d(s, *P["hs2"], padding=1)).float()
    yf = y.float()
    y_hat = torch.round(yf)
    sc = s.clamp_min(0.11)
    c = -(2 ** -0.5)
    y_lik = (0.5 * torch.erfc(c * ((0.5 - y_hat.abs()) / sc)) - 0.5 * torch.erfc(c * ((-0.5 - y_hat.abs()) / sc))).clamp_min(1e-9)
    t = y_hat.to(y.dtype)
    for i in range(4):
        t = F.conv_transpose2d(t, *P[f"gs{i}"], stride=2, padding=2, output_padding=1)
        if i < 3: t = gdn(t, P[f"gdn_s{i}"], inverse=True)
    return t, y_lik, z_lik

def timeit(fn, x, n=5):
    with torch.no_grad():
        for _ in range(3): fn(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn(x)
        e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def measure(B=64, steps=5, dev=None):
    """ms per batch of B images for the three stock configurations; restores the TF32 switches it touches"""
    dev = dev or torch.device("cuda", torch.cuda.current_device())
    build(dev)
    saved = (torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    out = {"batch": B, "steps": steps}
    try:
        torch.backends.cudnn.benchmark = True
        torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
        x = torch.rand(B, 3, H, W, device=dev)
        out["fp32_nchw_ms"] = timeit(forward, x, steps)
        torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
        out["tf32_nchw_ms"] = timeit(forward, x, steps)
        xcl = x.contiguous(memory_format=torch.channels_last)
        def fwd_bf16(t):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return forward(t)
        for k in P:
            if P[k][0].dim() == 4: P[k] = (P[k][0].contiguous(memory_format=torch.channels_last), P[k][1])
        out["bf16_autocast_channels_last_ms"] = timeit(fwd_bf16, xcl, steps)
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
        P.clear()
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    r = measure(B)
    for k in ("fp32_nchw_ms", "tf32_nchw_ms", "bf16_autocast_channels_last_ms"):
        print(f"torch {k[:-3]}: {r[k]:.2f} ms per {B} images = {B / r[k] * 1e3:.0f} img/s")

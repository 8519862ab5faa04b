"""Profiling aid: SM clock and board power while the g_a.2-class tensor-core conv runs back to back for ~1 s."""
import os, sys, subprocess, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "165-learning-based-multi-modality-image-and-video-compression_b200"))
import torch
from mmcodec import ops, _lib as L
dev = torch.device("cuda", 0)
B, cin, cout, h, w = 64, 128, int(os.environ.get("PROBE_COUT", "128")), 256, 384
x = torch.randn(B, h, w, cin, device=dev).to(torch.bfloat16)
wt = torch.randn(cout, cin, 5, 5, device=dev) * 0.02
b = torch.randn(cout, device=dev)
d = ops.conv_desc(False, B, h, w, cin, cout, 5, 2, L.BF16, L.NHWC, L.BF16, L.NHWC)
pk = ops.conv_pack_weights(d, wt)
ops.conv_forward_tc(d, x, pk, b); torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader", "-lms", "50"],
                     stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(600):
    ops.conv_forward_tc(d, x, pk, b)
e1.record(); torch.cuda.synchronize()
time.sleep(0.2)
p.terminate()
out = p.stdout.read().strip().splitlines()
print("ms/launch", e0.elapsed_time(e1) / 600)
print(" | ".join(out))

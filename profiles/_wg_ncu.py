"""ncu target: the backward kernels of one tran_conv-shaped layer (384 -> 192, 5x5, stride 1, 4 x 256 x 384)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "165-learning-based-multi-modality-image-and-video-compression_b200")]
import torch
import mmcodec
from mmcodec.layers import conv
from mmcodec.transforms import run_layers
dev = torch.device("cuda", 0)
m = conv(384, 192, kernel_size=5, stride=1).to(dev)
x = torch.randn(4, 256, 384, 384, device=dev).to(torch.bfloat16).requires_grad_(True)
for _ in range(2):
    y = run_layers([m], x, "nhwc_bf16", "nhwc_bf16")
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
print("ok")

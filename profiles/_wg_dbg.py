import sys, os
sys.path[:0]=['/root/repo','/root/repo/165-learning-based-multi-modality-image-and-video-compression_b200']
import torch, mmcodec
from mmcodec import ops
dev=torch.device('cuda',0)
x=torch.randn(2,16,24,64,device=dev).to(torch.bfloat16)
dy=torch.randn(2,8,12,128,device=dev).to(torch.bfloat16)
dw=ops.wgrad(dy,x,5,2); torch.cuda.synchronize(); print('wgrad ok', dw.abs().mean().item())
import torch.nn.functional as F
w=torch.zeros(128,64,5,5,dtype=torch.float64,requires_grad=True)
y=F.conv2d(x.double().cpu().permute(0,3,1,2),w,stride=2,padding=2); y.backward(dy.double().cpu().permute(0,3,1,2))
print('rel', float((dw.double().cpu()-w.grad).norm()/w.grad.norm()))

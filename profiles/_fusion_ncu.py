"""ncu target: the round-2 kernels that are not convolutions, at the sizes of the BASELINE configs --
the fusion layers' backward kernels (csrc/fusion_bwd.cu; cfg 4: 4 pairs of 768x512, N = 192) and the device coder
(csrc/rans_device.cu; cfg 3: 8 x 1088x1920, M = 320).  One launch of each after a warm-up launch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "165-learning-based-multi-modality-image-and-video-compression_b200")]
import torch
import mmcodec
from mmcodec import ops

dev = torch.device("cuda", 0)
gen = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=gen).bfloat16()
for rep in range(2):
    # ESA gate of the first fusion block: x 4 x 256 x 384 x 192, f = 48 channels, c1 127 x 191, pooled 41 x 62
    c1 = rnd(4, 127, 191, 48)
    v, idx = ops.maxpool_nhwc_bf16_idx(c1, 7, 3)
    ops.maxpool_nhwc_bf16_bwd(rnd(*v.shape), idx, c1.shape, 7, 3)
    ops.upsample_bilinear_bwd_bf16(rnd(4, 256, 384, 48), 41, 62)
    x, c4, g = rnd(4, 256, 384, 192), rnd(4, 256, 384, 192), rnd(4, 256, 384, 192)
    ops.sigmoid_gate_bwd_bf16(g, x, c4)
    # Spatial_aligner of the last decoder stage: token grid 4 x 128 x 192 x 96, 4 x 4 windows, 3 heads, shifted block
    tok, gt = rnd(4, 128, 192, 96), rnd(4, 128, 192, 96)
    w = torch.ones(96, device=dev)
    ops.layernorm_bwd_bf16(gt, tok, w, 1e-5, g_sum=gt)
    h = rnd(4, 128, 192, 384)
    ops.gelu_bwd_bf16(h, h)
    q, kv = rnd(4, 128, 192, 96), rnd(4, 128, 192, 192)
    table = torch.zeros(49, 3, device=dev)
    ops.window_attention(q, kv, table, 4, 2, 3, 32 ** -0.5)
    ops.window_attention_bwd(q, kv, table, gt, 4, 2, 3, 32 ** -0.5)
    # device coder: y of cfg 3 (320 x 68 x 120 symbols per image), Gaussian-conditional tables
    gc = mmcodec.GaussianConditional(None)
    gc.update_scale_table(mmcodec.models.get_scale_table())
    tabs = tuple(t.to(dev) for t in (gc._quantized_cdf, gc._cdf_length, gc._offset))
    n = 320 * 68 * 120
    cpu_gen = torch.Generator().manual_seed(1)
    idx_y = torch.randint(0, 64, (8, n), generator=cpu_gen, dtype=torch.int32)
    scale = torch.as_tensor(mmcodec.models.get_scale_table())[idx_y.long()]
    sym_y = torch.round(torch.randn(8, n, generator=cpu_gen) * scale).to(torch.int32)
    streams = ops.rans_encode_device(sym_y.to(dev), idx_y.to(dev), *tabs)
    out = ops.rans_decode_device(streams, idx_y.to(dev), *tabs)
    assert torch.equal(out.cpu(), sym_y)
torch.cuda.synchronize()
print("ok", sum(len(s) for s in streams) / 8 / n * 8, "bits per symbol")

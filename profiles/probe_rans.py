"""Host rANS coder timing on the box's CPU: C call only, cfg-3-sized streams (2.6 M symbols per image)."""
import glob, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, glob.glob(os.path.join(ROOT, "165-*"))[0]]
import numpy as np, torch, mmcodec
from mmcodec import _lib as L, ops
gc = mmcodec.GaussianConditional(None)
gc.update_scale_table(mmcodec.models.get_scale_table())
tab, lens, offs = (np.ascontiguousarray(t.numpy().astype(np.int32)) for t in (gc._quantized_cdf, gc._cdf_length, gc._offset))
g = torch.Generator().manual_seed(1)
n = 2611200
print("cpus", os.cpu_count())
for B in (1, 8, 16):
    idx = torch.randint(0, 40, (B, n), generator=g, dtype=torch.int32)
    scale = torch.as_tensor(mmcodec.models.get_scale_table())[idx.long()]
    sym = torch.round(torch.randn(B, n, generator=g) * scale).to(torch.int32).numpy()
    idx = idx.numpy()
    cap = 4 * n + 64
    out = np.empty((B, cap), dtype=np.uint8)
    nbytes = np.zeros(B, dtype=np.uint64)
    best = 1e9
    for _ in range(4):
        t = time.time()
        rc = L.lib().mmc_rans_encode_batch_host(sym.ctypes.data, idx.ctypes.data, B, n, tab.ctypes.data, tab.shape[0], tab.shape[1],
                                                lens.ctypes.data, offs.ctypes.data, out.ctypes.data, cap, nbytes.ctypes.data)
        best = min(best, time.time() - t)
    print(f"encode B={B}: {best * 1e3:.1f} ms = {best * 1e9 / n:.2f} ns/symbol/thread, {B / best:.0f} img/s, rc {rc}")
    if B == 8:
        strings = [out[b, : int(nbytes[b])].tobytes() for b in range(B)]
        t = time.time()
        d = ops.rans_decode(strings, torch.from_numpy(idx), gc._quantized_cdf, gc._cdf_length, gc._offset)
        print(f"decode B={B}: {(time.time() - t) * 1e3:.1f} ms")
        assert np.array_equal(d.numpy(), sym)

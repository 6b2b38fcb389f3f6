"""Dev: sensitivity of the two GEMMs to the batch (activation re-read traffic) at 8B shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arcticinference_b200 import _capi
lib = _capi.lib(); dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream
V, D = 128256, 4096
Wh = (torch.randn(V, D, device=dev) * 0.02).to(torch.bfloat16)
Wg = (torch.randn(4 * D, D, device=dev) * 0.02).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2] * 1e3, ts[0] * 1e3

for M in (8, 16, 32, 64, 128, 256):
    x = torch.randn(M, D, device=dev).to(torch.bfloat16)
    packed = torch.zeros(M, dtype=torch.int64, device=dev)
    gates = torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev)
    h = timeit(lambda: lib.ab200_head_argmax(x.data_ptr(), Wh.data_ptr(), M, V, D, V, D, D, packed.data_ptr(), None, 0, None, 2, None, 0, st))
    g = timeit(lambda: lib.ab200_linear(x.data_ptr(), Wg.data_ptr(), gates.data_ptr(), M, 4 * D, D, D, D, 4 * D, 2, st))
    cb = timeit(lambda: torch.nn.functional.linear(x, Wh))
    cg = timeit(lambda: torch.nn.functional.linear(x, Wg))
    print(f"M={M:4d} head {h[0]:7.1f} us (min {h[1]:6.1f}) {V*D*2/h[0]/1e3:6.0f} GB/s | gate {g[0]:6.1f} us (min {g[1]:5.1f}) {4*D*D*2/g[0]/1e3:6.0f} GB/s"
          f" | cuBLAS head {cb[0]:7.1f} gate {cg[0]:6.1f}", flush=True)
# pure read bandwidth reference: torch sum over the head weights (read-only stream)
t = timeit(lambda: Wh.view(torch.int16).sum())
print(f"torch read-only reduction over head weights: {t[0]:.1f} us -> {V*D*2/t[0]/1e3:.0f} GB/s")
a = torch.empty(1 << 29, dtype=torch.bfloat16, device=dev); b = torch.empty_like(a)
t = timeit(lambda: b.copy_(a))
print(f"torch copy 1 GiB+1 GiB: {t[0]:.1f} us -> {2*a.numel()*2/t[0]/1e3:.0f} GB/s")

import sys, os
sys.path.insert(0, '/root/repo')
import torch
from arcticinference_b200 import _capi
lib = _capi.lib(); dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(384 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, n=12):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        flush.zero_(); torch.cuda.synchronize(); torch.cuda._sleep(200000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
for (V, D) in ((128256, 4096), (201088, 2880)):
    Wh = (torch.randn(V, D, device=dev) * 0.02).to(torch.bfloat16)
    Wg = (torch.randn(4 * D, D, device=dev) * 0.02).to(torch.bfloat16)
    for M in (256, 384, 512):
        x = torch.randn(M, D, device=dev).to(torch.bfloat16)
        packed = torch.zeros(M + 1, dtype=torch.int64, device=dev); tok = torch.empty(M, dtype=torch.int64, device=dev)
        gates = torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev)
        def head():
            packed.zero_()
            _capi.check(lib.ab200_head_argmax(x.data_ptr(), Wh.data_ptr(), M, V, D, V, D, D, packed.data_ptr(), tok.data_ptr(), 1, packed[M:].data_ptr(), 2, None, 0, st))
        def gate():
            _capi.check(lib.ab200_linear(x.data_ptr(), Wg.data_ptr(), gates.data_ptr(), M, 4 * D, D, D, D, 4 * D, 2, st))
        th, tg = timeit(head), timeit(gate)
        ch = timeit(lambda: torch.nn.functional.linear(x, Wh).argmax(-1)); cg = timeit(lambda: torch.nn.functional.linear(x, Wg))
        ref = torch.nn.functional.linear(x, Wh).argmax(-1); head(); torch.cuda.synchronize()
        okg = (gates == torch.nn.functional.linear(x, Wg)).float().mean().item()
        print(f"V={V} D={D} M={M}: head {th:.1f} us ({2*M*V*D/th/1e6:.0f} TF/s) vs ATen linear+argmax {ch:.1f} | gate {tg:.1f} us ({2*M*4*D*D/tg/1e6:.0f} TF/s) vs cuBLAS {cg:.1f} | argmax agree {(tok==ref).float().mean().item():.4f} gate eq {okg:.4f}", flush=True)

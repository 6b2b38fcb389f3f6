"""ncu target (round 1b): kernels added after the first capture - CTA-pair GEMM at B = 256 (head + gate),
FP8 KV write at cfg3, fused cell at 16 384 rows (exact + faithful), MLP cell, FP8 head at B = 32."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from arcticinference_b200 import _capi, py_custom_ops
lib = _capi.lib(); dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
V, D = 128256, 4096
torch.manual_seed(0)
Wh = (torch.randn(V, D, device=dev) * 0.02).to(torch.bfloat16)
Wg = (torch.randn(4 * D, D, device=dev) * 0.02).to(torch.bfloat16)
for M in (256,):
    x = torch.randn(M, D, device=dev).to(torch.bfloat16)
    packed = torch.zeros(M + 1, dtype=torch.int64, device=dev); tok = torch.empty(M, dtype=torch.int64, device=dev)
    gates = torch.empty(M, 4 * D, dtype=torch.bfloat16, device=dev)
    for _ in range(2):
        packed.zero_()
        _capi.check(lib.ab200_head_argmax(x.data_ptr(), Wh.data_ptr(), M, V, D, V, D, D, packed.data_ptr(), tok.data_ptr(), 1,
                                          packed[M:].data_ptr(), 2, None, 0, st))
        _capi.check(lib.ab200_linear(x.data_ptr(), Wg.data_ptr(), gates.data_ptr(), M, 4 * D, D, D, D, 4 * D, 2, st))
# FP8 head, B = 32
Wq = torch.empty(V, D, dtype=torch.uint8, device=dev); sw = torch.zeros(1, dtype=torch.float32, device=dev)
_capi.check(lib.ab200_fp8_quantize_weight(Wh.data_ptr(), Wq.data_ptr(), sw.data_ptr(), V, D, D, D, st))
x = torch.randn(32, D, device=dev).to(torch.bfloat16); xq = torch.empty(32, D, dtype=torch.uint8, device=dev)
sx = torch.empty(32, dtype=torch.float32, device=dev)
_capi.check(lib.ab200_fp8_quantize_rows(x.data_ptr(), xq.data_ptr(), sx.data_ptr(), 32, D, D, D, st))
packed = torch.zeros(33, dtype=torch.int64, device=dev); tok = torch.empty(32, dtype=torch.int64, device=dev)
_capi.check(lib.ab200_head_argmax_fp8(xq.data_ptr(), sx.data_ptr(), Wq.data_ptr(), sw.data_ptr(), 32, V, D, V, D, D,
                                      packed.data_ptr(), tok.data_ptr(), 1, packed[32:].data_ptr(), st))
del Wh, Wg, Wq
torch.cuda.synchronize()
kv = bench.bench_kv_write(dev, iters=1, peak_gbs=6557.1)
print("kv", kv["value"], "fp8", kv["fp8_cache"]["value"])
cell = bench.bench_cell_hbm_regime(dev, 6557.1, iters=1)
print("cell", cell["exact"]["gbs"], cell["faithful"]["gbs"])

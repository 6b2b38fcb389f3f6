"""ncu target: the fused cell kernel in its HBM-bound regime (B rows x D=4096, bf16, emulation)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from arcticinference_b200 import _capi
lib = _capi.lib(); dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
D, V = 4096, 128256
torch.manual_seed(0)
emb = (torch.randn(V, D, device=dev) * 0.02).to(torch.bfloat16)
ln = [(torch.randn(D, device=dev) * 0.1 + (1 if i % 2 == 0 else 0)).to(torch.bfloat16) for i in range(4)]
gates = (torch.randn(B, 4 * D, device=dev) * 1.3).to(torch.bfloat16)
toks = torch.randint(0, V, (B,), device=dev)
cprev = torch.randn(B, D, device=dev).to(torch.bfloat16)
state = torch.empty(B, D, device=dev, dtype=torch.bfloat16); cell = torch.empty_like(state)
x = torch.randn(B, D, device=dev).to(torch.bfloat16); y = torch.empty_like(x)
for _ in range(3):
    for mode in (1, 2):   # exact, faithful: capture with  -k regex:sum_lstm_cell -s 2 -c 2
        _capi.check(lib.ab200_sum_lstm_cell(gates.data_ptr(), 4 * D, emb.data_ptr(), toks.data_ptr(), None, cprev.data_ptr(),
                                            ln[0].data_ptr(), ln[1].data_ptr(), ln[2].data_ptr(), ln[3].data_ptr(),
                                            state.data_ptr(), cell.data_ptr(), B, D, V, 23.07, 1e-6, 2, mode, None, None, None, None, st))
    _capi.check(lib.ab200_rms_scale(x.data_ptr(), y.data_ptr(), B, D, D, D, 1e-6, 0.70710678, 2, 1, st))
torch.cuda.synchronize()
print("ok", state.float().abs().mean().item())

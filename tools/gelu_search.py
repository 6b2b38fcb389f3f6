"""Offline search (B200) for polynomial-GELU constants that need no near-tie guard; see sigmoid_search.py."""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from arcticinference_b200 import _capi


def ulp_steps(v: float, lo: int, hi: int):
    b = np.float32(v).view(np.int32)
    return [float(np.int32(b + d).view(np.float32)) for d in range(lo, hi + 1)]


dev = torch.device("cuda:0")
c5s = ulp_steps(-4.5906913280e-01, -8, 8)
c6s = ulp_steps(-1.1511206627, -16, 16)
cm1s = ulp_steps(-1.0, -16, 16)
cands = list(itertools.product(c5s, c6s, cm1s))
ct = torch.tensor(cands, dtype=torch.float32, device=dev).contiguous()
st = torch.cuda.current_stream().cuda_stream
base = cands.index((c5s[8], c6s[16], -1.0))
for lo_bound in (-2.0, -2.5, -3.0, -3.5, -4.0):
    mism = torch.zeros(len(cands), dtype=torch.int32, device=dev)
    first = torch.full((len(cands),), -1, dtype=torch.int32, device=dev)
    _capi.check(_capi.debug_lib().ab200_debug_sigmoid_search(ct.data_ptr(), -len(cands), mism.data_ptr(), first.data_ptr(),
                                                       lo_bound, st), "search")
    torch.cuda.synchronize()
    m = mism.cpu().numpy()
    order = np.argsort(m, kind="stable")
    print(f"y > {lo_bound}: baseline mismatches = {m[base]}; zero-mismatch candidates = {(m == 0).sum()} of {len(cands)}; "
          f"histogram: {np.bincount(np.minimum(m, 12)).tolist()}")
    for i in order[:5]:
        c = cands[i]
        print(f"   mism={m[i]} c5={np.float32(c[0]).view(np.uint32):#x} c6={np.float32(c[1]).view(np.uint32):#x} "
              f"cm1={np.float32(c[2]).view(np.uint32):#x} first_bad={int(first[i]) & 0xffff:#06x}")

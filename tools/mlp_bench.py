"""Device-timed proposals of the MLP speculator (8B shape) next to the LSTM speculator."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from arcticinference_b200.mlp_speculator import ArcticMLPSpeculator
dev = torch.device("cuda:0")
V, H, D, S = (bench.CFG[k] for k in "VHDS")
for tie in (True, False):
    m = ArcticMLPSpeculator(vocab_size=V, emb_dim=H, inner_dim=D, n_predict=S, tie_weights=tie, max_num_seqs=256, device=dev)
    bench.synthetic_weights(m, 0)
    for B in (64, 256):
        g = torch.Generator(device="cpu").manual_seed(1)
        ids = torch.randint(0, V, (B,), generator=g).to(dev)
        hid = torch.randn(B, H, generator=g).to(torch.bfloat16).to(dev)
        for _ in range(10): m.generate_proposals(ids, hid, S)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): m.generate_proposals(ids, hid, S)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 200 * 1e3
        wbytes = 3 * (V * D * 2) + (H * D + 2 * D * D) * 2
        print(f"MLP speculator tie={tie} B={B}: {us:.1f} us/proposal -> {B*S/us*1e6:.0f} draft tok/s "
              f"({wbytes/us/1e3:.0f} GB/s of weights)", flush=True)
    del m

"""Opt-in deviation from the reference: FP8 head for ALL batch sizes (the reference uses it for <= 32 only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from arcticinference_b200.speculator import ArcticLSTMSpeculator
from oracle import speculator_ref as O
dev = torch.device("cuda:0")
V, H, D, S, B = (bench.CFG[k] for k in "VHDSB")
w = O.make_weights(V, H, D, torch.bfloat16, seed=0, device=dev)
ref = O.RefLSTMSpeculator(w, V, S)
for B in (64, 128):
    m = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=B, device=dev,
                             quantize_lm_head=True, fp8_head_max_batch=B)
    m.load_weights(w.items())
    agree = tot = 0
    for seed in range(8):
        ids, hid = O.make_inputs(B, V, H, torch.bfloat16, seed=500 + seed, device=dev)
        got, exp = m.generate_proposals(ids, hid, S), ref.generate_proposals(ids, hid, S)
        agree += (got == exp).sum().item(); tot += got.numel()
    for _ in range(20): m.generate_proposals(ids, hid, S)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200): m.generate_proposals(ids, hid, S)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    print(f"FP8 head for all batches, B={B}: {us:.1f} us/proposal -> {B*S/us*1e6:.0f} draft tok/s; token agreement with the bf16-head reference {agree/tot:.4f}")

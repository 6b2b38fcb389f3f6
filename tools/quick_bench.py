"""Dev: device time per proposal (graph replay) at cfg2 / other batch sizes; release library, compiled-in launch parameters."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, bench
from arcticinference_b200.speculator import ArcticLSTMSpeculator
dev = torch.device("cuda:0")
V, H, D, S = (bench.CFG[k] for k in "VHDS")
batches = [int(b) for b in (sys.argv[1] if len(sys.argv) > 1 else "64").split(",")]
model = ArcticLSTMSpeculator(vocab_size=V, input_hidden_dim=H, inner_dim=D, n_predict=S, max_num_seqs=max(batches), device=dev)
bench.synthetic_weights(model, 0)
for B in batches:
    g = torch.Generator(device="cpu").manual_seed(1)
    ids = torch.randint(0, V, (B,), generator=g).to(dev)
    hid = torch.randn(B, H, generator=g).to(torch.bfloat16).to(dev)
    for _ in range(5): out = model.generate_proposals(ids, hid, S)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): model.generate_proposals(ids, hid, S)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 200)
    print(f"B={B}: {best*1e3:.1f} us/proposal -> {B*S/best*1e3:.0f} tok/s  tokens[0]={out[0].tolist()}", flush=True)
